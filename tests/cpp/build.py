"""Compiles tests/cpp/test_dropin.cpp against a given build of the C-ABI library."""
import os
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build(lib_path: str, tag: str) -> str:
    out = os.path.join(HERE, "_build", f"test_dropin_{tag}")
    src = os.path.join(HERE, "test_dropin.cpp")
    deps = [src, lib_path, os.path.join(ROOT, "include", "apt", "phylogenetictree.hpp"),
            os.path.join(ROOT, "include", "apt_b200.h")]
    if os.path.exists(out) and all(os.path.getmtime(d) < os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json_inc = os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty")
    libdir, libname = os.path.split(lib_path)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", json_inc, "-o", out, src,
           "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True)
    return out
