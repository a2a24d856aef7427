"""Compiles tests/cpp/test_dropin.cpp against a given build of the C-ABI library."""
import os
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build(lib_path: str, tag: str, defines=()) -> str:
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
           "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir] + ["-D" + d for d in defines]
    subprocess.run(cmd, check=True)
    return out


REF = "/root/reference"


def build_reference() -> str | None:
    """The SAME test source against the reference's own header (+ the Tools stand-in of oracle/shim and the reference's
    four translation units, compiled where they lie).  Only where /root/reference exists; outputs under _build/."""
    if not os.path.isdir(os.path.join(REF, "src", "core")):
        return None
    out = os.path.join(HERE, "_build", "test_dropin_reference")
    src = os.path.join(HERE, "test_dropin.cpp")
    if os.path.exists(out) and os.path.getmtime(src) < os.path.getmtime(out):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json_inc = os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty")
    tus = [os.path.join(REF, "src", "core", "ptreeconfig.cpp")] + \
          [os.path.join(REF, "src", "core", "tree", f) for f in ("treetypes.cpp", "enveloppecriteria.cpp", "speciescontributors.cpp")]
    cmd = ["g++", "-std=c++17", "-O2", "-DNDEBUG", "-mpopcnt", "-DAPT_TEST_REFERENCE", "-I", os.path.join(ROOT, "oracle", "shim"),
           "-I", json_inc, "-I", os.path.join(REF, "src"), "-o", out, src] + tus
    subprocess.run(cmd, check=True)
    return out
