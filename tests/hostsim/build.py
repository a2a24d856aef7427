"""Builds tests/hostsim/_build/libapt_hostsim.so: the engine's host code over a stubbed CUDA runtime."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "apoget_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libapt_hostsim.so")


def build() -> str:
    srcs = [os.path.join(CSRC, "engine.cpp"), os.path.join(CSRC, "abi.cpp"), os.path.join(HERE, "cuda_stub.cpp")]
    deps = srcs + [os.path.join(CSRC, "engine.h"), os.path.join(CSRC, "kernels.cuh"),
                   os.path.join(ROOT, "include", "apt_b200.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) < os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-mpopcnt", "-Wall", "-pthread", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", OUT] + srcs
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
