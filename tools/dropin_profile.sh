#!/bin/bash
# Phase split of the C++ drop-in on the bench workload (encode / engine / callback replay), stderr lines "dropin: ..."
set -e
python - <<'PY'
import sys, os, subprocess, sysconfig, tempfile
ROOT = os.path.dirname(os.path.abspath("bench.py"))
sys.path.insert(0, ROOT)
import bench
from apoget_b200 import synth, build as apt_build
lib = apt_build.build()
tmp = tempfile.gettempdir()
synth.write_trace(os.path.join(tmp, "c2.trace"), bench.make_trace(100000, 6, 42))
json_inc = os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty")
libdir, libname = os.path.split(lib)
exe = os.path.join(tmp, "dropin_prof")
subprocess.run(["g++", "-std=c++17", "-O2", "-DAPT_DROPIN_PROFILE", "-I", ROOT + "/include", "-I", json_inc, "-o", exe,
                ROOT + "/tests/cpp/test_dropin.cpp", "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir], check=True)
r = subprocess.run([exe, os.path.join(tmp, "c2.trace"), os.path.join(tmp, "c2out"), "10"], capture_output=True, text=True)
print(r.stderr[-1200:]); print(r.stdout[-400:])
PY
