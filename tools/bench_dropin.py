#!/usr/bin/env python3
"""Times the C++ drop-in (include/apt/phylogenetictree.hpp) on the bench workload: the same C2 simulation as bench.py,
driven like a reference host program -- genome objects in, addGenomes(first, last, out), callbacks out -- by
tests/cpp/test_dropin.cpp.  What it adds to the engine's step: encoding every genome object into the staging arrays,
the InsertionResult array, replaying the event stream into Callbacks_t and keeping the Representative objects.

    python tools/bench_dropin.py [--pop 100000] [--steps 5]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from apoget_b200 import build as apt_build, synth  # noqa: E402
from cpp import build as cb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pop", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--profile", action="store_true", help="per-call split on stderr: encode / engine / event replay (APT_DROPIN_PROFILE)")
    args = ap.parse_args()
    warm = 3
    trace = bench.make_trace(args.pop, warm + args.steps, 42)
    path = os.path.join(tempfile.gettempdir(), "apt_dropin_c2.trace")
    synth.write_trace(path, trace)
    exe = cb.build(apt_build.build(), "cuda_prof" if args.profile else "cuda", ("APT_DROPIN_PROFILE",) if args.profile else ())
    r = subprocess.run([exe, path, os.path.join(tempfile.gettempdir(), "apt_dropin_c2"), str(bench.TREE_CFG["K"])],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stdout + r.stderr)
    if args.profile:
        sys.stderr.write(r.stderr)
    out = json.loads(r.stdout.strip().splitlines()[-1])
    ms = out["add_ms_per_generation"][bench.SETUP_GENS + warm:]
    print(json.dumps({"workload": "C2 through apt::PhylogeneticTree<GENOME,UDATA>::addGenomes (C++ drop-in, callbacks on)",
                      "births_per_step": args.pop, "ms_per_step": ms, "births_per_s": args.pop * len(ms) / (sum(ms) * 1e-3),
                      "species": out["species"]}))


if __name__ == "__main__":
    main()
