#!/usr/bin/env python3
"""Randomised dense sweeps on the GPU against the CPU oracle (tests-only code: imports oracle/).

    python tools/fuzz_dense.py --seeds 60

Every seed draws a genome kind and width (bit strings of 32..8192 bits -- popcount tile and tensor-core sweep -- or float
vectors), an enveloppe size, a number of species and births that are no multiple of any tile, ragged fills with free
slots, 1-3 compatibility classes, continuous or vote-count scoring, and compares apt_classify_dense with
oracle.score_dense over ALL (birth, species) pairs.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apoget_b200 import synth  # noqa: E402


def one(seed):
    import oracle
    from apoget_b200.kernels import Pool
    rng = np.random.default_rng(70_000 + seed)
    kind = synth.F32 if rng.random() < 0.2 else synth.BITS
    words = int(rng.choice([1, 3, 8, 32, 33, 64, 128, 256])) if kind == synth.BITS else int(rng.choice([8, 64]))
    K = int(rng.choice([1, 2, 3, 5, 7, 10]))
    S = int(rng.integers(1, 700))
    B = int(rng.integers(1, 1500))
    full = bool(rng.random() < 0.6) or kind == synth.F32
    if kind == synth.BITS:
        base = rng.integers(0, 2**32, size=(S, 1, words), dtype=np.uint32)
        noise = rng.integers(0, 2**32, size=(S, K, words), dtype=np.uint32)
        for _ in range(5):
            noise &= rng.integers(0, 2**32, size=(S, K, words), dtype=np.uint32)
        reps = np.ascontiguousarray(base ^ noise)
        sigma = 0.04 * words * 32
    else:
        base = rng.standard_normal((S, 1, words)).astype(np.float32)
        reps = np.ascontiguousarray((base + 0.05 * rng.standard_normal((S, K, words)).astype(np.float32)).view(np.uint32))
        sigma = 0.4
    fill = rng.integers(0 if rng.random() < 0.5 else K, K + 1, size=S).astype(np.uint32)      # ragged (with free slots) or all full
    ncls = int(rng.integers(1, 4))
    cls_of = rng.integers(0, ncls, size=(S, K))
    rc = np.zeros((S, K, 3), dtype=np.float32)
    rc[..., 1] = sigma * (1.0 + 0.25 * cls_of)
    rc[..., 2] = sigma * (1.25 + 0.25 * cls_of)
    src = rng.integers(0, S, size=B)
    births = reps[src, rng.integers(0, K, size=B)].copy()
    if kind == synth.BITS:
        births[np.arange(B), rng.integers(0, words, size=B)] ^= rng.integers(0, 2**32, size=B, dtype=np.uint32) & np.uint32(0x01010101)
    else:
        births = (births.view(np.float32) + 0.03 * rng.standard_normal((B, words)).astype(np.float32)).view(np.uint32)
    bcls = rng.integers(0, ncls, size=B)
    bc = np.zeros((B, 3), dtype=np.float32)
    bc[:, 1] = sigma * (1.0 + 0.25 * bcls)
    bc[:, 2] = sigma * (1.25 + 0.25 * bcls)
    pool = Pool(kind, words, reps, rc, fill, K=K, full_continuous=full)
    best, score, first, count = pool.classify_dense(births, bc)
    tc = pool.uses_tensor_cores()
    pool.close()
    ob, os_, of, oc = oracle.score_dense(kind, births, bc, reps, rc, fill, full_continuous=full)
    if kind == synth.BITS:
        ok = (best == ob) & (score == os_) & (first == of) & (count == oc)
    else:
        ok = np.abs(score - os_) <= 1e-5
    return dict(seed=seed, kind="f32" if kind == synth.F32 else "bits", words=words, K=K, S=S, B=B, full=full, classes=ncls, tc=bool(tc),
                pairs=int(B) * int(S), bad=int((~ok).sum()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=60)
    ap.add_argument("--first", type=int, default=0)
    args = ap.parse_args()
    t0 = time.time()
    rows = []
    for s in range(args.first, args.first + args.seeds):
        try:
            rows.append(one(s))
        except Exception as e:
            rows.append(dict(seed=s, pairs=0, tc=False, bad=-1, error=str(e)[:300]))
    failed = [r for r in rows if r["bad"]]
    print(json.dumps({"seeds": args.seeds, "first": args.first, "pairs": sum(r["pairs"] for r in rows),
                      "tensor_core_cases": sum(1 for r in rows if r["tc"]), "failed": failed, "seconds": round(time.time() - t0, 1)}))
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
