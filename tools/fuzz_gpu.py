#!/usr/bin/env python3
"""Randomised whole-tree parity runs on the GPU against the CPU oracle (tests-only code: imports oracle/).

    python tools/fuzz_gpu.py --seeds 40 [--first 0] [--pop-max 20000]

Every seed draws a genome kind and width, an enveloppe size and criterion, a population, hybrid rates, radiations and
compatibility classes, replays the trace through libapt_b200.so with a random subset of the engine's device-side paths
forced on for small batches (clustered streams, device-made list, device-side grouping), and compares species ids,
enveloppe entries, Stats, callback stream and (bit strings) the dump with oracle/apt_oracle.cpp.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apoget_b200 import synth  # noqa: E402


def one(seed, pop_max, allow_device_payload=True):
    import oracle
    from apoget_b200.tree import PhylogeneticTree
    rng = np.random.default_rng(90_000 + seed)
    kind = synth.F32 if rng.random() < 0.2 else synth.BITS
    words = int(rng.choice([4, 8, 16, 32, 128])) if kind == synth.BITS else int(rng.choice([16, 64]))
    K = int(rng.choice([1, 2, 3, 5, 5, 10, 10]))
    if seed >= 30_000:                          # (drawn from the same stream position: lower seeds keep their cases) other enveloppe sizes, up to the largest supported
        K = int([4, 7, 17, 32][K % 4])
    gen = dict(kind=kind, words=words, pop=int(rng.integers(300, pop_max)), gens=int(rng.integers(4, 9)), seed=seed,
               hybrid=float(rng.choice([0.0, 0.02, 0.1])), hybrid_far=float(rng.choice([0.0, 0.1, 1.0])),
               burst_levels=int(rng.integers(0, 4)), burst_fanout=int(rng.choice([3, 10, 40])),
               compat_classes=int(rng.integers(1, 4)), survive=float(rng.choice([0.0, 0.3])))
    # (dimensions added later: seeds from 10 000 on, drawn from a second generator -- a lower seed keeps the case it had,
    # two of them are regression tests)
    rng2 = np.random.default_rng(123_000 + seed)
    if seed >= 10_000:
        gen.update(founders=int(rng2.choice([1, 1, 2, 3])), parent_skew=float(rng2.choice([0.0, 0.0, 1.5])))
    if kind == synth.BITS:
        gen.update(flips=float(rng.uniform(0.3, 2.5)), sigma=float(rng.uniform(0.03, 0.12)) * words * 32, burst_bits=max(8, words * 6))
    else:
        gen.update(sigma=0.35, noise=0.02, noise_coords=4, burst_scale=1.0)
    t = synth.generate(**gen)
    cfg = dict(K=K, full_continuous=bool(rng.random() < 0.75), env_crit=int(rng.choice([0, 1, 1, 1, 1, 2, 3])))
    if rng.random() < 0.3:                      # stillborn trimming at work (step(), slot recycling)
        cfg.update(trim_period=int(rng.integers(2, 6)), trim_min_delay=int(rng.integers(1, 4)), trim_delay=1.0, trim_thr=float(rng.choice([0.3, 0.6])))
    forced = {}
    for var, p in (("APT_STREAM_CLUSTER_MIN", 0.5), ("APT_DEVLIST_MIN", 0.5), ("APT_DEVGROUP_MIN", 0.5)):
        os.environ.pop(var, None)
        if rng.random() < p:
            os.environ[var] = forced[var] = "1"
    os.environ.pop("APT_LUT_MAX_BYTES", None)
    if seed >= 20_000 and rng2.random() < 0.15:  # the compatibility table overflows somewhere in the run: generic path from there
        os.environ["APT_LUT_MAX_BYTES"] = forced["APT_LUT_MAX_BYTES"] = str(int(rng2.choice([1, 3000, 20000])))
    os.environ.pop("APT_DEVGROUP_PAGEABLE", None)
    if "APT_DEVGROUP_MIN" in forced:
        os.environ["APT_DEVGROUP_PAGEABLE"] = "1"
    path = os.path.join(tempfile.gettempdir(), "apt_fuzz.trace")
    synth.write_trace(path, t)
    o = oracle.Oracle(t.kind, t.words, **cfg)
    osid, oent, _ = o.replay(path, n_births=t.n_births)
    flags = 0
    if rng.random() < 0.2:                      # device exp() instead of the table: streams stop inside the guard band
        flags = 4
        forced["generic compatibility (flag 4)"] = "1"
    tree = PhylogeneticTree(t.kind, t.words, flags=flags, **cfg)
    to_device = None
    if allow_device_payload and rng.random() < 0.3:                      # genomes resident in HBM (apt_tree_add_genomes_dev)
        import torch

        def to_device(a):
            d = torch.from_numpy(a.view(np.int32)).cuda()
            torch.cuda.synchronize()
            return d, d.data_ptr(), torch.cuda.current_stream().cuda_stream
        forced["device-resident genomes"] = "1"
    cuts = None
    if seed >= 10_000 and rng2.random() < 0.3:  # a generation handed over in several calls of random sizes (the result may not depend on it)
        crng = np.random.default_rng(seed)
        cuts = lambda g, n: sorted(set(int(x) for x in crng.integers(1, max(2, n), size=int(crng.integers(1, 5)))))   # noqa: E731
        forced["random batch boundaries"] = "1"
    sids, slots = tree.replay(t, to_device=to_device, cuts=cuts)
    bad = []
    if not np.array_equal(sids, osid): bad.append("sids")
    if not np.array_equal(slots >= 0, oent > 0): bad.append("entries")
    if tree.stats() != o.stats(): bad.append("stats")
    if not np.array_equal(tree.events(), o.events()): bad.append("events")
    if kind == synth.BITS and tree.to_json() != o.to_json(): bad.append("dump")
    c = tree.counters()
    return dict(seed=seed, births=int(t.n_births), species=tree.stats()["next_sid"], gen={k: v for k, v in gen.items() if k != "kind"},
                kind="f32" if kind == synth.F32 else "bits", cfg=cfg, forced=sorted(forced), bad=bad,
                family_births=c["family_births"], stream_births=c["stream_births"], host_scores=c["host_scores"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--pop-max", type=int, default=20000)
    args = ap.parse_args()
    t0 = time.time()
    rows = []
    for s in range(args.first, args.first + args.seeds):
        try:
            rows.append(one(s, args.pop_max))
        except Exception as e:                   # an engine error is a failed case too
            rows.append(dict(seed=s, births=0, family_births=0, stream_births=0, bad=["exception: " + str(e)[:300]],
                             env={k: v for k, v in os.environ.items() if k.startswith("APT_")}))
    failed = [r for r in rows if r["bad"]]
    print(json.dumps({"seeds": args.seeds, "first": args.first, "births": sum(r["births"] for r in rows),
                      "family_births": sum(r["family_births"] for r in rows), "stream_births": sum(r["stream_births"] for r in rows),
                      "failed": failed, "seconds": round(time.time() - t0, 1)}))
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
