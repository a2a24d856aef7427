#!/usr/bin/env python3
"""Replays a synthetic host simulation through the CUDA tree and (bounded) through the reference.

    python tools/bench_trace.py --config c5 [--gens G] [--pop P] [--ref-births N]

Configs (BASELINE.json): c2 = 256-bit/K=10 steady ~1k species; c3 = 64 x f32 / L2 / K=10;
c4 = 4096-bit / K=10 (candidate mode); c5 = speciation stress (256-bit, K=5, many births found species).
Prints one JSON line: births/s of the CUDA path (whole replay incl. deaths/step bookkeeping excluded --
add_genomes only), the reference's births/s on the first --ref-births births of the same trace, parity of
the per-birth species ids on that prefix, and the engine's counters/timers.
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

B, F = synth.BITS, synth.F32
CONFIGS = {
    "c2": (dict(kind=B, words=8, pop=100_000, gens=10, seed=42, flips=0.5, sigma=12.0, hybrid=0.05, burst_levels=3,
                burst_fanout=10), dict(K=10)),
    "c3": (dict(kind=F, words=64, pop=65_536, gens=10, seed=43, sigma=0.35, noise=0.02, noise_coords=8, hybrid=0.05,
                burst_levels=3, burst_fanout=10, burst_scale=1.0), dict(K=10)),
    "c4": (dict(kind=B, words=128, pop=100_000, gens=8, seed=44, flips=4, sigma=100.0, hybrid=0.05, burst_levels=3,
                burst_fanout=10, burst_bits=700), dict(K=10)),
    "c5": (dict(kind=B, words=8, pop=100_000, gens=30, seed=45, flips=2.0, sigma=6.0, hybrid=0.05, hybrid_far=0.1,
                burst_levels=2, burst_fanout=10), dict(K=5)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--gens", type=int, default=0)
    ap.add_argument("--pop", type=int, default=0)
    ap.add_argument("--ref-births", type=int, default=400_000)
    ap.add_argument("--flags", type=int, default=1)
    args = ap.parse_args()
    gen, cfg = CONFIGS[args.config]
    gen = dict(gen)
    if args.gens:
        gen["gens"] = args.gens
    if args.pop:
        gen["pop"] = args.pop
    t = synth.generate(**gen)
    pop = gen["pop"]

    from apoget_b200.tree import PhylogeneticTree, INVALID
    tree = PhylogeneticTree(t.kind, t.words, flags=args.flags, **cfg)
    sid_of = np.full(t.n_births, INVALID, dtype=np.uint32)
    per_gen, gen_rows = [], []
    for g in range(t.n_gens):
        b0, b1 = int(t.gen_birth_off[g]), int(t.gen_birth_off[g + 1])
        m, f = t.mother[b0:b1], t.father[b0:b1]
        hm, hf = m != INVALID, f != INVALID
        ids = [np.arange(b0 + 1, b1 + 1, dtype=np.uint32), np.where(hm, m + 1, INVALID).astype(np.uint32),
               np.where(hm, sid_of[np.where(hm, m, 0)], INVALID).astype(np.uint32),
               np.where(hf, f + 1, INVALID).astype(np.uint32),
               np.where(hf, sid_of[np.where(hf, f, 0)], INVALID).astype(np.uint32)]
        tm0, cn0 = tree.timers(), tree.counters()
        t0 = time.perf_counter()
        s, _, _ = tree.add_genomes(t.payload[b0:b1], t.compat[b0:b1], *ids)
        per_gen.append(time.perf_counter() - t0)
        tm1, cn1 = tree.timers(), tree.counters()
        gen_rows.append([round(1e3 * per_gen[-1], 3)] +
                        [round(1e3 * (tm1[k] - tm0[k]), 3) for k in ("prepare_s", "device_s", "walk_s", "refresh_s")] +
                        [round(tm1["stream_kernel_ms"] - tm0["stream_kernel_ms"], 3)] +
                        [cn1[k] - cn0[k] for k in ("stream_launches", "host_scores", "stream_entries")])
        sid_of[b0:b1] = s
        tree.del_genomes(sid_of[t.deaths_of(g)])
        tree.step(g + 1, sid_of[t.alive_of(g)])
    st = tree.stats()
    out = {"config": args.config, "births": t.n_births, "gens": t.n_gens, "pop": pop, "species": st["next_sid"],
           "width": st["width"], "comparisons": st["comparisons"], "branching": st["branching"],
           "cuda_seconds": sum(per_gen), "cuda_births_per_s": t.n_births / sum(per_gen),
           "cuda_births_per_s_last_gen": pop / per_gen[-1], "counters": tree.counters(), "timers": tree.timers(),
           "per_gen": {"columns": ["call_ms", "prepare_ms", "device_wait_ms", "commit_ms", "refresh_ms", "stream_kernel_ms",
                                   "device_passes", "host_scores", "stream_entries"], "rows": gen_rows}}

    if args.ref_births:
        import oracle
        path = os.path.join(tempfile.gettempdir(), "apt_bench_trace.trace")
        synth.write_trace(path, t)
        nref = min(args.ref_births, t.n_births)
        if oracle.have_reference():
            prefix = os.path.join(tempfile.gettempdir(), "apt_bench_trace_ref")
            r = oracle.run_reference(path, prefix, cfg, extra=["--no-json", "--no-events", "--max-births", str(nref)])
            res = oracle.read_res(prefix + ".res")
            n = res["sids"].size
            out.update(ref_kind="reference", ref_births=n, ref_births_per_s=r["births_per_s"],
                       ref_last_gen_births_per_s=pop / r["gen_seconds"][-1],
                       sid_parity=bool(np.array_equal(res["sids"], sid_of[:n])))
        else:
            o = oracle.Oracle(t.kind, t.words, **cfg)
            o.log_events(False)
            osid, _, sec = o.replay(path, max_births=nref)
            out.update(ref_kind="port", ref_births=int(osid.size), ref_births_per_s=osid.size / sec,
                       sid_parity=bool(np.array_equal(osid, sid_of[:osid.size])))
        g_ref = out["ref_births"] // pop
        out["cuda_births_per_s_same_prefix"] = out["ref_births"] / sum(per_gen[:g_ref])
        out["speedup_same_prefix"] = out["cuda_births_per_s_same_prefix"] / out["ref_births_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
