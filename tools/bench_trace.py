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
    # BASELINE.json configs at their STATED sizes: 10 k species (four 10-way radiations: 11 111 species) for configs[2]
    # (1 M births in calls of 65 536) and configs[3] (candidate mode; K = 10 -> 100 k enveloppe members), 10 M births for configs[4]
    "c3full": (dict(kind=F, words=64, pop=65_536, gens=4 + 16, seed=43, sigma=0.35, noise=0.02, noise_coords=8, hybrid=0.05,
                    burst_levels=4, burst_fanout=10, burst_scale=1.0), dict(K=10)),
    "c4full": (dict(kind=B, words=128, pop=100_000, gens=4 + 10, seed=44, flips=4, sigma=100.0, hybrid=0.05, burst_levels=4,
                    burst_fanout=10, burst_bits=700), dict(K=10)),
    "c5full": (dict(kind=B, words=8, pop=100_000, gens=100, seed=45, flips=2.0, sigma=6.0, hybrid=0.05, hybrid_far=0.1,
                    burst_levels=2, burst_fanout=10), dict(K=5)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--gens", type=int, default=0)
    ap.add_argument("--pop", type=int, default=0)
    ap.add_argument("--ref-births", type=int, default=400_000, help="births replayed through the reference (-1: the whole stream, 0: none)")
    ap.add_argument("--ref", default="unmodified", choices=["unmodified", "cref"],
                    help="cref: the reference with std::cref(nodes) in its eligibility binder (node.hpp:177) -- identical "
                         "output (tests/test_oracle.py), no O(#species) map copy per insertion: the one that can replay a whole stream")
    ap.add_argument("--check-json", action="store_true", help="also compare the final tree dump with the reference's")
    ap.add_argument("--flags", type=int, default=1)
    ap.add_argument("--pageable", action="store_true",
                    help="feed the genomes from ordinary (pageable) memory: the engine then stages them through its own page-locked "
                         "buffer.  Default: the trace's genomes sit in page-locked memory (apt_host_alloc), as a host program's do")
    args = ap.parse_args()
    gen, cfg = CONFIGS[args.config]
    gen = dict(gen)
    if args.gens:
        gen["gens"] = args.gens
    if args.pop:
        gen["pop"] = args.pop
    t = synth.generate(**gen)
    pop = gen["pop"]

    from apoget_b200.tree import PhylogeneticTree, INVALID, pinned_array
    tree = PhylogeneticTree(t.kind, t.words, flags=args.flags, **cfg)
    payload, compat = t.payload, t.compat
    if not args.pageable:
        payload = pinned_array(t.payload.shape, t.payload.dtype);  payload[...] = t.payload
        compat = pinned_array(t.compat.shape, t.compat.dtype);  compat[...] = t.compat
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
        s, _, _ = tree.add_genomes(payload[b0:b1], compat[b0:b1], *ids)
        per_gen.append(time.perf_counter() - t0)
        tm1, cn1 = tree.timers(), tree.counters()
        gen_rows.append([round(1e3 * per_gen[-1], 3)] +
                        [round(1e3 * (tm1[k] - tm0[k]), 3) for k in ("prepare_s", "device_s", "walk_s", "refresh_s")] +
                        [round(tm1["stream_kernel_ms"] - tm0["stream_kernel_ms"], 3)] +
                        [round(tm1["family_kernel_ms"] - tm0["family_kernel_ms"], 3), round(1e3 * (tm1["alloc_s"] - tm0["alloc_s"]), 3)] +
                        [cn1[k] - cn0[k] for k in ("stream_launches", "host_scores", "stream_entries")])
        sid_of[b0:b1] = s
        tree.del_genomes(sid_of[t.deaths_of(g)])
        tree.step(g + 1, sid_of[t.alive_of(g)])
    st = tree.stats()
    out = {"config": args.config, "births": t.n_births, "gens": t.n_gens, "pop": pop, "species": st["next_sid"],
           "genomes_in": "pageable memory" if args.pageable else "page-locked memory", "width": st["width"], "comparisons": st["comparisons"], "branching": st["branching"],
           "cuda_seconds": sum(per_gen), "cuda_births_per_s": t.n_births / sum(per_gen),
           "cuda_births_per_s_last_gen": pop / per_gen[-1],
           # first-use cost: growth of the engine's device / page-locked buffers inside the calls (a long simulation pays it once)
           "buffer_growth_seconds": tree.timers()["alloc_s"],
           "cuda_births_per_s_excl_buffer_growth": t.n_births / max(1e-9, sum(per_gen) - tree.timers()["alloc_s"]),
           "counters": tree.counters(), "timers": tree.timers(),
           "per_gen": {"columns": ["call_ms", "prepare_ms", "device_wait_ms", "commit_ms", "refresh_ms", "stream_kernel_ms",
                                   "family_kernel_ms", "alloc_ms", "device_passes", "host_scores", "stream_entries"], "rows": gen_rows}}

    if args.ref_births:
        import oracle
        path = os.path.join(tempfile.gettempdir(), "apt_bench_trace.trace")
        synth.write_trace(path, t)
        nref = t.n_births if args.ref_births < 0 else min(args.ref_births, t.n_births)
        cref = args.ref == "cref"
        if oracle.have_reference(cref):
            prefix = os.path.join(tempfile.gettempdir(), "apt_bench_trace_ref")
            whole = nref == t.n_births
            extra = ["--no-events", "--max-births", str(nref)] + ([] if (args.check_json and whole) else ["--no-json"])
            r = oracle.run_reference(path, prefix, cfg, cref=cref, extra=extra)
            res = oracle.read_res(prefix + ".res")
            n = res["sids"].size
            same = res["sids"] == sid_of[:n]
            out.update(ref_kind="reference (std::cref in the eligibility binder)" if cref else "reference (unmodified)", ref_births=n,
                       ref_births_per_s=r["births_per_s"], ref_last_gen_births_per_s=pop / r["gen_seconds"][-1],
                       sid_parity=bool(same.all()), sid_mismatches=int((~same).sum()))
            if whole:
                out["stats_parity"] = all(res["stats"][k] == st[k] for k in ("insertions", "comparisons", "branching", "next_sid", "width"))
                if args.check_json:
                    ours, theirs = tree.to_json(), json.load(open(prefix + ".json"))
                    if t.kind == B:
                        out["json_parity"] = ours == theirs
                    else:
                        sys.path.insert(0, os.path.join(ROOT, "tests"))
                        from conftest import json_close
                        try:
                            json_close(ours, theirs, 1e-5)
                            out["json_parity"] = True
                        except AssertionError as e:
                            out["json_parity"] = False
                            out["json_first_difference"] = str(e)[:200]
                    out["json_tolerance"] = 0.0 if t.kind == B else 1e-5
        else:
            o = oracle.Oracle(t.kind, t.words, **cfg)
            o.log_events(False)
            osid, _, sec = o.replay(path, max_births=nref)
            same = osid == sid_of[:osid.size]
            out.update(ref_kind="port", ref_births=int(osid.size), ref_births_per_s=osid.size / sec,
                       sid_parity=bool(same.all()), sid_mismatches=int((~same).sum()))
        g_ref = max(1, out["ref_births"] // pop)
        out["cuda_births_per_s_same_prefix"] = out["ref_births"] / sum(per_gen[:g_ref])
        out["speedup_same_prefix"] = out["cuda_births_per_s_same_prefix"] / out["ref_births_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
