#!/usr/bin/env python3
"""Benchmark of the species-assignment hot path (BASELINE.json metric: births classified / s).

Workload (BASELINE.json configs[1], "C2"): 256-bit bit-string genomes, Hamming distance,
100 000 births per step against ~1 000 live species with enveloppes of K = 10 representatives.
The synthetic host simulation (apoget_b200.synth, seed 42) first radiates 1 000 species below the
root in three generations; every later generation is one STEP: 100 000 births whose parents are
drawn from the previous generation, classified by one batched addGenome
(reference: src/core/tree/phylogenetictree.hpp:268-302, 553-650, 657-738), with the reference's exact
sequential semantics.  Timed per step: the add_genomes call only (the hot path); deaths and step()
bookkeeping run between timed regions, as in oracle/ref_driver.cpp which times the same loop.

  value : births/s with each step's payloads already resident in HBM (apt_tree_add_genomes_dev, read in place)
  e2e   : births/s through apt_tree_add_genomes with pinned HOST buffers (H2D + kernels + D2H + the
          ordered commit on the host inside the timed region)
  roofline : the dominant kernel (resolve_streams: the ordered resolution on the device), timed with
          CUDA events on the engine's own stream
  cpu_baseline / --impl reference : the UNMODIFIED reference (oracle/_ref/apt_ref_driver, one
          thread -- the reference is single-threaded and the tree is sequential) on the same trace.

N > 1 (torchrun): the candidate-mode tree is one sequential stream, so ranks are independent
replicas ("islands": rank r simulates seed 42 + r); no data-path collective; weak scaling.
"""
from __future__ import annotations

import argparse
import ctypes
import gc
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from apoget_b200 import synth  # noqa: E402

EPOCH_STEPS = 5         # timed generations per simulation (the steady window, see main)
SETUP_GENS = 4          # generations 0..3: root + three 10-way radiations -> 1 000 species (+111 ancestors)
WORKLOAD = dict(kind=synth.BITS, words=8, flips=0.5, sigma=12.0, hybrid=0.05, burst_levels=3, burst_fanout=10)
TREE_CFG = dict(K=10)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu captures (profiles/)
TRAFFIC = {"score_pairs_kernel": 6.10e6, "resolve_streams_kernel": 6.15e6}
TRAFFIC_SOURCE = "ncu --set full capture of an earlier run of this workload (profiles/r02c_resolve_streams_ncu.txt: dram read 6.15 MB, written 0), not measured in this run"
METRIC = "births classified/sec (species assignment, reference-exact)"
UNIT = "births/s"


def make_trace(pop: int, steps: int, seed: int) -> synth.Trace:
    return synth.generate(pop=pop, gens=SETUP_GENS + steps, seed=seed, **WORKLOAD)


def peaks() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks and throttle reasons during the timed region (B200_PROFILING.md's clocks line).  One nvidia-smi query
    is made BETWEEN timed steps (after the first, the middle and the last one), not concurrently with them: any NVML
    query that overlapped a timed call -- a polling thread, nvidia-smi -lms, in-process NVML -- stalled the CUDA
    driver for 2-100 ms on some boxes and wrecked the step it hit."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples = index, []

    def sample(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
            f = [x.strip() for x in out.stdout.strip().split(",")]
            if len(f) >= 6:
                self.samples.append(f)
        except Exception:
            pass

    def stop(self) -> dict:
        samples = self.samples
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "reasons": reasons, "samples": len(sm),
                "how": "nvidia-smi between timed steps (first, middle, last)"}


def run_reference_arm(trace_path: str, pop: int, warmup: int, steps: int, cref: bool = False) -> dict:
    """Times the reference's own CPU implementation on the same trace (addGenome loop only)."""
    import oracle
    if oracle.have_reference(cref):
        r = oracle.run_reference(trace_path, os.path.join(tempfile.gettempdir(), "apt_bench_ref"), TREE_CFG, cref=cref,
                                 extra=["--no-json", "--no-res", "--no-events",
                                        "--max-births", str(pop * (SETUP_GENS + warmup + steps))])
        gs, kind = r["gen_seconds"], "reference"
        evals = r["comparisons"]
    else:           # the reference binary did not travel: time our CPU restatement instead
        tr = synth.read_trace(trace_path)
        o = oracle.Oracle(tr.kind, tr.words, **TREE_CFG)
        o.log_events(False)
        o.replay(trace_path, max_births=pop * (SETUP_GENS + warmup + steps))
        gs, kind = list(o.gen_seconds), "port"
        evals = o.stats()["comparisons"]
    timed = gs[SETUP_GENS + warmup:SETUP_GENS + warmup + steps]
    sec = float(sum(timed))
    return dict(value=pop * len(timed) / sec, seconds=sec, kind=kind, cores=1, gens_timed=len(timed),
                gen_seconds=[float(x) for x in timed], comparisons=int(evals))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pop", type=int, default=100_000, help="births per step (C2: 100 000)")
    ap.add_argument("--cpu-sample-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-sweep measurement")
    ap.add_argument("--dense-births", type=int, default=1_000_000, help="births of the dense sweep (configs[3]: 1 M)")
    ap.add_argument("--dense-check", type=int, default=10240, help="births of the dense sweep verified against the oracle (x all species)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                    # timing rule: W >= 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_steps = args.warmup + args.steps
    config = {"workload": "C2: 256-bit Hamming genomes, 100k births/step, ~1k species x enveloppe K=10 "
                          "(candidate mode, reference-exact sequential semantics)",
              "births_per_step": args.pop, "species": "1000 live (+111 ancestors)", "K": TREE_CFG["K"],
              "window": f"steady regime: generations 7-11 of each simulation ({EPOCH_STEPS} timed steps per simulation, seeds 42 + 1000 e); "
                        "the speciation that sets in from generation 12 on is configs[4] (tools/bench_trace.py)",
              "genome_bits": 256, "seed": 42, "l2": "payload+outputs of each step are new data (never re-read); "
              "256 MiB L2 flush between timed steps", "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
              "barriers": "in front of each simulation's timed steps + max over ranks at the end (independent replicas)"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        t0 = time.time()
        r = None
        for e in range((args.steps + EPOCH_STEPS - 1) // EPOCH_STEPS):          # the same simulations as our arm
            n_warm, n_timed = (args.warmup if e == 0 else 3), min(EPOCH_STEPS, args.steps - e * EPOCH_STEPS)
            trace = make_trace(args.pop, n_warm + n_timed, 42 + 1000 * e)
            path = os.path.join(tempfile.gettempdir(), "apt_bench_c2.trace")
            synth.write_trace(path, trace)
            re = run_reference_arm(path, args.pop, n_warm, n_timed)
            if r is None:
                r = re
            else:
                r["seconds"] += re["seconds"];  r["gens_timed"] += re["gens_timed"];  r["comparisons"] += re["comparisons"]
        r["value"] = args.pop * r["gens_timed"] / r["seconds"]
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(1, r["gens_timed"]),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 popcount + f64 compat",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"],
                                 "sample": f"{r['gens_timed']} timed steps x {args.pop} births after "
                                           f"{SETUP_GENS}+{args.warmup} untimed generations, addGenome loop only; "
                                           f"unmodified reference, -O2 -DNDEBUG, 1 thread (single-threaded code)"},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": time.time() - t0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm (B200)
    import torch
    import torch.distributed as dist
    from apoget_b200.tree import PhylogeneticTree, INVALID, FLAG_NO_EVENTS

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner / debug lines would go to stdout, in front of the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # the ranks share the host: each engine gets its share of the cores for its grouping / commit workers
        os.environ.setdefault("APT_HOST_THREADS", str(max(2, min(12, (os.cpu_count() or 8) // world))))     # measured on 8 GPUs / 32 cores: 2: 405 M, 3: 501 M, 4: 535 M, 8: 232 M births/s

    # The workload is the STEADY regime of the simulation: generations 7..11 (after 4 set-up and 3 warm-up
    # generations), where ~1 000 species take 100 000 births each step.  Left running, the drifting population
    # starts to speciate from generation 12 on -- that is BASELINE.json's configs[4], measured by
    # tools/bench_trace.py, not this metric -- so more than EPOCH_STEPS timed steps are taken from further
    # simulations (seeds 42 + 1000 e), each grown through its own untimed set-up and warm-up generations.
    pop, W = args.pop, 8
    # Ranks are independent replicas: they meet at a barrier in front of the timed steps of every simulation and in the
    # reduction behind the last one, not in front of every timed call (BENCH_STEP_BARRIER=1 does that: it lines up the
    # host-side bursts of all ranks on the shared cores -- 631 M instead of 765 M births/s on 8 GPUs / 32 cores)
    step_barrier = os.environ.get("BENCH_STEP_BARRIER", "0") == "1"
    dc, dt, dtd, n_launches, per_step, span = {}, {}, {}, [0], [], [0.0, 0.0]
    sampler = ClockSampler(local_rank)

    def run_epoch(e, n_warm, n_timed):
        trace = make_trace(args.pop, n_warm + n_timed, 42 + rank + 1000 * e)
        B = trace.n_births

        def fresh_tree():
            t = PhylogeneticTree(trace.kind, W, device=local_rank, flags=FLAG_NO_EVENTS, **TREE_CFG)
            sid_of = np.full(B, INVALID, dtype=np.uint32)
            return t, sid_of

        def ids_for(g, sid_of):
            b0, b1 = g * pop, (g + 1) * pop
            m, f = trace.mother[b0:b1], trace.father[b0:b1]
            hm, hf = m != INVALID, f != INVALID
            return [np.arange(b0 + 1, b1 + 1, dtype=np.uint32),
                    np.where(hm, m + 1, INVALID).astype(np.uint32),
                    np.where(hm, sid_of[np.where(hm, m, 0)], INVALID).astype(np.uint32),
                    np.where(hf, f + 1, INVALID).astype(np.uint32),
                    np.where(hf, sid_of[np.where(hf, f, 0)], INVALID).astype(np.uint32)]

        def finish_gen(t, g, sid_of):
            t.del_genomes(sid_of[trace.deaths_of(g)])
            t.step(g + 1, sid_of[trace.alive_of(g)])

        tree_h, sid_h = fresh_tree()       # fed from pinned host memory (e2e)
        tree_d, sid_d = fresh_tree()       # fed from device-resident payloads (value)

        # untimed setup: grow both trees through the radiation generations (same product path)
        for g in range(SETUP_GENS):
            for t, sid_of in ((tree_h, sid_h), (tree_d, sid_d)):
                b0, b1 = g * pop, (g + 1) * pop
                s, _, _ = t.add_genomes(trace.payload[b0:b1], trace.compat[b0:b1], *ids_for(g, sid_of))
                sid_of[b0:b1] = s
                finish_gen(t, g, sid_of)

        pin_payload = torch.empty((pop, W), dtype=torch.int32).pin_memory()
        pin_compat = torch.empty((pop, 3), dtype=torch.float32).pin_memory()
        pin_sid = torch.empty(pop, dtype=torch.int32).pin_memory()
        pin_slot = torch.empty(pop, dtype=torch.int32).pin_memory()
        pin_rep = torch.empty(pop, dtype=torch.int32).pin_memory()
        pin_ids = [torch.empty(pop, dtype=torch.int32).pin_memory() for _ in range(5)]      # gid, mother gid / sid, father gid / sid

        def pinned_ids(ids):
            for dst, src in zip(pin_ids, ids):
                dst.numpy()[:] = src.view(np.int32)
            return [vp(x.data_ptr()) for x in pin_ids]
        dev_payload = torch.empty((pop, W), dtype=torch.int32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        vp = ctypes.c_void_p
        stream = torch.cuda.current_stream()

        def timed_step(g, timed):
            b0, b1 = g * pop, (g + 1) * pop
            out = {}
            # --- e2e: host buffers through the public C ABI -------------------------------------------
            pin_payload.numpy()[:] = trace.payload[b0:b1].view(np.int32)
            pin_compat.numpy()[:] = trace.compat[b0:b1]
            args_h = (pop, vp(pin_payload.data_ptr()), vp(pin_compat.data_ptr()), pinned_ids(ids_for(g, sid_h)),
                      vp(pin_sid.data_ptr()), vp(pin_slot.data_ptr()), vp(pin_rep.data_ptr()))      # marshalled outside the timed region
            flush.fill_(g & 0xFF)
            if world > 1 and step_barrier:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tm0, cn0 = tree_h.timers(), tree_h.counters()
            w0 = time.perf_counter()
            e0.record(stream)
            tree_h.add_genomes_raw(*args_h)
            w1 = time.perf_counter()
            e1.record(stream)
            torch.cuda.synchronize()
            out["e2e_ms"] = e0.elapsed_time(e1)
            out["e2e_wall_ms"] = 1e3 * (w1 - w0)
            tm1, cn1 = tree_h.timers(), tree_h.counters()
            out["host"] = [round(1e3 * (tm1[k] - tm0[k]), 3) for k in ("prepare_s", "device_s", "walk_s", "refresh_s")] + \
                          [round(tm1["stream_kernel_ms"] - tm0["stream_kernel_ms"], 3), cn1["kernel_launches"] - cn0["kernel_launches"],
                           cn1["host_scores"] - cn0["host_scores"], cn1["stream_launches"] - cn0["stream_launches"]]
            sid_h[b0:b1] = pin_sid.numpy().view(np.uint32)
            finish_gen(tree_h, g, sid_h)
            # --- value: payloads already resident in HBM ------------------------------------------------
            dev_payload.copy_(pin_payload, non_blocking=True)
            args_d = (pop, vp(dev_payload.data_ptr()), vp(pin_compat.data_ptr()), vp(stream.cuda_stream),
                      pinned_ids(ids_for(g, sid_d)), vp(pin_sid.data_ptr()), vp(pin_slot.data_ptr()), vp(pin_rep.data_ptr()))
            flush.fill_((g + 1) & 0xFF)
            if world > 1 and step_barrier:
                dist.barrier()
            torch.cuda.synchronize()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tmd0 = tree_d.timers()
            d0.record(stream)
            tree_d.add_genomes_dev(*args_d)
            d1.record(stream)
            torch.cuda.synchronize()
            out["dev_ms"] = d0.elapsed_time(d1)
            tmd1 = tree_d.timers()
            out["host_dev"] = [round(1e3 * (tmd1[k] - tmd0[k]), 3) for k in ("prepare_s", "device_s", "walk_s", "refresh_s")]
            sid_d[b0:b1] = pin_sid.numpy().view(np.uint32)
            finish_gen(tree_d, g, sid_d)
            assert np.array_equal(sid_h[b0:b1], sid_d[b0:b1]), "host-fed and device-fed trees diverged"
            return out

        for w in range(n_warm):
            timed_step(SETUP_GENS + w, False)
        gc.collect()
        if world > 1:
            dist.barrier()                   # all ranks enter the timed steps of this simulation together        # nothing is left for the collector to do inside a timed call (it is disabled below, see main)

        c0, t0 = tree_h.counters(), tree_h.timers()
        cd0, td0 = tree_d.counters(), tree_d.timers()
        sp0 = tree_h.kernel_span()
        steps_here = []
        for k in range(n_timed):
            steps_here.append(timed_step(SETUP_GENS + n_warm + k, True))
            if len(per_step) + len(steps_here) in (1, args.steps // 2 + 1, args.steps):
                sampler.sample()
        c1, t1 = tree_h.counters(), tree_h.timers()
        cd1, td1 = tree_d.counters(), tree_d.timers()
        sp1 = tree_h.kernel_span()
        span[0] += sp1[0] - sp0[0];  span[1] += sp1[1] - sp0[1]
        for k in td1:
            dtd[k] = dtd.get(k, 0.0) + td1[k] - td0[k]
        for k in c1:
            dc[k] = dc.get(k, 0) + c1[k] - c0[k]
        for k in t1:
            dt[k] = dt.get(k, 0.0) + t1[k] - t0[k]
        n_launches[0] += (c1["kernel_launches"] - c0["kernel_launches"]) + (cd1["kernel_launches"] - cd0["kernel_launches"])
        per_step.extend(steps_here)
        width = tree_h.stats()["width"]
        tree_h.close();  tree_d.close()      # engines (streams, ~70 device / pinned buffers, worker threads) go now, not whenever the collector finds them
        return trace, width


    # Python's cyclic collector runs when allocation counts say so -- also inside a timed call, where freeing an old
    # simulation's engine (dozens of cudaFree / cudaFreeHost, each a device-wide synchronisation, milliseconds apiece in
    # a process that sees several GPUs) showed up as a 2-5 ms step.  Collected explicitly between the timed regions instead.
    gc.disable()
    trace, width = None, 0
    for e in range((args.steps + EPOCH_STEPS - 1) // EPOCH_STEPS):
        gc.collect()
        n_timed = min(EPOCH_STEPS, args.steps - e * EPOCH_STEPS)
        tr, width = run_epoch(e, args.warmup if e == 0 else 3, n_timed)
        if e == 0:
            trace = tr
    clocks = sampler.stop()

    e2e_ms = sum(s["e2e_ms"] for s in per_step)
    dev_ms = sum(s["dev_ms"] for s in per_step)
    if world > 1:
        tt = torch.tensor([e2e_ms, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms, dev_ms = float(tt[0]), float(tt[1])
    births = pop * args.steps * world

    # ---- roofline of the dominant kernel, timed by the engine's own CUDA events on its stream ------
    G = 4 * W
    K = TREE_CFG["K"]
    peak, peak_src = peaks()
    stream_ms, score_ms = dt["stream_kernel_ms"], dt["score_kernel_ms"]
    pairs = dc["device_pairs"] / max(1, dc["score_launches"])
    if stream_ms >= score_ms:
        kname, launches = "resolve_streams_kernel", max(1, dc["stream_launches"])
        kernel_ms = stream_ms / launches
        nb, reps, ent = dc["stream_births"] / launches, dc["device_reps_read"] / launches, dc["stream_entries"] / launches
        # algorithmic bytes per launch: every stream birth once (payload + 16 B compat meta + 4 B order
        # entry), every touched enveloppe once (rows + 16 B meta per representative + K*K*4 B matrix),
        # the write-back of the representatives that changed (row + meta + matrix), a 4-byte result word
        # per birth and one 4K-byte distance row per enveloppe entry
        alg_bytes = nb * (G + 16 + 4) + reps * (G + 16 + 4 * K) + ent * (G + 16 + 4 * K * K) + nb * 4 + ent * 4 * K
        units = nb
    else:
        kname, launches = "score_pairs_kernel", max(1, dc["score_launches"])
        kernel_ms = score_ms / launches
        reps = dc["device_reps_read"] / launches
        alg_bytes = pop * (G + 16) + reps * (G + 16) + pairs * (8 + 4 + 4 * K)
        units = pairs
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0

    # bytes the engine copies per step (engine.cpp, runChunk): genomes, the streams' order list and the three per-stream
    # tables go up (the births' compat metadata is one triple here and is filled on the device); result word, species id
    # and slot per birth, the per-stream totals and the speculatively fetched distance rows (last step's count + 25 % + 64)
    # come back; two-parent births add their pair lists / scores
    n_streams = int(dc["stream_launches"] and width)
    ent = int(dc["stream_entries"] / args.steps)
    # (id arrays in page-locked memory: the births are grouped into streams on the device -- the two species-id arrays go up
    # instead of the order list; the per-species stream lengths, the birth -> stream table and the order list come back,
    # and so does the list of the births the host has to look at: 8 bytes each, last step's count + 25 % + 64)
    listed = ent + int(pairs)
    h2d = pop * G + 2 * pop * 4 + (3 * n_streams + 1) * 4 + int(pairs) * 8
    d2h = (3 * pop * 4 + n_streams * 16 + 8 + (ent + ent // 4 + 64) * 4 * K + int(pairs) * (4 + 4 * K)
           + 2 * pop * 4 + int(width) * 4 + 24 + (listed + listed // 4 + 64) * 8)
    line = {
        "metric": METRIC, "value": births / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "ms_per_step_median": float(np.median([s["dev_ms"] for s in per_step])),
        "ms_per_step_all": [round(s["dev_ms"], 4) for s in per_step],
        "per_step_host": {"columns": ["prepare_ms", "device_wait_ms", "commit_ms", "refresh_ms"], "rows": [s["host_dev"] for s in per_step]},
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 popcount + f64 compat (bit-exact)", "data": "synthetic", "config": config,
        "e2e": {"value": births / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "ms_per_step_median": float(np.median([s["e2e_ms"] for s in per_step])),
                "ms_per_step_all": [round(s["e2e_ms"], 4) for s in per_step],
                "host_wall_ms_per_step_all": [round(s["e2e_wall_ms"], 4) for s in per_step],
                "per_step_host": {"columns": ["prepare_ms", "device_wait_ms", "commit_ms", "refresh_ms", "stream_kernel_ms",
                                              "kernel_launches", "host_scores", "device_passes"],
                                  "rows": [s["host"] for s in per_step]},
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(n_launches[0]),
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": TRAFFIC.get(kname), "traffic_source": TRAFFIC_SOURCE, "peak_source": peak_src,
                     "kernel_ms_per_launch": kernel_ms, "kernel_ms_timed_by": "CUDA events on the engine's stream, around the launch",
                     # the same launches as the kernel itself sees them (global nanosecond timer, first CTA in to last CTA out): the
                     # events also contain the host's gap between recording and launching when the stream is idle at that point
                     "kernel_span_ms_per_launch_on_device": (span[0] / span[1]) if (span[1] and kname == "resolve_streams_kernel") else None,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "births_per_launch": units, "evals_per_s_kernel": units * K / (kernel_ms * 1e-3) if kernel_ms else 0,
                     "births_per_s_kernel": units / (kernel_ms * 1e-3) if kernel_ms else 0,
                     "note": "latency-bound by construction: one CTA per species stream, the launch lasts as long as the stream with most "
                             "enveloppe entries (~8 us fixed + ~1.0 us per entry, up to 13 per step; DESIGN.md 4.1) and moves 9 MB, 1.4 us worth "
                             "of HBM time; the throughput kernel of this path is the dense sweep (dense_sweep.roofline)",
                     "other_kernels_of_the_step": "group_births_kernel (one cooperative launch on the upload stream, beside the genomes' H2D copy) lasts "
                                                  "58 us under ncu (profiles/r02g_launches_summary.txt) -- longer than this kernel -- but is PCIe-bound, not "
                                                  "HBM-bound: it reads the 0.8 MB of parent species ids in place from page-locked host memory and writes "
                                                  "0.8 MB of tables back (DESIGN.md 2.2, 4); unsettled_count/list 4 + 5 us"},
        "clocks": clocks,
        "breakdown_e2e_ms_per_step": {
            "prepare": 1e3 * dt["prepare_s"] / args.steps,
            "device_wait": 1e3 * dt["device_s"] / args.steps,
            "ordered_resolution": 1e3 * dt["walk_s"] / args.steps,
            "pool_refresh": 1e3 * dt["refresh_s"] / args.steps},
        "breakdown_resident_ms_per_step": {
            "prepare": 1e3 * dtd["prepare_s"] / args.steps, "device_wait": 1e3 * dtd["device_s"] / args.steps,
            "ordered_resolution": 1e3 * dtd["walk_s"] / args.steps, "pool_refresh": 1e3 * dtd["refresh_s"] / args.steps},
        "resolution": {k: dc[k] / args.steps for k in ("device_scores", "patched_scores", "host_scores",
                                                         "host_distance_evals", "guard_rechecks", "stream_births",
                                                         "stream_entries", "streams_taken_over")},
        "species": width, "host_cores": os.cpu_count(), "host_threads_per_tree": os.environ.get("APT_HOST_THREADS", "default (12 on >= 16 cores)"),
    }

    if not args.no_dense:
        # second headline of BASELINE.json ("genome-distance evals/sec vs ... roofline", north_star target): the dense
        # sweep of configs[3] at its FULL size -- 1 M births x 100 000 enveloppe members of 4096 bits -- on the tensor
        # cores; with N > 1 GPUs the species are sharded over the ranks, every rank sweeps all births against its shard
        # and the per-birth records are exchanged by peer-memory stores from the sweep kernel itself (NVLink), rank 0
        # also timing the unsharded sweep for the speed-up.  `check` births are verified against the oracle, each
        # against ALL species.
        import argparse as _ap
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_dense
        gc.enable()
        d = bench_dense.run(_ap.Namespace(config="c4", births=args.dense_births, species=0, iters=2, check=args.dense_check,
                                          exchange="p2p", also_single=True))
        if rank == 0:
            dense = {
                "workload": f"C4: 4096-bit Hamming genomes, {d['births']} births x 10 000 species x K=10 = 100 000 enveloppe members"
                            + (f", species sharded over {world} GPUs" if world > 1 else ""),
                "path": d["path"], "exchange": d["exchange"], "n_gpus": world,
                "ms": d["ms"], "evals_per_s": d["evals_per_s"], "births_per_s": d["births_per_s"],
                "roofline": d["tensor_roofline"] or {
                    "bound": "int_pipe (popc32)", "achieved": d["pipe_ops_per_s"], "peak": d["pipe_peak_measured"], "unit": "popc32/s",
                    "frac": d["frac_of_pipe_peak"], "int_pipe_model": d["int_pipe_model"]},
                "popc32_equivalent": {"achieved": d["pipe_ops_per_s"], "peak_measured": d["pipe_peak_measured"],
                                      "frac": d["frac_of_pipe_peak"],
                                      "note": "algorithmic popc32 (one per genome word per eval) against the measured POPC-pipe "
                                              "peak: what a popcount kernel at 100 % of the integer pipe would reach is 1.0"},
                "oracle_checked_births": d["oracle_checked_births"], "oracle_checked_pairs": d["oracle_checked_pairs"],
                "oracle_mismatches": d["oracle_mismatches"]}
            if "speedup_vs_1gpu" in d:
                dense["single_gpu_ms"] = d["single_gpu_ms"]
                dense["speedup_vs_1gpu"] = d["speedup_vs_1gpu"]
            if "per_rank_shard_sweep_ms" in d:
                dense["per_rank_shard_sweep_ms"] = d["per_rank_shard_sweep_ms"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        path = os.path.join(tempfile.gettempdir(), "apt_bench_c2.trace")
        synth.write_trace(path, trace)
        n = max(1, min(args.cpu_sample_steps, args.steps, EPOCH_STEPS))
        r = run_reference_arm(path, pop, args.warmup, n)         # `trace` = the first simulation (seed 42)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"],
                                "sample": f"first {n} timed steps ({n * pop} births) of the same trace after the same "
                                          f"{SETUP_GENS}+{args.warmup} untimed generations; addGenome loop only; "
                                          f"1 thread (the reference is single-threaded); host has {os.cpu_count()} cores"}
        rc = run_reference_arm(path, pop, args.warmup, n, cref=True)
        if rc["kind"] == "reference":
            line["cpu_baseline_cref"] = {
                "value": rc["value"], "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": "same sample; the reference with ONE token changed -- std::cref(nodes) in the eligibility binder "
                          "(src/core/tree/node.hpp:177), which otherwise copies the whole species map on every insertion; "
                          "identical output (tests/test_oracle.py).  The honest CPU figure for a speed-up"}
    if rank == 0:
        if not args.no_dense:
            line["dense_sweep"] = dense          # last: the north-star numbers survive a `tail` of the line
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
