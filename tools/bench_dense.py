#!/usr/bin/env python3
"""Dense sweep benchmark (BASELINE.json configs[3] / north_star: 1 M births x 100 k enveloppe members).

Every birth is scored against every species of a synthetic pool by classify_dense_tiled_kernel
(apt_classify_dense_dev, inputs and outputs resident in HBM); the kernel is integer-pipe bound, so the
roofline denominator is the popc32 rate measured on this GPU by apt_microbench.
With torchrun (N ranks) the species are sharded across the GPUs, every rank sweeps ALL births against
its shard, and the per-birth partial winners are combined by one NCCL all-gather.

    python tools/bench_dense.py --config c4 [--births N] [--species S]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_dense.py --config c4
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {  # name: (kind, words, K, species, births, sigma)
    "c2": (0, 8, 10, 1_000, 100_000, 12.0),
    "c2x": (0, 8, 10, 10_000, 1_000_000, 12.0),
    "c3": (1, 64, 10, 10_000, 65_536, 0.4),
    "c4": (0, 128, 10, 10_000, 1_000_000, 60.0),
}


def make_pool(rng, kind, words, S, K, sigma):
    if kind == 0:
        base = rng.integers(0, 2**32, size=(S, 1, words), dtype=np.uint32)
        noise = rng.integers(0, 2**32, size=(S, K, words), dtype=np.uint32)
        for _ in range(4 if words <= 8 else 6):     # AND of random words: ~3 % (short genomes) / ~0.8 % of the bits differ
            noise &= rng.integers(0, 2**32, size=(S, K, words), dtype=np.uint32)
        reps = base ^ noise
    else:
        base = rng.standard_normal((S, 1, words)).astype(np.float32)
        reps = (base + 0.03 * rng.standard_normal((S, K, words)).astype(np.float32)).view(np.uint32)
    compat = np.zeros((S, K, 3), dtype=np.float32)
    compat[..., 1:] = sigma
    fill = np.full(S, K, dtype=np.uint32)
    return np.ascontiguousarray(reps), compat, fill


def run(args):
    """One dense sweep measurement; returns the result dict on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from apoget_b200 import _lib
    from apoget_b200.kernels import Pool
    from apoget_b200 import shard

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    own_group = world > 1 and not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kind, words, K, S, B, sigma = CONFIGS[args.config]
    S, B = args.species or S, args.births or B
    rng = np.random.default_rng(44)
    reps, rc, fill = make_pool(rng, kind, words, S, K, sigma)
    # births: mutated copies of random representatives (so that some species accept them)
    src = rng.integers(0, S, size=B)
    if kind == 0:
        births = reps[src, rng.integers(0, K, size=B)].copy()
        births[np.arange(B), rng.integers(0, words, size=B)] ^= rng.integers(0, 2**32, size=B, dtype=np.uint32) & np.uint32(0x00010101)
    else:
        births = (reps[src, 0].view(np.float32) + 0.03 * rng.standard_normal((B, words)).astype(np.float32)).view(np.uint32)
    bc = np.zeros((B, 3), dtype=np.float32)
    bc[:, 1:] = sigma

    d_births = torch.from_numpy(births.view(np.int32)).cuda()
    d_compat = torch.from_numpy(bc).cuda()
    d_best = torch.empty(B, dtype=torch.int32, device="cuda")
    d_score = torch.empty(B, dtype=torch.float32, device="cuda")
    d_first = torch.empty(B, dtype=torch.int32, device="cuda")
    d_count = torch.empty(B, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()
    vp = ctypes.c_void_p

    # N > 1: rank 0 first sweeps the WHOLE pool alone (the other ranks wait), so that the line carries the sharded
    # sweep's speed-up over one GPU measured in the same run on the same inputs
    single_ms = None
    if world > 1 and getattr(args, "also_single", False):
        if rank == 0:
            whole = Pool(kind, words, reps, rc, fill, device=local, K=K)
            ts = []
            for it in range(3):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                whole.classify_dense_dev(B, vp(d_births.data_ptr()), vp(d_compat.data_ptr()), vp(d_best.data_ptr()),
                                         vp(d_score.data_ptr()), vp(d_first.data_ptr()), vp(d_count.data_ptr()), vp(stream.cuda_stream))
                e1.record(stream)
                torch.cuda.synchronize()
                if it:
                    ts.append(e0.elapsed_time(e1))
            single_ms = min(ts)
            whole.close()
            del whole
        dist.barrier()

    # species shard of this rank
    s_lo, s_hi = shard.shard_range(S, rank, world)
    pool = Pool(kind, words, reps[s_lo:s_hi], rc[s_lo:s_hi], fill[s_lo:s_hi], device=local, K=K)

    peaks = np.zeros(4, dtype=np.float64)
    _lib.check(_lib.lib().apt_microbench(local, peaks.ctypes.data_as(_lib.f64p)))
    cvt = np.zeros(1, dtype=np.float64)
    _lib.check(_lib.lib().apt_microbench_cvt(local, cvt.ctypes.data_as(_lib.f64p)))

    def sweep():
        pool.classify_dense_dev(B, vp(d_births.data_ptr()), vp(d_compat.data_ptr()), vp(d_best.data_ptr()),
                                vp(d_score.data_ptr()), vp(d_first.data_ptr()), vp(d_count.data_ptr()),
                                vp(stream.cuda_stream))

    def combine():
        return shard.combine(d_best, d_score, d_first, d_count, s_lo)

    p2p = None
    if world > 1 and args.exchange == "p2p":
        # fused sweep + exchange over NVLink peer memory; checked against the NCCL all-gather path first
        p2p = shard.PeerExchange(B)
        sweep()
        ref = [x.clone() for x in combine()]
        got = p2p.sweep_and_combine(pool, d_births, d_compat, s_lo)
        torch.cuda.synchronize()
        for a, b, name in zip(got, ref, ("best", "score", "first", "count")):
            assert torch.equal(a, b), f"peer-memory exchange differs from the all-gather path: {name}"
        _sweep, _combine = sweep, combine
        sweep = lambda: None                                       # noqa: E731
        combine = lambda: p2p.sweep_and_combine(pool, d_births, d_compat, s_lo)   # noqa: E731

    sweep(); combine()                                    # warm-up
    torch.cuda.synchronize()
    times = []
    for _ in range(args.iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sweep()
        res = combine()
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])

    # where the sharded time goes: every rank's sweep of its own shard alone (no exchange, no cross-rank wait), all ranks
    # sweeping at the same time as in the measurement above
    per_rank_ms = None
    if world > 1 and p2p is not None:
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _sweep()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [round(float(x[0]), 3) for x in allt]

    best, score, first, count = [x.cpu().numpy() for x in res]
    # ---- verification against the CPU oracle: `check` births x ALL species (threaded scorer of oracle/) ----
    bad, checked = 0, 0
    if rank == 0 and args.check:
        import oracle
        idx = np.sort(rng.choice(B, size=min(args.check, B), replace=False))
        checked = int(idx.size)
        ob, os_, of, oc = oracle.score_dense(kind, births[idx], bc[idx], reps, rc, fill)
        if kind == 0:
            ok = (best[idx].view(np.uint32) == ob) & (score[idx] == os_) & (first[idx].view(np.uint32) == of) \
                & (count[idx].view(np.uint32) == oc)
        else:
            ok = np.abs(score[idx] - os_) <= 1e-5
        bad = int((~ok).sum())
    evals = float(B) * S * K
    ops = evals * (words if kind == 0 else 3 * words)
    peak = peaks[0] if kind == 0 else peaks[3] * 2
    # Integer-pipe model of the bit-string tile (dense.cu): per 128-bit step of a (birth, representative) pair the
    # carry-save accumulation issues 2 POPC (16-lane pipe) and 8 ALU operations (LOP3, 64-lane pipe; the two count
    # additions are IMADs on the otherwise idle FMA pipe);
    # the slower of the two pipes bounds the sweep.  The algorithmic count (one popc32 per word) can therefore exceed
    # the measured POPC peak.
    model = None
    if kind == 0:
        steps = evals * ((words + 3) // 4)
        t_popc, t_alu = 2 * steps / peaks[0], 8 * steps / peaks[1]
        model = {"popc_per_step": 2, "alu_ops_per_step": 8, "bound": "alu (lop3)" if t_alu >= t_popc else "popc",
                 "min_ms": 1e3 * max(t_popc, t_alu) / world, "frac": max(t_popc, t_alu) / world / (ms * 1e-3)}
    tc = pool.uses_tensor_cores()
    tensor = None
    if tc:
        # tensor-core sweep: the distances are an int8 GEMM of 2 * births * members * bits operations.  Denominator:
        # twice the bf16 rate the driver measured on this pool (int8 runs at twice the bf16 rate on sm_100; nominal
        # 4.5 vs 2.25 PFLOP/s dense) -- the sustained figure for sweeps of a second, the burst one for short ones.
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                mp = json.load(f)
            bf16 = float(mp["bf16_tflops_sustained" if ms > 100 else "bf16_tflops"])
            src = "2 x measured cuBLAS bf16 (MEASURED_PEAKS.json, %s)" % ("sustained" if ms > 100 else "burst")
        except Exception:
            bf16, src = 1590.0, "2 x fallback bf16 (B200_PROFILING.md)"
        tops = 2.0 * evals * words * 32 / (ms * 1e-3) / 1e12
        tensor = {"bound": "tensor", "achieved": tops, "peak": 2 * bf16 * world, "unit": "TOP/s (int8)",
                  "frac": tops / (2 * bf16 * world), "peak_source": src}
    out = None
    if rank == 0:
        out = ({
            "bench": "dense sweep", "config": args.config, "n_gpus": world,
            "path": ("tensor cores: int8 tcgen05 GEMM, (L - a.b)/2, scoring epilogue on TMEM accumulators" if tc
                     else "integer pipes: carry-save popcount tile"),
            "exchange": ("peer-memory stores from the sweep kernel + device barrier + combine kernel" if p2p is not None
                         else ("NCCL all-gather + torch reduction" if world > 1 else "none")), "births": B, "species": S, "K": K,
            "genome_words": words, "kind": "bits/hamming" if kind == 0 else "f32/l2", "ms": ms,
            "births_per_s": B / (ms * 1e-3), "evals_per_s": evals / (ms * 1e-3),
            "pipe_ops_per_s": ops / (ms * 1e-3), "pipe": "popc32 (+xor)" if kind == 0 else "fp32 flop (sub, fma)",
            "pipe_peak_measured": peak, "frac_of_pipe_peak": ops / (ms * 1e-3) / (peak * world),
            "int_pipe_model": None if tc else model, "tensor_roofline": tensor,
            "microbench": {"popc32_per_s": peaks[0], "lop3_per_s": peaks[1], "dfma_per_s": peaks[2], "ffma_per_s": peaks[3],
                           "f2f_cvt_per_s": cvt[0]},
            "accepting_births_frac": float((count > 0).mean()), "oracle_checked_births": checked,
            "oracle_checked_pairs": checked * S, "oracle_mismatches": int(bad)})
        if single_ms is not None:
            out["single_gpu_ms"] = single_ms
            out["speedup_vs_1gpu"] = single_ms / ms
        if per_rank_ms is not None:
            out["per_rank_shard_sweep_ms"] = per_rank_ms
    if own_group:
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--births", type=int, default=0)
    ap.add_argument("--species", type=int, default=0)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--check", type=int, default=10240, help="births verified against the CPU oracle, each against ALL species")
    ap.add_argument("--also-single", action="store_true", help="N > 1: rank 0 also times the unsharded sweep (speedup_vs_1gpu)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: records stored into the peers' buffers by the sweep kernel (p2p) or NCCL all-gather")
    out = run(ap.parse_args())
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
