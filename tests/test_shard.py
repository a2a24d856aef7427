"""N > 1 path of the dense sweep on CPU: world_size-2 gloo processes exchange per-shard winners and
must agree with the single-pool answer (computed by the CPU oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from apoget_b200 import shard, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_winners(births, bc, reps, rc, fill, lo, hi):
    B = births.shape[0]
    best = np.full(B, -1, np.int32); score = np.full(B, -np.inf, np.float32)
    first = np.full(B, -1, np.int32); count = np.zeros(B, np.int32)
    for b in range(B):
        for s in range(lo, hi):
            if not fill[s]:
                continue
            sc = np.float32(oracle.score(synth.BITS, births[b], bc[b], reps[s, :fill[s]], rc[s, :fill[s]])[0])
            if best[b] < 0 or sc > score[b]:
                best[b], score[b] = s - lo, sc
            if sc > 0:
                count[b] += 1
                if first[b] < 0:
                    first[b] = s - lo
    return best, score, first, count


def _case():
    rng = np.random.default_rng(5)
    S, K, W, B = 11, 4, 4, 40
    reps = rng.integers(0, 2**32, size=(S, K, W), dtype=np.uint32)
    reps[:, 1:] = reps[:, :1] ^ (rng.integers(0, 2**32, size=(S, K - 1, W), dtype=np.uint32) & np.uint32(0x11))
    reps[7] = reps[2]                                    # duplicate species -> tie on the best score
    fill = rng.integers(1, K + 1, size=S).astype(np.uint32)
    fill[7] = fill[2]
    fill[4] = 0
    rc = np.zeros((S, K, 3), np.float32); rc[..., 1:] = 6.0
    births = reps[rng.integers(0, S, size=B), 0] ^ (rng.integers(0, 2**32, size=(B, W), dtype=np.uint32) & np.uint32(0x3))
    bc = np.zeros((B, 3), np.float32); bc[:, 1:] = 6.0
    return births, bc, reps, rc, fill


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    births, bc, reps, rc, fill = _case()
    lo, hi = shard.shard_range(fill.size, rank, world)
    loc = _oracle_winners(births, bc, reps, rc, fill, lo, hi)
    res = shard.combine(*[torch.from_numpy(x) for x in loc], s_lo=lo)
    out[rank] = [x.numpy().copy() for x in res]
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_winners_equal_single_pool(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    births, bc, reps, rc, fill = _case()
    ref = _oracle_winners(births, bc, reps, rc, fill, 0, fill.size)
    for r in range(world):
        for got, exp, name in zip(out[r], ref, ("best", "score", "first", "count")):
            assert np.array_equal(got, exp), f"rank {r}: {name}"
    assert (ref[0] == 2).any() and not (ref[0] == 7).any(), "ties must go to the lowest species index"


def test_shard_ranges_partition():
    for S in (1, 7, 100, 10_000):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(S, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == S and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


@pytest.mark.gpu
@pytest.mark.parametrize("config,births,species", [("c2", "8192", "301"), ("c4", "4096", "613")])
def test_peer_memory_exchange_matches_allgather_on_two_gpus(config, births, species):
    """The fused sweep + exchange (records stored into the peers' buffers over NVLink by the sweep kernel) against
    the NCCL all-gather path, on 2 GPUs (c2: popcount sweep, c4: tensor-core sweep); tools/bench_dense.py asserts that
    both give identical winners, and checks them against the oracle."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(root, "tools", "bench_dense.py"), "--config", config, "--births", births,
                        "--species", species, "--exchange", "p2p", "--check", "64"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"oracle_mismatches": 0' in r.stdout and "peer-memory" in r.stdout
