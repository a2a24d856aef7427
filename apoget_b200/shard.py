"""Species-enveloppe sharding of the dense sweep over several GPUs (north_star (5)).

Every rank holds a contiguous range of the species [s_lo, s_hi) in its own pool, sweeps ALL births
against it (classify_dense), and the per-birth partial winners are exchanged with ONE all-gather of
16 bytes per birth per rank; every rank then reduces them identically:

    best species  = strictly greatest score, lowest species index on ties
                    (ranks own increasing index ranges -> first rank holding the maximum)
    first accept  = lowest species index with score > 0
    accept count  = sum

Two exchanges exist on the GPUs:

  * `PeerExchange` (the product path on one NVSwitch node): the sweep kernel itself stores every
    birth's 16-byte record into ALL ranks' gather buffers through peer-mapped memory (torch symmetric
    memory hands out the peer pointers; the stores travel over NVLink) -- compute and exchange are one
    kernel, followed by a device-side barrier and our combine kernel.  No collective call, no pack or
    reduce kernels of torch's.
  * `combine` (NCCL all-gather + torch reduction): the baseline the above is measured against, and
    what the gloo CPU tests exercise.

`torch.distributed` is plumbing only (rendezvous, NCCL over NVLink, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

NONE = -1      # APT_INVALID (0xFFFFFFFF) viewed as int32


def shard_range(n_species: int, rank: int, world: int) -> tuple[int, int]:
    return rank * n_species // world, (rank + 1) * n_species // world


def pack(best: torch.Tensor, score: torch.Tensor, first: torch.Tensor, count: torch.Tensor, s_lo: int) -> torch.Tensor:
    """Local winners (int32 species indices relative to the shard) -> [B, 4] int32 record, global indices."""
    off = torch.tensor(s_lo, dtype=torch.int32, device=best.device)
    return torch.stack([torch.where(best != NONE, best + off, best), score.view(torch.int32),
                        torch.where(first != NONE, first + off, first), count], dim=1).contiguous()


def reduce_records(rec: torch.Tensor):
    """rec: [world, B, 4] int32, rank-major -> (best, score, first, count) over the whole pool."""
    best, score, first, count = rec[:, :, 0], rec[:, :, 1].view(torch.float32), rec[:, :, 2], rec[:, :, 3]
    masked = torch.where(best != NONE, score, torch.full_like(score, float("-inf")))
    win = torch.argmax(masked, dim=0, keepdim=True)              # first maximal value = lowest rank
    big = torch.iinfo(torch.int32).max
    f = torch.where(first != NONE, first, torch.full_like(first, big)).amin(dim=0)
    return (best.gather(0, win).squeeze(0), masked.gather(0, win).squeeze(0),
            torch.where(f == big, torch.full_like(f, NONE), f), count.sum(dim=0, dtype=torch.int32))


def combine(best, score, first, count, s_lo: int, group=None):
    """All-gathers the per-shard winners and returns the pool-wide winners (identical on all ranks)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return best, score, first, count
    rec = pack(best, score, first, count, s_lo)
    world = dist.get_world_size(group)
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)          # rank-major concatenation
    return reduce_records(out.view(world, rec.shape[0], rec.shape[1]))


class PeerExchange:
    """Fused sweep + exchange of the species-sharded dense sweep over NVLink peer memory.

    One symmetric buffer of [world, n_births, 4] int32 records per rank; rank r's sweep writes its
    records into slice r of EVERY rank's buffer (apt_classify_dense_shard_dev), a device-side barrier
    orders the peers' stores before the combine kernel (apt_combine_shards_dev) reads them, and a second
    barrier keeps a fast peer's next sweep from overwriting records a slow rank is still reading."""

    def __init__(self, n_births: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n = n_births
        self.records = symm_mem.empty((self.world, n_births, 4), dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        self.handle = symm_mem.rendezvous(self.records, self.group)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        dev = self.records.device
        self.best = torch.empty(n_births, dtype=torch.int32, device=dev)
        self.score = torch.empty(n_births, dtype=torch.float32, device=dev)
        self.first = torch.empty(n_births, dtype=torch.int32, device=dev)
        self.count = torch.empty(n_births, dtype=torch.int32, device=dev)

    def sweep_and_combine(self, pool, d_births: torch.Tensor, d_compat: torch.Tensor, s_lo: int):
        """pool: this rank's shard (apoget_b200.kernels.Pool); returns pool-wide (best, score, first, count)."""
        import ctypes
        from .kernels import combine_shards_dev
        vp = ctypes.c_void_p
        stream = vp(torch.cuda.current_stream().cuda_stream)
        pool.classify_dense_shard_dev(self.n, vp(d_births.data_ptr()), vp(d_compat.data_ptr()), s_lo, self.rank, self.world,
                                      self.peer_ptrs, stream)
        self.handle.barrier(channel=0)
        combine_shards_dev(self.n, self.world, vp(self.records.data_ptr()), vp(self.best.data_ptr()), vp(self.score.data_ptr()),
                           vp(self.first.data_ptr()), vp(self.count.data_ptr()), stream, device=self.records.device.index)
        self.handle.barrier(channel=1)
        return self.best, self.score, self.first, self.count
