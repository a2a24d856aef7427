"""Species-enveloppe sharding of the dense sweep over several GPUs (north_star (5)).

Every rank holds a contiguous range of the species [s_lo, s_hi) in its own pool, sweeps ALL births
against it (classify_dense), and the per-birth partial winners are exchanged with ONE all-gather of
16 bytes per birth per rank; every rank then reduces them identically:

    best species  = strictly greatest score, lowest species index on ties
                    (ranks own increasing index ranges -> first rank holding the maximum)
    first accept  = lowest species index with score > 0
    accept count  = sum

`torch.distributed` is plumbing only (NCCL over NVLink on the GPUs, gloo in the CPU tests).
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
