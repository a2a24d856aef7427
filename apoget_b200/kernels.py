"""Python harness over the stateless kernels of the C ABI (include/apt_b200.h, second half)."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib
from .tree import make_params, _ptr, _u32, INVALID  # noqa: F401


class Pool:
    """Device-resident enveloppe pool: rep_payload [S,K,words] u32, rep_compat [S,K,3] f32, fill [S]."""

    def __init__(self, kind, words, rep_payload, rep_compat, fill, device=0, flags=0, **cfg):
        self.params = make_params(kind, words, device, flags, **cfg)
        self.K, self.words = self.params.rset_size, words
        fill = _u32(fill)
        self.n_species = fill.size
        rp = _u32(rep_payload).reshape(self.n_species, self.K, words)
        rc = np.ascontiguousarray(rep_compat, dtype=np.float32).reshape(self.n_species, self.K, 3)
        h = ctypes.c_void_p()
        check(lib().apt_pool_create(ctypes.byref(self.params), self.n_species, _ptr(rp), _ptr(rc), _ptr(fill),
                                    ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().apt_pool_destroy(self._h)
            self._h = None

    __del__ = close

    def uses_tensor_cores(self) -> bool:
        """True when the dense sweeps of this pool run as an int8 GEMM on the tensor cores (tc_dense.cu)."""
        return bool(lib().apt_pool_uses_tensor_cores(self._h))

    def score_pairs(self, payload, compat, pair_birth, pair_species):
        payload = _u32(payload).reshape(-1, self.words)
        n = payload.shape[0]
        compat = np.ascontiguousarray(compat, dtype=np.float32).reshape(n, 3)
        pb, ps = _u32(pair_birth), _u32(pair_species)
        score = np.zeros(pb.size, dtype=np.float32)
        dist = np.zeros((pb.size, self.K), dtype=np.float32)
        check(lib().apt_score_pairs(self._h, n, _ptr(payload), _ptr(compat), pb.size, _ptr(pb), _ptr(ps),
                                    _ptr(score), _ptr(dist)))
        return score, dist

    def classify_dense(self, payload, compat):
        payload = _u32(payload).reshape(-1, self.words)
        n = payload.shape[0]
        compat = np.ascontiguousarray(compat, dtype=np.float32).reshape(n, 3)
        best = np.zeros(n, dtype=np.uint32)
        best_score = np.zeros(n, dtype=np.float32)
        first = np.zeros(n, dtype=np.uint32)
        count = np.zeros(n, dtype=np.uint32)
        check(lib().apt_classify_dense(self._h, n, _ptr(payload), _ptr(compat), _ptr(best), _ptr(best_score),
                                       _ptr(first), _ptr(count)))
        return best, best_score, first, count

    def classify_dense_dev(self, n, d_payload, d_compat, d_best, d_best_score, d_first, d_count, stream):
        """All arguments after n are raw device pointers (ints); `stream` a CUstream handle (int)."""
        check(lib().apt_classify_dense_dev(self._h, n, d_payload, d_compat, d_best, d_best_score, d_first,
                                           d_count, stream))

    def classify_dense_shard_dev(self, n, d_payload, d_compat, species_offset, rank, world, peer_records, stream):
        """Sharded sweep: writes this shard's records into every rank's gather buffer (peer_records: `world`
        device addresses of world*n 16-byte records each; peer-mapped memory)."""
        ptrs = np.ascontiguousarray(peer_records, dtype=np.uint64)
        assert ptrs.size == world
        check(lib().apt_classify_dense_shard_dev(self._h, n, d_payload, d_compat, species_offset, rank, world,
                                                 ptrs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), stream))


def combine_shards_dev(n, world, d_records, d_best, d_best_score, d_first, d_count, stream, device=0):
    """Pool-wide winners from the gathered [world][n] records (all arguments raw device pointers)."""
    check(lib().apt_combine_shards_dev(device, n, world, d_records, d_best, d_best_score, d_first, d_count, stream))


def env_contribution(K, edist, gdist, crit=1, device=0):
    e = np.ascontiguousarray(edist, dtype=np.float32).reshape(-1, K, K)
    n = e.shape[0]
    g = np.ascontiguousarray(gdist, dtype=np.float32).reshape(n, K)
    p = make_params(0, 1, device, 0, K=K, env_crit=crit)
    better = np.zeros(n, dtype=np.uint8)
    than = np.zeros(n, dtype=np.uint32)
    value = np.zeros(n, dtype=np.float32)
    check(lib().apt_env_contribution(ctypes.byref(p), n, _ptr(e), _ptr(g), _ptr(better), _ptr(than), _ptr(value)))
    return better.astype(bool), than, value


def distance_matrix(payload_a, payload_b, device=0):
    """Hamming distances of every genome of payload_a [n_a, words] to every genome of payload_b [n_b, words]
    (u32 [n_a, n_b]) through the tensor-core pipeline of the dense sweep (genomes of 1024..32768 bits)."""
    a, b = _u32(payload_a), _u32(payload_b)
    assert a.ndim == 2 and b.ndim == 2 and a.shape[1] == b.shape[1]
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.uint32)
    p = make_params(0, a.shape[1], device, 0)
    check(lib().apt_distance_matrix(ctypes.byref(p), a.shape[0], _ptr(a), b.shape[0], _ptr(b), _ptr(out)))
    return out


def pairwise(kind, payload, idx_a, idx_b, device=0):
    payload = _u32(payload)
    n, words = payload.shape
    a, b = _u32(idx_a), _u32(idx_b)
    out = np.zeros(a.size, dtype=np.float32)
    p = make_params(kind, words, device, 0)
    check(lib().apt_pairwise(ctypes.byref(p), n, _ptr(payload), a.size, _ptr(a), _ptr(b), _ptr(out)))
    return out


def mate_compat(kind, payload, compat, idx_mother, idx_father, device=0):
    """bailOutCrossver's pre-filter: (distance(mother, father), mother.compatibility(distance)) per pair."""
    payload = _u32(payload)
    n, words = payload.shape
    compat = np.ascontiguousarray(compat, dtype=np.float32).reshape(n, 3)
    a, b = _u32(idx_mother), _u32(idx_father)
    dist = np.zeros(a.size, dtype=np.float32)
    comp = np.zeros(a.size, dtype=np.float64)
    p = make_params(kind, words, device, 0)
    check(lib().apt_mate_compat(ctypes.byref(p), n, _ptr(payload), _ptr(compat), a.size, _ptr(a), _ptr(b),
                                _ptr(dist), _ptr(comp)))
    return dist, comp
