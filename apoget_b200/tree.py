"""Python harness over the C ABI: the reference's tree interface for device-encoded genomes.

Mirrors ``phylogeny::PhylogeneticTree`` (reference ``src/core/tree/phylogenetictree.hpp``):
``add_genomes`` is the batched ``addGenome`` (:268-302), ``del_genome`` (:305-317), ``step`` (:245-262),
``register_candidate`` / ``unregister_candidate`` (:320-329), ``stats`` (:343-365), ``to_json`` (:934-941).
All computation happens inside libapt_b200.so; this file only marshals numpy buffers.
"""
from __future__ import annotations

import ctypes
import json

import numpy as np

from . import _lib
from ._lib import AptParams, check, lib

INVALID = 0xFFFFFFFF
BITS, F32 = 0, 1
FLAG_NO_EVENTS, FLAG_NO_CHILD_KERNEL, FLAG_FORCE_GENERIC, FLAG_NO_STREAM_KERNEL, FLAG_NO_FAMILY_KERNEL = 1, 2, 4, 8, 16

# config::PTree defaults, src/core/ptreeconfig.cpp:7-19
DEFAULTS = dict(K=5, full_continuous=True, env_crit=1, compat_thr=.1, sim_thr=.5, avg_thr=.25,
                trim_period=100, trim_thr=.25, trim_delay=4., trim_min_delay=200)


def make_params(kind: int, words: int, device: int = 0, flags: int = 0, **cfg) -> AptParams:
    c = dict(DEFAULTS, **cfg)
    p = AptParams()
    lib().apt_params_default(ctypes.byref(p))
    p.kind, p.words, p.rset_size = kind, words, c["K"]
    p.full_continuous, p.env_crit = int(c["full_continuous"]), c["env_crit"]
    p.compat_thr, p.sim_thr, p.avg_thr = c["compat_thr"], c["sim_thr"], c["avg_thr"]
    p.trim_period, p.trim_thr, p.trim_delay, p.trim_min_delay = \
        c["trim_period"], c["trim_thr"], c["trim_delay"], c["trim_min_delay"]
    p.device, p.flags = device, flags
    return p


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def pinned_array(shape, dtype) -> np.ndarray:
    """numpy array in page-locked host memory (apt_host_alloc): what a host program keeps its encoded genomes and
    result arrays in (the header binding's apt::PinnedVector) -- add_genomes then reads / writes it by
    DMA, no staging copy.  Freed with the array."""
    import weakref
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = ctypes.c_void_p()
    check(lib().apt_host_alloc(max(n, 1), ctypes.byref(p)))
    buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib().apt_host_free, p)
    return arr


class PhylogeneticTree:
    def __init__(self, kind: int, words: int, device: int = 0, flags: int = 0, **cfg):
        self.kind, self.words = kind, words
        self.K = dict(DEFAULTS, **cfg)["K"]
        self._params = make_params(kind, words, device, flags, **cfg)
        h = ctypes.c_void_p()
        check(lib().apt_tree_create(ctypes.byref(self._params), ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().apt_tree_destroy(self._h)
            self._h = None

    __del__ = close

    # -- core API -------------------------------------------------------------------------------
    def add_genomes(self, payload, compat, gid, mother_gid, mother_sid, father_gid, father_sid):
        """Batched addGenome; returns (sid[n], slot[n] (-1 = not in the enveloppe), replaced_gid[n])."""
        payload = _u32(payload).reshape(-1, self.words)
        n = payload.shape[0]
        compat = np.ascontiguousarray(compat, dtype=np.float32).reshape(n, 3)
        ids = [_u32(x).reshape(n) for x in (gid, mother_gid, mother_sid, father_gid, father_sid)]
        sid = np.full(n, INVALID, dtype=np.uint32)
        slot = np.full(n, -1, dtype=np.int32)
        rep = np.full(n, INVALID, dtype=np.uint32)
        done = ctypes.c_uint32()
        check(lib().apt_tree_add_genomes(self._h, n, _ptr(payload), _ptr(compat), *[_ptr(x) for x in ids],
                                         _ptr(sid), _ptr(slot), _ptr(rep), ctypes.byref(done)))
        return sid, slot, rep

    def add_genomes_raw(self, n, payload_ptr, compat_ptr, ids, sid, slot, rep):
        """Pointer-level entry for the benchmark (pinned host buffers, preallocated outputs)."""
        done = ctypes.c_uint32()
        check(lib().apt_tree_add_genomes(self._h, n, payload_ptr, compat_ptr, *ids, sid, slot, rep,
                                         ctypes.byref(done)))

    def add_genomes_dev(self, n, d_payload_ptr, compat_ptr, stream, ids, sid, slot, rep):
        """Payloads resident in device memory (raw pointer), everything else host pointers."""
        done = ctypes.c_uint32()
        check(lib().apt_tree_add_genomes_dev(self._h, n, d_payload_ptr, compat_ptr, stream, *ids, sid, slot,
                                             rep, ctypes.byref(done)))

    def del_genome(self, sid: int):
        check(lib().apt_tree_del_genome(self._h, int(sid)))

    def del_genomes(self, sids):
        a = _u32(sids)
        check(lib().apt_tree_del_genomes(self._h, a.size, _ptr(a)))

    def register_candidate(self, mother_sid: int, father_sid: int):
        check(lib().apt_tree_register_candidate(self._h, mother_sid, father_sid, +1))

    def unregister_candidate(self, mother_sid: int, father_sid: int):
        check(lib().apt_tree_register_candidate(self._h, mother_sid, father_sid, -1))

    def set_step(self, step: int):
        check(lib().apt_tree_set_step(self._h, step))

    def step(self, step: int, alive_sids):
        a = _u32(alive_sids)
        check(lib().apt_tree_step(self._h, step, _ptr(a), a.size))

    # -- observers ------------------------------------------------------------------------------
    def stats(self) -> dict:
        out = np.zeros(8, dtype=np.uint32)
        check(lib().apt_tree_stats(self._h, out.ctypes.data_as(_lib.u32p)))
        keys = ["insertions", "deletions", "comparisons", "branching", "next_sid", "width", "step", "stillborns"]
        return dict(zip(keys, (int(x) for x in out)))

    def reset_stats(self):
        check(lib().apt_tree_reset_stats(self._h))

    def events(self, drain: bool = True) -> np.ndarray:
        n = lib().apt_tree_event_count(self._h)
        out = np.zeros((n, 4), dtype=np.uint32)
        if n:
            lib().apt_tree_events(self._h, _ptr(out), n, int(drain))
        return out

    def to_json(self) -> dict:
        n = lib().apt_tree_to_json(self._h, None, 0)
        buf = ctypes.create_string_buffer(int(n))
        lib().apt_tree_to_json(self._h, buf, n)
        return json.loads(buf.raw[:n].decode())

    def from_json(self, dump):
        """fromJson (:946-967) into this (empty) tree; `dump` is a dict as returned by to_json, or its text."""
        text = dump if isinstance(dump, (str, bytes)) else json.dumps(dump)
        if isinstance(text, str):
            text = text.encode()
        check(lib().apt_tree_from_json(self._h, text, len(text)))

    def counters(self) -> dict:
        out = np.zeros(16, dtype=np.uint64)
        check(lib().apt_tree_counters(self._h, out.ctypes.data_as(_lib.u64p)))
        keys = ["births", "device_scores", "patched_scores", "host_scores", "kernel_launches",
                "host_distance_evals", "guard_rechecks", "device_pairs", "device_reps_read",
                "score_launches", "refresh_rows", "streams_taken_over", "stream_launches", "stream_births",
                "stream_entries", "family_births"]
        return dict(zip(keys, (int(x) for x in out)))

    def timers(self) -> dict:
        out = np.zeros(8, dtype=np.float64)
        check(lib().apt_tree_timers(self._h, out.ctypes.data_as(_lib.f64p)))
        return dict(zip(["prepare_s", "device_s", "walk_s", "refresh_s", "score_kernel_ms", "stream_kernel_ms", "family_kernel_ms", "alloc_s"],
                        (float(x) for x in out)))

    def kernel_span(self) -> tuple[float, int]:
        """Stream kernel as timed on the device: (ms summed over its launches: first CTA in to last CTA out, launches)."""
        out = np.zeros(2, dtype=np.float64)
        check(lib().apt_tree_kernel_span(self._h, out.ctypes.data_as(_lib.f64p)))
        return float(out[0]), int(out[1])

    def node_info(self, sid: int) -> dict:
        info = _lib.AptNodeInfo()
        check(lib().apt_tree_node_info(self._h, sid, ctypes.byref(info)))
        return dict(sid=info.sid, parent_sid=info.parent_sid, data=list(info.data), fill=info.fill,
                    n_children=info.n_children, n_contributors=info.n_contributors)

    # -- trace replay (same loop as oracle/ref_driver.cpp) -----------------------------------------
    def replay(self, trace, max_gens: int | None = None, to_device=None, cuts=None):
        """Feeds a synth.Trace generation by generation: one add_genomes batch per generation (or, with
        ``cuts(g, n) -> increasing offsets inside (0, n)``, several calls per generation), then
        the deaths, then step().  Returns (sid per birth, slot per birth)."""
        B = trace.n_births
        sid_of = np.full(B, INVALID, dtype=np.uint32)
        slot_of = np.full(B, -1, dtype=np.int32)
        gens = trace.n_gens if max_gens is None else min(max_gens, trace.n_gens)
        for g in range(gens):
            g0, g1 = int(trace.gen_birth_off[g]), int(trace.gen_birth_off[g + 1])
            edges = [g0] + ([g0 + int(c) for c in cuts(g, g1 - g0) if 0 < int(c) < g1 - g0] if cuts else []) + [g1]
            for b0, b1 in zip(edges[:-1], edges[1:]):
                self._replay_call(trace, b0, b1, sid_of, slot_of, to_device)
            self.del_genomes(sid_of[trace.deaths_of(g)])
            self.step(g + 1, sid_of[trace.alive_of(g)])
        return sid_of, slot_of

    def _replay_call(self, trace, b0, b1, sid_of, slot_of, to_device):
        if True:
            m, f = trace.mother[b0:b1], trace.father[b0:b1]
            hm, hf = m != INVALID, f != INVALID
            mgid = np.where(hm, m + 1, INVALID).astype(np.uint32)
            fgid = np.where(hf, f + 1, INVALID).astype(np.uint32)
            msid = np.where(hm, sid_of[np.where(hm, m, 0)], INVALID).astype(np.uint32)
            fsid = np.where(hf, sid_of[np.where(hf, f, 0)], INVALID).astype(np.uint32)
            gid = np.arange(b0 + 1, b1 + 1, dtype=np.uint32)
            if to_device is None:
                s, sl, _ = self.add_genomes(trace.payload[b0:b1], trace.compat[b0:b1], gid, mgid, msid, fgid, fsid)
            else:
                # payloads handed over as a device pointer: to_device(ndarray) -> (keep-alive object, pointer, stream)
                keep, dptr, stream = to_device(np.ascontiguousarray(trace.payload[b0:b1]))
                compat = np.ascontiguousarray(trace.compat[b0:b1], dtype=np.float32)
                s = np.full(b1 - b0, INVALID, dtype=np.uint32)
                sl = np.full(b1 - b0, -1, dtype=np.int32)
                rp = np.full(b1 - b0, INVALID, dtype=np.uint32)
                self.add_genomes_dev(b1 - b0, ctypes.c_void_p(dptr), _ptr(compat), ctypes.c_void_p(stream),
                                     [_ptr(x) for x in (gid, mgid, msid, fgid, fsid)], _ptr(s), _ptr(sl), _ptr(rp))
                del keep
            sid_of[b0:b1], slot_of[b0:b1] = s, sl
