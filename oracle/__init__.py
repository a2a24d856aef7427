"""TEST INFRASTRUCTURE -- Python access to the parity oracle.

Two oracles live here ("Oracle"):

* ``Oracle``: ctypes binding of ``libapt_oracle.so`` (``apt_oracle.cpp``), our CPU restatement of the
  reference's species-assignment path;
* ``run_reference``: runs ``_ref/apt_ref_driver``, the UNMODIFIED reference compiled by
  ``oracle/Makefile`` from the sources under /root/reference (binary only; it travels to the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product (``apoget_b200``) never does.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libapt_oracle.so")
REF_DRIVER = os.path.join(_HERE, "_ref", "apt_ref_driver")
REF_DRIVER_CREF = os.path.join(_HERE, "_ref", "apt_ref_driver_cref")

_u32p = ctypes.POINTER(ctypes.c_uint32)
_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_lib = None


def build(ref: bool = True) -> None:
    """Compile the restatement and, when /root/reference is present, the reference driver."""
    subprocess.run(["make", "-C", _HERE, "restate"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = ctypes.CDLL(LIB_PATH)
        L.apto_create.restype = ctypes.c_void_p
        L.apto_create.argtypes = [ctypes.c_uint32] * 3 + [ctypes.c_int] * 2 + [ctypes.c_double] * 3 + \
            [ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_uint32]
        L.apto_destroy.argtypes = [ctypes.c_void_p]
        L.apto_error.restype = ctypes.c_char_p
        L.apto_error.argtypes = [ctypes.c_void_p]
        L.apto_log_events.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.apto_add_genome.argtypes = [ctypes.c_void_p, _u32p, _f32p] + [ctypes.c_uint32] * 5 + [_u32p, _u8p]
        L.apto_del_genome.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.apto_candidacy.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
        L.apto_set_step.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.apto_step.argtypes = [ctypes.c_void_p, ctypes.c_uint32, _u32p, ctypes.c_uint32]
        L.apto_stats.argtypes = [ctypes.c_void_p, _u32p]
        L.apto_event_count.restype = ctypes.c_uint32
        L.apto_event_count.argtypes = [ctypes.c_void_p]
        L.apto_events.argtypes = [ctypes.c_void_p, _u32p]
        L.apto_to_json.restype = ctypes.c_uint64
        L.apto_to_json.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64]
        L.apto_replay.restype = ctypes.c_double
        L.apto_replay.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, _u32p, _u8p, _u32p,
                                  ctypes.POINTER(ctypes.c_double)]
        L.apto_distance.restype = ctypes.c_double
        L.apto_distance.argtypes = [ctypes.c_uint32, ctypes.c_uint32, _u32p, _u32p]
        L.apto_compatibility.restype = ctypes.c_double
        L.apto_compatibility.argtypes = [ctypes.c_float] * 3 + [ctypes.c_double]
        L.apto_score.restype = ctypes.c_float
        L.apto_score.argtypes = [ctypes.c_uint32] * 3 + [ctypes.c_int] + [ctypes.c_double] * 3 + \
            [_u32p, _f32p, _u32p, _f32p, _f32p]
        L.apto_score_dense.restype = None
        L.apto_score_dense.argtypes = [ctypes.c_uint32] * 3 + [ctypes.c_int] + [ctypes.c_double] * 3 + \
            [ctypes.c_uint32, _u32p, _f32p, ctypes.c_uint32, _u32p, _f32p, _u32p, ctypes.c_uint32, _u32p, _f32p, _u32p, _u32p]
        L.apto_env_contribution.restype = ctypes.c_int
        L.apto_env_contribution.argtypes = [ctypes.c_int, ctypes.c_uint32, _f32p, _f32p, _u32p, _f32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


# Defaults of src/core/ptreeconfig.cpp:7-19
DEFAULTS = dict(K=5, full_continuous=True, env_crit=1, compat_thr=.1, sim_thr=.5, avg_thr=.25,
                trim_period=100, trim_thr=.25, trim_delay=4., trim_min_delay=200)


class Oracle:
    """Sequential CPU tree with the reference's semantics (one call per birth)."""

    def __init__(self, kind: int, words: int, **cfg):
        c = dict(DEFAULTS, **cfg)
        self.kind, self.words, self.cfg = kind, words, c
        self._h = lib().apto_create(kind, words, c["K"], int(c["full_continuous"]), c["env_crit"],
                                    c["compat_thr"], c["sim_thr"], c["avg_thr"], c["trim_period"],
                                    c["trim_thr"], c["trim_delay"], c["trim_min_delay"])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().apto_destroy(self._h)
            self._h = None

    def _check(self, rc):
        if rc:
            raise RuntimeError(lib().apto_error(self._h).decode())

    def log_events(self, on: bool):
        lib().apto_log_events(self._h, int(on))

    def add_genome(self, payload, compat, gid, mgid, msid, fgid, fsid):
        payload = np.ascontiguousarray(payload, dtype=np.uint32)
        compat = np.ascontiguousarray(compat, dtype=np.float32)
        sid, ent = ctypes.c_uint32(), ctypes.c_uint8()
        self._check(lib().apto_add_genome(self._h, _p(payload, _u32p), _p(compat, _f32p), gid, mgid, msid,
                                          fgid, fsid, ctypes.byref(sid), ctypes.byref(ent)))
        return sid.value, bool(ent.value)

    def del_genome(self, sid):
        self._check(lib().apto_del_genome(self._h, sid))

    def register_candidate(self, msid, fsid):
        self._check(lib().apto_candidacy(self._h, msid, fsid, +1))

    def unregister_candidate(self, msid, fsid):
        self._check(lib().apto_candidacy(self._h, msid, fsid, -1))

    def set_step(self, step):
        lib().apto_set_step(self._h, step)

    def step(self, step, alive_sids):
        a = np.ascontiguousarray(alive_sids, dtype=np.uint32)
        self._check(lib().apto_step(self._h, step, _p(a, _u32p), a.size))

    def stats(self) -> dict:
        out = np.zeros(8, dtype=np.uint32)
        lib().apto_stats(self._h, _p(out, _u32p))
        keys = ["insertions", "deletions", "comparisons", "branching", "next_sid", "width", "step", "stillborns"]
        return dict(zip(keys, (int(x) for x in out)))

    def events(self) -> np.ndarray:
        n = lib().apto_event_count(self._h)
        out = np.zeros((n, 4), dtype=np.uint32)
        if n:
            lib().apto_events(self._h, _p(out, _u32p))
        return out

    def to_json(self) -> dict:
        n = lib().apto_to_json(self._h, None, 0)
        buf = ctypes.create_string_buffer(n)
        lib().apto_to_json(self._h, buf, n)
        return json.loads(buf.raw[:n].decode())

    def replay(self, trace_path: str, max_births: int = 0, n_births: int | None = None):
        """Replays a trace file with the loop of oracle/ref_driver.cpp.
        Returns (sids, entered, seconds spent in the addGenome loop)."""
        from apoget_b200.synth import read_trace
        tr = read_trace(trace_path)
        n_births = tr.n_births
        self.gen_seconds = np.zeros(tr.n_gens, dtype=np.float64)
        sids = np.full(n_births, 0xFFFFFFFF, dtype=np.uint32)
        ent = np.zeros(n_births, dtype=np.uint8)
        done = ctypes.c_uint32()
        sec = lib().apto_replay(self._h, trace_path.encode(), max_births, _p(sids, _u32p), _p(ent, _u8p),
                                ctypes.byref(done), _p(self.gen_seconds, ctypes.POINTER(ctypes.c_double)))
        if sec < 0:
            raise RuntimeError(lib().apto_error(self._h).decode())
        return sids[:done.value], ent[:done.value], sec


def distance(kind, a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    return lib().apto_distance(kind, a.size, _p(a, _u32p), _p(b, _u32p))


def compatibility(mu, sin, sout, d) -> float:
    return lib().apto_compatibility(mu, sin, sout, d)


def score(kind, birth, birth_compat, reps, rep_compat, full_continuous=True,
          compat_thr=.1, sim_thr=.5, avg_thr=.25):
    """Score of one birth against one species' k representatives -> (score, float distances[k])."""
    birth = np.ascontiguousarray(birth, dtype=np.uint32)
    reps = np.ascontiguousarray(reps, dtype=np.uint32)
    bc = np.ascontiguousarray(birth_compat, dtype=np.float32)
    rc = np.ascontiguousarray(rep_compat, dtype=np.float32)
    k = reps.shape[0]
    d = np.zeros(k, dtype=np.float32)
    s = lib().apto_score(kind, birth.size, k, int(full_continuous), compat_thr, sim_thr, avg_thr,
                         _p(birth, _u32p), _p(bc, _f32p), _p(reps, _u32p), _p(rc, _f32p), _p(d, _f32p))
    return float(s), d


def score_dense(kind, births, birth_compat, reps, rep_compat, fill, full_continuous=True,
                compat_thr=.1, sim_thr=.5, avg_thr=.25, threads=0):
    """Every birth [B, words] against every species (reps [S, K, words], rep_compat [S, K, 3], fill [S]) with the
    reference's arithmetic -> (best species, best score, first accepting species, accept count), the winners the
    CUDA dense sweep reports.  Threaded over births."""
    births = np.ascontiguousarray(births, dtype=np.uint32)
    reps = np.ascontiguousarray(reps, dtype=np.uint32)
    B, words = births.shape
    S, K = reps.shape[0], reps.shape[1]
    bc = np.ascontiguousarray(birth_compat, dtype=np.float32).reshape(B, 3)
    rc = np.ascontiguousarray(rep_compat, dtype=np.float32).reshape(S, K, 3)
    fill = np.ascontiguousarray(fill, dtype=np.uint32)
    best, first, count = (np.zeros(B, dtype=np.uint32) for _ in range(3))
    best_score = np.zeros(B, dtype=np.float32)
    lib().apto_score_dense(kind, words, K, int(full_continuous), compat_thr, sim_thr, avg_thr, B, _p(births, _u32p),
                           _p(bc, _f32p), S, _p(reps, _u32p), _p(rc, _f32p), _p(fill, _u32p),
                           threads or (os.cpu_count() or 1), _p(best, _u32p), _p(best_score, _f32p), _p(first, _u32p),
                           _p(count, _u32p))
    return best, best_score, first, count


def env_contribution(crit, edist, gdist):
    """Enveloppe criterion on a full symmetric k x k matrix -> (better, than, value)."""
    e = np.ascontiguousarray(edist, dtype=np.float32)
    g = np.ascontiguousarray(gdist, dtype=np.float32)
    than, value = ctypes.c_uint32(), ctypes.c_float()
    better = lib().apto_env_contribution(crit, g.size, _p(e, _f32p), _p(g, _f32p), ctypes.byref(than),
                                         ctypes.byref(value))
    return bool(better), int(than.value), float(value.value)


def ref_args(cfg: dict) -> list:
    c = dict(DEFAULTS, **cfg)
    a = ["--rset", str(c["K"]), "--compat-thr", repr(c["compat_thr"]), "--sim-thr", repr(c["sim_thr"]),
         "--avg-thr", repr(c["avg_thr"]), "--crit", str(c["env_crit"]),
         "--trim-period", str(c["trim_period"]), "--trim-thr", repr(c["trim_thr"]),
         "--trim-delay", repr(c["trim_delay"]), "--trim-min-delay", str(c["trim_min_delay"])]
    if not c["full_continuous"]:
        a.append("--semi")
    return a


def read_res(path: str) -> dict:
    """Parses the .res file written by the reference driver."""
    buf = open(path, "rb").read()
    n = int(np.frombuffer(buf, "<u4", 1, 0)[0])
    off = 4
    sids = np.frombuffer(buf, "<u4", n, off).copy(); off += 4 * n
    ent = np.frombuffer(buf, "u1", n, off).copy(); off += n
    ne = int(np.frombuffer(buf[off:off + 4], "<u4")[0]); off += 4
    ev = np.frombuffer(buf[off:off + 16 * ne], "<u4").reshape(ne, 4).copy(); off += 16 * ne
    tail = np.frombuffer(buf[off:off + 32], "<u4")
    keys = ["insertions", "deletions", "comparisons", "branching", "next_sid", "width", "step"]
    return dict(sids=sids, entered=ent, events=ev, stats=dict(zip(keys, (int(x) for x in tail[:7]))))


def have_reference(cref: bool = False) -> bool:
    return os.path.exists(REF_DRIVER_CREF if cref else REF_DRIVER)


def run_reference(trace_path: str, prefix: str, cfg: dict, cref: bool = False, extra=()) -> dict:
    """Runs the compiled reference on a trace; returns its stdout summary (dict)."""
    exe = REF_DRIVER_CREF if cref else REF_DRIVER
    out = subprocess.run([exe] + ref_args(cfg) + list(extra) + [trace_path, prefix], check=True,
                         capture_output=True, text=True)
    return json.loads(out.stdout.strip().splitlines()[-1])
