"""Synthetic host simulations ("birth traces") for the species-assignment path.

APOGeT has no simulation of its own: a host program calls ``addGenome`` / ``delGenome`` / ``step``
on the tree (reference ``src/core/tree/phylogenetictree.hpp:245-317``).  A *trace* records what such
a host would feed it -- every birth (fixed-length payload, the three BOC compatibility genes of
``src/core/crossover.h:43-49``, parents), every death and the per-generation alive list -- so that the
same input can be replayed through the unmodified reference, the CPU oracle and the CUDA path.

Generators follow SURVEY.md section 8(d): generational populations, clone-with-flips or per-word
uniform crossover for bit-strings, gaussian perturbation for float vectors, all seeded.
The binary layout is documented in ``oracle/trace.h``.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

BITS, F32 = 0, 1
NONE = 0xFFFFFFFF
_MAGIC = b"APTTRACE"


@dataclasses.dataclass
class Trace:
    kind: int                 # BITS | F32
    words: int                # 32-bit words per payload
    gen_birth_off: np.ndarray  # u32 [G+1]
    gen_death_off: np.ndarray  # u32 [G+1]
    gen_alive_off: np.ndarray  # u32 [G+1]
    mother: np.ndarray        # u32 [B] birth index or NONE
    father: np.ndarray        # u32 [B]
    compat: np.ndarray        # f32 [B,3] (mu, sigma_in, sigma_out)
    payload: np.ndarray       # u32 [B,words]
    deaths: np.ndarray        # u32 [D] birth indices
    alive: np.ndarray         # u32 [A] birth indices

    @property
    def n_births(self) -> int:
        return int(self.mother.shape[0])

    @property
    def n_gens(self) -> int:
        return int(self.gen_birth_off.shape[0]) - 1

    def births_of(self, gen: int) -> range:
        return range(int(self.gen_birth_off[gen]), int(self.gen_birth_off[gen + 1]))

    def deaths_of(self, gen: int) -> np.ndarray:
        return self.deaths[int(self.gen_death_off[gen]):int(self.gen_death_off[gen + 1])]

    def alive_of(self, gen: int) -> np.ndarray:
        return self.alive[int(self.gen_alive_off[gen]):int(self.gen_alive_off[gen + 1])]


def write_trace(path: str, t: Trace) -> None:
    u32 = lambda a: np.ascontiguousarray(a, dtype="<u4").tobytes()
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(struct.pack("<8I", 1, t.kind, t.words, t.n_births, t.n_gens,
                            int(t.deaths.shape[0]), int(t.alive.shape[0]), 0))
        f.write(u32(t.gen_birth_off))
        f.write(u32(t.gen_death_off))
        f.write(u32(t.gen_alive_off))
        f.write(u32(t.mother))
        f.write(u32(t.father))
        f.write(np.ascontiguousarray(t.compat, dtype="<f4").tobytes())
        f.write(u32(t.payload))
        f.write(u32(t.deaths))
        f.write(u32(t.alive))


def read_trace(path: str) -> Trace:
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != _MAGIC:
        raise ValueError(f"{path}: not a trace file")
    version, kind, words, B, G, D, A, _ = struct.unpack_from("<8I", buf, 8)
    if version != 1:
        raise ValueError(f"{path}: unsupported trace version {version}")
    off = 40

    def take(n, dtype="<u4"):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=n, offset=off).copy()
        off += 4 * n
        return a

    gbo, gdo, gao = take(G + 1), take(G + 1), take(G + 1)
    mother, father = take(B), take(B)
    compat = take(3 * B, "<f4").reshape(B, 3)
    payload = take(B * words).reshape(B, words)
    deaths, alive = take(D), take(A)
    return Trace(kind, words, gbo, gdo, gao, mother, father, compat, payload, deaths, alive)


def _flip_bits(rng: np.random.Generator, payload: np.ndarray, flips: float) -> None:
    """XOR ``flips`` uniformly chosen bit positions (with replacement) into every row; a fractional
    part is the probability of one more flip."""
    n, words = payload.shape
    rows = np.arange(n)
    whole = int(flips)
    for _ in range(whole):
        pos = rng.integers(0, words * 32, size=n)
        payload[rows, pos >> 5] ^= (np.uint32(1) << (pos & 31).astype(np.uint32))
    frac = float(flips) - whole
    if frac > 0:
        pos = rng.integers(0, words * 32, size=n)
        bit = np.where(rng.random(n) < frac, np.uint32(1) << (pos & 31).astype(np.uint32), np.uint32(0))
        payload[rows, pos >> 5] ^= bit.astype(np.uint32)


def generate(kind: int = BITS, words: int = 8, pop: int = 1000, gens: int = 100, seed: int = 42,
             flips: float = 2, init_flips: int = 0, hybrid: float = 0.0, survive: float = 0.0,
             mu: float = 0.0, sigma: float = 8.0, compat_classes: int = 1,
             noise: float = 0.05, noise_coords: int = 8, founders: int = 1,
             parent_skew: float = 0.0, burst_levels: int = 0, burst_fanout: int = 10,
             burst_bits: int = 48, burst_scale: float = 1.0, hybrid_far: float = -1.0) -> Trace:
    """Generational toy simulation.

    Generation 0 holds ``pop`` primordial genomes (no parents) derived from ``founders`` random
    ancestors (``init_flips`` bit flips, or gaussian noise, each).  Every later generation has
    ``pop`` births: the mother (and, with probability ``hybrid``, a distinct-index father) is drawn
    from the individuals alive after the previous generation; the child is the mother's payload
    (or a per-word uniform crossover of both parents) mutated by ``flips`` bit flips (BITS) or by
    N(0, noise^2) on ``noise_coords`` coordinates (F32).  After the births a fraction
    ``1 - survive`` of the previous population dies.  ``parent_skew`` > 0 draws parents with
    probability proportional to rank^-skew (a few prolific lineages).  With ``compat_classes`` > 1
    each founder lineage gets its own (mu, sigma_in, sigma_out) triple, inherited from the mother.
    ``burst_levels`` > 0 adds an adaptive radiation: in each of the generations 1..burst_levels every
    child jumps to one of ``burst_fanout`` niches (XOR of a ``burst_bits``-bit mask, or a shift of
    ``burst_scale`` sigma for float genomes), which grows a tree of fanout^levels well separated
    species below the root before the steady regime starts.  Fathers are drawn from the mother's own
    lineage/niche (mates must be compatible, cf. the bail-out crossover of crossover.h:222-248); a
    fraction ``hybrid_far`` of the hybrids (default: all of them when no bursts/founders are used)
    takes a father from anywhere in the population.
    """
    rng = np.random.default_rng(seed)
    B = pop * gens
    payload = np.zeros((B, words), dtype=np.uint32)
    compat = np.zeros((B, 3), dtype=np.float32)
    mother = np.full(B, NONE, dtype=np.uint32)
    father = np.full(B, NONE, dtype=np.uint32)
    gbo, gdo, gao = [0], [0], [0]
    deaths, alive_all = [], []

    classes = np.zeros((max(1, compat_classes), 3), dtype=np.float32)
    for c in range(classes.shape[0]):
        scale = 1.0 if c == 0 else float(rng.uniform(0.75, 1.5))
        classes[c] = (mu, sigma * scale, sigma * (scale if c == 0 else float(rng.uniform(0.75, 1.5))))

    # generation 0
    if kind == BITS:
        anc = rng.integers(0, 2**32, size=(founders, words), dtype=np.uint32)
    else:
        anc = rng.standard_normal((founders, words)).astype(np.float32).view(np.uint32)
    lineage = rng.integers(0, founders, size=pop) if founders > 1 else np.zeros(pop, dtype=np.int64)
    payload[:pop] = anc[lineage]
    if kind == BITS:
        _flip_bits(rng, payload[:pop], init_flips)
    elif init_flips:
        f = payload[:pop].view(np.float32)
        f += (rng.standard_normal(f.shape) * noise).astype(np.float32)
    compat[:pop] = classes[lineage % classes.shape[0]]
    cluster = np.zeros(B, dtype=np.int64)       # lineage / niche path of every individual
    cluster[:pop] = lineage
    if hybrid_far < 0:
        hybrid_far = 1.0 if (burst_levels == 0 and founders == 1) else 0.0
    alive = np.arange(pop, dtype=np.uint32)
    gbo.append(pop)
    gdo.append(0)
    alive_all.append(alive.copy())
    gao.append(gao[-1] + alive.size)

    for g in range(1, gens):
        b0 = g * pop
        n_al = alive.size
        if parent_skew > 0:
            w = (np.arange(1, n_al + 1, dtype=np.float64)) ** (-parent_skew)
            w /= w.sum()
            m_idx = rng.choice(n_al, size=pop, p=w)
            f_idx = rng.choice(n_al, size=pop, p=w)
        else:
            m_idx = rng.integers(0, n_al, size=pop)
            f_idx = rng.integers(0, n_al, size=pop)
        m = alive[m_idx]
        f = alive[f_idx]
        if hybrid > 0 and hybrid_far < 1.0:
            # same-cluster father: a random member of the mother's cluster
            order = np.argsort(cluster[alive], kind="stable")
            sorted_cl = cluster[alive][order]
            lo = np.searchsorted(sorted_cl, cluster[m], side="left")
            hi = np.searchsorted(sorted_cl, cluster[m], side="right")
            near = alive[order[lo + (rng.random(pop) * (hi - lo)).astype(np.int64)]]
            f = np.where(rng.random(pop) < hybrid_far, f, near)
        is_h = (rng.random(pop) < hybrid) & (f != m)
        child = payload[m].copy()
        if is_h.any():
            mask = rng.integers(0, 2, size=(pop, words)).astype(bool) & is_h[:, None]
            child = np.where(mask, payload[f], child)
        cluster[b0:b0 + pop] = cluster[m]
        if g <= burst_levels:
            niche = rng.integers(0, burst_fanout, size=pop)
            cluster[b0:b0 + pop] = cluster[m] * burst_fanout + niche
            if kind == BITS:
                masks = np.zeros((burst_fanout, words), dtype=np.uint32)
                _flip_bits(rng, masks, burst_bits)
                child ^= masks[niche]
            else:
                shifts = (rng.standard_normal((burst_fanout, words)) * burst_scale).astype(np.float32)
                child = (child.view(np.float32) + shifts[niche]).view(np.uint32)
        if kind == BITS:
            _flip_bits(rng, child, flips)
        else:
            cf = child.view(np.float32)
            rows = np.arange(pop)
            for _ in range(noise_coords):
                pos = rng.integers(0, words, size=pop)
                cf[rows, pos] += (rng.standard_normal(pop) * noise).astype(np.float32)
        payload[b0:b0 + pop] = child
        compat[b0:b0 + pop] = compat[m]
        mother[b0:b0 + pop] = m
        father[b0:b0 + pop] = np.where(is_h, f, NONE).astype(np.uint32)

        keep = rng.random(n_al) < survive
        deaths.append(alive[~keep])
        alive = np.concatenate([alive[keep], np.arange(b0, b0 + pop, dtype=np.uint32)])
        gbo.append(b0 + pop)
        gdo.append(gdo[-1] + int((~keep).sum()))
        alive_all.append(alive.copy())
        gao.append(gao[-1] + alive.size)

    return Trace(kind, words,
                 np.asarray(gbo, dtype=np.uint32), np.asarray(gdo, dtype=np.uint32),
                 np.asarray(gao, dtype=np.uint32), mother, father, compat, payload,
                 np.concatenate(deaths).astype(np.uint32) if deaths else np.zeros(0, np.uint32),
                 np.concatenate(alive_all).astype(np.uint32))
