"""GPU parity tests (-m gpu): the CUDA path, through the C ABI, against the oracle.

* tree level: every golden trace replayed through libapt_b200.so must reproduce the UNMODIFIED
  reference (per-birth species ids, enveloppe entries, callback stream, Stats, JSON dump).
  Bit-exact for bit-string genomes; float-vector genomes: identical structure and assignments,
  distances within 1e-5 relative (north_star tolerance; measured error is ~1e-7, one float ulp).
* kernel level: the stateless kernels against oracle/apt_oracle.cpp on seeded random inputs.
"""
import os

import numpy as np
import pytest

import oracle
from apoget_b200 import synth
from conftest import golden_cases, golden_json, golden_trace_path, GOLDEN

pytestmark = pytest.mark.gpu
CASES = sorted(golden_cases())
F32_RTOL = 1e-5


def _json_close(a, b, rtol):
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.keys() == b.keys()
        for k in a:
            _json_close(a[k], b[k], rtol)
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b)
        for x, y in zip(a, b):
            _json_close(x, y, rtol)
    elif isinstance(a, float) or isinstance(b, float):
        assert abs(a - b) <= rtol * max(abs(a), abs(b)), (a, b)
    else:
        assert a == b, (a, b)


def _compare(name, flags=0, device_scores=True):
    from apoget_b200.tree import PhylogeneticTree
    cfg = golden_cases()[name]["cfg"]
    t = synth.read_trace(golden_trace_path(name))
    tree = PhylogeneticTree(t.kind, t.words, flags=flags, **cfg)
    sids, slots = tree.replay(t)
    ref = oracle.read_res(os.path.join(GOLDEN, name + ".res"))
    assert np.array_equal(sids, ref["sids"]), "per-birth species ids"
    assert np.array_equal(slots >= 0, ref["entered"] > 0), "enveloppe entries"
    st = tree.stats()
    for k, v in ref["stats"].items():
        assert st[k] == v, f"Stats::{k}"
    assert np.array_equal(tree.events(), ref["events"]), "callback stream"
    if t.kind == synth.BITS:
        assert tree.to_json() == golden_json(name), "tree dump"
    else:
        _json_close(tree.to_json(), golden_json(name), F32_RTOL)
    c = tree.counters()
    assert c["kernel_launches"] > 0 and (c["device_scores"] > 0 or not device_scores), "the CUDA path did not run"
    return tree


@pytest.mark.parametrize("name", CASES)
def test_tree_matches_unmodified_reference(name):
    _compare(name)


@pytest.mark.parametrize("name", CASES)
def test_tree_with_the_device_made_list_of_unsettled_births(name, monkeypatch):
    """Large batches read the births the host has to look at from a list the device compacts in arrival order
    (launch_list_unsettled) instead of scanning every result word; forced on for the golden cases' small batches."""
    monkeypatch.setenv("APT_DEVLIST_MIN", "1")
    _compare(name)


@pytest.mark.parametrize("name", CASES)
def test_tree_with_births_grouped_on_the_device(name, monkeypatch):
    """Large batches whose id arrays are page-locked are grouped into per-species streams by the device
    (launch_group_births: keys, scan, scatter, per-stream sort) instead of the host's counting sort; forced on for the
    golden cases' small batches and ordinary arrays."""
    monkeypatch.setenv("APT_DEVGROUP_MIN", "1")
    monkeypatch.setenv("APT_DEVGROUP_PAGEABLE", "1")
    monkeypatch.setenv("APT_DEVLIST_MIN", "1")
    _compare(name)


@pytest.mark.parametrize("name", CASES)
def test_tree_with_streams_on_thread_block_clusters(name, monkeypatch):
    """A pass of few streams one of which is very long runs every stream on a cluster of CTAs (replicated enveloppe,
    events elected through distributed shared memory); forced on for the golden cases' short streams, whenever a pass
    has few enough streams for clusters to fit."""
    monkeypatch.setenv("APT_STREAM_CLUSTER_MIN", "1")
    _compare(name)


@pytest.mark.parametrize("seed", range(4))
def test_long_streams_on_clusters_equal_single_cta_streams(seed, monkeypatch):
    """One species holding thousands of births per call (the start of a simulation): the clustered stream kernel against
    the single-CTA one on the same trace -- ids, entries, callback stream, Stats and dump must be identical."""
    from apoget_b200.tree import PhylogeneticTree
    rng = np.random.default_rng(300 + seed)
    kind = synth.F32 if seed == 3 else synth.BITS
    gen = dict(kind=kind, words=[8, 128, 4, 16][seed], pop=int(rng.integers(3000, 9000)), gens=3, seed=seed, hybrid=0.02)
    if kind == synth.BITS:
        gen.update(flips=1.5, sigma=[12.0, 100.0, 6.0, 0.0][seed])
    else:
        gen.update(sigma=0.35, noise=0.02, noise_coords=4)
    t = synth.generate(**gen)
    cfg = dict(K=[10, 10, 5, 10][seed])
    out = []
    for cluster_min in ("1000000000", "1"):
        monkeypatch.setenv("APT_STREAM_CLUSTER_MIN", cluster_min)
        tree = PhylogeneticTree(t.kind, t.words, **cfg)
        sids, slots = tree.replay(t)
        out.append((sids, slots, tree.stats(), tree.events(), tree.to_json()))
    a, b = out
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[2] == b[2]
    assert np.array_equal(a[3], b[3])
    assert a[4] == b[4]
    assert int(np.bincount(a[0]).max()) >= 2048, "the trace has no long stream"


def test_variants_of_the_wide_stream_kernel_in_one_process(monkeypatch, tmp_path):
    """The variants of a stream kernel that need more than 48 KB of shared memory (long streams) each need their own
    opt-in before their first launch.  All variants have the same function-pointer type and go through one launcher
    lambda: a cache of "size already set" that was not keyed by the kernel let the second variant with an EQUAL size
    (bit strings with device exp(), then float vectors of the same width) skip its opt-in -- launch: invalid argument.
    Found by tools/fuzz_gpu.py.  Both traces must run, and match the oracle."""
    from apoget_b200.tree import PhylogeneticTree
    monkeypatch.setenv("APT_STREAM_CLUSTER_MIN", "1000000000")         # the 16-warp single-CTA variant, not clusters
    cases = [(dict(kind=synth.BITS, words=4, pop=5000, gens=3, seed=77, flips=1.0, sigma=9.0, hybrid=0.02), 4),
             (dict(kind=synth.F32, words=4, pop=5000, gens=3, seed=78, sigma=0.35, noise=0.02, noise_coords=2, hybrid=0.02), 0),
             (dict(kind=synth.BITS, words=4, pop=5000, gens=3, seed=79, flips=1.0, sigma=9.0, hybrid=0.02), 4)]
    for gen, flags in cases:
        t = synth.generate(**gen)
        path = str(tmp_path / "t.trace")
        synth.write_trace(path, t)
        o = oracle.Oracle(t.kind, t.words, K=10)
        osid, oent, _ = o.replay(path, n_births=t.n_births)
        tree = PhylogeneticTree(t.kind, t.words, flags=flags, K=10)
        sids, slots = tree.replay(t)
        assert np.array_equal(sids, osid) and np.array_equal(slots >= 0, oent > 0)
        assert tree.stats() == o.stats()


@pytest.mark.parametrize("seed", range(1000, 1012))
def test_randomised_trees_with_random_device_paths(seed):
    """tools/fuzz_gpu.py's cases (random genome kind / width / enveloppe size / criterion / hybrids / radiations, a random
    subset of the device-side paths forced on for small batches) against the oracle; profiles/fuzz_r02e.json is a run of
    250 larger ones (30 M births)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_gpu", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "fuzz_gpu.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    saved = {k: os.environ.get(k) for k in ("APT_STREAM_CLUSTER_MIN", "APT_DEVLIST_MIN", "APT_DEVGROUP_MIN", "APT_DEVGROUP_PAGEABLE")}
    try:
        row = fuzz.one(seed, 5000)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert not row["bad"], row


@pytest.mark.parametrize("name", ["c2_b256_k10", "hybrids_b256_k5", "c4_b4096_k10"])
def test_tree_generic_compat_path(name):
    """Device exp() instead of the libm-built table: decisions must still be identical."""
    _compare(name, flags=4)


@pytest.mark.parametrize("name,cap", [("semi_classes_b160_k5", 3000), ("founders_orphans_classes", 5000)])
def test_tree_compat_table_overflow(name, cap, monkeypatch):
    """The compatibility table overflows mid-run (APT_LUT_MAX_BYTES stands in for the 256 MiB limit): device exp()
    from then on; vote-count scoring then leaves the device kernels (no guard band can protect a vote)."""
    monkeypatch.setenv("APT_LUT_MAX_BYTES", str(cap))
    # (vote counting without the table: every device score is re-taken on the host -- none may be trusted)
    _compare(name, device_scores=not name.startswith("semi"))


@pytest.mark.parametrize("seed", range(10))
def test_radiations_resolved_by_the_family_kernel(seed, tmp_path):
    """Adaptive radiations: every birth is rejected at top level and joins or founds a sub-species, in order
    (findBestDerived :493-548, insertInto :657-726, makeNode :398-426).  resolve_families_kernel must reproduce the oracle
    bit for bit -- species ids, enveloppe entries, callback order, Stats and the dump."""
    from apoget_b200.tree import PhylogeneticTree
    rng = np.random.default_rng(7000 + seed)
    K = int(rng.choice([1, 3, 5, 10]))
    kind = synth.F32 if seed % 5 == 4 else synth.BITS
    words = int(rng.choice([4, 8, 16, 128])) if kind == synth.BITS else 64
    gen = dict(kind=kind, words=words, pop=int(rng.integers(300, 3000)), gens=int(rng.integers(5, 8)), seed=seed,
               hybrid=float(rng.choice([0.0, 0.0, 0.05])), burst_levels=int(rng.integers(2, 5)),
               burst_fanout=int(rng.choice([3, 10, 40, 300])), compat_classes=int(rng.integers(1, 3)))
    if kind == synth.BITS:
        gen.update(flips=float(rng.uniform(0.3, 2.0)), sigma=float(rng.uniform(0.05, 0.12)) * words * 32, burst_bits=max(8, words * 6))
    else:
        gen.update(sigma=0.35, noise=0.02, noise_coords=4, burst_scale=1.0)
    t = synth.generate(**gen)
    cfg = dict(K=K, full_continuous=bool(seed % 4) or kind == synth.F32, env_crit=int(seed % 2))
    path = str(tmp_path / "t.trace")
    synth.write_trace(path, t)
    o = oracle.Oracle(t.kind, t.words, **cfg)
    osid, oent, _ = o.replay(path, n_births=t.n_births)
    tree = PhylogeneticTree(t.kind, t.words, **cfg)
    to_dev = None
    if seed % 2:
        import torch
        def to_dev(a):
            d = torch.from_numpy(a.view(np.int32)).cuda()
            return d, d.data_ptr(), torch.cuda.current_stream().cuda_stream
    sids, slots = tree.replay(t, to_device=to_dev)
    assert np.array_equal(sids, osid), "per-birth species ids"
    assert np.array_equal(slots >= 0, oent > 0), "enveloppe entries"
    assert tree.stats() == o.stats()
    assert np.array_equal(tree.events(), o.events()), "callback stream"
    if kind == synth.BITS:
        assert tree.to_json() == o.to_json()
    else:
        _json_close(tree.to_json(), o.to_json(), F32_RTOL)
    assert tree.counters()["family_births"] > 0, "the family kernel never ran"


def _random_pool(rng, kind, words, S, K, sigma):
    if kind == synth.BITS:
        base = rng.integers(0, 2**32, size=(S, 1, words), dtype=np.uint32)
        flips = np.zeros((S, K, words), dtype=np.uint32)
        for _ in range(6):
            pos = rng.integers(0, words * 32, size=(S, K))
            np.put_along_axis(flips, (pos >> 5)[..., None],
                              np.take_along_axis(flips, (pos >> 5)[..., None], 2)
                              ^ (np.uint32(1) << (pos & 31).astype(np.uint32))[..., None], 2)
        reps = base ^ flips
    else:
        base = rng.standard_normal((S, 1, words)).astype(np.float32)
        reps = (base + 0.05 * rng.standard_normal((S, K, words)).astype(np.float32)).view(np.uint32)
    fill = rng.integers(1, K + 1, size=S).astype(np.uint32)
    compat = np.zeros((S, K, 3), dtype=np.float32)
    compat[..., 1] = sigma
    compat[..., 2] = sigma * 1.25
    compat[::3, :, 1] = sigma * 0.75          # a second compatibility class
    return reps, compat, fill


@pytest.mark.parametrize("kind,words,K,sigma,full", [
    (synth.BITS, 8, 10, 8.0, True), (synth.BITS, 8, 5, 8.0, False), (synth.BITS, 128, 10, 60.0, True),
    (synth.BITS, 3, 7, 4.0, True), (synth.BITS, 1, 1, 2.0, True), (synth.BITS, 40, 32, 20.0, True),
    (synth.F32, 64, 10, 0.4, True), (synth.F32, 5, 4, 0.2, True)])
def test_score_pairs_kernel(kind, words, K, sigma, full):
    from apoget_b200.kernels import Pool
    rng = np.random.default_rng(7 + words)
    S, B = 37, 211
    reps, rc, fill = _random_pool(rng, kind, words, S, K, sigma)
    pool = Pool(kind, words, reps, rc, fill, K=K, full_continuous=full)
    src = rng.integers(0, S, size=B)
    if kind == synth.BITS:
        births = reps[src, 0].copy()
        births[np.arange(B), rng.integers(0, words, size=B)] ^= rng.integers(0, 2**32, size=B, dtype=np.uint32) \
            & np.uint32(0x0101)
    else:
        births = (reps[src, 0].view(np.float32) + 0.05 * rng.standard_normal((B, words)).astype(np.float32)).view(np.uint32)
    bc = rc[src, 0].copy()
    pb = np.repeat(np.arange(B), 2).astype(np.uint32)
    ps = np.stack([src, rng.integers(0, S, size=B)], 1).reshape(-1).astype(np.uint32)
    score, dist = pool.score_pairs(births, bc, pb, ps)
    for p in range(pb.size):
        b, s = int(pb[p]), int(ps[p])
        k = int(fill[s])
        so, do = oracle.score(kind, births[b], bc[b], reps[s, :k], rc[s, :k], full_continuous=full)
        if kind == synth.BITS:
            assert np.array_equal(dist[p, :k], do), "Hamming distances must be exact"
            assert score[p] == np.float32(so), "score must be bit-exact (libm-built compat table)"
        else:
            assert np.allclose(dist[p, :k], do, rtol=F32_RTOL, atol=0)
            assert abs(score[p] - so) <= 1e-5


def test_score_pairs_empty_and_bad_indices():
    from apoget_b200.kernels import Pool
    from apoget_b200._lib import AptError
    rng = np.random.default_rng(1)
    reps, rc, fill = _random_pool(rng, synth.BITS, 8, 4, 5, 8.0)
    pool = Pool(synth.BITS, 8, reps, rc, fill, K=5)
    s, d = pool.score_pairs(reps[:, 0], rc[:, 0], [], [])
    assert s.size == 0 and d.shape == (0, 5)
    with pytest.raises(AptError):
        pool.score_pairs(reps[:, 0], rc[:, 0], [0], [99])


@pytest.mark.parametrize("kind,words,K,sigma", [(synth.BITS, 8, 10, 8.0), (synth.BITS, 128, 10, 60.0),
                                                (synth.BITS, 3, 4, 4.0), (synth.F32, 64, 10, 0.4)])
def test_dense_sweep_kernel(kind, words, K, sigma):
    from apoget_b200.kernels import Pool
    rng = np.random.default_rng(11 + words)
    S, B = 53, 97
    reps, rc, fill = _random_pool(rng, kind, words, S, K, sigma)
    fill[5] = 0                                    # a free slot must be skipped
    pool = Pool(kind, words, reps, rc, fill, K=K)
    src = rng.integers(0, S, size=B)
    births = reps[src, 0].copy()
    bc = rc[src, 0].copy()
    best, best_score, first, count = pool.classify_dense(births, bc)
    for b in range(B):
        scores = np.full(S, -np.inf, dtype=np.float32)
        for s in range(S):
            if fill[s]:
                scores[s] = oracle.score(kind, births[b], bc[b], reps[s, :fill[s]], rc[s, :fill[s]])[0]
        if kind == synth.BITS:
            assert best[b] == int(np.argmax(scores)) and best_score[b] == scores.max()
            acc = np.nonzero(scores > 0)[0]
            assert count[b] == acc.size
            assert first[b] == (acc[0] if acc.size else 0xFFFFFFFF)
        else:
            assert abs(best_score[b] - scores.max()) <= 1e-5
            sure = np.abs(scores) > 1e-5
            assert abs(int(count[b]) - int((scores > 0).sum())) <= int((~sure).sum())


@pytest.mark.parametrize("crit", [0, 1, 2, 3])
def test_env_contribution_kernel(crit):
    from apoget_b200.kernels import env_contribution
    rng = np.random.default_rng(3)
    K, n = 10, 300
    e = rng.integers(1, 40, size=(n, K, K)).astype(np.float32)
    e = np.triu(e, 1) + np.transpose(np.triu(e, 1), (0, 2, 1))
    g = rng.integers(0, 40, size=(n, K)).astype(np.float32)
    better, than, value = env_contribution(K, e, g, crit=crit)
    for i in range(n):
        ob, ot, ov = oracle.env_contribution(crit, e[i], g[i])
        if crit == 3:       # device exp(): the value may be an ulp off, the decision only inside 1e-13 of a tie
            assert (bool(better[i]), int(than[i])) == (ob, ot) and abs(float(value[i]) - ov) <= 1e-5 * max(1.0, abs(ov))
        else:
            assert (bool(better[i]), int(than[i]), float(value[i])) == (ob, ot, ov)


@pytest.mark.parametrize("kind,words", [(synth.BITS, 8), (synth.BITS, 128), (synth.BITS, 5), (synth.F32, 64)])
def test_pairwise_kernel(kind, words):
    from apoget_b200.kernels import pairwise
    rng = np.random.default_rng(words)
    n = 64
    pl = rng.integers(0, 2**32, size=(n, words), dtype=np.uint32) if kind == synth.BITS else \
        rng.standard_normal((n, words)).astype(np.float32).view(np.uint32)
    a, b = rng.integers(0, n, size=500), rng.integers(0, n, size=500)
    out = pairwise(kind, pl, a, b)
    exp = np.float32([oracle.distance(kind, pl[i], pl[j]) for i, j in zip(a, b)])
    if kind == synth.BITS:
        assert np.array_equal(out, exp)
    else:
        assert np.allclose(out, exp, rtol=F32_RTOL, atol=0)


@pytest.mark.parametrize("device_passes", [False, True])
@pytest.mark.parametrize("name", ["c2_b256_k10", "hybrids_b256_k5", "ragged_b96_k3", "c4_b4096_k10", "c5_speciation_trimming"])
def test_tree_device_resident_payloads(name, device_passes, monkeypatch):
    """apt_tree_add_genomes_dev: payloads stay in HBM, the host fetches only the rows it needs.  (device_passes: with
    the device-side grouping, list and clustered streams forced on for these small batches as well.)"""
    import torch
    from apoget_b200.tree import PhylogeneticTree
    if device_passes:
        for var in ("APT_DEVGROUP_MIN", "APT_DEVGROUP_PAGEABLE", "APT_DEVLIST_MIN", "APT_STREAM_CLUSTER_MIN"):
            monkeypatch.setenv(var, "1")
    cfg = golden_cases()[name]["cfg"]
    t = synth.read_trace(golden_trace_path(name))
    tree = PhylogeneticTree(t.kind, t.words, **cfg)

    def to_device(a):
        d = torch.from_numpy(a.view(np.int32)).cuda()
        torch.cuda.synchronize()
        return d, d.data_ptr(), torch.cuda.current_stream().cuda_stream

    sids, _ = tree.replay(t, to_device=to_device)
    ref = oracle.read_res(os.path.join(GOLDEN, name + ".res"))
    assert np.array_equal(sids, ref["sids"])
    assert np.array_equal(tree.events(), ref["events"])
    assert tree.to_json() == golden_json(name)


@pytest.mark.parametrize("name", ["c2_b256_k10", "hybrids_b256_k5", "c5_speciation_trimming"])
def test_tree_parallel_host_paths(name, monkeypatch):
    """Worker-pool versions of the host loops forced on for the small golden chunks."""
    monkeypatch.setenv("APT_PAR_MIN", "1")
    monkeypatch.setenv("APT_HOST_THREADS", "4")
    _compare(name)


@pytest.mark.parametrize("kind,words", [(synth.BITS, 8), (synth.BITS, 128), (synth.F32, 64)])
def test_mate_compat_kernel(kind, words):
    """bailOutCrossver's pre-filter (crossover.h:234-240): distance + the mother's compatibility."""
    from apoget_b200.kernels import mate_compat
    rng = np.random.default_rng(words + 1)
    n = 80
    if kind == synth.BITS:
        pl = rng.integers(0, 2**32, size=(1, words), dtype=np.uint32) ^ \
            (rng.integers(0, 2**32, size=(n, words), dtype=np.uint32) & np.uint32(0x01010101))
        cp = np.stack([rng.uniform(0, 20, n), rng.uniform(3, 12, n), rng.uniform(3, 12, n)], 1).astype(np.float32)
    else:
        pl = (rng.standard_normal((1, words)) + 0.1 * rng.standard_normal((n, words))).astype(np.float32).view(np.uint32)
        cp = np.stack([rng.uniform(0, 1, n), rng.uniform(.2, 1, n), rng.uniform(.2, 1, n)], 1).astype(np.float32)
    m, f = rng.integers(0, n, size=400), rng.integers(0, n, size=400)
    dist, comp = mate_compat(kind, pl, cp, m, f)
    for p in range(m.size):
        d = oracle.distance(kind, pl[m[p]], pl[f[p]])
        c = oracle.compatibility(*[float(x) for x in cp[m[p]]], d)
        if kind == synth.BITS:
            assert dist[p] == np.float32(d)
            assert abs(comp[p] - c) <= 4e-16 * max(c, 1e-300) + 1e-300 or abs(comp[p] - c) <= 1e-15
        else:
            assert abs(dist[p] - d) <= F32_RTOL * d
            assert abs(comp[p] - c) <= 1e-9


def test_tree_json_checkpoint_resume():
    """Dump after half the generations, load into a fresh tree (pool rebuilt from the dump), continue."""
    from apoget_b200.tree import PhylogeneticTree, INVALID
    name = "c2_b256_k10"
    cfg = golden_cases()[name]["cfg"]
    t = synth.read_trace(golden_trace_path(name))
    half = t.n_gens // 2
    a = PhylogeneticTree(t.kind, t.words, **cfg)
    sids, _ = a.replay(t, max_gens=half)
    b = PhylogeneticTree(t.kind, t.words, **cfg)
    b.from_json(a.to_json())
    for g in range(half, t.n_gens):
        b0, b1 = int(t.gen_birth_off[g]), int(t.gen_birth_off[g + 1])
        m, f = t.mother[b0:b1], t.father[b0:b1]
        hm, hf = m != INVALID, f != INVALID
        s, _, _ = b.add_genomes(t.payload[b0:b1], t.compat[b0:b1], np.arange(b0 + 1, b1 + 1, dtype=np.uint32),
                                np.where(hm, m + 1, INVALID), np.where(hm, sids[np.where(hm, m, 0)], INVALID),
                                np.where(hf, f + 1, INVALID), np.where(hf, sids[np.where(hf, f, 0)], INVALID))
        sids[b0:b1] = s
        b.del_genomes(sids[t.deaths_of(g)])
        b.step(g + 1, sids[t.alive_of(g)])
    ref = oracle.read_res(os.path.join(GOLDEN, name + ".res"))
    assert np.array_equal(sids, ref["sids"])
    assert b.to_json() == golden_json(name)
    assert b.counters()["stream_births"] > 0


@pytest.mark.parametrize("kind,words,sigma", [(synth.BITS, 8, 8.0), (synth.BITS, 128, 60.0), (synth.F32, 64, 0.4)])
def test_sharded_dense_sweep_records_and_combine(kind, words, sigma):
    """Species-sharded sweep: the sweep kernel writes per-birth records into the gather
    buffer of every rank and combine_shards reduces them.  Here three shards run one after the other on one
    GPU into one buffer (peer pointers = the same buffer): the result must equal the unsharded sweep, and
    the torch reduction of the NCCL path must agree with the combine kernel."""
    import ctypes
    import torch
    from apoget_b200 import shard
    from apoget_b200.kernels import Pool, combine_shards_dev
    rng = np.random.default_rng(5 + words)
    S, B, K, world = 53, 197, 10, 3
    reps, rc, fill = _random_pool(rng, kind, words, S, K, sigma)
    rc[...] = rc[1, 0]                              # device-resident births: one compatibility class
    fill[7] = 0
    src = rng.integers(0, S, size=B)
    births, bc = reps[src, 0].copy(), rc[src, 0].copy()
    whole = Pool(kind, words, reps, rc, fill, K=K)
    want = whole.classify_dense(births, bc)
    vp = ctypes.c_void_p
    d_births = torch.from_numpy(births.view(np.int32)).cuda()
    d_compat = torch.from_numpy(bc).cuda()
    rec = torch.full((world, B, 4), 0x55, dtype=torch.int32, device="cuda")
    stream = vp(torch.cuda.current_stream().cuda_stream)
    for r in range(world):
        lo, hi = shard.shard_range(S, r, world)
        part = Pool(kind, words, reps[lo:hi], rc[lo:hi], fill[lo:hi], K=K)
        part.classify_dense_shard_dev(B, vp(d_births.data_ptr()), vp(d_compat.data_ptr()), lo, r, world,
                                      [rec.data_ptr()] * world, stream)
        torch.cuda.synchronize()
        part.close()
    out = [torch.empty(B, dtype=dt, device="cuda") for dt in (torch.int32, torch.float32, torch.int32, torch.int32)]
    combine_shards_dev(B, world, vp(rec.data_ptr()), *[vp(o.data_ptr()) for o in out], stream)
    torch.cuda.synchronize()
    got = [o.cpu().numpy() for o in out]
    assert np.array_equal(got[0].view(np.uint32), want[0]), "best species"
    assert np.array_equal(got[1], want[1]), "best score"
    assert np.array_equal(got[2].view(np.uint32), want[2]), "first accepting species"
    assert np.array_equal(got[3].view(np.uint32), want[3]), "accept count"
    ref = shard.reduce_records(rec)
    for a, b in zip(out, ref):
        assert torch.equal(a, b), "combine kernel and the torch reduction of the all-gather path disagree"


@pytest.mark.parametrize("kind,words,K,sigma", [(synth.BITS, 8, 10, 8.0), (synth.BITS, 8, 5, 8.0), (synth.BITS, 128, 10, 60.0),
                                                (synth.F32, 64, 10, 0.4), (synth.F32, 16, 5, 0.3)])
def test_dense_sweep_full_uniform_pool(kind, words, K, sigma):
    """The unrolled epilogues of the dense sweep (full enveloppes, one compatibility class, K = 5 / 10): same
    contract as test_dense_sweep_kernel, plus one partially filled and one foreign-class species that must take
    the general path inside the same launch."""
    from apoget_b200.kernels import Pool
    rng = np.random.default_rng(23 + words + K)
    S, B = 131, 203
    reps, rc, fill = _random_pool(rng, kind, words, S, K, sigma)
    rc[...] = rc[1, 0]
    fill[:] = K
    fill[40] = max(1, K - 2)
    rc[77, :, 1] *= 0.5
    pool = Pool(kind, words, reps, rc, fill, K=K)
    src = rng.integers(0, S, size=B)
    births, bc = reps[src, 0].copy(), rc[np.full(B, 1), 0].copy()
    best, best_score, first, count = pool.classify_dense(births, bc)
    for b in range(0, B, 3):
        scores = np.array([oracle.score(kind, births[b], bc[b], reps[s, :fill[s]], rc[s, :fill[s]])[0] for s in range(S)],
                          dtype=np.float32)
        if kind == synth.BITS:
            assert best[b] == int(np.argmax(scores)) and best_score[b] == scores.max()
            acc = np.nonzero(scores > 0)[0]
            assert count[b] == acc.size
            assert first[b] == (acc[0] if acc.size else 0xFFFFFFFF)
        else:
            assert abs(best_score[b] - scores.max()) <= 1e-5
            sure = np.abs(scores) > 1e-5
            assert abs(int(count[b]) - int((scores > 0).sum())) <= int((~sure).sum())
