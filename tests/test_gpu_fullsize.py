"""GPU tests at BASELINE.json's full sizes (-m gpu).

* candidate mode at C2 size (100 000 births per generation, ~1 000 species, K = 10): the whole tree must equal
  the CPU oracle's (species ids, enveloppe entries, dump), and the result must not depend on how a generation is
  cut into add_genomes calls (the batched call is DEFINED as n sequential addGenome calls);
* dense sweep at 1 M births x 100 k enveloppe members: size-independent properties (permutation of the births,
  shard additivity of the accept counts, argmax merge) plus an oracle check on a sample.
"""
import ctypes
import os

import numpy as np
import pytest

import oracle
from apoget_b200 import synth

pytestmark = pytest.mark.gpu


def _c2_trace(gens):
    return synth.generate(kind=synth.BITS, words=8, pop=100_000, gens=gens, seed=42, flips=0.5, sigma=12.0, hybrid=0.05,
                          burst_levels=3, burst_fanout=10)


def test_c2_full_size_tree_equals_oracle_and_is_batch_invariant(tmp_path):
    from apoget_b200.tree import PhylogeneticTree, INVALID
    t = _c2_trace(6)
    path = str(tmp_path / "c2.trace")
    synth.write_trace(path, t)
    o = oracle.Oracle(t.kind, t.words, K=10)
    o.log_events(False)
    osids, oent, _ = o.replay(path, n_births=t.n_births)

    whole = PhylogeneticTree(t.kind, t.words, K=10)
    sids, slots = whole.replay(t)
    assert np.array_equal(sids, osids), "per-birth species ids differ from the oracle at full size"
    assert np.array_equal(slots >= 0, oent > 0), "enveloppe entries differ from the oracle at full size"
    assert whole.to_json() == o.to_json(), "tree dump differs from the oracle at full size"
    c = whole.counters()
    assert c["stream_births"] >= 200_000 and c["kernel_launches"] > 0, "the stream kernel did not carry the steady generations"

    # the same generations cut into uneven calls (ragged: 1, 7, 33 333, the rest)
    cut = PhylogeneticTree(t.kind, t.words, K=10)
    sid_of = np.full(t.n_births, INVALID, dtype=np.uint32)
    for g in range(t.n_gens):
        b0, b1 = int(t.gen_birth_off[g]), int(t.gen_birth_off[g + 1])
        edges = [b0, b0 + 1, b0 + 8, b0 + 33_341, b1]
        for lo, hi in zip(edges[:-1], edges[1:]):
            m, f = t.mother[lo:hi], t.father[lo:hi]
            hm, hf = m != INVALID, f != INVALID
            s, _, _ = cut.add_genomes(t.payload[lo:hi], t.compat[lo:hi], np.arange(lo + 1, hi + 1, dtype=np.uint32),
                                      np.where(hm, m + 1, INVALID), np.where(hm, sid_of[np.where(hm, m, 0)], INVALID),
                                      np.where(hf, f + 1, INVALID), np.where(hf, sid_of[np.where(hf, f, 0)], INVALID))
            sid_of[lo:hi] = s
        cut.del_genomes(sid_of[t.deaths_of(g)])
        cut.step(g + 1, sid_of[t.alive_of(g)])
    assert np.array_equal(sid_of, sids), "result depends on the batch boundaries"
    assert cut.to_json() == whole.to_json(), "tree depends on the batch boundaries"

    # the same generations with everything a host program hands over in page-locked memory (apt_host_alloc): the births
    # are then grouped into streams by the device (one cooperative kernel reading the id arrays in place), the unsettled
    # births are listed by it, and the results land in the caller's arrays directly
    from apoget_b200.tree import pinned_array
    pop = 100_000
    pin = PhylogeneticTree(t.kind, t.words, K=10)
    p_payload = pinned_array((pop, t.words), np.uint32)
    p_compat = pinned_array((pop, 3), np.float32)
    p_ids = [pinned_array((pop,), np.uint32) for _ in range(5)]
    p_out = [pinned_array((pop,), np.uint32), pinned_array((pop,), np.int32), pinned_array((pop,), np.uint32)]
    vp = ctypes.c_void_p
    psid = np.full(t.n_births, INVALID, dtype=np.uint32)
    launches_before_steady = None
    for g in range(t.n_gens):
        b0, b1 = int(t.gen_birth_off[g]), int(t.gen_birth_off[g + 1])
        assert b1 - b0 == pop
        m, f = t.mother[b0:b1], t.father[b0:b1]
        hm, hf = m != INVALID, f != INVALID
        p_payload[...] = t.payload[b0:b1]
        p_compat[...] = t.compat[b0:b1]
        for dst, src in zip(p_ids, (np.arange(b0 + 1, b1 + 1, dtype=np.uint32), np.where(hm, m + 1, INVALID),
                                    np.where(hm, psid[np.where(hm, m, 0)], INVALID), np.where(hf, f + 1, INVALID),
                                    np.where(hf, psid[np.where(hf, f, 0)], INVALID))):
            dst[...] = src.astype(np.uint32)
        if g == t.n_gens - 1:
            launches_before_steady = pin.counters()["kernel_launches"]
        pin.add_genomes_raw(pop, vp(p_payload.ctypes.data), vp(p_compat.ctypes.data), [vp(x.ctypes.data) for x in p_ids],
                            vp(p_out[0].ctypes.data), vp(p_out[1].ctypes.data), vp(p_out[2].ctypes.data))
        psid[b0:b1] = p_out[0]
        pin.del_genomes(psid[t.deaths_of(g)])
        pin.step(g + 1, psid[t.alive_of(g)])
    assert np.array_equal(psid, sids), "page-locked inputs (device-side grouping) change the result"
    assert pin.to_json() == whole.to_json(), "page-locked inputs (device-side grouping) change the tree"
    # the last (steady) generation: grouping kernel + stream kernel + the two list kernels + the fill of the class table at most
    assert pin.counters()["kernel_launches"] - launches_before_steady <= 8


@pytest.mark.parametrize("words,sigma", [(8, 12.0), (128, 60.0)])
def test_dense_sweep_full_size_properties(words, sigma):
    """1 M births x 10 000 species x K = 10 (256-bit: the whole batch; 4096-bit: a 131 072-birth slice keeps the
    test under a few seconds -- the properties do not depend on the number of births)."""
    import torch
    from apoget_b200 import shard
    from apoget_b200.kernels import Pool, combine_shards_dev
    rng = np.random.default_rng(44)
    S, K = 10_000, 10
    B = 1_000_000 if words == 8 else 131_072
    base = rng.integers(0, 2**32, size=(S, 1, words), dtype=np.uint32)
    noise = rng.integers(0, 2**32, size=(S, K, words), dtype=np.uint32)
    for _ in range(4 if words <= 8 else 6):
        noise &= rng.integers(0, 2**32, size=(S, K, words), dtype=np.uint32)
    reps = np.ascontiguousarray(base ^ noise)
    rc = np.zeros((S, K, 3), dtype=np.float32)
    rc[..., 1:] = sigma
    fill = np.full(S, K, dtype=np.uint32)
    fill[rng.integers(0, S, size=50)] = rng.integers(0, K, size=50)        # ragged and empty enveloppes
    src = rng.integers(0, S, size=B)
    births = reps[src, rng.integers(0, K, size=B)].copy()
    births[np.arange(B), rng.integers(0, words, size=B)] ^= rng.integers(0, 2**32, size=B, dtype=np.uint32) & np.uint32(0x00010101)
    bc = np.zeros((B, 3), dtype=np.float32)
    bc[:, 1:] = sigma
    vp = ctypes.c_void_p
    stream = vp(torch.cuda.current_stream().cuda_stream)
    d_compat = torch.from_numpy(bc).cuda()

    def sweep(pool, d_births, n):
        out = [torch.empty(n, dtype=dt, device="cuda") for dt in (torch.int32, torch.float32, torch.int32, torch.int32)]
        pool.classify_dense_dev(n, vp(d_births.data_ptr()), vp(d_compat.data_ptr()), *[vp(x.data_ptr()) for x in out], stream)
        torch.cuda.synchronize()
        return out

    whole = Pool(synth.BITS, words, reps, rc, fill, K=K)
    d_births = torch.from_numpy(births.view(np.int32)).cuda()
    best, score, first, count = sweep(whole, d_births, B)

    # 1. permuting the births permutes the results
    perm = torch.from_numpy(rng.permutation(B)).cuda()
    pb, ps, pf, pc = sweep(whole, d_births[perm].contiguous(), B)
    assert torch.equal(pb, best[perm]) and torch.equal(ps, score[perm]) and torch.equal(pf, first[perm]) and torch.equal(pc, count[perm])

    # 2. shard additivity / argmax merge: two species shards into one record buffer, combined on the device
    world = 2
    rec = torch.empty((world, B, 4), dtype=torch.int32, device="cuda")
    for r in range(world):
        lo, hi = shard.shard_range(S, r, world)
        part = Pool(synth.BITS, words, reps[lo:hi], rc[lo:hi], fill[lo:hi], K=K)
        part.classify_dense_shard_dev(B, vp(d_births.data_ptr()), vp(d_compat.data_ptr()), lo, r, world, [rec.data_ptr()] * world, stream)
        torch.cuda.synchronize()
        part.close()
    out = [torch.empty(B, dtype=dt, device="cuda") for dt in (torch.int32, torch.float32, torch.int32, torch.int32)]
    combine_shards_dev(B, world, vp(rec.data_ptr()), *[vp(o.data_ptr()) for o in out], stream)
    torch.cuda.synchronize()
    for a, b, name in zip(out, (best, score, first, count), ("best", "score", "first", "count")):
        assert torch.equal(a, b), f"sharded sweep differs from the whole pool: {name}"

    # 3. every birth is a near-copy of a representative of species src[b]: nearly all are accepted somewhere
    cnt = count.cpu().numpy()
    ok_src = fill[src] == K
    assert (cnt[ok_src] >= 1).mean() > 0.95
    # 4. the oracle on 10 240 births x ALL 10 000 species (1e9 distance evaluations, threaded scorer of oracle/)
    bst, sc, fst = best.cpu().numpy(), score.cpu().numpy(), first.cpu().numpy()
    idx = np.sort(rng.choice(B, size=10_240, replace=False))
    ob, osc, of, oc = oracle.score_dense(synth.BITS, births[idx], bc[idx], reps, rc, fill)
    assert np.array_equal(bst[idx].view(np.uint32), ob), "best species"
    assert np.array_equal(sc[idx], osc), "best score"
    assert np.array_equal(fst[idx].view(np.uint32), of), "first accepting species"
    assert np.array_equal(cnt[idx].view(np.uint32), oc), "accept count"


@pytest.mark.parametrize("kind,words,K,full", [
    (synth.BITS, 8, 10, True), (synth.BITS, 8, 10, False), (synth.BITS, 8, 7, True), (synth.BITS, 128, 10, True),
    (synth.BITS, 128, 10, False), (synth.BITS, 128, 5, True), (synth.F32, 64, 10, True), (synth.F32, 64, 10, False)])
def test_dense_sweep_10k_births_all_species_against_the_oracle(kind, words, K, full):
    """Every dense-sweep schedule at scale -- multi-chunk / double-buffered popcount tiles, the tensor-core pipeline,
    the float tile; continuous and vote-count epilogues -- with ragged fills (empty enveloppes included) AND four
    compatibility classes in the same pool: 10 240 births, each against all 3 000 species, all verified."""
    from apoget_b200.kernels import Pool
    rng = np.random.default_rng(1000 * kind + 10 * words + K + int(full))
    S, B = 3_000, 10_240
    if kind == synth.BITS:
        base = rng.integers(0, 2**32, size=(S, 1, words), dtype=np.uint32)
        noise = rng.integers(0, 2**32, size=(S, K, words), dtype=np.uint32)
        for _ in range(4 if words <= 8 else 6):
            noise &= rng.integers(0, 2**32, size=(S, K, words), dtype=np.uint32)
        reps = np.ascontiguousarray(base ^ noise)
        sigma = 12.0 if words == 8 else 60.0
    else:
        base = rng.standard_normal((S, 1, words)).astype(np.float32)
        reps = np.ascontiguousarray((base + 0.03 * rng.standard_normal((S, K, words)).astype(np.float32)).view(np.uint32))
        sigma = 0.4
    fill = rng.integers(1, K + 1, size=S).astype(np.uint32)
    fill[rng.random(S) < 0.6] = K
    fill[rng.integers(0, S, size=20)] = 0
    rc = np.zeros((S, K, 3), dtype=np.float32)
    rc[..., 1] = sigma
    rc[..., 2] = sigma * 1.25
    rc[1::4, :, 1] = sigma * 0.75                    # classes per species ...
    rc[2::9, :, 2] = sigma * 2.0
    rc[3::50, K // 2, 1] = sigma * 0.5               # ... and a foreign representative inside an enveloppe
    src = rng.integers(0, S, size=B)
    if kind == synth.BITS:
        births = reps[src, 0].copy()
        births[np.arange(B), rng.integers(0, words, size=B)] ^= rng.integers(0, 2**32, size=B, dtype=np.uint32) & np.uint32(0x00010101)
    else:
        births = (reps[src, 0].view(np.float32) + 0.03 * rng.standard_normal((B, words)).astype(np.float32)).view(np.uint32)
    bc = rc[src, 0].copy()
    pool = Pool(kind, words, reps, rc, fill, K=K, full_continuous=full)
    best, score, first, count = pool.classify_dense(births, bc)
    ob, osc, of, oc = oracle.score_dense(kind, births, bc, reps, rc, fill, full_continuous=full)
    assert (oc > 0).any() and (oc == 0).any()
    if kind == synth.BITS:
        assert np.array_equal(best, ob) and np.array_equal(score, osc), "winners must be bit-exact"
        assert np.array_equal(first, of) and np.array_equal(count, oc)
    else:
        # float vectors: distances and compatibilities in FP32 on the device (tolerance 1e-5, BASELINE.json north_star);
        # vote counts can differ where a compatibility sits within that tolerance of its threshold
        tol = 1e-5 if full else 1.0
        assert np.abs(score - osc).max() <= tol
        assert (best == ob).mean() > 0.999 and (np.abs(count.astype(np.int64) - oc) <= 1).mean() > 0.999
