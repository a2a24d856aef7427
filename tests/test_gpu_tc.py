"""GPU parity tests of the tensor-core dense sweep (apoget_b200/csrc/tc_dense.cu): Hamming distance as an exact
int8 GEMM (tcgen05.mma, accumulators in TMEM), scoring in the epilogue.  Checked against numpy, the oracle's
threaded dense scorer (every birth x every species) and the popcount sweep of dense.cu."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from apoget_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _popcount(x):
    return np.unpackbits(x.view(np.uint8), axis=-1).sum(-1)


@pytest.mark.parametrize("words,na,nb", [(32, 130, 300), (128, 257, 700), (36, 64, 513), (33, 5, 9)])
def test_distance_matrix_on_tensor_cores(words, na, nb):
    """births x members Hamming distances straight from the accumulators: every tile, every K block, ragged edges
    (row counts that are no multiple of the tile, a genome length that is no multiple of the 128-bit K block)."""
    from apoget_b200.kernels import distance_matrix
    rng = np.random.default_rng(words + na)
    a = rng.integers(0, 2**32, size=(na, words), dtype=np.uint32)
    b = rng.integers(0, 2**32, size=(nb, words), dtype=np.uint32)
    b[:na // 2] = a[:na // 2]                       # zero and small distances too
    b[1, 0] ^= 1
    got = distance_matrix(a, b)
    want = np.stack([_popcount(a[i][None, :] ^ b) for i in range(na)])
    assert np.array_equal(got, want.astype(np.uint32))


def _pool(rng, words, S, K, sigma, classes, ragged):
    base = rng.integers(0, 2**32, size=(S, 1, words), dtype=np.uint32)
    flips = np.zeros((S, K, words), dtype=np.uint32)
    for _ in range(40):
        pos = rng.integers(0, words * 32, size=(S, K))
        np.put_along_axis(flips, (pos >> 5)[..., None],
                          np.take_along_axis(flips, (pos >> 5)[..., None], 2)
                          ^ (np.uint32(1) << (pos & 31).astype(np.uint32))[..., None], 2)
    reps = base ^ flips
    fill = np.full(S, K, dtype=np.uint32)
    if ragged:
        fill = rng.integers(1, K + 1, size=S).astype(np.uint32)
        fill[::7] = K
        fill[5] = 0
    rc = np.zeros((S, K, 3), dtype=np.float32)
    rc[..., 1] = sigma
    rc[..., 2] = sigma * 1.25
    for c in range(1, classes):                     # further compatibility classes, whole species and single representatives
        rc[c::5 + c, :, 1] = sigma * (0.5 + 0.2 * c)
        rc[2 * c::11, K // 2, 2] = sigma * (1.5 + 0.1 * c)
    return reps, rc, fill


@pytest.mark.parametrize("words,K,classes,ragged,full", [
    (128, 10, 1, False, True), (128, 10, 3, True, True), (128, 10, 3, True, False), (128, 5, 2, True, True),
    (128, 7, 3, True, True), (32, 10, 3, True, True), (64, 1, 2, True, True), (128, 32, 2, True, False)])
def test_tensor_core_sweep_equals_oracle_and_popcount_sweep(words, K, classes, ragged, full, monkeypatch):
    from apoget_b200.kernels import Pool
    rng = np.random.default_rng(words * 7 + K)
    S, B = 613, 1500                                # several representative tiles and birth tiles, ragged ends
    sigma = words * 32 / 60.0
    reps, rc, fill = _pool(rng, words, S, K, sigma, classes, ragged)
    src = rng.integers(0, S, size=B)
    births = reps[src, 0].copy()
    births[np.arange(B), rng.integers(0, words, size=B)] ^= rng.integers(0, 2**32, size=B, dtype=np.uint32)
    births[:16] = rng.integers(0, 2**32, size=(16, words), dtype=np.uint32)      # strangers: no species accepts them
    bc = rc[src, 0].copy()
    monkeypatch.setenv("APT_DENSE_TC", "1")
    pool = Pool(synth.BITS, words, reps, rc, fill, K=K, full_continuous=full)
    got = pool.classify_dense(births, bc)
    want = oracle.score_dense(synth.BITS, births, bc, reps, rc, fill, full_continuous=full)
    for g, w, what in zip(got, want, ("best species", "best score", "first accepting species", "accept count")):
        assert np.array_equal(g, w), what + " differs from the oracle"
    assert (want[3] > 0).any() and (want[3] == 0).any(), "the case must exercise both outcomes"
    monkeypatch.setenv("APT_DENSE_TC", "0")
    pop = pool.classify_dense(births, bc)
    for g, w in zip(got, pop):
        assert np.array_equal(g, w), "tensor-core and popcount sweeps disagree"
