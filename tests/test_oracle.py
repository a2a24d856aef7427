"""Pins the CPU oracle (oracle/apt_oracle.cpp) against the UNMODIFIED reference.

The golden files were written by oracle/_ref/apt_ref_driver (the reference's own headers and TUs,
see tests/golden/make_golden.py).  The reference ships no tests or vectors of its own (SURVEY.md
section 4), so "every golden vector the reference holds" is the empty set; these fixtures are
outputs of the reference itself.
"""
import json
import os

import numpy as np
import pytest

import oracle
from apoget_b200 import synth
from conftest import golden_cases, golden_json, golden_trace_path, GOLDEN

CASES = sorted(golden_cases())


def _replay(name, cfg):
    t = synth.read_trace(golden_trace_path(name))
    o = oracle.Oracle(t.kind, t.words, **cfg)
    sids, ent, _ = o.replay(golden_trace_path(name), n_births=t.n_births)
    return t, o, sids, ent


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_golden(name):
    cfg = golden_cases()[name]["cfg"]
    t, o, sids, ent = _replay(name, cfg)
    ref = oracle.read_res(os.path.join(GOLDEN, name + ".res"))
    assert np.array_equal(sids, ref["sids"]), "per-birth species ids differ"
    assert np.array_equal(ent, ref["entered"]), "enveloppe-entry flags differ"
    st = o.stats()
    for k, v in ref["stats"].items():
        assert st[k] == v, f"Stats::{k}"
    assert np.array_equal(o.events(), ref["events"]), "callback stream differs"
    assert o.to_json() == golden_json(name), "tree dump differs"


def test_per_birth_api_equals_replay():
    """The one-call-per-birth interface (what the GPU tests drive) equals the bulk replay."""
    name = "hybrids_b256_k5"
    cfg = golden_cases()[name]["cfg"]
    t = synth.read_trace(golden_trace_path(name))
    o = oracle.Oracle(t.kind, t.words, **cfg)
    sid_of = np.full(t.n_births, synth.NONE, dtype=np.uint32)
    for g in range(t.n_gens):
        for b in t.births_of(g):
            m, f = int(t.mother[b]), int(t.father[b])
            sid_of[b], _ = o.add_genome(t.payload[b], t.compat[b], b + 1,
                                        synth.NONE if m == synth.NONE else m + 1,
                                        synth.NONE if m == synth.NONE else int(sid_of[m]),
                                        synth.NONE if f == synth.NONE else f + 1,
                                        synth.NONE if f == synth.NONE else int(sid_of[f]))
        for d in t.deaths_of(g):
            o.del_genome(int(sid_of[d]))
        o.step(g + 1, sid_of[t.alive_of(g)])
    ref = oracle.read_res(os.path.join(GOLDEN, name + ".res"))
    assert np.array_equal(sid_of, ref["sids"])
    assert o.to_json() == golden_json(name)


@pytest.mark.skipif(not (oracle.have_reference() and os.path.isdir("/root/reference")),
                    reason="live reference run needs the build container")
def test_live_reference_on_fresh_seed(tmp_path):
    """A trace that is NOT in the fixtures, through the reference binary and the restatement."""
    t = synth.generate(kind=synth.BITS, words=8, pop=120, gens=35, seed=2026, flips=3, sigma=9.0,
                       hybrid=0.25, survive=0.2)
    tp = str(tmp_path / "fresh.trace")
    synth.write_trace(tp, t)
    cfg = dict(K=7)
    for cref in (False, True):
        oracle.run_reference(tp, str(tmp_path / f"ref{int(cref)}"), cfg, cref=cref)
    o = oracle.Oracle(t.kind, t.words, **cfg)
    sids, ent, _ = o.replay(tp, n_births=t.n_births)
    r0 = oracle.read_res(str(tmp_path / "ref0.res"))
    r1 = oracle.read_res(str(tmp_path / "ref1.res"))
    assert np.array_equal(r0["sids"], r1["sids"]) and np.array_equal(r0["events"], r1["events"]), \
        "the std::cref timing variant must not change results"
    assert json.load(open(tmp_path / "ref0.json")) == json.load(open(tmp_path / "ref1.json"))
    assert np.array_equal(sids, r0["sids"]) and np.array_equal(ent, r0["entered"])
    assert o.to_json() == json.load(open(tmp_path / "ref0.json"))


def test_stateless_score_and_criterion():
    rng = np.random.default_rng(5)
    reps = rng.integers(0, 2**32, size=(4, 8), dtype=np.uint32)
    birth = reps[0].copy()
    birth[0] ^= np.uint32(0b1011)
    rc = np.tile(np.float32([0, 8, 8]), (4, 1))
    s, d = oracle.score(synth.BITS, birth, np.float32([0, 8, 8]), reps, rc)
    assert d[0] == 3.0
    ds = [oracle.distance(synth.BITS, birth, r) for r in reps]
    assert np.array_equal(d, np.float32(ds))
    avg = np.float32(0)
    for x in ds:
        avg = np.float32(np.float64(avg) + oracle.compatibility(0, 8, 8, x))
    assert s == np.float32(np.float64(avg / np.float32(4)) - .25)
    # maxMinDist on a hand-checkable enveloppe (enveloppecriteria.cpp:87-133)
    e = np.float32([[0, 2, 4], [2, 0, 6], [4, 6, 0]])
    better, than, value = oracle.env_contribution(1, e, np.float32([5, 5, 5]))
    assert (better, than, value) == (True, 0, 3.0)      # ties -> first index wins
    better, than, value = oracle.env_contribution(1, e, np.float32([1, 1, 1]))
    assert (better, than, value) == (False, 0, -1.0)
