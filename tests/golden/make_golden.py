"""Regenerates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference): ``python tests/golden/make_golden.py``.
For every case below it writes ``<name>.trace`` (the input, produced by apoget_b200.synth with a
fixed seed), ``<name>.res`` (per-birth species ids, enveloppe-entry flags, the callback stream and
Stats as written by oracle/_ref/apt_ref_driver) and ``<name>.json.gz`` (the reference's own
``saveTo`` dump, phylogenetictree.hpp:934-941, 997-1001); ``<name>.rt.json.gz`` is the dump after the
reference's own toJson -> fromJson -> toJson when that is not a fixed point.  ``cases.json`` lists the cases with the
PTree configuration each one was run with.
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from apoget_b200 import synth  # noqa: E402

B, F = synth.BITS, synth.F32

# name -> (generator kwargs, PTree config overrides)
CASES = {
    # configs[0]: compiled-in defaults of ptreeconfig.cpp:7-19 (rsetSize 5, continuous, maxMinDist)
    "c1_defaults_b256": (dict(kind=B, words=8, pop=100, gens=40, seed=42, flips=2, sigma=8.0), dict()),
    "hybrids_b256_k5": (dict(kind=B, words=8, pop=100, gens=40, seed=1, flips=3, sigma=8.0, hybrid=0.3,
                             survive=0.3), dict(K=5)),
    "c2_b256_k10": (dict(kind=B, words=8, pop=150, gens=30, seed=2, flips=3, sigma=12.0, hybrid=0.1),
                    dict(K=10)),
    "semicontinuous_b256": (dict(kind=B, words=8, pop=100, gens=30, seed=3, flips=3, sigma=8.0, hybrid=0.2),
                            dict(K=5, full_continuous=False)),
    "crit0_b256": (dict(kind=B, words=8, pop=80, gens=30, seed=4, flips=3, sigma=8.0, hybrid=0.2),
                   dict(K=6, env_crit=0)),
    "crit2_b256": (dict(kind=B, words=8, pop=80, gens=30, seed=6, flips=3, sigma=8.0, hybrid=0.2),
                   dict(K=6, env_crit=2)),
    "crit3_b256": (dict(kind=B, words=8, pop=80, gens=30, seed=7, flips=3, sigma=8.0, hybrid=0.2),
                   dict(K=6, env_crit=3)),
    "c3_f32x64_k10": (dict(kind=F, words=64, pop=100, gens=30, seed=43, sigma=0.35, noise=0.05,
                           noise_coords=8, hybrid=0.1), dict(K=10)),
    "c4_b4096_k10": (dict(kind=B, words=128, pop=60, gens=25, seed=44, flips=20, sigma=60.0, hybrid=0.1),
                     dict(K=10)),
    "founders_orphans_classes": (dict(kind=B, words=8, pop=100, gens=30, seed=10, flips=2, sigma=8.0,
                                      founders=4, hybrid=0.2, compat_classes=4), dict(K=5)),
    "c5_speciation_trimming": (dict(kind=B, words=8, pop=60, gens=260, seed=45, flips=3, sigma=6.0,
                                    hybrid=0.1),
                               dict(K=8, trim_period=50, trim_min_delay=20, trim_delay=1.0, trim_thr=0.5)),
    # 96-bit payloads: three words per genome, padded to four on the device
    "ragged_b96_k3": (dict(kind=B, words=3, pop=37, gens=20, seed=12, flips=1, sigma=3.0, hybrid=0.5,
                           survive=0.5), dict(K=3)),
    # vote-count scoring with four compatibility classes on 160-bit payloads (five words, padded to eight)
    "semi_classes_b160_k5": (dict(kind=B, words=5, pop=100, gens=30, seed=14, flips=2, sigma=6.0, founders=4,
                                  hybrid=0.2, compat_classes=4), dict(K=5, full_continuous=False)),
    "k1_b64": (dict(kind=B, words=2, pop=40, gens=20, seed=13, flips=1, sigma=2.0), dict(K=1)),
}


def main():
    if not oracle.have_reference():
        oracle.build(ref=True)
    index = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (gen, cfg) in CASES.items():
            t = synth.generate(**gen)
            tp = os.path.join(HERE, name + ".trace")
            synth.write_trace(tp, t)
            prefix = os.path.join(tmp, name)
            summary = oracle.run_reference(tp, prefix, cfg, extra=["--roundtrip"])
            shutil.copy(prefix + ".res", os.path.join(HERE, name + ".res"))
            with open(prefix + ".json", "rb") as f, \
                    gzip.GzipFile(os.path.join(HERE, name + ".json.gz"), "wb", mtime=0) as g:
                g.write(f.read())
            # the reference's fromJson re-attaches sub-species in ascending id (phylogenetictree.hpp:960-962):
            # where that changes the dump, keep what the reference itself produces
            rt = os.path.join(HERE, name + ".rt.json.gz")
            if summary["roundtrip_ok"]:
                if os.path.exists(rt):
                    os.remove(rt)
            else:
                with open(prefix + ".rt.json", "rb") as f, gzip.GzipFile(rt, "wb", mtime=0) as g:
                    g.write(f.read())
            index[name] = dict(roundtrip_fixed_point=bool(summary["roundtrip_ok"]), gen=gen, cfg=cfg, births=summary["births"], species=summary["next_sid"],
                               width=summary["width"], comparisons=summary["comparisons"],
                               branching=summary["branching"])
            print(name, summary)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
