import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)


def golden_json(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rb") as f:
        return json.loads(f.read().decode())


def golden_roundtrip_json(name):
    """What the reference's own toJson -> fromJson -> toJson yields (differs from the dump where
    re-parented sub-species get re-attached in ascending id, phylogenetictree.hpp:960-962)."""
    p = os.path.join(GOLDEN, name + ".rt.json.gz")
    if not os.path.exists(p):
        return golden_json(name)
    with gzip.open(p, "rb") as f:
        return json.loads(f.read().decode())


def golden_trace_path(name):
    return os.path.join(GOLDEN, name + ".trace")


@pytest.fixture(scope="session")
def cases():
    return golden_cases()


def json_close(a, b, rtol):
    """Structural equality of two dumps with floats compared to a relative tolerance (float-vector genomes)."""
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.keys() == b.keys()
        for k in a:
            json_close(a[k], b[k], rtol)
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b)
        for x, y in zip(a, b):
            json_close(x, y, rtol)
    elif isinstance(a, float) or isinstance(b, float):
        assert abs(a - b) <= rtol * max(abs(a), abs(b)), (a, b)
    else:
        assert a == b, (a, b)
