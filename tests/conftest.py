import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per stream (see stochastic_gcn_b200/__init__.py)
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def dec(obj):
    """inverse of tests/golden/generate_golden.py:enc"""
    if "f32bits" in obj:
        return np.array(obj["f32bits"], dtype=np.uint32).view(np.float32)
    return np.array(obj["i32"], dtype=np.int32)


@pytest.fixture(scope="session")
def sampler_golden():
    with open(os.path.join(GOLDEN, "sampler_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def mult_slice_golden():
    with open(os.path.join(GOLDEN, "mult_slice_golden.json")) as f:
        return json.load(f)


def graph_from_tag(tag):
    from tests.graphs_small import random_graph, tree11
    if tag == "tree11":
        return tree11()
    _, n, avg, seed = tag.split(":")
    return random_graph(int(n), int(avg), int(seed))


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    if a.dtype == np.float32 or b.dtype == np.float32:
        a = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        b = np.ascontiguousarray(b, dtype=np.float32).view(np.uint32)
    assert np.array_equal(a, b), "%s differs at %s" % (what, np.nonzero(a != b)[0][:10])
