"""CPU: property-based fuzz of the oracle restatement (oracle/sgcn_oracle.c) against the compiled,
unmodified reference (oracle/_ref) -- random graphs (empty rows, hubs), seeds, degrees (below, at and
above the maximum degree), batch sizes from one node to every node, several consecutive batches (the
sampler's state -- RNG stream and in-place row permutation -- persists), uniform / control-variate /
importance modes.  Every output array is compared bit for bit."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import native
from tests.conftest import assert_bits_equal
from tests.graphs_small import random_graph

pytestmark = pytest.mark.skipif(not native.have_ref(), reason="compiled reference (oracle/_ref) not present")

VEC_NAMES = ["field", "ffield", "edg_s", "edg_t", "fedg_s", "fedg_t", "scales", "edg_w", "medg_w", "fedg_w"]


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(n=st.integers(2, 120), avg=st.integers(1, 12), gseed=st.integers(0, 10_000), seed=st.integers(0, 2**31 - 1),
       mode=st.sampled_from(["ns", "cv", "is"]), degrees=st.lists(st.integers(1, 40), min_size=1, max_size=2),
       batches=st.lists(st.integers(1, 120), min_size=1, max_size=4), pick=st.integers(0, 2**31 - 1))
def test_oracle_equals_compiled_reference(n, avg, gseed, seed, mode, degrees, batches, pick):
    g = random_graph(n, avg, gseed)
    cv, importance = mode == "cv", mode == "is"
    o = native.OracleSampler(g.data, g.indices, g.indptr, cv=cv, importance=importance)
    r = native.RefSampler(g.data, g.indices, g.indptr, cv=cv, importance=importance, L=len(degrees))
    o.seed(seed)
    r.seed(seed)
    rng = np.random.RandomState(pick)
    for b in batches:
        ids = rng.choice(n, size=min(b, n), replace=False).astype(np.int32)
        o.start_batch(ids)
        r.start_batch(ids)
        for d in degrees:
            rc_o, rc_r = o.expand(d), r.expand(d)
            assert (rc_o != 0) == (rc_r != 0)  # where the reference throws ("Prob is empty", "nan") the oracle errors
            if rc_o != 0:
                return
            so, sr = o.snapshot(), r.snapshot()
            for k in VEC_NAMES:
                assert_bits_equal(so[k], sr[k], k)
    assert_bits_equal(o.vec("adj_i"), r.vec("adj_i"), "permuted adj_i")
    assert_bits_equal(o.vec("adj_w"), r.vec("adj_w"), "permuted adj_w")
