"""CPU: the host side of PyScheduler (feed-dict assembly, minibatch cursor, hop order) over a stand-in for the
device sampler -- the pinned C oracle behind the DeviceSampler methods PyScheduler uses -- against the oracle's own
restatement of gcn/_scheduler.pyx:28-151.  (The same comparison with the real device sampler is a GPU test.)"""
import numpy as np
import pytest

from oracle import native
from oracle.pyscheduler import OraclePyScheduler, default_placeholders
from tests.graphs_small import random_graph


class HostSampler:
    """start_batch / expand / snapshot of DeviceSampler, served by the CPU oracle"""

    def __init__(self, g, cv, seed):
        self.o = native.OracleSampler(g.data, g.indices, g.indptr, cv=cv)
        self.o.seed(seed)

    def start_batch(self, ids):
        self.o.start_batch(ids)

    def expand(self, degree, materialize_full=False):
        self.o.expand(degree)

    def snapshot(self):
        return self.o.snapshot()


@pytest.mark.parametrize("cv", [False, True])
@pytest.mark.parametrize("degrees", [[1, 2], [3, 1]])
def test_feed_dict_layout_matches_the_reference_restatement(cv, degrees):
    from stochastic_gcn_b200.scheduler import PyScheduler
    g = random_graph(300, 8, 5)
    labels = np.arange(900, dtype=np.float32).reshape(300, 3)
    ph = default_placeholders(2)
    a = PyScheduler.__new__(PyScheduler)                      # no GPU: the sampler is injected
    a.c_sch, a.labels, a.data, a.degrees, a.L = HostSampler(g, cv, 3), labels, np.arange(40, dtype=np.int32), degrees, 2
    a.start, a.placeholders, a.t, a.cv = 0, ph, 0, cv
    b = OraclePyScheduler(g, labels, 2, degrees, ph, 3, data=np.arange(40, dtype=np.int32), cv=cv)
    seen = 0
    for _ in range(4):                                        # 16 + 16 + 8 ids, then None
        fa, fb = a.minibatch(16), b.minibatch(16)
        if fb is None:
            assert fa is None
            continue
        assert set(fa) == set(fb)
        for key, want in fb.items():
            got = fa[key]
            if isinstance(want, tuple):                       # COO triple: (int32 [ne, 2], float32 [ne], shape)
                assert got[0].dtype == np.int32 and got[0].ndim == 2 and got[0].shape[1] == 2
                for x, y in zip(got, want):
                    assert np.array_equal(np.asarray(x), np.asarray(y)), key
            else:
                assert got.dtype == want.dtype and np.array_equal(got, want), key
        seen += 1
    assert seen == 3 and a.get_t() == 0
    a.shuffle()
    assert a.start == 0 and sorted(a.data.tolist()) == list(range(40))
