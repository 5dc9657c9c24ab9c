"""GPU: trains of batches (sgcn_sampler_expand_train: n batches sampled by ONE launch, one thread block per
batch) must leave exactly what n sequential start_batch + expand calls of the reference's Scheduler leave
(gcn/scheduler.cpp:41-61,125-189): every output vector of every batch, the permuted adjacency, the mt19937
state -- also when batches of one train, or of consecutive trains, share nodes."""
import numpy as np
import pytest
import torch

from oracle import native
from tests.conftest import assert_bits_equal
from tests.graphs_small import random_graph

pytestmark = pytest.mark.gpu

LEVEL_VECS = ["field", "edg_s", "edg_t", "edg_w", "scales"]


def make(g, cv, seed):
    from stochastic_gcn_b200.sampler import DeviceSampler
    s = DeviceSampler(g.data, g.indices, g.indptr, L=1, cv=cv)
    s.seed(seed)
    return s


def check_set(s, o_snap, cv, what):
    z = s.sizes()
    assert z.n_in == len(o_snap["field"]) and z.nnz_s == len(o_snap["edg_s"]), what + " sizes"
    for v in LEVEL_VECS + (["medg_w"] if cv else []):
        n = {"field": z.n_in, "scales": z.n_out}.get(v, z.nnz_s)
        assert_bits_equal(s.host(v, n), o_snap[v], what + " " + v)
    if cv:
        assert z.nnz_f == len(o_snap["fedg_s"]), what + " nnz_f"
        rowptr_f = s.host("rowptr_f", z.n_out + 1)
        want = np.searchsorted(o_snap["fedg_s"], np.arange(z.n_out + 1)).astype(np.int32)
        assert np.array_equal(rowptr_f, want), what + " rowptr_f"
    tgt = s.host("tgt", z.nnz_s)
    assert np.array_equal(tgt, o_snap["field"][o_snap["edg_t"]]), what + " tgt"


@pytest.mark.parametrize("cv", [False, True])
@pytest.mark.parametrize("degree", [1, 2, 5])
def test_trains_match_sequential_oracle(cv, degree):
    n, B, n_sets = 3000, 64, 12
    g = random_graph(n, 12, 77)
    o = native.OracleSampler(g.data, g.indices, g.indptr, cv=cv)
    s = make(g, cv, 21)
    o.seed(21)
    s.reserve_sets(n_sets, B, degree)
    rng = np.random.RandomState(degree)
    first = 0
    total_draws = 0
    for train, length in enumerate((5, 1, 6, 3, 6)):
        # disjoint batches (an epoch-style shuffle) ...
        perm = rng.permutation(n).astype(np.int32)
        table = perm[:length * B].reshape(length, B).copy()
        if train == 2:          # ... except here: batches 3 and 5 re-use nodes of batches 0 and 3 of the same train
            table[3, :10] = table[0, 5:15]
            table[5, 20:30] = table[3, 40:50]
            table[5, :5] = table[0, 50:55]
        t = torch.from_numpy(table).cuda()
        s.expand_train(t, first_set=first)
        torch.cuda.synchronize()
        for j in range(length):
            o.start_batch(table[j])
            assert o.expand(degree) == 0
            snap = o.snapshot()
            total_draws += len(snap["edg_s"])
            s.set_slot((first + j) % n_sets)
            check_set(s, snap, cv, "train %d batch %d" % (train, j))
        first = (first + length) % n_sets
    assert_bits_equal(s.host("adj_i", s.num_edges), o.vec("adj_i"), "permuted adj_i")
    assert_bits_equal(s.host("adj_w", s.num_edges), o.vec("adj_w"), "permuted adj_w")
    # the engine is where std::mt19937 is after exactly that many draws
    mt = native.MT19937(21)
    for _ in range(total_draws):
        mt.next()
    state, pos = s.get_rng()
    assert pos == mt._s.pos and np.array_equal(state, np.ctypeslib.as_array(mt._s.x))
    # and a plain expand continues the same stream / permutation
    ids = rng.choice(n, size=50, replace=False).astype(np.int32)
    s.set_slot(0)
    o.start_batch(ids); s.start_batch(ids)
    o.expand(degree); s.expand(degree)
    assert_bits_equal(s.snapshot()["edg_t"], o.snapshot()["edg_t"], "plain expand after the trains")
    s.close()


def test_train_shapes_of_the_bench():
    """batch 512, degree 2, 16 batches per launch on a graph with high-degree rows; two consecutive trains."""
    n, B, degree = 20000, 512, 2
    g = random_graph(n, 60, 5)
    o = native.OracleSampler(g.data, g.indices, g.indptr, cv=True)
    s = make(g, True, 1)
    o.seed(1)
    s.reserve_sets(32, B, degree)
    rng = np.random.RandomState(0)
    perm = rng.permutation(n).astype(np.int32)
    for train in range(2):
        table = perm[train * 16 * B:(train + 1) * 16 * B].reshape(16, B).copy()
        s.expand_train(torch.from_numpy(table).cuda(), first_set=16 * train)
        torch.cuda.synchronize()
        for j in range(16):
            o.start_batch(table[j])
            o.expand(degree)
            s.set_slot(16 * train + j)
            check_set(s, o.snapshot(), True, "train %d batch %d" % (train, j))
    assert_bits_equal(s.host("adj_i", s.num_edges), o.vec("adj_i"), "permuted adj_i")
    s.close()


def test_train_waits_for_consumers_of_the_previous_train():
    """Pipeline mode: a batch that shares a node with the previous train must not permute that row before the
    previous train's consumer passes are marked finished.  As in the step drivers, the marks are work that is
    already queued when the train is launched (here: on another stream, behind a ~1 ms sleep kernel), so the
    train's thread block really waits on the device -- and the permuted rows come out as the sequential
    reference leaves them."""
    n, B, degree = 3000, 64, 2
    g = random_graph(n, 12, 78)
    o = native.OracleSampler(g.data, g.indices, g.indptr, cv=True)
    s = make(g, True, 9)
    o.seed(9)
    s.reserve_sets(8, B, degree)
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    s.mark_consumed(b)                          # first launch of the mark kernel (module load) outside the test
    torch.cuda.synchronize()
    s.pipeline(True)                            # arms the guard, counters back to zero
    rng = np.random.RandomState(3)
    perm = rng.permutation(n).astype(np.int32)
    t0 = perm[:4 * B].reshape(4, B).copy()
    t1 = perm[4 * B:8 * B].reshape(4, B).copy()
    t1[2, :8] = t0[1, :8]                       # shared with the previous train
    d0, d1 = torch.from_numpy(t0).cuda(), torch.from_numpy(t1).cuda()
    torch.cuda.synchronize()
    s.expand_train(d0, first_set=0, stream=a)
    with torch.cuda.stream(b):                  # (not the legacy default stream: it would wait for stream a)
        torch.cuda._sleep(2_000_000)            # the marks land ~1 ms later
    for _ in range(4):
        s.mark_consumed(b)
    s.expand_train(d1, first_set=4, prev=d0, stream=a)
    torch.cuda.synchronize()
    for j, ids in enumerate(list(t0) + list(t1)):
        o.start_batch(ids)
        o.expand(degree)
        s.set_slot(j)
        check_set(s, o.snapshot(), True, "batch %d" % j)
    assert_bits_equal(s.host("adj_i", s.num_edges), o.vec("adj_i"), "permuted adj_i")
    s.close()
