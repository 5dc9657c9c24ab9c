"""CPU: pins the numeric (TensorFlow-side) oracle to the only reference-authored SpMM there is.

The reference evaluates its aggregators with TensorFlow ops that are not under /root/reference, so
oracle/aggregators.py is "parity unpinned" by rule.  The authors' own CSR aggregation loop, `compute_history`
(gcn/history.cpp:10-37), IS in the reference -- commented out.  oracle/Makefile un-comments exactly those lines
into oracle/_ref at build time; here the oracle's full-neighbour term dot(fadj, gather(history, ffield))
(gcn/layers.py:354-357) is checked against it: bit for bit for the fp32 storage-order port, to rounding for
the float64 reference value, with torch.sparse.mm in float64 as a third opinion."""
import numpy as np
import pytest
import torch

from oracle import aggregators as agg
from oracle import native
from tests.graphs_small import random_graph

needs_ref = pytest.mark.skipif(not native.have_ref(), reason="oracle/_ref not built (reference sources absent)")


def _case(seed, n=1500, avg=14, B=40, D=24, degree=2):
    g = random_graph(n, avg, seed, normalise="rand" if seed % 2 else "row")
    rng = np.random.RandomState(seed)
    hist = rng.standard_normal((n, D)).astype(np.float32)
    cls = native.RefSampler if native.have_ref() else native.OracleSampler
    o = cls(g.data, g.indices, g.indptr, cv=True)
    o.seed(seed)
    snaps = []
    for _ in range(3):                       # several batches: the stored rows get permuted in between
        ids = rng.choice(n, size=B, replace=False).astype(np.int32)
        o.start_batch(ids)
        assert o.expand(degree) == 0
        snaps.append((ids, o.snapshot(), o.vec("adj_i"), o.vec("adj_w")))
    return g, hist, snaps


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_full_neighbour_term_matches_the_references_own_csr_loop(seed):
    g, hist, snaps = _case(seed)
    adj_p = np.asarray(g.indptr, np.int32)
    for ids, s, adj_i, adj_w in snaps:
        B = len(ids)
        fadj = (np.stack([s["fedg_s"], s["fedg_t"]], 1).astype(np.int32), s["fedg_w"], (B, len(s["ffield"])))
        ref = native.ref_compute_history(adj_w, adj_i, adj_p, ids, hist)          # the reference's loop, fp32
        port = agg._coo_matmul(fadj, hist[s["ffield"]], np.float32)                # oracle, fp32 storage order
        assert np.array_equal(port.view(np.uint32), ref.view(np.uint32)), "fp32 port differs from compute_history"
        want = agg._coo_matmul(fadj, hist[s["ffield"]], np.float64)                # oracle, float64 reference value
        scale = np.abs(want).max()
        assert np.abs(ref - want).max() <= 2e-6 * scale
        # third opinion: torch.sparse.mm in float64 over the same COO
        a = torch.sparse_coo_tensor(torch.from_numpy(fadj[0].T.astype(np.int64)),
                                    torch.from_numpy(fadj[1].astype(np.float64)), size=fadj[2])
        third = torch.sparse.mm(a, torch.from_numpy(hist[s["ffield"]].astype(np.float64))).numpy()
        assert np.abs(third - want).max() <= 1e-12 * scale


@needs_ref
def test_cv_forward_oracle_is_compute_history_plus_the_sampled_terms():
    """whole CV estimator (gcn/layers.py:350-362) with its dominant term replaced by the reference's loop"""
    g, hist, snaps = _case(7)
    adj_p = np.asarray(g.indptr, np.int32)
    rng = np.random.RandomState(0)
    for ids, s, adj_i, adj_w in snaps:
        B, n_in = len(ids), len(s["field"])
        x = rng.standard_normal((n_in, hist.shape[1])).astype(np.float32)
        adj = (np.stack([s["edg_s"], s["edg_t"]], 1).astype(np.int32), s["edg_w"], (B, n_in))
        fadj = (np.stack([s["fedg_s"], s["fedg_t"]], 1).astype(np.int32), s["fedg_w"], (B, len(s["ffield"])))
        out, _ = agg.cv_forward(adj, fadj, s["field"], s["ffield"], hist, x, False)
        mean = native.ref_compute_history(adj_w, adj_i, adj_p, ids, hist).astype(np.float64)
        cur = agg._coo_matmul(adj, x, np.float64) - agg._coo_matmul(adj, hist[s["field"]], np.float64)
        assert np.abs(out - (cur + mean)).max() <= 2e-6 * max(np.abs(out).max(), 1.0)
