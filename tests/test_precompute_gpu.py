"""GPU: the pre-processing product A_hat @ X over the whole graph (gcn/utils.py:168-169,321-322,
gcn/models.py:230-239) through the C ABI vs SciPy -- the library the reference itself calls --
in float64 (1e-4 of scale, north_star's bound) and in float32 (the reference's own dtype)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import aggregators as agg
from tests.graphs_small import random_graph

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, want64, what, rtol=1e-4):
    got = got.detach().cpu().numpy().astype(np.float64)
    scale = max(np.abs(want64).max(), 1e-30)
    err = np.abs(got - want64).max() / scale
    assert err <= rtol, "%s: max error %.3e of scale" % (what, err)


def skewed_csr(n, seed, hubs=3, hub_deg=4000, deg=12):
    """power-law-ish rows: a few hub rows, many short rows, some empty rows (also at both ends)"""
    rng = np.random.RandomState(seed)
    lens = rng.poisson(deg, n)
    lens[rng.choice(n, hubs, replace=False)] = min(hub_deg, n)
    lens[rng.choice(n, n // 10, replace=False)] = 0
    lens[0] = 0
    lens[-1] = 0
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(lens)
    indices = np.concatenate([rng.choice(n, k, replace=False) for k in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    data = rng.rand(indptr[-1]).astype(np.float32)
    return sp.csr_matrix((data, indices, indptr), shape=(n, n))


@pytest.mark.parametrize("d,tile", [(602, 0), (602, 64), (128, 0), (256, 256), (500, 128), (33, 0), (7, 0), (2, 0)])
def test_csr_spmm_matches_scipy(d, tile):
    from stochastic_gcn_b200 import ops
    a = skewed_csr(5000, d)
    rng = np.random.RandomState(d + 1)
    x = rng.randn(5000, d).astype(np.float32)
    want = agg.preprocess_features(a, x, False)
    got = ops.csr_spmm(dev(a.indptr), dev(a.indices), dev(a.data), dev(x), tile_cols=tile)
    close(got, want, "A @ X d=%d" % d)
    close(got, a.dot(x).astype(np.float64), "vs scipy float32 (the reference's own call)", rtol=2e-5)


@pytest.mark.parametrize("normalization", ["graphsage", "gcn"])
def test_preprocess_features_layout(normalization):
    """[X | A X] for graphsage (X copied bit-exactly into the left half), A X for gcn"""
    from stochastic_gcn_b200 import ops
    g = random_graph(3000, 20, 5)
    a = sp.csr_matrix((g.data, g.indices, g.indptr), shape=(3000, 3000))
    x = np.random.RandomState(2).randn(3000, 602).astype(np.float32)
    want = agg.preprocess_features(a, x, normalization == "graphsage")
    got = ops.preprocess_features(dev(g.indptr), dev(g.indices), dev(g.data), dev(x), normalization)
    assert tuple(got.shape) == want.shape
    close(got, want, "PP input")
    if normalization == "graphsage":
        assert np.array_equal(got[:, :602].cpu().numpy(), x)


def test_csr_spmm_empty_matrix_and_views():
    from stochastic_gcn_b200 import ops
    n = 257
    indptr = np.zeros(n + 1, np.int32)
    y = ops.csr_spmm(dev(indptr), dev(np.zeros(0, np.int32)), dev(np.zeros(0, np.float32)),
                     torch.randn((n, 16), device="cuda"))
    assert float(y.abs().max()) == 0.0
    # x and y as column blocks of wider matrices (unaligned start: scalar path)
    a = skewed_csr(1000, 3, hubs=1, hub_deg=500)
    wide = np.random.RandomState(0).randn(1000, 50).astype(np.float32)
    xt = dev(wide)
    out = torch.full((1000, 70), 7.0, device="cuda")
    ops.csr_spmm(dev(a.indptr), dev(a.indices), dev(a.data), xt[:, 3:40], out=out[:, 5:42])
    close(out[:, 5:42], agg.preprocess_features(a, wide[:, 3:40], False), "column views")
    assert float(out[:, :5].min()) == 7.0 and float(out[:, 42:].max()) == 7.0
