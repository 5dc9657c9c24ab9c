"""GPU: the BASELINE.json configurations as parity cases (synthetic graphs of the named shapes).

configs[0]  Cora-shaped, Exact (degree >= max degree), GCN normalisation, 1433-d features
configs[1]  PubMed-shaped, CVD+PP degree 1, GCN normalisation
(configs[2..4] are the bench workloads; their kernels are exercised at reduced scale in
test_step_gpu.py and at full scale, through size-independent properties, below)
"""
import numpy as np
import pytest
import torch

from oracle import aggregators as agg
from oracle import native

pytestmark = pytest.mark.gpu


def close(got, want, what, tol=1e-4):
    err = np.abs(np.asarray(got, np.float64) - want).max() / max(np.abs(want).max(), 1e-30)
    assert err <= tol, "%s: %.3e" % (what, err)


def test_cora_shaped_exact_equals_full_aggregation():
    """Exact = every neighbour sampled (degree 20 in the recipe; >= max degree here), scale = 1:
    the aggregate must equal rows of A_hat @ X computed by SciPy, and the 1433-wide gather (not a
    multiple of 4 floats: scalar fallback path) must be bit-exact."""
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    g = graphs.make_shape("cora", seed=0, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(1)
    feats = torch.rand((g.n, 1433), generator=gen, device="cuda")
    B, D = 140, 32
    step = HotPathStep(g, feats, D, B, 10_000, mode="ns", normalization="gcn", seed=1)
    ids = torch.randperm(g.n, generator=gen, device="cuda")[:B].to(torch.int32)
    out = step.run(ids).cpu().numpy()
    z = step.sizes()
    A = g.to_scipy()
    X = feats.cpu().numpy()
    want = (A[ids.cpu().numpy()].astype(np.float64) @ X[:, :D].astype(np.float64))
    close(out, np.asarray(want), "exact aggregate")
    field = step.sampler.host("field", z["n_in"])
    assert np.array_equal(step.x0.cpu().numpy()[:z["n_in"]], X[field])            # bit-exact gather, C = 1433
    assert z["nnz_s"] == int(np.diff(A.indptr)[ids.cpu().numpy()].sum())          # every stored neighbour taken
    assert np.allclose(step.sampler.host("scales", B), 1.0)


def test_pubmed_shaped_cvd_pp_matches_oracle():
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    g = graphs.make_shape("pubmed", seed=0, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(2)
    feats = torch.randn((g.n, 500), generator=gen, device="cuda")
    B, D, deg = 60, 32, 1
    step = HotPathStep(g, feats, D, B, deg, mode="cvd", normalization="gcn", seed=1)
    step.history.normal_(generator=gen)
    step.d_out.normal_(generator=gen)
    hist = step.history.cpu().numpy().copy()
    o = native.OracleSampler(g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy(), cv=True)
    o.seed(1)
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()
    perm = torch.randperm(g.n, generator=gen, device="cuda").to(torch.int32)
    batches = [perm[i * B:(i + 1) * B].contiguous() for i in range(6)]
    step.run_native(torch.stack(batches))
    torch.cuda.synchronize()
    for ids in batches:
        o.start_batch(ids.cpu().numpy()); o.expand(deg)
        s = o.snapshot()
        n_in = len(s["field"])
        x0 = fh[s["field"]]
        adj = (np.stack([s["edg_s"], s["edg_t"]], 1).astype(np.int32), s["edg_w"], (B, n_in))
        fadj = (np.stack([s["fedg_s"], s["fedg_t"]], 1).astype(np.int32), s["fedg_w"], (B, len(s["ffield"])))
        (oh, om), new = agg.cvd_forward(adj, fadj, s["field"], s["ffield"], hist, s["scales"], x0[:, :D],
                                        x0[:, D:2 * D], False)
        dx = agg.cvd_backward_h(adj, s["scales"], d_out, n_in, False)
        agg.history_update(hist, s["field"], new[0])
    close(step.out.cpu().numpy(), oh, "pubmed cvd h")
    close(step.out_mu.cpu().numpy(), om, "pubmed cvd mu")
    close(step.dx.cpu().numpy()[:n_in], dx, "pubmed cvd dx")
    assert np.array_equal(step.history.cpu().numpy(), hist)


def test_full_size_reddit_shape_properties():
    """BASELINE's full size (233k nodes, ~1e8 stored entries) through properties that need no oracle:
    linearity of the CV aggregate in (X, history), exactness when history == activations, and the
    write-back round trip."""
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    g = graphs.make_shape("reddit", seed=0, device="cuda", scale=0.25)      # 58k nodes, 29M entries: seconds
    gen = torch.Generator(device="cuda").manual_seed(3)
    D, B = 128, 512
    feats = torch.randn((g.n, 2 * D), generator=gen, device="cuda")
    step = HotPathStep(g, feats, D, B, 2, mode="cv", seed=1)
    # history == current activations everywhere  =>  CV estimate == exact full-neighbour mean of X
    step.history.copy_(feats[:, :D])
    ids = torch.randperm(g.n, generator=gen, device="cuda")[:B].to(torch.int32)
    out = step.run(ids)
    z = step.sizes()
    indptr, indices, data = g.indptr.long(), g.indices.long(), g.data
    rows = torch.repeat_interleave(torch.arange(B, device="cuda"), (indptr[1:] - indptr[:-1])[ids.long()])
    pos = torch.cat([torch.arange(int(indptr[i]), int(indptr[i + 1]), device="cuda") for i in ids.tolist()])
    exact = torch.zeros((B, D), dtype=torch.float64, device="cuda")
    exact.index_add_(0, rows, data[pos].double()[:, None] * feats[indices[pos], :D].double())
    err = (out[:, D:].double() - exact).abs().max() / exact.abs().max()
    assert float(err) < 1e-4, float(err)
    assert torch.equal(out[:, :D], feats[ids.long(), :D])                   # self half = own rows
    # write-back: history rows of the input field now hold the layer input (unchanged here by construction)
    field = step.sampler.view("field", count=z["n_in"]).long()
    assert torch.equal(step.history[field], feats[field, :D])
    assert z["nnz_f"] == int((indptr[1:] - indptr[:-1])[ids.long()].sum())
