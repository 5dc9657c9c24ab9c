"""GPU: aggregation kernels through the C ABI vs the oracle (NumPy float64 reference value and the
fp32 storage-order C restatement).  Tolerance: north_star's 1e-4 relative, measured against the
magnitude of the float64 result (the CV estimator subtracts nearly equal terms, so an elementwise
relative bound is ill-posed at the cancelling entries)."""
import numpy as np
import pytest
import torch

from oracle import aggregators as agg
from oracle import native
from tests.graphs_small import random_graph

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, want64, what):
    got = got.detach().cpu().numpy().astype(np.float64)
    scale = max(np.abs(want64).max(), 1e-30)
    err = np.abs(got - want64).max() / scale
    assert err <= RTOL, "%s: max error %.3e of scale" % (what, err)


def sample(g, ids, degree, cv, seed=3):
    """reference-format pieces from the CPU oracle sampler"""
    s = native.OracleSampler(g.data, g.indices, g.indptr, cv=cv)
    s.seed(seed)
    s.start_batch(ids)
    s.expand(degree)
    z = s.snapshot()
    n_out, n_in = len(ids), len(z["field"])
    adj = (np.stack([z["edg_s"], z["edg_t"]], 1).astype(np.int32), z["edg_w"], (n_out, n_in))
    fadj = None
    if cv:
        fadj = (np.stack([z["fedg_s"], z["fedg_t"]], 1).astype(np.int32), z["fedg_w"], (n_out, len(z["ffield"])))
    return z, adj, fadj, s


@pytest.mark.parametrize("d", [4, 32, 64, 100, 128, 256, 512, 1100, 30, 7])
@pytest.mark.parametrize("graphsage", [False, True])
def test_plain_forward_backward(d, graphsage):
    from stochastic_gcn_b200.layers import DeviceAdj, PlainAggregator
    g = random_graph(600, 10, d)
    rng = np.random.RandomState(d)
    ids = rng.choice(600, size=150, replace=False).astype(np.int32)
    z, adj, _, _ = sample(g, ids, 3, False)
    x = rng.randn(len(z["field"]), d).astype(np.float32)
    want = agg.plain_forward(adj, x, graphsage)
    xt = dev(x).requires_grad_(True)
    layer = PlainAggregator(DeviceAdj.from_coo(adj), normalization="graphsage" if graphsage else "gcn")
    out = layer(xt)
    close(out, want, "plain fwd")
    close(out, agg.plain_forward(adj, x, graphsage, dtype=np.float32).astype(np.float64), "plain fwd vs fp32 port")
    dy = rng.randn(*want.shape).astype(np.float32)
    out.backward(dev(dy))
    close(xt.grad, agg.plain_backward(adj, dy, x.shape[0], graphsage), "plain bwd")


@pytest.mark.parametrize("d", [32, 128, 256, 100, 6])
@pytest.mark.parametrize("graphsage", [False, True])
@pytest.mark.parametrize("in_place", [False, True])
def test_cv_forward_backward(d, graphsage, in_place):
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, VRAggregator
    g = random_graph(700, 25, 100 + d)
    rng = np.random.RandomState(d)
    ids = rng.choice(700, size=120, replace=False).astype(np.int32)
    z, adj, fadj, s = sample(g, ids, 2, True)
    n_in = len(z["field"])
    hist = rng.randn(700, d).astype(np.float32)
    x = (hist[z["field"]] + 0.1 * rng.randn(n_in, d)).astype(np.float32)   # activations close to history
    want, new_hist = agg.cv_forward(adj, fadj, z["field"], z["ffield"], hist, x, graphsage)
    hist_t = dev(hist)
    ifield = dev(z["field"])
    if in_place:   # rows read straight from the (permuted) CSR, nothing materialised
        rowptr_f = np.concatenate([[0], np.cumsum(np.diff(g.indptr)[ids])]).astype(np.int32)
        full = FullNeighbours.in_place(dev(ids), dev(rowptr_f), dev(s.vec("adj_p")), dev(s.vec("adj_i")),
                                       dev(s.vec("adj_w")))
    else:
        full = FullNeighbours.from_coo(fadj, z["ffield"])
    layer = VRAggregator(DeviceAdj.from_coo(adj), full, None, ifield, None, [hist_t], None, False,
                         normalization="graphsage" if graphsage else "gcn")
    xt = dev(x).requires_grad_(True)
    out = layer(xt)
    close(out, want, "cv fwd")
    dy = rng.randn(*want.shape).astype(np.float32)
    out.backward(dev(dy))
    close(xt.grad, agg.plain_backward(adj, dy, n_in, graphsage), "cv bwd")
    # write-back: history rows of the input field are overwritten by the layer input
    layer.write_back()
    want_hist = agg.history_update(hist.copy(), z["field"], x)
    assert np.array_equal(hist_t.cpu().numpy(), want_hist)


@pytest.mark.parametrize("d", [32, 128, 20])
@pytest.mark.parametrize("graphsage", [False, True])
def test_cvd_forward_backward(d, graphsage):
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, VRAggregator
    g = random_graph(500, 20, 200 + d)
    rng = np.random.RandomState(d)
    ids = rng.choice(500, size=90, replace=False).astype(np.int32)
    z, adj, fadj, s = sample(g, ids, 1, True)
    n_in = len(z["field"])
    hist = rng.randn(500, d).astype(np.float32)
    mu = (hist[z["field"]] + 0.1 * rng.randn(n_in, d)).astype(np.float32)
    h = (mu + 0.3 * rng.randn(n_in, d)).astype(np.float32)
    (want_h, want_mu), _ = agg.cvd_forward(adj, fadj, z["field"], z["ffield"], hist, z["scales"], h, mu, graphsage)
    hist_t = dev(hist)
    layer = VRAggregator(DeviceAdj.from_coo(adj), FullNeighbours.from_coo(fadj, z["ffield"]), None,
                         dev(z["field"]), None, [hist_t], dev(z["scales"]), True,
                         normalization="graphsage" if graphsage else "gcn")
    ht, mut = dev(h).requires_grad_(True), dev(mu)
    out_h, out_mu = layer((ht, mut))
    close(out_h, want_h, "cvd h"); close(out_mu, want_mu, "cvd mu")
    dy = rng.randn(*want_h.shape).astype(np.float32)
    out_h.backward(dev(dy))
    close(ht.grad, agg.cvd_backward_h(adj, z["scales"], dy, n_in, graphsage), "cvd bwd")
    layer.write_back()
    assert np.array_equal(hist_t.cpu().numpy(), agg.history_update(hist.copy(), z["field"], mu))


def test_cvd_mu_gradient_against_torch_autograd():
    """d/d mu (unused by the reference, stop_gradient at gcn/layers.py:412) vs a dense torch fp32 model"""
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, VRAggregator
    g = random_graph(300, 15, 9)
    rng = np.random.RandomState(0)
    ids = rng.choice(300, size=40, replace=False).astype(np.int32)
    z, adj, fadj, _ = sample(g, ids, 2, True)
    n_in, d = len(z["field"]), 32
    hist = rng.randn(300, d).astype(np.float32)
    h, mu = rng.randn(n_in, d).astype(np.float32), rng.randn(n_in, d).astype(np.float32)
    A = torch.zeros(len(ids), n_in, dtype=torch.float64)
    A.index_put_((torch.from_numpy(adj[0][:, 0]).long(), torch.from_numpy(adj[0][:, 1]).long()),
                 torch.from_numpy(adj[1]).double(), accumulate=True)
    hr, mr = torch.from_numpy(h).double().requires_grad_(True), torch.from_numpy(mu).double().requires_grad_(True)
    sc = torch.from_numpy(z["scales"]).double()[:, None]
    mu_nb = A @ (mr - torch.from_numpy(hist[z["field"]]).double())
    h_nb = (A @ (hr - mr)) * sc + mu_nb
    gy_h, gy_mu = torch.randn(len(ids), d, dtype=torch.float64), torch.randn(len(ids), d, dtype=torch.float64)
    ((h_nb * gy_h).sum() + (mu_nb * gy_mu).sum()).backward()
    layer = VRAggregator(DeviceAdj.from_coo(adj), FullNeighbours.from_coo(fadj, z["ffield"]), None,
                         dev(z["field"]), None, [dev(hist)], dev(z["scales"]), True)
    ht, mt = dev(h).requires_grad_(True), dev(mu).requires_grad_(True)
    oh, om = layer((ht, mt))
    ((oh * gy_h.float().cuda()).sum() + (om * gy_mu.float().cuda()).sum()).backward()
    close(ht.grad, hr.grad.numpy(), "dh"); close(mt.grad, mr.grad.numpy(), "dmu")


@pytest.mark.parametrize("d", [32, 128, 20, 260])
@pytest.mark.parametrize("graphsage", [False, True])
def test_plain_det_dropout_pair(d, graphsage):
    """PlainAggregator on a (mu, var) pair: adj @ mu and tf.square(adj) @ var (gcn/layers.py:238-247)"""
    from stochastic_gcn_b200.layers import DeviceAdj, PlainAggregator
    g = random_graph(400, 12, 300 + d)
    rng = np.random.RandomState(d)
    ids = rng.choice(400, size=70, replace=False).astype(np.int32)
    z, adj, _, _ = sample(g, ids, 3, False)
    n_in = len(z["field"])
    mu = rng.randn(n_in, d).astype(np.float32)
    var = (rng.rand(n_in, d) + 0.05).astype(np.float32)
    want_mu, want_var = agg.plain_forward_det(adj, mu, var, graphsage)
    layer = PlainAggregator(DeviceAdj.from_coo(adj), normalization="graphsage" if graphsage else "gcn")
    mt, vt = dev(mu).requires_grad_(True), dev(var).requires_grad_(True)
    om, ov = layer((mt, vt))
    close(om, want_mu, "plain det mu"); close(ov, want_var, "plain det var")
    gm, gv = rng.randn(*want_mu.shape).astype(np.float32), rng.randn(*want_var.shape).astype(np.float32)
    ((om * dev(gm)).sum() + (ov * dev(gv)).sum()).backward()
    close(mt.grad, agg.plain_backward(adj, gm, n_in, graphsage), "plain det dmu")
    close(vt.grad, agg.plain_backward(agg._sq(adj), gv, n_in, graphsage), "plain det dvar")


@pytest.mark.parametrize("d", [32, 128, 20, 260])
@pytest.mark.parametrize("graphsage", [False, True])
@pytest.mark.parametrize("in_place", [False, True])
def test_vr_det_dropout_forward_backward(d, graphsage, in_place):
    """VRAggregator det-dropout branch (gcn/layers.py:320-349): two histories, madj, relu + 1e-10"""
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, VRAggregator
    g = random_graph(500, 18, 400 + d)
    rng = np.random.RandomState(d + 1)
    ids = rng.choice(500, size=80, replace=False).astype(np.int32)
    z, adj, fadj, s = sample(g, ids, 2, True)
    n_in = len(z["field"])
    madj = (adj[0], z["medg_w"], adj[2])
    mu_hist = rng.randn(500, d).astype(np.float32)
    var_hist = (rng.rand(500, d) + 0.05).astype(np.float32)
    mu = (mu_hist[z["field"]] + 0.1 * rng.randn(n_in, d)).astype(np.float32)
    var = (var_hist[z["field"]] * (0.5 + rng.rand(n_in, d))).astype(np.float32)
    (want_mu, want_var), _, pre = agg.det_forward(adj, fadj, madj, z["field"], z["ffield"], mu_hist, var_hist,
                                                  mu, var, graphsage)
    mh, vh = dev(mu_hist), dev(var_hist)
    if in_place:
        rowptr_f = np.concatenate([[0], np.cumsum(np.diff(g.indptr)[ids])]).astype(np.int32)
        full = FullNeighbours.in_place(dev(ids), dev(rowptr_f), dev(s.vec("adj_p")), dev(s.vec("adj_i")),
                                       dev(s.vec("adj_w")))
    else:
        full = FullNeighbours.from_coo(fadj, z["ffield"])
    layer = VRAggregator(DeviceAdj.from_coo(adj), full, madj, dev(z["field"]), None, [mh, vh], None, False,
                         normalization="graphsage" if graphsage else "gcn")
    mt, vt = dev(mu).requires_grad_(True), dev(var).requires_grad_(True)
    om, ov = layer((mt, vt))
    close(om, want_mu, "det mu"); close(ov, want_var, "det var")
    assert float(ov.min()) >= 1e-10
    gm, gv = rng.randn(*want_mu.shape).astype(np.float32), rng.randn(*want_var.shape).astype(np.float32)
    ((om * dev(gm)).sum() + (ov * dev(gv)).sum()).backward()
    close(mt.grad, agg.plain_backward(adj, gm, n_in, graphsage), "det dmu")
    close(vt.grad, agg.det_backward_var(adj, madj, z["field"], var_hist, var, pre, gv, graphsage), "det dvar")
    layer.write_back()
    assert np.array_equal(mh.cpu().numpy(), agg.history_update(mu_hist.copy(), z["field"], mu))
    assert np.array_equal(vh.cpu().numpy(), agg.history_update(var_hist.copy(), z["field"], var))


def test_vr_det_dropout_relu_clamps():
    """entries whose control-variate sum is negative are clamped to 1e-10 and pass no gradient.  With the
    sampler's own madj the sum is a sum of squares (w ds + fw sigma_bar)^2 >= 0, so the clamp only ever
    catches rounding; the test feeds a madj three times too large to drive it negative on purpose."""
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, VRAggregator
    g = random_graph(300, 10, 77)
    rng = np.random.RandomState(4)
    ids = rng.choice(300, size=50, replace=False).astype(np.int32)
    z, adj, fadj, s = sample(g, ids, 2, True)
    n_in, d = len(z["field"]), 32
    madj = (adj[0], (3.0 * z["medg_w"]).astype(np.float32), adj[2])
    mu_hist = rng.randn(300, d).astype(np.float32)
    var_hist = np.full((300, d), 1e-4, np.float32)
    var_hist[z["field"]] = 4.0                              # big sigma_bar on the sampled rows only
    mu = mu_hist[z["field"]].copy()
    var = np.full((n_in, d), 0.01, np.float32)              # sigma << sigma_bar: the cross term is negative
    (want_mu, want_var), _, pre = agg.det_forward(adj, fadj, madj, z["field"], z["ffield"], mu_hist, var_hist,
                                                  mu, var, False)
    assert (pre < 0).any() and (pre > 0).any()
    layer = VRAggregator(DeviceAdj.from_coo(adj), FullNeighbours.from_coo(fadj, z["ffield"]), madj,
                         dev(z["field"]), None, [dev(mu_hist), dev(var_hist)], None, False)
    vt = dev(var).requires_grad_(True)
    om, ov = layer((dev(mu), vt))
    close(ov, want_var, "clamped var")
    gv = rng.randn(*want_var.shape).astype(np.float32)
    (ov * dev(gv)).sum().backward()
    close(vt.grad, agg.det_backward_var(adj, madj, z["field"], var_hist, var, pre, gv, False), "clamped dvar")


def test_coo_kernel_and_gather_layer():
    from stochastic_gcn_b200 import ops
    from stochastic_gcn_b200.layers import GatherAggregator
    rng = np.random.RandomState(5)
    nnz, n_r, n_c, d = 5000, 300, 400, 64
    idx = np.stack([rng.randint(0, n_r, nnz), rng.randint(0, n_c, nnz)], 1).astype(np.int32)   # unsorted
    val = rng.randn(nnz).astype(np.float32)
    x = rng.randn(n_c, d).astype(np.float32)
    want = agg._coo_matmul((idx, val, (n_r, n_c)), x, np.float64)
    close(ops.spmm_coo(dev(idx), dev(val), dev(x), n_r), want, "coo")
    dy = rng.randn(n_r, d).astype(np.float32)
    want_t = agg._coo_matmul_t((idx, val, (n_r, n_c)), dy, np.float64)
    close(ops.spmm_coo(dev(idx), dev(val), dev(dy), n_c, transpose=True), want_t, "coo^T")
    field = dev(rng.randint(0, n_c, 77).astype(np.int32))
    xt = dev(x).requires_grad_(True)
    out = GatherAggregator(field)(xt)
    assert np.array_equal(out.detach().cpu().numpy(), x[field.cpu().numpy()])
    out.backward(torch.ones_like(out))
    want_g = np.zeros_like(x); np.add.at(want_g, field.cpu().numpy(), 1.0)
    close(xt.grad, want_g.astype(np.float64), "gather bwd")


def test_full_mean_skewed_rows_and_empty_rows():
    """edge-balanced kernel: one 40k-entry row next to empty and single-entry rows, D=128 and D=32"""
    from stochastic_gcn_b200 import ops
    rng = np.random.RandomState(11)
    n = 50_000
    deg = np.array([40_000, 0, 1, 0, 0, 700, 3, 0, 64, 65, 31, 0], dtype=np.int64)
    indptr = np.zeros(len(deg) + 1, np.int32); indptr[1:] = np.cumsum(deg)
    cols = np.concatenate([rng.choice(n, size=k, replace=False) for k in deg]).astype(np.int32)
    w = rng.rand(indptr[-1]).astype(np.float32)
    nodes = np.arange(len(deg), dtype=np.int32)
    for d in (128, 32, 8, 36):
        hist = rng.randn(n, d).astype(np.float32)
        want = np.zeros((len(deg), d))
        for r in range(len(deg)):
            sl = slice(indptr[r], indptr[r + 1])
            want[r] = (w[sl, None].astype(np.float64) * hist[cols[sl]].astype(np.float64)).sum(0)
        y0 = torch.zeros((len(deg), d), device="cuda"); y1 = torch.ones((len(deg), d), device="cuda")
        ops.full_history_mean(dev(nodes), dev(indptr), len(deg), dev(indptr), dev(cols), dev(w), dev(hist), y0, y1)
        close(y0, want, "full mean d=%d" % d); close(y1, want + 1.0, "full mean second output")


@pytest.mark.skipif(not native.have_ref(), reason="oracle/_ref not built (reference sources absent)")
def test_full_mean_kernel_against_the_references_own_csr_loop():
    """full_mean_kernel vs `compute_history` (gcn/history.cpp:10-37, the reference's commented-out CSR loop,
    compiled into oracle/_ref by oracle/Makefile): same rows, same permuted adjacency; fp32 sums in a different
    order, so within 1e-4 of the row scale element-wise; float64 torch.sparse.mm as the third opinion."""
    from stochastic_gcn_b200 import graphs, ops
    from stochastic_gcn_b200.sampler import DeviceSampler
    g = graphs.powerlaw_graph(4000, 200_000, seed=6, device="cuda", max_degree=700)
    D, B = 128, 200
    gen = torch.Generator(device="cuda").manual_seed(5)
    hist = torch.randn((g.n, D), generator=gen, device="cuda")
    s = DeviceSampler(g.data, g.indices, g.indptr, L=1, cv=True)
    s.seed(3)
    for _ in range(2):                                   # second batch: rows already permuted once
        ids = torch.randperm(g.n, generator=gen, device="cuda")[:B].to(torch.int32)
        s.start_batch(ids); s.expand(2)
    y = torch.zeros((B, D), device="cuda")
    ops.full_history_mean(s.view("field"), s.view("rowptr_f"), B, s.view("adj_p"), s.view("adj_i"), s.view("adj_w"),
                          hist, y)
    adj_i, adj_w = s.host("adj_i", s.num_edges), s.host("adj_w", s.num_edges)
    adj_p = g.indptr.cpu().numpy()
    ref = native.ref_compute_history(adj_w, adj_i, adj_p, ids.cpu().numpy(), hist.cpu().numpy()).astype(np.float64)
    got = y.cpu().numpy().astype(np.float64)
    rowscale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) <= 1e-4 * np.abs(ref) + 1e-4 * rowscale).all()
    rows = torch.repeat_interleave(torch.arange(B), torch.from_numpy(np.diff(adj_p)[ids.cpu().numpy()]))
    pos = np.concatenate([np.arange(adj_p[i], adj_p[i + 1]) for i in ids.cpu().numpy()])
    a = torch.sparse_coo_tensor(torch.stack([rows, torch.from_numpy(adj_i[pos].astype(np.int64))]),
                                torch.from_numpy(adj_w[pos].astype(np.float64)), size=(B, g.n))
    third = torch.sparse.mm(a, hist.cpu().double()).numpy()
    assert np.abs(got - third).max() <= 1e-5 * np.abs(third).max()
