"""GPU: the dense layers, loss and optimiser around the aggregate through the C ABI vs
the float64 restatement in oracle/dense.py (1e-4 of scale, north_star's bound), and one whole training
step of the pre-processed model -- sampler, input rows, dense, aggregate, dense, loss, backward, Adam,
history write-back -- against the same model in float64 with dense adjacency matrices."""
import numpy as np
import pytest
import torch

from oracle import aggregators as agg
from oracle import dense as od
from oracle import native
from tests.graphs_small import random_graph

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, want, what, rtol=RTOL):
    got = got.detach().cpu().numpy().astype(np.float64) if isinstance(got, torch.Tensor) else np.asarray(got, np.float64)
    want = want.detach().numpy() if isinstance(want, torch.Tensor) else np.asarray(want, np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max() / scale
    assert err <= rtol, "%s: max error %.3e of scale" % (what, err)


@pytest.mark.parametrize("d", [7, 32, 128, 256, 602, 1024])
@pytest.mark.parametrize("affine", [False, True])
@pytest.mark.parametrize("relu", [False, True])
def test_layer_norm_act(d, affine, relu):
    from stochastic_gcn_b200 import nn
    rng = np.random.RandomState(d)
    x = (rng.randn(300, d) * 2 + 0.5).astype(np.float32)
    sc = (rng.rand(d) + 0.5).astype(np.float32) if affine else None
    off = rng.randn(d).astype(np.float32) if affine else None
    gy = rng.randn(300, d).astype(np.float32)
    xr = od.t64(x).requires_grad_(True)
    scr = od.t64(sc).requires_grad_(True) if affine else None
    offr = od.t64(off).requires_grad_(True) if affine else None
    want = od.layer_norm_act(xr, scr, offr, 1e-9, relu)
    (want * od.t64(gy)).sum().backward()
    xt = dev(x).requires_grad_(True)
    sct = dev(sc).requires_grad_(True) if affine else None
    offt = dev(off).requires_grad_(True) if affine else None
    got = nn.layer_norm_act(xt, sct, offt, 1e-9, relu)
    close(got, want, "ln fwd")
    (got * dev(gy)).sum().backward()
    close(xt.grad, xr.grad, "ln dx")
    if affine:
        close(sct.grad, scr.grad, "ln dscale")
        close(offt.grad, offr.grad, "ln doffset")


def test_dropout_injected_mask_and_philox():
    from stochastic_gcn_b200 import nn
    rng = np.random.RandomState(0)
    x = rng.randn(257, 131).astype(np.float32)
    mask = (rng.rand(257, 131) < 0.7).astype(np.uint8)
    xt = dev(x).requires_grad_(True)
    y = nn.dropout(xt, 0.7, mask=dev(mask))
    want = np.where(mask != 0, x * np.float32(1.0 / np.float32(0.7)), np.float32(0))
    assert np.array_equal(y.detach().cpu().numpy(), want.astype(np.float32))
    y.sum().backward()
    assert np.array_equal(xt.grad.cpu().numpy(), np.where(mask != 0, np.float32(1.0 / np.float32(0.7)), 0).astype(np.float32))
    # generated masks: reproducible from (seed, offset), different across offsets, keep rate ~ keep_prob
    big = torch.ones((2000, 500), device="cuda")
    a = nn.dropout(big, 0.8, state=nn.DropoutState(5))
    b = nn.dropout(big, 0.8, state=nn.DropoutState(5))
    st = nn.DropoutState(5); st.take(4)
    c = nn.dropout(big, 0.8, state=st)
    assert torch.equal(a, b) and not torch.equal(a, c)
    rate = float((a != 0).float().mean())
    assert abs(rate - 0.8) < 5e-3
    assert torch.all((a == 0) | (torch.abs(a - 1.25) < 1e-6))
    assert nn.dropout(big, 1.0) is big


@pytest.mark.parametrize("multitask", [False, True])
@pytest.mark.parametrize("c", [3, 41, 7, 121])
def test_cross_entropy(multitask, c):
    from stochastic_gcn_b200 import nn
    rng = np.random.RandomState(c)
    n = 333
    z = (rng.randn(n, c) * 3).astype(np.float32)
    if multitask:
        t = (rng.rand(n, c) < 0.3).astype(np.float32)
    else:
        t = np.eye(c, dtype=np.float32)[rng.randint(0, c, n)]
        t[:5] = 0                                  # unlabeled rows (all-zero one-hot) contribute 0
    zr = od.t64(z).requires_grad_(True)
    want = od.cross_entropy(zr, od.t64(t), multitask)
    (want * 1.7).backward()
    zt = dev(z).requires_grad_(True)
    got = nn.cross_entropy(zt, dev(t), multitask)
    close(got, want, "loss")
    (got * 1.7).backward()
    close(zt.grad, zr.grad, "dlogits")


def test_adam_matches_tf_formula():
    from stochastic_gcn_b200 import nn
    rng = np.random.RandomState(1)
    p0 = rng.randn(1204, 128).astype(np.float32)
    par = nn.Parameter(dev(p0), weight_decay=5e-4)
    opt = nn.Adam([par], learning_rate=0.01, beta1=0.9, beta2=0.999)
    p, m, v = p0.astype(np.float64), np.zeros_like(p0, np.float64), np.zeros_like(p0, np.float64)
    for t in range(1, 6):
        g = rng.randn(*p0.shape).astype(np.float32)
        par.data.grad = dev(g)
        opt.step()
        od.adam_step(p, g.astype(np.float64), m, v, t, weight_decay=5e-4)
        close(par.data, p, "adam step %d" % t, rtol=2e-6)


def sample(g, ids, degree, seed=3):
    s = native.OracleSampler(g.data, g.indices, g.indptr, cv=True)
    s.seed(seed)
    s.start_batch(ids)
    s.expand(degree)
    z = s.snapshot()
    n_out, n_in = len(ids), len(z["field"])
    adj = (np.stack([z["edg_s"], z["edg_t"]], 1).astype(np.int32), z["edg_w"], (n_out, n_in))
    fadj = (np.stack([z["fedg_s"], z["fedg_t"]], 1).astype(np.int32), z["fedg_w"], (n_out, len(z["ffield"])))
    return z, adj, fadj


@pytest.mark.parametrize("mode,nfc,graphsage,ln", [("cv", 1, True, True), ("cv", 2, True, True), ("cvd", 1, True, True),
                                                   ("cvd", 2, False, True), ("cv", 1, False, False), ("ns", 2, True, True)])
def test_pp_model_training_step(mode, nfc, graphsage, ln):
    """one full training step (dropout with injected masks) vs the float64 model"""
    from stochastic_gcn_b200 import nn
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, PlainAggregator, VRAggregator
    n, f, hid, ncls, keep = 400, 50, 32, 5, 0.8
    g = random_graph(n, 15, 21)
    rng = np.random.RandomState(7)
    feats = rng.randn(n, f).astype(np.float32)
    labels = np.eye(ncls, dtype=np.float32)[rng.randint(0, ncls, n)]
    ids = rng.choice(n, size=60, replace=False).astype(np.int32)
    z, adj, fadj = sample(g, ids, 2 if mode == "cv" else 1)
    field = z["field"]
    n_in = len(field)
    hist = (rng.randn(n, hid) * 0.5).astype(np.float32)
    cvd, norm = mode == "cvd", "graphsage" if graphsage else "gcn"
    model = nn.PPModel(f, hid, ncls, num_fc_layers=nfc, normalization=norm, cvd=cvd, layer_norm=ln,
                       dropout=1 - keep, weight_decay=5e-4, seed=3)
    weights = [p.data.detach().cpu().numpy().copy() for p in model.parameters()]
    # dropout sites in call order: (cvd: one per AugmentedDropoutDense on h) / (one per Dropout layer)
    shapes = []
    d_agg = hid * (2 if graphsage else 1)
    for l in range(nfc):
        shapes.append((n_in, f if l == 0 else hid))
    for l2 in range(nfc):
        shapes.append((len(ids), d_agg if l2 == 0 else hid))
    masks = [(rng.rand(*s) < keep).astype(np.uint8) for s in shapes]
    sites = [l for l in model.pre + model.post if isinstance(l, (nn.Dropout, nn.AugmentedDropoutDense))]
    assert len(sites) == len(masks)
    for l, m in zip(sites, masks):
        l.mask = dev(m)

    ref = od.PPReference(weights, nfc, graphsage, cvd, ln, 5e-4)
    logits_ref = ref.forward(feats[field], adj, fadj if mode != "ns" else None, field, z["ffield"],
                             hist if mode != "ns" else None, z["scales"], keep, masks)
    loss_ref = ref.loss(logits_ref, labels[ids])
    loss_ref.backward()

    hist_t = dev(hist)
    dadj = DeviceAdj.from_coo(adj)
    if mode == "ns":
        aggr = PlainAggregator(dadj, normalization=norm)
    else:
        aggr = VRAggregator(dadj, FullNeighbours.from_coo(fadj, z["ffield"]), None, dev(field), None, [hist_t],
                            dev(z["scales"]), cvd, normalization=norm)
    x = dev(feats[field])
    logits = model.forward(x, aggr)
    close(logits, logits_ref, "logits")
    loss = model.loss(logits, dev(labels[ids]))
    close(float(loss) + model.l2_term(), float(loss_ref), "loss")
    opt = nn.Adam(model.parameters(), learning_rate=0.01)
    opt.zero_grad()
    loss.backward()
    for i, (p, wr) in enumerate(zip(model.parameters(), ref.w)):
        gpu_grad = p.data.grad.cpu().numpy().astype(np.float64) + p.weight_decay * p.data.detach().cpu().numpy()
        close(gpu_grad, wr.grad, "grad of parameter %d" % i)
    opt.step()                                                    # gcn/models.py:186-194: Adam, then the write-back
    for i, (p, wr) in enumerate(zip(model.parameters(), ref.w)):
        w64 = wr.detach().numpy().copy()
        od.adam_step(w64, wr.grad.numpy(), np.zeros_like(w64), np.zeros_like(w64), 1, weight_decay=0.0)
        # the first Adam step is lr * g / (|g| + eps): entries whose gradient is ~eps amplify fp32 rounding
        close(p.data, w64, "parameter %d after Adam" % i, rtol=1e-4)
    if mode != "ns":
        aggr.write_back()
        want_hist = agg.history_update(hist.astype(np.float64).copy(), field, ref.new_history)
        close(hist_t, want_hist, "history after write-back", rtol=1e-5)


def test_pp_model_learns_a_separable_task():
    """loss of the assembled model falls on a synthetic task where the label is a function of the
    neighbourhood mean (dropout 0.2, CV aggregator, device sampler in the loop)"""
    from stochastic_gcn_b200 import graphs, nn, ops
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, VRAggregator
    from stochastic_gcn_b200.sampler import DeviceSampler
    torch.manual_seed(0)
    g = graphs.powerlaw_graph(3000, 60_000, seed=0, device="cuda", max_degree=300)
    n, f, hid, ncls, B = g.n, 16, 32, 4, 256
    x = torch.randn((n, f), device="cuda")
    pp = ops.preprocess_features(g.indptr, g.indices, g.data, x, "graphsage")           # [X | A X]
    proto = torch.randn((2 * f, ncls), device="cuda")
    y = torch.nn.functional.one_hot((pp @ proto).argmax(1), ncls).float()
    model = nn.PPModel(2 * f, hid, ncls, num_fc_layers=1, normalization="graphsage", cvd=False, layer_norm=True,
                       dropout=0.2, weight_decay=5e-4, seed=1)
    opt = nn.Adam(model.parameters(), learning_rate=0.01)
    sampler = DeviceSampler(g.data, g.indices, g.indptr, L=1, cv=True)
    sampler.seed(1)
    hist = torch.zeros((n, hid), device="cuda")
    losses = []
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    for it in range(60):
        ids = torch.randperm(n, generator=gen, device="cuda")[:B].to(torch.int32)
        sampler.start_batch(ids)
        sampler.expand(2, materialize_full=False)
        n_in = sampler.sizes().n_in
        field = sampler.view("field")[:n_in]
        adj = DeviceAdj(sampler.view("rowptr_s"), sampler.view("edg_t"), sampler.view("edg_w"), B, n_in,
                        tgt=sampler.view("tgt"))
        full = FullNeighbours.in_place(ids, sampler.view("rowptr_f"), sampler.view("adj_p"), sampler.view("adj_i"),
                                       sampler.view("adj_w"))
        aggr = VRAggregator(adj, full, None, field, None, [hist], None, False, normalization="graphsage")
        logits = model.forward(ops.gather_rows(pp, field), aggr)
        loss = model.loss(logits, y[ids.long()])
        opt.zero_grad()
        loss.backward()
        opt.step()
        aggr.write_back()
        losses.append(float(loss))
    assert np.mean(losses[-10:]) < 0.6 * np.mean(losses[:5]), losses
