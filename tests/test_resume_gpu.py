"""GPU: a run interrupted by a checkpoint and resumed in a fresh set of objects repeats the uninterrupted run --
sampler engine AND permuted adjacency rows (DeviceSampler.get_state / set_state), history table, model variables,
Adam slots, the dropout Philox counter, the NumPy RNG behind PyScheduler.shuffle and the epoch cursor
(stochastic_gcn_b200/io.py; the reference's TF Saver keeps variables + history only, gcn/models.py:204-220).
Everything integer (batches, sampled fields, permuted adjacency, dropout masks) repeats bit for bit; the
floating-point state repeats up to the order of the aggregate's vector reductions (1e-5), which no two runs share."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(g, pp, y, seed):
    from stochastic_gcn_b200 import nn
    from stochastic_gcn_b200.sampler import DeviceSampler
    model = nn.PPModel(pp.shape[1], 32, y.shape[1], num_fc_layers=1, normalization="graphsage", cvd=False,
                       layer_norm=True, dropout=0.3, weight_decay=5e-4, seed=seed)
    opt = nn.Adam(model.parameters(), learning_rate=0.01)
    sampler = DeviceSampler(g.data, g.indices, g.indptr, L=1, cv=True)
    sampler.seed(seed)
    hist = torch.zeros((g.n, 32), device="cuda")
    return model, opt, sampler, hist


def train(g, pp, y, model, opt, sampler, hist, data, start, n_steps, B=128):
    from stochastic_gcn_b200 import ops
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, VRAggregator
    losses = []
    for _ in range(n_steps):
        if start + B > data.shape[0]:
            np.random.shuffle(data)                              # PyScheduler.shuffle: the global NumPy RNG
            start = 0
        ids = torch.from_numpy(data[start:start + B]).cuda()
        start += B
        sampler.start_batch(ids)
        sampler.expand(2, materialize_full=False)
        n_in = sampler.sizes().n_in
        field = sampler.view("field")[:n_in]
        adj = DeviceAdj(sampler.view("rowptr_s"), sampler.view("edg_t"), sampler.view("edg_w"), B, n_in,
                        tgt=sampler.view("tgt"))
        full = FullNeighbours.in_place(ids, sampler.view("rowptr_f"), sampler.view("adj_p"), sampler.view("adj_i"),
                                       sampler.view("adj_w"))
        aggr = VRAggregator(adj, full, None, field, None, [hist], None, False, normalization="graphsage")
        loss = model.loss(model.forward(ops.gather_rows(pp, field), aggr), y[ids.long()])
        opt.zero_grad()
        loss.backward()
        opt.step()
        aggr.write_back()
        losses.append(float(loss.detach()))
    return losses, start


def test_resumed_run_repeats_the_uninterrupted_one(tmp_path):
    from stochastic_gcn_b200 import graphs, io
    g = graphs.powerlaw_graph(2000, 40_000, seed=0, device="cuda", max_degree=200)
    gen = torch.Generator(device="cuda").manual_seed(1)
    pp = torch.randn((g.n, 24), generator=gen, device="cuda")
    y = torch.nn.functional.one_hot(torch.randint(0, 5, (g.n,), generator=gen, device="cuda"), 5).float()

    def fresh_data():
        np.random.seed(77)
        d = np.arange(1024, dtype=np.int32)
        np.random.shuffle(d)
        return d

    # ---- uninterrupted: 14 steps (crosses an epoch boundary: 1024 / 128 = 8 steps per epoch) ----
    data = fresh_data()
    model, opt, sampler, hist = make(g, pp, y, 3)
    want, _ = train(g, pp, y, model, opt, sampler, hist, data, 0, 14)
    want_vars = [p.data.detach().cpu().numpy().copy() for p in model.parameters()]
    want_hist = hist.cpu().numpy().copy()
    want_adj = sampler.host("adj_i", sampler.num_edges).copy()
    sampler.close()

    # ---- 6 steps, checkpoint, everything rebuilt from scratch, 8 more steps ----
    data = fresh_data()
    model, opt, sampler, hist = make(g, pp, y, 3)
    got, start = train(g, pp, y, model, opt, sampler, hist, data, 0, 6)
    params = model.parameters()
    path = str(tmp_path / "ck.npz")
    io.save_checkpoint(path, [p.data.detach().cpu().numpy() for p in params], [hist.cpu().numpy()],
                       {"t": opt.t, "m": [p.m.cpu().numpy() for p in params], "v": [p.v.cpu().numpy() for p in params]},
                       sampler.get_state(),
                       {"dropout_seed": model.drop_state.seed, "dropout_offset": model.drop_state.offset,
                        "numpy_rng": np.random.get_state(), "epoch_data": data, "epoch_start": start})
    sampler.close()
    del model, opt, sampler, hist
    np.random.seed(999)                                          # the resumed process starts from other state

    ck = io.load_checkpoint(path)
    model, opt, sampler, hist = make(g, pp, y, 12345)            # different seed: everything comes from the file
    for p, v, m, vv in zip(model.parameters(), ck["variables"], ck["optimizer"]["m"], ck["optimizer"]["v"]):
        p.data.data.copy_(torch.from_numpy(v))
        p.m.copy_(torch.from_numpy(m))
        p.v.copy_(torch.from_numpy(vv))
    opt.t = ck["optimizer"]["t"]
    hist.copy_(torch.from_numpy(ck["history"][0]))
    sampler.set_state(ck["sampler_state"])
    hs = ck["host_state"]
    model.drop_state.seed, model.drop_state.offset = hs["dropout_seed"], hs["dropout_offset"]
    np.random.set_state(hs["numpy_rng"])
    more, _ = train(g, pp, y, model, opt, sampler, hist, hs["epoch_data"].copy(), hs["epoch_start"], 8)

    assert np.array_equal(sampler.host("adj_i", sampler.num_edges), want_adj), "permuted adjacency (bit for bit)"
    np.testing.assert_allclose(got + more, want, rtol=1e-5, err_msg="loss sequence differs after the resume")
    np.testing.assert_allclose(hist.cpu().numpy(), want_hist, rtol=1e-4, atol=1e-5, err_msg="history table")
    for p, w in zip(model.parameters(), want_vars):
        np.testing.assert_allclose(p.data.detach().cpu().numpy(), w, rtol=1e-4, atol=1e-6, err_msg="model variables")
    # and the resume is NOT a no-op: without the sampler state the sequence diverges at once
    model2, opt2, sampler2, hist2 = make(g, pp, y, 12345)
    other, _ = train(g, pp, y, model2, opt2, sampler2, hist2, hs["epoch_data"].copy(), hs["epoch_start"], 2)
    assert abs(other[0] - want[6]) > 1e-4
    sampler.close()
    sampler2.close()
