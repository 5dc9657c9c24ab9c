"""GPU: the whole pass (HotPathStep) -- eager and CUDA-graph replay -- against the oracle, for the
three estimator modes, over several consecutive batches (history and sampler state carry over)."""
import numpy as np
import pytest
import torch

from oracle import aggregators as agg
from oracle import native

pytestmark = pytest.mark.gpu


def oracle_step(o, mode, deg, ids, feats, hist, D, d_out, graphsage=True):
    o.start_batch(ids)
    o.expand(deg)
    s = o.snapshot()
    B, n_in = len(ids), len(s["field"])
    x0 = feats[s["field"]]
    adj = (np.stack([s["edg_s"], s["edg_t"]], 1).astype(np.int32), s["edg_w"], (B, n_in))
    if mode == "ns":
        out = agg.plain_forward(adj, x0[:, :D], graphsage)
        dx = agg.plain_backward(adj, d_out, n_in, graphsage)
        return out, None, dx, s
    fadj = (np.stack([s["fedg_s"], s["fedg_t"]], 1).astype(np.int32), s["fedg_w"], (B, len(s["ffield"])))
    if mode == "cv":
        out, new = agg.cv_forward(adj, fadj, s["field"], s["ffield"], hist, x0[:, :D], graphsage)
        dx = agg.plain_backward(adj, d_out, n_in, graphsage)
        agg.history_update(hist, s["field"], new[0])
        return out, None, dx, s
    (oh, om), new = agg.cvd_forward(adj, fadj, s["field"], s["ffield"], hist, s["scales"], x0[:, :D],
                                    x0[:, D:2 * D], graphsage)
    dx = agg.cvd_backward_h(adj, s["scales"], d_out, n_in, graphsage)
    agg.history_update(hist, s["field"], new[0])
    return oh, om, dx, s


def close(got, want, what):
    err = np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30)
    assert err <= 1e-4, "%s: %.3e" % (what, err)


@pytest.mark.parametrize("mode,deg", [("ns", 1), ("cv", 2), ("cvd", 1)])
@pytest.mark.parametrize("use_graph", [False, True])
def test_pass_matches_oracle_over_batches(mode, deg, use_graph):
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    g = graphs.powerlaw_graph(3000, 120_000, seed=4, device="cuda", max_degree=600)
    D, B = 32, 48
    gen = torch.Generator(device="cuda").manual_seed(0)
    feats = torch.randn((g.n, 80), generator=gen, device="cuda")
    step = HotPathStep(g, feats, D, B, deg, mode=mode, seed=5)
    step.history.normal_(generator=gen)
    step.d_out.normal_(generator=gen)
    hist = step.history.cpu().numpy().copy()
    o = native.OracleSampler(g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy(), cv=mode != "ns")
    o.seed(5)
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()
    perm = torch.randperm(g.n, generator=gen, device="cuda").to(torch.int32)
    batches = [perm[i * B:(i + 1) * B].contiguous() for i in range(5)]
    for i, ids in enumerate(batches):
        if use_graph and i == 0:
            step.capture(ids)            # eager warm-up pass on this batch, then capture (= second pass)
            # the capture pass does not execute; replay once so that history/sampler state advance
            oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out)
            continue
        out = (step.replay(ids) if use_graph else step.run(ids)).cpu().numpy()
        oh, om, dx, s = oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out)
        z = step.sizes()
        assert z["n_in"] == len(s["field"]) and z["nnz_s"] == len(s["edg_s"])
        assert np.array_equal(step.sampler.host("field", z["n_in"]), s["field"])
        close(out, oh, "batch %d out" % i)
        if om is not None:
            close(step.out_mu.cpu().numpy(), om, "batch %d out_mu" % i)
        close(step.dx.cpu().numpy()[:z["n_in"]], dx, "batch %d dx" % i)
        if mode != "ns":
            assert np.array_equal(step.history.cpu().numpy(), hist), "history after batch %d" % i


def test_host_api_round_trip():
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    g = graphs.powerlaw_graph(2000, 60_000, seed=1, device="cuda", max_degree=300)
    feats = torch.randn((g.n, 64), device="cuda")
    step = HotPathStep(g, feats, 32, 32, 2, mode="cv", seed=2)
    ids = torch.arange(32, dtype=torch.int32).pin_memory()
    out = step.step_host(ids)
    assert out.is_pinned() and out.shape == (32, 64)
    assert torch.equal(out, step.out.cpu())
    # graph with the H2D / D2H copies as memcpy nodes: same answer as an eager pass on the next batch
    step2 = HotPathStep(g, feats, 32, 32, 2, mode="cv", seed=2)
    step2.run(ids.cuda())
    ids_b = torch.arange(100, 132, dtype=torch.int32).pin_memory()
    want = step2.run(ids_b.cuda()).cpu()
    step.capture_host()
    got = step.step_host(ids_b)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    b = step.algorithmic_bytes()
    assert b["total"] == sum(v for k, v in b.items() if k != "total") and b["aggregate_full"] > 0


@pytest.mark.parametrize("mode,deg", [("cv", 2), ("cvd", 1), ("ns", 1)])
def test_pipelined_driver_matches_oracle_incl_overlapping_batches(mode, deg):
    """One-batch sampler lookahead (two graphs per buffer set, two streams) must give exactly the
    sequential results -- also when consecutive batches share nodes (the device guard serialises
    the in-place row permutation against the previous batch's full-neighbour reads)."""
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    g = graphs.powerlaw_graph(3000, 120_000, seed=4, device="cuda", max_degree=600)
    D, B = 32, 48
    gen = torch.Generator(device="cuda").manual_seed(0)
    feats = torch.randn((g.n, 80), generator=gen, device="cuda")
    step = HotPathStep(g, feats, D, B, deg, mode=mode, seed=5)
    step.history.normal_(generator=gen)
    step.d_out.normal_(generator=gen)
    hist = step.history.cpu().numpy().copy()
    o = native.OracleSampler(g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy(), cv=mode != "ns")
    o.seed(5)
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()
    perm = torch.randperm(g.n, generator=gen, device="cuda").to(torch.int32)
    batches = [perm[i * B:(i + 1) * B].contiguous() for i in range(9)]
    batches[4] = torch.cat((batches[3][:20], batches[4][20:])).contiguous()     # shares 20 nodes with batch 3
    batches[5] = batches[4].flip(0).contiguous()                                # same nodes as batch 4
    # warm-up passes of capture_pipelined execute batches 0 and 1 eagerly
    step.capture_pipelined(batches[0], batches[1], host_io=True, steps_per_graph=4)
    for ids in batches[:2]:
        oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out)
    got = []

    def grab(first, count, st, done):
        done.synchronize()
        if count > 1:      # a chunk graph: per-step rows were copied to pinned memory by the graph
            c = (first // st._pipe["S"]) & 1
            for k in range(count):
                got.append(st._pipe["pin_out"][c][k].numpy().copy())
        else:              # eager tail step
            got.append(st.out.cpu().numpy().copy())
    host_batches = [b.cpu() for b in batches[2:]]            # 7 steps = one chunk of 4 + a tail of 3
    step.run_pipelined(host_batches, on_chunk=grab)
    torch.cuda.synchronize()
    for i, ids in enumerate(batches[2:]):
        oh, om, dx, s = oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out)
        close(got[i], oh, "pipelined batch %d out" % i)
    z = step.sizes()
    assert z["n_in"] == len(s["field"]) and z["nnz_s"] == len(s["edg_s"])
    close(step.dx.cpu().numpy()[:z["n_in"]], dx, "pipelined last dx")
    if mode != "ns":
        assert np.array_equal(step.history.cpu().numpy(), hist)
    assert np.array_equal(step.sampler.host("adj_i", step.sampler.num_edges), o.vec("adj_i"))
    # closed chunks (n a multiple of S), a second run and a plain eager pass keep working afterwards
    seq2 = host_batches[:4] + host_batches[3:7]              # 8 steps = an open + a closed chunk
    for ids in seq2:
        oracle_step(o, mode, deg, ids.numpy(), fh, hist, D, d_out)
    step.run_pipelined(seq2)
    torch.cuda.synchronize()
    if mode != "ns":
        assert np.array_equal(step.history.cpu().numpy(), hist)
    assert np.array_equal(step.sampler.host("adj_i", step.sampler.num_edges), o.vec("adj_i"))
    step.run(batches[6])
    torch.cuda.synchronize()
    assert step.sizes()["n_out"] == B


@pytest.mark.parametrize("mode,deg,norm", [("cv", 2, "graphsage"), ("cvd", 1, "graphsage"), ("ns", 1, "graphsage"),
                                           ("cvd", 1, "gcn"), ("cv", 2, "gcn")])
@pytest.mark.parametrize("host_io", [False, True])
def test_native_driver_matches_oracle(mode, deg, norm, host_io):
    """csrc/step.cu: n passes from one C call (three streams, sampler lookahead) == n sequential
    oracle passes, incl. batches that share nodes with their predecessor; 'gcn' normalisation is the
    PubMed / Cora form (no self-concat, BASELINE configs[0-1])."""
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    gs = norm != "gcn"
    g = graphs.powerlaw_graph(3000, 120_000, seed=4, device="cuda", max_degree=600)
    D, B = 32, 48
    gen = torch.Generator(device="cuda").manual_seed(0)
    feats = torch.randn((g.n, 80), generator=gen, device="cuda")
    step = HotPathStep(g, feats, D, B, deg, mode=mode, seed=5, normalization=norm)
    step.history.normal_(generator=gen)
    step.d_out.normal_(generator=gen)
    hist = step.history.cpu().numpy().copy()
    o = native.OracleSampler(g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy(), cv=mode != "ns")
    o.seed(5)
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()
    perm = torch.randperm(g.n, generator=gen, device="cuda").to(torch.int32)
    batches = [perm[i * B:(i + 1) * B].contiguous() for i in range(7)]
    batches[3] = torch.cat((batches[2][:20], batches[3][20:])).contiguous()     # shares 20 nodes with batch 2
    batches[4] = batches[3].flip(0).contiguous()                                # same nodes as batch 3
    table = torch.stack(batches)
    width = step.outs[0].shape[1]
    outs = torch.empty((len(batches), B, width), dtype=torch.float32).pin_memory() if host_io else None
    step.run_native(table.cpu() if host_io else table, out_host=outs)
    torch.cuda.synchronize()
    for i, ids in enumerate(batches):
        oh, om, dx, s = oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out, graphsage=gs)
        if host_io:
            close(outs[i].numpy(), oh, "native batch %d out" % i)
    z = step.sizes()
    assert z["n_in"] == len(s["field"]) and z["nnz_s"] == len(s["edg_s"])
    close(step.out.cpu().numpy(), oh, "native last out")
    if om is not None:
        close(step.out_mu.cpu().numpy(), om, "native last out_mu")
    close(step.dx.cpu().numpy()[:z["n_in"]], dx, "native last dx")
    if mode != "ns":
        assert np.array_equal(step.history.cpu().numpy(), hist)
    assert np.array_equal(step.sampler.host("adj_i", step.sampler.num_edges), o.vec("adj_i"))
    # a second native run, then a plain eager pass, continue the same sequence
    step.run_native(batches[:2])
    step.run(batches[5])
    torch.cuda.synchronize()
    for ids in batches[:2] + [batches[5]]:
        oh, om, dx, s = oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out, graphsage=gs)
    close(step.out.cpu().numpy(), oh, "eager pass after native runs")
    if mode != "ns":
        assert np.array_equal(step.history.cpu().numpy(), hist)
