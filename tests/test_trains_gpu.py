"""GPU: the trains schedule (csrc/step.cu:sgcn_step_run_trains, HotPathStep.run_trains / capture_trains /
replay_trains) -- trains of batches sampled by one launch, gather one pass ahead, full-neighbour means back to
back with the history write-back carried by their tails -- against the CPU oracle pass after pass: eager and as
CUDA graphs, device and host buffers, CV / CVD / NS, both normalisations, batches that share nodes within a
train and across trains."""
import numpy as np
import pytest
import torch

from oracle import native
from tests.test_step_gpu import close, oracle_step

pytestmark = pytest.mark.gpu


def setup(mode, deg, norm, n_batches, train, fuse=False, share=True, seed=4):
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    g = graphs.powerlaw_graph(3000, 120_000, seed=seed, device="cuda", max_degree=600)
    D, B = 32, 48
    gen = torch.Generator(device="cuda").manual_seed(0)
    feats = torch.randn((g.n, 80), generator=gen, device="cuda")
    step = HotPathStep(g, feats, D, B, deg, mode=mode, normalization=norm, seed=5)
    step.train, step.fuse_write_back = train, fuse
    step.history.normal_(generator=gen)
    step.d_out.normal_(generator=gen)
    o = native.OracleSampler(g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy(), cv=mode != "ns")
    o.seed(5)
    batches = []
    for i in range(n_batches):
        ids = torch.randperm(g.n, generator=gen, device="cuda")[:B].to(torch.int32)
        if share and i > 0 and i % 3 != 1:          # shares nodes with the previous batch / the one before
            ids[:12] = batches[i - 1][20:32]
            if i > 1:
                ids[12:16] = batches[i - 2][40:44]
            ids = torch.unique(ids)
            extra = torch.randperm(g.n, generator=gen, device="cuda").to(torch.int32)
            extra = extra[~torch.isin(extra, ids)][:B - ids.numel()]
            ids = torch.cat((ids, extra))[torch.randperm(B, generator=gen, device="cuda")]
        batches.append(ids.contiguous())
    return g, step, o, feats, torch.stack(batches).contiguous(), D


def check_run(step, o, mode, deg, norm, feats, table, D, got_rows, first=0):
    """oracle replay of the passes table[first:]; got_rows[i] = rows the library produced for pass first + i"""
    hist = check_run.hist
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()
    last = None
    for i in range(first, table.shape[0]):
        oh, om, dx, s = oracle_step(o, mode, deg, table[i].cpu().numpy(), fh, hist, D, d_out, graphsage=norm != "gcn")
        if got_rows is not None and got_rows[i - first] is not None:
            close(got_rows[i - first], oh, "pass %d rows" % i)
        last = (oh, om, dx, s)
    return last


@pytest.mark.parametrize("mode,deg,norm", [("cv", 2, "graphsage"), ("cvd", 1, "graphsage"), ("ns", 1, "gcn"),
                                           ("cv", 1, "gcn"), ("ns", 2, "graphsage")])
@pytest.mark.parametrize("fuse", [True, False])
def test_eager_trains_match_oracle(mode, deg, norm, fuse):
    if mode == "ns" and not fuse:
        pytest.skip("plain sampling keeps no history: the flag has no effect")
    n = 23
    g, step, o, feats, table, D = setup(mode, deg, norm, n, train=4, fuse=fuse)
    check_run.hist = step.history.cpu().numpy().copy()
    width = step.outs[0].shape[1]
    rows = torch.empty((n, step.B, width), dtype=torch.float32).pin_memory()
    step.run_trains(table, out_host=rows, first_train=2)
    torch.cuda.synchronize()
    step.check_flags()
    oh, om, dx, s = check_run(step, o, mode, deg, norm, feats, table, D, [r.numpy() for r in rows])
    z = step.sizes()
    assert z["n_in"] == len(s["field"]) and z["nnz_s"] == len(s["edg_s"])
    close(step.out.cpu().numpy(), oh, "last rows on the device")
    if om is not None:
        close(step.out_mu.cpu().numpy(), om, "last mu rows")
    close(step.last_dx.cpu().numpy()[:z["n_in"]], dx, "last dx")
    assert np.array_equal(step.last_x0.cpu().numpy()[:z["n_in"]], feats.cpu().numpy()[s["field"]])
    if mode != "ns":
        assert np.array_equal(step.history.cpu().numpy(), check_run.hist), "history after the run"
    assert np.array_equal(step.sampler.host("adj_i", step.sampler.num_edges), o.vec("adj_i")), "permuted adjacency"


def test_pinned_host_ids_and_a_second_run_continue_the_state():
    mode, deg, norm = "cv", 2, "graphsage"
    g, step, o, feats, table, D = setup(mode, deg, norm, 18, train=4)
    check_run.hist = step.history.cpu().numpy().copy()
    width = step.outs[0].shape[1]
    rows = torch.empty((18, step.B, width), dtype=torch.float32).pin_memory()
    host_tab = table.cpu().pin_memory()
    step.run_trains(host_tab[:11], out_host=rows[:11])
    step.run_trains(host_tab[11:], out_host=rows[11:], first_train=1)
    torch.cuda.synchronize()
    check_run(step, o, mode, deg, norm, feats, table, D, [r.numpy() for r in rows])
    assert np.array_equal(step.history.cpu().numpy(), check_run.hist)


@pytest.mark.parametrize("mode,deg,fuse", [("cv", 2, False), ("cvd", 1, False), ("ns", 1, False), ("cv", 2, True),
                                           ("cvd", 1, True)])
def test_captured_trains_device_tables(mode, deg, fuse):
    S = 10
    g, step, o, feats, table, D = setup(mode, deg, "graphsage", 4 * S, train=4, fuse=fuse)
    check_run.hist = step.history.cpu().numpy().copy()
    step.capture_trains(S, table[:S], first_train=2)           # eager warm-up run = passes 0 .. S-1
    step.replay_trains(table[S:])                              # three replays
    torch.cuda.synchronize()
    step.check_flags()
    oh, om, dx, s = check_run(step, o, mode, deg, "graphsage", feats, table, D, None)
    z = step.sizes()
    assert z["n_in"] == len(s["field"])
    close(step.out.cpu().numpy(), oh, "last rows")
    close(step.last_dx.cpu().numpy()[:z["n_in"]], dx, "last dx")
    if mode != "ns":
        assert np.array_equal(step.history.cpu().numpy(), check_run.hist)


def test_captured_trains_host_buffers():
    mode, deg, S = "cv", 2, 6
    g, step, o, feats, table, D = setup(mode, deg, "graphsage", 5 * S, train=4)
    check_run.hist = step.history.cpu().numpy().copy()
    step.capture_trains(S, table[:S], host_io=True, first_train=2)
    got, pending = {}, []

    def drain():
        f0, c0, r0, e0 = pending.pop()
        e0.synchronize()
        for j in range(c0):
            got[f0 + j] = r0[j].clone().numpy()

    def on_chunk(first, count, rows, done):
        if pending:                                            # consume one replay behind the launches
            drain()
        pending.append((first, count, rows, done))

    step.replay_trains(table[S:].cpu().pin_memory(), on_chunk=on_chunk)
    drain()
    torch.cuda.synchronize()
    rows = [None] * S + [got[i] for i in range(4 * S)]
    check_run(step, o, mode, deg, "graphsage", feats, table, D, rows)
    assert np.array_equal(step.history.cpu().numpy(), check_run.hist)


@pytest.mark.parametrize("D,ld_rows", [(128, 256), (32, 80), (30, 30), (300, 304)])
def test_fused_write_back_equals_mean_then_write_back(D, ld_rows):
    """sgcn_full_history_mean_wb == sgcn_full_history_mean followed by sgcn_history_update: the aggregated rows
    come from the table BEFORE the write-back (every read precedes every store), the table afterwards holds the
    new rows bit for bit, the consumer counter moved by one; three passes back to back on one stream (the
    counters are handed from launch to launch), the sampled aggregate's signal given by a real launch."""
    from stochastic_gcn_b200 import _lib, graphs, ops
    from stochastic_gcn_b200.sampler import DeviceSampler
    lib = _lib.load()
    g = graphs.powerlaw_graph(5000, 300_000, seed=2, device="cuda", max_degree=900)
    B = 256
    gen = torch.Generator(device="cuda").manual_seed(1)
    hist = torch.randn((g.n, D), generator=gen, device="cuda")
    s = DeviceSampler(g.data, g.indices, g.indptr, L=1, cv=True)
    counters = torch.ones(8, dtype=torch.int32, device="cuda")          # garbage: the reset must clear it
    consumed = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.sgcn_wb_counters_reset(_lib.ptr(counters), _lib.stream_ptr()))
    ref = hist.clone()
    for it in range(3):
        ids = torch.randperm(g.n, generator=gen, device="cuda")[:B].to(torch.int32)
        s.start_batch(ids); s.expand(2)
        z = s.sizes()
        field, rowptr_f = s.view("field").clone(), s.view("rowptr_f").clone()
        n_in = torch.tensor([z.n_in], dtype=torch.int32, device="cuda")
        rows = torch.randn((B * 3, ld_rows), generator=gen, device="cuda")
        x = torch.randn((B * 3, D), generator=gen, device="cuda")
        # the pass's sampled aggregate (reads hist[tgt]) signals the tail
        y_s = torch.zeros((B, D), device="cuda")
        _lib.check(lib.sgcn_sampled_done_attach(_lib.ptr(counters)))
        ops.cv_sampled_fwd(s.view("rowptr_s"), s.view("edg_t"), s.view("edg_w"), s.view("tgt"), B, x, hist, y_s)
        y = torch.zeros((B, D), device="cuda")
        _lib.check(lib.sgcn_full_history_mean_wb(
            _lib.ptr(field), _lib.ptr(rowptr_f), B, None, _lib.ptr(s.view("adj_p")), _lib.ptr(s.view("adj_i")),
            _lib.ptr(s.view("adj_w")), _lib.ptr(hist), hist.stride(0), D, _lib.ptr(y), D, None, 0,
            _lib.ptr(field), _lib.ptr(n_in), B * 3, _lib.ptr(rows), ld_rows, _lib.ptr(counters), _lib.ptr(consumed),
            _lib.stream_ptr()))
        want = torch.zeros((B, D), device="cuda")
        ops.full_history_mean(field, rowptr_f, B, s.view("adj_p"), s.view("adj_i"), s.view("adj_w"), ref, want)
        ref[field[:z.n_in].long()] = rows[:z.n_in, :D]
        torch.cuda.synchronize()
        err = (y - want).abs().max() / want.abs().max()
        assert float(err) < 1e-6, (it, float(err))
        assert torch.equal(hist, ref), "table after the fused write-back, pass %d" % it
        assert int(consumed.item()) == it + 1
    c = counters.cpu().tolist()
    assert c[0] == 0 and c[1] == 0 and c[2] == 3 and c[3] == 3 and c[4] == 0, c
