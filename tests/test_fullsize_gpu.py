"""GPU: the BENCHED configurations at full size, checked pass by pass against the CPU oracle.

bench.py times BASELINE.json configs[2] (Reddit-shaped CV+PP degree 2), reports configs[3] (CVD+PP degree 1)
and configs[4] (power-law 2M, NS degree 1) beside it, all through HotPathStep.capture_trains / replay_trains:
CUDA graphs of the trains schedule, with host buffers (pinned ids in, every pass's rows out) for the e2e leg.
Here exactly those graphs run at B = 512, D = 128 on the full-size graphs and every pass is compared with

  * the sampler oracle -- the compiled unmodified reference (oracle/_ref) when it travelled with the repo, the
    pinned C restatement otherwise: input fields and the permuted adjacency bit for bit;
  * the float64 aggregate oracle (oracle/aggregators.py, gcn/layers.py:298-319,350-362): rows and dX within
    1e-4 of the result's magnitude, reported as max-norm error AND element-wise
    |err| <= 1e-4 |want| + 1e-4 max|want_row|;
  * the history table after the last write-back: bit for bit.
A summary goes to gpurun_out/parity_fullsize.json (copied to profiles/ by hand).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import native
from tests.test_step_gpu import oracle_step

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUMMARY = {}


def errors(got, want):
    got = np.asarray(got, np.float64)
    err = np.abs(got - want)
    maxnorm = float(err.max() / max(np.abs(want).max(), 1e-30))
    rowscale = np.abs(want).max(axis=1, keepdims=True)
    bound = 1e-4 * np.abs(want) + 1e-4 * rowscale
    worst = float((err / np.maximum(bound, 1e-300)).max())       # <= 1 means the element-wise bound holds
    return maxnorm, worst


def record(name, entry):
    SUMMARY[name] = entry
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_fullsize.json"), "w") as f:
        json.dump(SUMMARY, f, indent=1)


def run_case(name, workload, S, n_graphs, host_io, fuse=False):
    import bench
    from stochastic_gcn_b200.step import HotPathStep
    w = bench.WORKLOADS[workload]
    dev = torch.device("cuda", 0)
    g, feats = bench.build_inputs(w, 1, dev, 1.0)
    B, D, deg, mode = w["batch"], w["hidden"], w["degree"], w["mode"]
    step = HotPathStep(g, feats, D, B, deg, mode=mode, seed=1)
    step.fuse_write_back = fuse
    gen = torch.Generator(device=dev).manual_seed(11)
    step.history.normal_(generator=gen)
    step.d_out.normal_(generator=gen)
    n = S * (1 + n_graphs)
    table = torch.stack(bench.make_batches(g.n, B, n, 1, dev)).contiguous()

    cls = native.RefSampler if native.have_ref() else native.OracleSampler
    o = cls(g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy(), cv=mode != "ns")
    o.seed(1)
    hist = step.history.cpu().numpy().copy()
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()

    step.capture_trains(S, table[:S], host_io=host_io)           # eager warm-up run = passes 0 .. S-1
    got, pending = {}, []

    def drain():
        f0, c0, r0, e0 = pending.pop()
        e0.synchronize()
        for j in range(c0):
            got[S + f0 + j] = r0[j].clone().numpy()

    def on_chunk(first, count, rows, done):
        if pending:
            drain()
        pending.append((first, count, rows, done))

    if host_io:
        step.replay_trains(table[S:].cpu().pin_memory(), on_chunk=on_chunk)
        drain()
    else:
        step.replay_trains(table[S:])
    torch.cuda.synchronize()
    step.check_flags()

    worst_max, worst_elem, checked = 0.0, 0.0, 0
    for i in range(n):
        oh, om, dx, s = oracle_step(o, mode, deg, table[i].cpu().numpy(), fh, hist, D, d_out)
        if i in got:
            m, e = errors(got[i], oh)
            worst_max, worst_elem, checked = max(worst_max, m), max(worst_elem, e), checked + 1
    z = step.sizes()
    assert z["n_in"] == len(s["field"]) and z["nnz_s"] == len(s["edg_s"])
    assert np.array_equal(step.sampler.host("field", z["n_in"]), s["field"]), "last input field"
    m_out, e_out = errors(step.out.cpu().numpy(), oh)
    m_dx, e_dx = errors(step.last_dx.cpu().numpy()[:z["n_in"]], dx)
    entry = {"workload": workload, "passes": n, "passes_with_rows_checked": checked, "graph_passes": S,
             "host_buffers": host_io, "write_back_fused_into_the_mean": bool(fuse and mode != "ns"), "sampler_oracle": cls.__name__, "nodes": g.n, "stored_edges": g.nnz,
             "rows_max_norm_err": max(worst_max, m_out), "rows_elementwise_bound_ratio": max(worst_elem, e_out),
             "dx_max_norm_err": m_dx, "dx_elementwise_bound_ratio": e_dx, "last_sizes": z}
    if om is not None:
        entry["mu_rows_max_norm_err"], entry["mu_rows_elementwise_bound_ratio"] = errors(step.out_mu.cpu().numpy(), om)
    adj_equal = bool(np.array_equal(step.sampler.host("adj_i", step.sampler.num_edges), o.vec("adj_i")))
    hist_equal = mode == "ns" or bool(np.array_equal(step.history.cpu().numpy(), hist))
    entry["permuted_adj_i_bit_equal"], entry["history_bit_equal"] = adj_equal, hist_equal
    record(name, entry)
    assert adj_equal, "permuted adjacency differs"
    assert hist_equal, "history table differs"
    for k in ("rows_max_norm_err", "dx_max_norm_err", "mu_rows_max_norm_err"):
        assert entry.get(k, 0.0) <= 1e-4, (k, entry[k])
    for k in ("rows_elementwise_bound_ratio", "dx_elementwise_bound_ratio", "mu_rows_elementwise_bound_ratio"):
        assert entry.get(k, 0.0) <= 1.0, (k, entry[k])


def test_reddit_cv_d2_host_buffer_graphs():
    """configs[2], the e2e leg of the bench: two alternating host-buffer graphs of 16 passes, 64 passes."""
    run_case("reddit_cv_e2e_graphs", "reddit_cv", 16, 3, True)


def test_reddit_cv_d2_device_graph_of_20():
    """configs[2], the device-resident leg at the driver's own K = 20: one graph of 20 passes, replayed."""
    run_case("reddit_cv_device_graph20", "reddit_cv", 20, 2, False)


def test_reddit_cv_d2_write_back_fused_into_the_mean():
    """same, with the history write-back carried by the tail of the full-neighbour mean's launch (the opt-in form)"""
    run_case("reddit_cv_device_graph20_fused", "reddit_cv", 20, 1, False, fuse=True)


def test_reddit_cvd_d1_host_buffer_graphs():
    """configs[3] on one GPU (the sharded form is checked by tests/mgpu_check.py)."""
    run_case("reddit_cvd_e2e_graphs", "reddit_cvd", 16, 3, True)


def test_powerlaw_2m_ns_d1_host_buffer_graphs():
    """configs[4] on one GPU: 2M nodes / 100M stored entries, NS degree 1."""
    run_case("powerlaw_ns_e2e_graphs", "powerlaw_ns", 16, 1, True)
