"""Run as a script (tests/test_zzz_ahead_schedule_gpu.py spawns it in a process of its own): the
EXPERIMENTAL "gather ahead" schedule (csrc/step.cu:sgcn_step_run_ahead, HotPathStep.run_ahead /
capture_ahead) against the CPU oracle, pass after pass, incl. batches that share nodes.
Prints "ahead_check ok" when every case agrees."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import native                                   # noqa: E402
from tests.test_step_gpu import close, oracle_step          # noqa: E402


def _setup(mode, deg, norm, n_batches):
    from stochastic_gcn_b200 import graphs
    from stochastic_gcn_b200.step import HotPathStep
    g = graphs.powerlaw_graph(3000, 120_000, seed=4, device="cuda", max_degree=600)
    D, B = 32, 48
    gen = torch.Generator(device="cuda").manual_seed(0)
    feats = torch.randn((g.n, 80), generator=gen, device="cuda")
    step = HotPathStep(g, feats, D, B, deg, mode=mode, seed=5, normalization=norm)
    step.history.normal_(generator=gen)
    step.d_out.normal_(generator=gen)
    o = native.OracleSampler(g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy(), cv=mode != "ns")
    o.seed(5)
    perm = torch.randperm(g.n, generator=gen, device="cuda").to(torch.int32)
    batches = [perm[i * B:(i + 1) * B].contiguous() for i in range(n_batches)]
    batches[3] = torch.cat((batches[2][:20], batches[3][20:])).contiguous()     # shares 20 nodes with batch 2
    batches[4] = batches[3].flip(0).contiguous()                                # same nodes as batch 3
    return g, step, o, feats, batches, D


def _check(step, o, mode, deg, norm, feats, batches, hist, D):
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()
    for ids in batches:
        oh, om, dx, s = oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out, graphsage=norm != "gcn")
    close(step.out.cpu().numpy(), oh, "last out")
    if om is not None:
        close(step.out_mu.cpu().numpy(), om, "last out_mu")
    if mode != "ns":
        assert np.array_equal(step.history.cpu().numpy(), hist)
    assert np.array_equal(step.sampler.host("adj_i", step.sampler.num_edges), o.vec("adj_i"))


def run_ahead_matches_oracle(mode, deg, norm):
    g, step, o, feats, batches, D = _setup(mode, deg, norm, 7)
    hist = step.history.cpu().numpy().copy()
    step.run_ahead(torch.stack(batches))
    torch.cuda.synchronize()
    _check(step, o, mode, deg, norm, feats, batches, hist, D)


def captured_ahead_graph_matches_oracle():
    mode, deg, norm, S = "cv", 2, "graphsage", 6
    g, step, o, feats, batches, D = _setup(mode, deg, norm, 3 * S)
    hist = step.history.cpu().numpy().copy()
    table = torch.stack(batches)
    step.capture_ahead(table[:S], steps_per_graph=S)            # eager warm-up run = passes 0 .. S-1
    step.replay_ahead(table[S:])                                # two graph replays = passes S .. 3S-1
    torch.cuda.synchronize()
    _check(step, o, mode, deg, norm, feats, batches, hist, D)


if __name__ == "__main__":
    for case in (("cv", 2, "graphsage"), ("cvd", 1, "graphsage"), ("ns", 1, "graphsage"), ("cv", 2, "gcn")):
        run_ahead_matches_oracle(*case)
        print("run_ahead %s ok" % (case,), flush=True)
    captured_ahead_graph_matches_oracle()
    print("captured graph ok", flush=True)
    print("ahead_check ok", flush=True)
