"""Run as a script (tests/test_zzz_ahead_schedule_gpu.py spawns it in a process of its own): the
HOST-BUFFER forms of the gather-ahead schedule -- pinned ids in, every pass's aggregated rows out to
pinned memory -- eager (HotPathStep.run_ahead) and as two alternating CUDA graphs
(capture_ahead(host_io=True) / replay_ahead), against the CPU oracle pass after pass.
Prints "ahead_host_check ok" when every case agrees."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.ahead_check import _setup                        # noqa: E402
from tests.test_step_gpu import close, oracle_step          # noqa: E402


def eager_host_io(mode, deg, norm):
    g, step, o, feats, batches, D = _setup(mode, deg, norm, 7)
    hist = step.history.cpu().numpy().copy()
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()
    width = step.outs[0].shape[1]
    outs = torch.empty((len(batches), len(batches[0]), width), dtype=torch.float32).pin_memory()
    step.run_ahead(torch.stack(batches).cpu().pin_memory(), out_host=outs)
    torch.cuda.synchronize()
    for i, ids in enumerate(batches):
        oh, om, dx, s = oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out, graphsage=norm != "gcn")
        close(outs[i].numpy(), oh, "pass %d rows on the host" % i)
    if mode != "ns":
        assert np.array_equal(step.history.cpu().numpy(), hist)


def graphs_host_io():
    mode, deg, norm, S = "cv", 2, "graphsage", 6
    g, step, o, feats, batches, D = _setup(mode, deg, norm, 5 * S)
    hist = step.history.cpu().numpy().copy()
    fh, d_out = feats.cpu().numpy(), step.d_out.cpu().numpy()
    table = torch.stack(batches)
    step.capture_ahead(table[:S], steps_per_graph=S, host_io=True)      # eager warm-up run = passes 0 .. S-1
    got, pending = {}, []

    def on_chunk(first, count, rows, done):
        if pending:                                                      # consume one chunk behind the launches
            f0, c0, r0, e0 = pending.pop()
            e0.synchronize()
            for j in range(c0):
                got[f0 + j] = r0[j].clone().numpy()
        pending.append((first, count, rows, done))

    step.replay_ahead(table[S:].cpu().pin_memory(), on_chunk=on_chunk)   # four graph replays = passes S .. 5S-1
    f0, c0, r0, e0 = pending.pop()
    e0.synchronize()
    for j in range(c0):
        got[f0 + j] = r0[j].clone().numpy()
    torch.cuda.synchronize()
    for i, ids in enumerate(batches):
        oh, om, dx, s = oracle_step(o, mode, deg, ids.cpu().numpy(), fh, hist, D, d_out, graphsage=True)
        if i >= S:
            close(got[i - S], oh, "pass %d rows from the graph's staging set" % i)
    assert np.array_equal(step.history.cpu().numpy(), hist)


if __name__ == "__main__":
    for case in (("cv", 2, "graphsage"), ("cvd", 1, "graphsage"), ("ns", 1, "gcn")):
        eager_host_io(*case)
        print("eager host io %s ok" % (case,), flush=True)
    graphs_host_io()
    print("graphs host io ok", flush=True)
    print("ahead_host_check ok", flush=True)
