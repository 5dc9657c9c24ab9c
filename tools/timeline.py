"""Device timeline of one step (globaltimer stamps written by the kernels themselves, see
sgcn_trace_set): start / end of every kernel class relative to the first kernel of the graph.
Usage: python tools/timeline.py [serial|pipelined] [n_steps]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from stochastic_gcn_b200 import _lib
from stochastic_gcn_b200.step import HotPathStep

NAMES = ["sampler", "full_mean", "gather", "sampled_fwd", "spmm_bwd", "history_update", "copy/zero", "exchange",
         "wb_push", "wb_claim", "wb_copy", "full_last_body", "full_tail_end", "sampled_end"]


def dump_events(trace, S, what):
    t = trace.cpu().tolist()
    n = min(t[16], 1024)
    ev = sorted((t[18 + 2 * i], t[17 + 2 * i]) for i in range(n) if t[18 + 2 * i] > 0)
    t0 = ev[0][0]
    print("%s of %d steps: %d events, span %.1f us (%.1f us / step)" % (
        what, S, n, (ev[-1][0] - t0) / 1e3, (ev[-1][0] - t0) / 1e3 / S))
    open_at = {}
    for tm, code in ev:
        cls, is_end = code >> 1, code & 1
        if not is_end:
            open_at.setdefault(cls, []).append(tm)
        else:
            st = open_at[cls].pop(0) if open_at.get(cls) else tm
            print("    %-15s %7.1f -> %7.1f  (%.1f us)" % (NAMES[cls] if cls < len(NAMES) else "class%d" % cls,
                                                           (st - t0) / 1e3, (tm - t0) / 1e3, (tm - st) / 1e3))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "serial"
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    w = bench.WORKLOADS["reddit_cv"]
    dev = torch.device("cuda", 0)
    g, feats = bench.build_inputs(w, 1, dev, 1.0)
    step = HotPathStep(g, feats, w["hidden"], w["batch"], w["degree"], mode=w["mode"], seed=1)
    batches = bench.make_batches(g.n, w["batch"], 64, 1, dev)
    step.d_out.normal_()
    LOG = 17 + 2 * 1024
    trace = torch.zeros(LOG, dtype=torch.int64, device=dev)
    _lib.load().sgcn_trace_set(C.c_void_p(trace.data_ptr()))       # before capture: baked into the graphs
    init = torch.zeros(LOG, dtype=torch.int64, device=dev)           # min slots = ~0 (as uint64), max slots = 0
    init[0:16:2] = -1
    step.capture(batches[0])
    if mode == "pipelined":
        step.capture_pipelined(batches[0], batches[1], steps_per_graph=n_steps)
    if mode == "trains":                                             # the bench's schedule, one graph of n_steps
        step.fuse_write_back = os.environ.get("FUSE", "1") != "0"
        step.train = int(os.environ.get("TRAIN", "16"))
        ft = int(os.environ.get("FIRST_TRAIN", "4"))
        step.capture_trains(n_steps, torch.stack(batches[:n_steps]), first_train=ft)
        step.replay_trains(torch.stack(batches[12:12 + n_steps]))
        torch.cuda.synchronize()
        step._trains["tab"].copy_(torch.stack(batches[30:30 + n_steps]))
        trace.copy_(init); torch.cuda.synchronize()
        step._trains["graph"].replay(); torch.cuda.synchronize()
        dump_events(trace, n_steps, "one trains graph (train %d, first %d, fused write-back %s)" % (
            step.train, ft, step.fuse_write_back))
        _lib.load().sgcn_trace_set(None)
        return
    for b in batches[2:12]:
        step.replay(b) if mode == "serial" else None
    torch.cuda.synchronize()
    rows = []
    if mode == "serial":
        for b in batches[12:12 + n_steps]:
            trace.copy_(init); torch.cuda.synchronize()
            step.replay(b); torch.cuda.synchronize()
            rows.append(trace.cpu().tolist())
    else:
        pipe = step._pipe
        S = pipe["S"]
        step.run_pipelined(batches[12:12 + S])                       # settle
        tab = pipe["tab"]
        tab[1][S - 1].copy_(batches[30]); pipe["first"].replay()
        for k in range(S):
            tab[0][k].copy_(batches[31 + k])
        trace.copy_(init); torch.cuda.synchronize()
        pipe["open"][0].replay(); torch.cuda.synchronize()
        t = trace.cpu().tolist()
        n = min(t[16], 1024)
        ev = sorted((t[18 + 2 * i], t[17 + 2 * i]) for i in range(n))
        t0 = ev[0][0]
        print("one open chunk of %d steps: %d events, span %.1f us (%.1f us / step)" % (
            S, n, (ev[-1][0] - t0) / 1e3, (ev[-1][0] - t0) / 1e3 / S))
        open_at = {}
        for tm, code in ev:
            cls, is_end = code >> 1, code & 1
            if not is_end:
                open_at[cls] = tm
            else:
                print("    %-15s %7.1f -> %7.1f  (%.1f us)" % (NAMES[cls], (open_at.get(cls, tm) - t0) / 1e3,
                                                               (tm - t0) / 1e3, (tm - open_at.get(cls, tm)) / 1e3))
        _lib.load().sgcn_trace_set(None)
        return
    for k, t in enumerate(rows):
        starts = [t[2 * i] for i in range(8) if t[2 * i + 1] > 0]
        t0 = min(starts)
        end = max(t[2 * i + 1] for i in range(8))
        print("step %d: graph span %.1f us" % (k, (end - t0) / 1e3))
        for i in range(8):
            if t[2 * i + 1] > 0:
                print("    %-15s start %6.1f  end %6.1f  (%.1f us)" % (NAMES[i], (t[2 * i] - t0) / 1e3,
                                                                      (t[2 * i + 1] - t0) / 1e3,
                                                                      (t[2 * i + 1] - t[2 * i]) / 1e3))
    _lib.load().sgcn_trace_set(None)


if __name__ == "__main__":
    main()
