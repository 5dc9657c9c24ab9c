"""Scratch: host enqueue cost vs device time of the native step driver."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from stochastic_gcn_b200.step import HotPathStep
w = bench.WORKLOADS["reddit_cv"]
dev = torch.device("cuda", 0)
g, feats = bench.build_inputs(w, 1, dev, 1.0)
step = HotPathStep(g, feats, w["hidden"], w["batch"], w["degree"], mode=w["mode"], seed=1)
batches = torch.stack(bench.make_batches(g.n, w["batch"], 420, 1, dev))
step.d_out.normal_()
step.run_native(batches[:10]); torch.cuda.synchronize()
for n in (200, 400):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record(); step.run_native(batches[:n]); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("n=%d: host enqueue %.1f us/step, device %.1f us/step, wall %.1f us/step" % (
        n, (t1 - t0) / n * 1e6, e0.elapsed_time(e1) / n * 1e3, (t2 - t0) / n * 1e6))

# device timeline of a short native run (block-0 stamps of every launch)
import ctypes as C
from stochastic_gcn_b200 import _lib
NAMES = ["sampler", "full_mean", "gather", "sampled_fwd", "spmm_bwd", "history_update", "copy/zero", "exchange"]
LOG = 17 + 2 * 1024
trace = torch.zeros(LOG, dtype=torch.int64, device=dev)
trace[0:16:2] = -1
torch.cuda.synchronize()
_lib.load().sgcn_trace_set(C.c_void_p(trace.data_ptr()))
step.run_native(batches[20:32]); torch.cuda.synchronize()
_lib.load().sgcn_trace_set(None)
t = trace.cpu().tolist()
n = min(t[16], 1024)
ev = sorted((t[18 + 2 * i], t[17 + 2 * i]) for i in range(n))
t0 = ev[0][0]
open_at = {}
for tm, code in ev:
    cls, is_end = code >> 1, code & 1
    if not is_end:
        open_at[cls] = tm
    else:
        print("    %-15s %7.1f -> %7.1f  (%.1f us)" % (NAMES[cls], (open_at.get(cls, tm) - t0) / 1e3, (tm - t0) / 1e3,
                                                       (tm - open_at.get(cls, tm)) / 1e3))
