"""Scratch probe: where does the pipelined step's time go?  (rest-only, sampler-only, both)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from stochastic_gcn_b200.step import HotPathStep

w = bench.WORKLOADS["reddit_cv"]
dev = torch.device("cuda", 0)
g, feats = bench.build_inputs(w, 1, dev, 1.0)
step = HotPathStep(g, feats, w["hidden"], w["batch"], w["degree"], mode=w["mode"], seed=1)
batches = bench.make_batches(g.n, w["batch"], 260, 1, dev)
step.d_out.normal_()
dyn = len(sys.argv) > 1 and sys.argv[1] == "dyn"
step.dynamic_full = dyn
step.capture(batches[0])
step.capture_pipelined(batches[0], batches[1], steps_per_graph=int(os.environ.get('SPG', '8')))
pipe = step._pipe

def timeit(fn, n=200):
    for _ in range(5): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def serial(i):
    step.replay(batches[i % 250])
print("serial one-graph  %.1f us" % timeit(serial))
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); step.run_pipelined(batches[:200]); e1.record(); torch.cuda.synchronize()
    print("pipelined         %.1f us" % (e0.elapsed_time(e1) / 200 * 1e3))
t0 = time.perf_counter(); step.run_pipelined(batches[:200]); t1 = time.perf_counter(); torch.cuda.synchronize()
print("host enqueue time per step %.1f us" % ((t1 - t0) / 200 * 1e6))
