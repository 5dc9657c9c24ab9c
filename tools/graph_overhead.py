"""Scratch: CUDA-graph replay cost model on this box -- linear chains vs fork/join topologies."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochastic_gcn_b200 import ops
dev = torch.device("cuda", 0)
buf = [torch.zeros((64, 128), device=dev) for _ in range(8)]
def k(i): ops.copy_rows_pad(None, 0, buf[i % 8])

def capture(fn):
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=side): fn(side)
    return g
def timeit(g, n=300):
    for _ in range(20): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def linear(n):
    def fn(side):
        for i in range(n): k(i)
    return fn
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def forkjoin(reps, branches, per_branch):
    streams = [s1, s2, s3][:branches - 1]
    def fn(side):
        for r in range(reps):
            ev = torch.cuda.Event(); ev.record(side)
            evs = []
            for st in streams:
                with torch.cuda.stream(st):
                    st.wait_event(ev)
                    for i in range(per_branch): k(i)
                    e = torch.cuda.Event(); e.record(st); evs.append(e)
            for i in range(per_branch): k(i + 4)
            for e in evs: side.wait_event(e)
            k(7)
    return fn
for n in (1, 5, 10, 20, 40, 80):
    print("linear %3d kernels: %.1f us/replay  (%.2f us/kernel)" % (n, timeit(capture(linear(n))), timeit(capture(linear(n))) / n))
for reps, br, pb in ((1, 2, 3), (1, 3, 3), (4, 2, 3), (4, 3, 3), (8, 3, 3), (8, 2, 4)):
    t = timeit(capture(forkjoin(reps, br, pb)))
    nk = reps * (br * pb + 1)
    print("fork/join reps=%d branches=%d per_branch=%d (%d kernels): %.1f us/replay (%.2f us/kernel, %.1f us/rep)" % (reps, br, pb, nk, t, t / nk, t / reps))
