"""debug aid: the cross-train wait of sgcn_sampler_expand_train (prints every set's meta block)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.graphs_small import random_graph
from stochastic_gcn_b200.sampler import DeviceSampler

n, B, degree = 3000, 64, 2
g = random_graph(n, 12, 78)
s = DeviceSampler(g.data, g.indices, g.indptr, L=1, cv=True)
s.seed(9)
s.reserve_sets(8, B, degree)
a, b = torch.cuda.Stream(), torch.cuda.Stream()
s.mark_consumed(b)
torch.cuda.synchronize()
s.pipeline(True)
rng = np.random.RandomState(3)
perm = rng.permutation(n).astype(np.int32)
t0 = perm[:4 * B].reshape(4, B).copy()
t1 = perm[4 * B:8 * B].reshape(4, B).copy()
mode = sys.argv[1] if len(sys.argv) > 1 else "late"
if mode != "noshare":
    t1[2, :8] = t0[1, :8]
d0, d1 = torch.from_numpy(t0).cuda(), torch.from_numpy(t1).cuda()
torch.cuda.synchronize()
t = time.time()
s.expand_train(d0, first_set=0, stream=a)
if mode == "early":
    for _ in range(4):
        s.mark_consumed(b)
    torch.cuda.synchronize()
s.expand_train(d1, first_set=4, prev=d0, stream=a)
if mode == "probe":
    # does ANY kernel on another stream run while the train's thread block spins?
    x = torch.zeros(4, device="cuda")
    ev = torch.cuda.Event()
    with torch.cuda.stream(b):
        x.fill_(1.0)
        ev.record(b)
    t1_ = time.time()
    while not ev.query() and time.time() - t1_ < 8:
        time.sleep(0.001)
    print("a kernel on stream b finished after %.4f s while the train waits" % (time.time() - t1_))
    for _ in range(4):
        s.mark_consumed(b)
elif mode == "hiprio":
    c = torch.cuda.Stream(priority=-1)
    with torch.cuda.stream(c):
        torch.cuda._sleep(2_000_000)
    for _ in range(4):
        s.mark_consumed(c)
elif mode != "early":
    with torch.cuda.stream(b):
        torch.cuda._sleep(2_000_000)
    for _ in range(4):
        s.mark_consumed(b)
torch.cuda.synchronize()
print(mode, "elapsed %.3f s" % (time.time() - t), "pipe", s.view("pipe").cpu().tolist())
for j in range(8):
    s.set_slot(j)
    print(j, s.view("meta").cpu().tolist())
