"""A few launches of the tcgen05 dense layer at the headline shape (for `ncu -k regex:gather_gemm`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochastic_gcn_b200 import ops

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
n_nodes, k, n = 232965, 1204, 1510
feats = torch.randn((n_nodes, k), generator=gen, device=dev)
w = torch.randn((k, 128), generator=gen, device=dev) / np.sqrt(k)
packed = ops.pack_dense_weights(w)
for it in range(6):
    idx = torch.randint(0, n_nodes, (n,), generator=gen, device=dev, dtype=torch.int32)
    out = ops.gathered_dense(feats, idx, packed, k, epilogue="ln_relu")
torch.cuda.synchronize()
want = torch.relu(torch.nn.functional.layer_norm(feats[idx.long()].double() @ w.double(), (128,), eps=1e-9))
print("max abs diff vs float64: %.3e" % float((out - want).abs().max()))
