#!/usr/bin/env python
"""Time the pre-processing product A_hat @ X on the bench graph (Reddit shape: 233k nodes, 104M stored
edges, 602 features) for several column tiles; checks one tile against torch.sparse (cuSPARSE)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochastic_gcn_b200 import graphs, ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    g = graphs.make_shape("reddit", seed=1, device=dev, scale=float(os.environ.get("SCALE", "1.0")))
    f = 602
    x = torch.randn((g.n, f), device=dev)
    out = torch.empty((g.n, 2 * f), device=dev)
    out[:, :f].copy_(x)
    alg = g.nnz * (8 + 4 * f) + 8 * g.n * f
    ref = None
    try:
        a = torch.sparse_csr_tensor(g.indptr.long(), g.indices.long(), g.data, size=(g.n, g.n))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ref = a @ x
        e0.record(); ref = a @ x; e1.record(); torch.cuda.synchronize()
        print(json.dumps({"impl": "torch.sparse (cuSPARSE) csr @ dense", "ms": e0.elapsed_time(e1)}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "torch.sparse", "error": str(e)[:200]}), flush=True)
    for tile in [int(t) for t in os.environ.get("TILES", "64,128,192,256").split(",")]:
        ops.csr_spmm(g.indptr, g.indices, g.data, x, out=out[:, f:], tile_cols=tile)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.csr_spmm(g.indptr, g.indices, g.data, x, out=out[:, f:], tile_cols=tile)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        err = None
        if ref is not None:
            err = float((out[:, f:] - ref).abs().max() / ref.abs().max())
        print(json.dumps({"impl": "sgcn_csr_spmm", "tile_cols": tile, "ms": ms, "nodes": g.n, "nnz": g.nnz,
                          "algorithmic_GB": alg / 1e9, "algorithmic_GBs": alg / ms / 1e6,
                          "edges_per_s": g.nnz / ms * 1e3, "max_rel_diff_vs_cusparse": err}), flush=True)


if __name__ == "__main__":
    main()
