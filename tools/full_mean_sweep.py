#!/usr/bin/env python
"""A/B sweep of the full-neighbour history-mean variants on the bench workload (one GPU visit).

For every configuration: correctness against the register variant on the first batches (REDs commute
up to fp32 rounding -> allclose 2e-5 of scale), then CUDA-event time of back-to-back launches over
`--n` different batches (same method as bench.py's roofline leg).  Prints one JSON line per config.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from stochastic_gcn_b200 import _lib, graphs, ops  # noqa: E402
from stochastic_gcn_b200.step import HotPathStep  # noqa: E402

VARIANT, WARPS, ROWS, DEPTH = 0, 1, 2, 3


def tune(variant, warps=12, rows=16, depth=2):
    lib = _lib.load()
    _lib.check(lib.sgcn_tune_set(VARIANT, variant))
    if variant == 1:
        _lib.check(lib.sgcn_tune_set(WARPS, warps))
        _lib.check(lib.sgcn_tune_set(ROWS, rows))
        _lib.check(lib.sgcn_tune_set(DEPTH, depth))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=400)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--configs", default="0;1,12,16,2;1,8,32,1;1,16,16,1;1,12,32,1;1,16,8,2;1,8,16,2;1,12,8,4;1,16,12,1")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = graphs.make_shape("reddit", seed=1, device=dev, scale=args.scale)
    B, H = args.batch, args.hidden
    feats = torch.randn((g.n, H), device=dev)
    step = HotPathStep(g, feats, H, B, 2, mode="cv", seed=1)
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    step.run(torch.randperm(g.n, generator=gen, device=dev)[:B].to(torch.int32).contiguous())   # sizes the views
    step.history.normal_()
    v = step._ensure_views(0)
    deg = v["adj_p"][1:] - v["adj_p"][:-1]
    calls, total_bytes = [], 0
    for _ in range(args.n):
        b = torch.randperm(g.n, generator=gen, device=dev)[:B].to(torch.int32).contiguous()
        rp = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        rp[1:] = torch.cumsum(deg[b.long()], 0)
        nnz = int(rp[-1])
        total_bytes += 8 * nnz + 4 * (B + 1) + 4 * H * nnz
        calls.append((b, rp))
    out = torch.zeros((B, H), device=dev)

    def launch(c, y=out):
        ops.full_history_mean(c[0], c[1], B, v["adj_p"], v["adj_i"], v["adj_w"], step.history, y)

    tune(0)
    refs = []
    for c in calls[:4]:
        y = torch.zeros((B, H), device=dev)
        launch(c, y)
        refs.append(y)
    torch.cuda.synchronize()
    for cfg in args.configs.split(";"):
        f = [int(x) for x in cfg.split(",")]
        tune(*f)
        err = 0.0
        try:
            for c, ref in zip(calls[:4], refs):
                y = torch.zeros((B, H), device=dev)
                launch(c, y)
                torch.cuda.synchronize()
                err = max(err, float((y - ref).abs().max() / ref.abs().max()))
            for c in calls[:5]:
                launch(c)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0.record()
                for c in calls:
                    launch(c)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / len(calls))
            print(json.dumps({"cfg": cfg, "us_per_launch": round(best, 3), "rel_err_vs_reg": err,
                              "alg_GBs": round(total_bytes / len(calls) / best / 1e3, 1)}), flush=True)
        except Exception as e:  # noqa: BLE001 -- report and go on with the next configuration
            print(json.dumps({"cfg": cfg, "error": str(e)[:300]}), flush=True)
            break
    tune(0)


if __name__ == "__main__":
    main()
