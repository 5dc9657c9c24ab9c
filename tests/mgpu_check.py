"""Run under torchrun (one rank per GPU): the sharded pass against R oracle samplers sharing one
history table (SURVEY.md 8e parity definition).  Usage:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_check.py peer cv
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import aggregators as agg          # noqa: E402
from oracle import native                      # noqa: E402
from stochastic_gcn_b200 import graphs         # noqa: E402
from stochastic_gcn_b200.sharding import ShardedHotPathStep, row_range   # noqa: E402


def main():
    transport, mode = sys.argv[1], sys.argv[2]
    use_graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
    pipelined = len(sys.argv) > 3 and sys.argv[3] == "pipelined"
    trains = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] in ("trains", "trains-graph") else None
    tables = sys.argv[4] if len(sys.argv) > 4 else "replicated"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    deg = 2 if mode == "cv" else 1
    D, B, steps, seed = 32, 24, (8 if pipelined else 15 if trains else 5), 3
    fullsize = os.environ.get("MGPU_FULLSIZE", "") == "1"      # the BENCHED shapes: BASELINE configs[2] / configs[3]
    gen = torch.Generator(device=dev).manual_seed(0)
    if fullsize:
        import bench
        w = bench.WORKLOADS["reddit_cv" if mode == "cv" else "reddit_cvd"]
        g, feats = bench.build_inputs(w, 1, dev, 1.0)
        D, B, deg, steps = w["hidden"], w["batch"], w["degree"], 8
        assert trains == "trains-graph", "the full-size check runs the benched form: graphs of the trains schedule"
    else:
        g = graphs.powerlaw_graph(1500, 60_000, seed=4, device=dev, max_degree=300)
        feats = torch.randn((g.n, 80), generator=gen, device=dev)
    step = ShardedHotPathStep(g, feats, D, B, deg, mode=mode, seed=seed + rank, rank=rank, world=world,
                              transport=transport, tables=tables)
    hist0 = torch.randn((g.n, D), generator=gen, device=dev)
    if tables == "sharded":
        step.history[:step.hi - step.lo].copy_(hist0[step.lo:step.hi])      # this rank's shard
    else:
        step.history.copy_(hist0)
    torch.cuda.synchronize()
    dist.barrier()

    # every rank's batches (all ranks need them to replay the oracle)
    batches = []
    for r in range(world):
        if tables == "sharded":
            from stochastic_gcn_b200.sharding import shard_rows
            S = shard_rows(g.n, world)
            lo, hi = min(g.n, r * S), min(g.n, (r + 1) * S)
        else:
            lo, hi = row_range(g.n, r, world)
        rng = np.random.RandomState(50 + r)
        batches.append([(rng.permutation(hi - lo)[:B] + lo).astype(np.int32) for _ in range(steps)])

    # oracle: R reference samplers on the global CSR, one shared history, synchronous steps
    gw, gi, gp = g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy()
    samplers = []
    for r in range(world):
        o = native.OracleSampler(gw, gi, gp, cv=True)
        o.seed(seed + r)
        samplers.append(o)
    hist = hist0.cpu().numpy().copy()
    fh = feats.cpu().numpy()
    want_out = []
    for s in range(steps):
        updates, outs = [], []
        for r in range(world):
            o = samplers[r]
            o.start_batch(batches[r][s]); o.expand(deg)
            z = o.snapshot()
            n_in = len(z["field"])
            x0 = fh[z["field"]]
            adj = (np.stack([z["edg_s"], z["edg_t"]], 1).astype(np.int32), z["edg_w"], (B, n_in))
            fadj = (np.stack([z["fedg_s"], z["fedg_t"]], 1).astype(np.int32), z["fedg_w"], (B, len(z["ffield"])))
            if mode == "cv":
                out, new = agg.cv_forward(adj, fadj, z["field"], z["ffield"], hist, x0[:, :D], True)
            else:
                (out, _), new = agg.cvd_forward(adj, fadj, z["field"], z["ffield"], hist, z["scales"], x0[:, :D],
                                                x0[:, D:2 * D], True)
            outs.append(out)
            updates.append((z["field"], new[0]))
        for ids, rows in updates:          # rank order: the highest rank wins a contended row
            hist[ids] = rows
        want_out.append(outs[rank])

    if pipelined:
        # multi-step graphs (2 steps each) with sampler lookahead and the peer exchange inside the graphs;
        # capture_pipelined runs the first two batches as eager warm-up passes
        dev_batches = [torch.from_numpy(b).to(dev) for b in batches[rank]]
        step.capture_pipelined(dev_batches[0], dev_batches[1], steps_per_graph=2)
        step.run_pipelined(dev_batches[2:])
        torch.cuda.synchronize()
        step.check_exchange()
        out = step.out.cpu().numpy()
        err = np.abs(out - want_out[-1]).max() / max(np.abs(want_out[-1]).max(), 1e-30)
        assert err < 1e-4, "rank %d: last pipelined out differs by %g" % (rank, err)
        steps = 0
    if trains:
        # trains schedule (sgcn_step_run_trains) with the peer exchange: eager with every pass's rows read back,
        # or 5 passes per CUDA graph; trains of 4 batches so that several trains are in flight
        step.train = 4
        table = torch.from_numpy(np.stack(batches[rank])).to(dev)
        if trains == "trains":
            rows = torch.empty((steps, B, step.outs[0].shape[1]), dtype=torch.float32).pin_memory()
            step.run_trains(table, out_host=rows, first_train=2)
            torch.cuda.synchronize()
            for s in range(steps):
                err = np.abs(rows[s].numpy() - want_out[s]).max() / max(np.abs(want_out[s]).max(), 1e-30)
                assert err < 1e-4, "rank %d: trains pass %d differs by %g" % (rank, s, err)
        else:
            S = 4 if fullsize else 5
            step.capture_trains(S, table[:S], first_train=2)      # eager warm-up run = passes 0 .. S-1
            step.replay_trains(table[S:])
            torch.cuda.synchronize()
        step.check_exchange()
        out = step.out.cpu().numpy()
        err = np.abs(out - want_out[-1]).max() / max(np.abs(want_out[-1]).max(), 1e-30)
        assert err < 1e-4, "rank %d: last trains out differs by %g" % (rank, err)
        steps = 0
    for s in range(steps):
        ids = torch.from_numpy(batches[rank][s]).to(dev)
        if use_graph and s == 0:
            step.capture(ids)              # eager pass on this batch (incl. exchange), then capture
            out = step.out.cpu().numpy()
        elif use_graph:
            out = step.replay(ids).cpu().numpy()
        else:
            out = step.run(ids).cpu().numpy()
        torch.cuda.synchronize()
        step.check_exchange()
        err = np.abs(out - want_out[s]).max() / max(np.abs(want_out[s]).max(), 1e-30)
        assert err < 1e-4, "rank %d step %d: out differs by %g" % (rank, s, err)
    dist.barrier()
    torch.cuda.synchronize()
    got_hist = step.full_history().cpu().numpy()
    assert np.array_equal(got_hist, hist), "rank %d: history replica differs from the oracle in %d rows" % (
        rank, int((got_hist != hist).any(1).sum()))
    dist.barrier()
    if rank == 0:
        print("mgpu_check ok: transport=%s mode=%s form=%s tables=%s world=%d nodes=%d batch=%d width=%d passes=%d" % (
            transport, mode, sys.argv[3] if len(sys.argv) > 3 else "eager", tables, world, g.n, B, D,
            15 if trains and not fullsize else 8))
    step.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
