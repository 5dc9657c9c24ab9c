"""Device timeline of the sharded step (trains schedule, peer exchange) on EVERY rank (run under torchrun, one
rank per GPU): globaltimer stamps of block 0 of every kernel of one graph of S passes.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/timeline_mgpu.py [S] [workload]"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from stochastic_gcn_b200 import _lib
from stochastic_gcn_b200.sharding import ShardedHotPathStep

NAMES = ["sampler", "full_mean", "gather", "sampled_fwd", "spmm_bwd", "history_update", "copy/zero", "exchange",
         "wb_push", "wb_claim", "wb_copy"]


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    w = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "reddit_cv"]
    g, feats = bench.build_inputs(w, 1, dev, float(os.environ.get("SCALE", "1.0")))
    step = ShardedHotPathStep(g, feats, w["hidden"], w["batch"], w["degree"], mode=w["mode"], seed=1 + rank,
                              rank=rank, world=world, transport="peer")
    batches = bench.make_batches(g.n, w["batch"], 4 * S, 1 + rank, dev, step.lo, step.hi)
    step.d_out.normal_()
    step.history.normal_()
    LOG = 17 + 2 * 1024
    trace = torch.zeros(LOG, dtype=torch.int64, device=dev)
    _lib.load().sgcn_trace_set(C.c_void_p(trace.data_ptr()))       # before capture: baked into the graphs
    step.capture_trains(S, torch.stack(batches[:S]), first_train=4)
    step.replay_trains(torch.stack(batches[S:3 * S]))
    torch.cuda.synchronize(); dist.barrier()
    trace.zero_()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step.replay_trains(torch.stack(batches[3 * S:4 * S]))
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    step.check_exchange()
    for r in range(world):                      # every rank prints its own timeline, in rank order
        if r == rank:
            t = trace.cpu().tolist()
            n = min(t[16], 1024)
            ev = sorted((t[18 + 2 * i], t[17 + 2 * i]) for i in range(n))
            t0 = ev[0][0]
            print("rank %d of %d: %d passes in %.1f us (%.1f us / pass by CUDA events), %d events" % (
                rank, world, S, e0.elapsed_time(e1) * 1e3, e0.elapsed_time(e1) * 1e3 / S, n), flush=True)
            open_at = {}
            for tm, code in ev:
                cls, is_end = code >> 1, code & 1
                if not is_end:
                    open_at.setdefault(cls, []).append(tm)
                else:
                    st = open_at[cls].pop(0) if open_at.get(cls) else tm
                    print("    %-15s %7.1f -> %7.1f  (%.1f us)" % (NAMES[cls], (st - t0) / 1e3, (tm - t0) / 1e3,
                                                                   (tm - st) / 1e3), flush=True)
        dist.barrier()
    _lib.load().sgcn_trace_set(None)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
