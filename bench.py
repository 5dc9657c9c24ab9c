#!/usr/bin/env python
"""bench.py -- throughput of the variance-reduced GCN hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path, host cores

A step = one pass of the hot path over one batch (sampler -> feature gather -> aggregate forward
-> aggregate backward -> history write-back) on the synthetic Reddit-shaped graph of BASELINE.json
configs[2] (233k nodes, ~115M stored edges, 602-d features -> 1204-d PP input, hidden 128, CV+PP,
degree 2, batch 512).  Metric: aggregated edges / s = (nnz(adj) + nnz(fadj)) per step / step time
(SURVEY.md 8d); the reference's own `amt_data` counter (sampled edges only) is reported beside it.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per stream (see stochastic_gcn_b200/__init__.py)
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: graph shape, mode, degree, batch, hidden, PP feature width            BASELINE.json
    "reddit_cv": dict(shape="reddit", mode="cv", degree=2, batch=512, hidden=128, feat=1204),       # configs[2]
    "reddit_cvd": dict(shape="reddit", mode="cvd", degree=1, batch=512, hidden=128, feat=1204),     # configs[3]
    "powerlaw_ns": dict(shape="powerlaw2m", mode="ns", degree=1, batch=512, hidden=128, feat=512),  # configs[4]
    "pubmed_cvd": dict(shape="pubmed", mode="cvd", degree=1, batch=60, hidden=32, feat=500),        # configs[1]
    # large-batch stress points of SURVEY 8d (the headline batch of 512 is latency-bound by construction)
    "reddit_cv_b4096": dict(shape="reddit", mode="cv", degree=2, batch=4096, hidden=128, feat=1204),
    "reddit_cv_b32768": dict(shape="reddit", mode="cv", degree=2, batch=32768, hidden=128, feat=1204),
}
METRIC = "aggregated edges/s (sampled + full-neighbour edges through SpMM fwd+bwd per step)"


def load_traffic(kernel_name):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/rNN_*_ncu.json: dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured
    launches); None when no capture of that kernel is committed."""
    import glob
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu.json"))):
        try:
            with open(p) as f:
                d = json.load(f)
        except (OSError, ValueError):
            continue
        if kernel_name.split("<")[0].split(" ")[0] in d.get("kernel", "") and d.get("traffic_bytes_per_launch"):
            best = {"bytes": float(d["traffic_bytes_per_launch"]), "source": os.path.relpath(p, ROOT)}
    return best


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(w, seed, device, scale):
    from stochastic_gcn_b200 import graphs
    g = graphs.make_shape(w["shape"], seed=seed, device=device, scale=scale)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    feats = torch.randn((g.n, w["feat"]), generator=gen, device=device, dtype=torch.float32)
    return g, feats


def make_batches(n_nodes, batch, n_steps, seed, device, lo=0, hi=None):
    """Distinct ids per batch, drawn without replacement from [lo, hi) (an epoch-style shuffle)."""
    hi = n_nodes if hi is None else hi
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 2)
    out, pool, pos = [], None, 0
    for _ in range(n_steps):
        if pool is None or pos + batch > pool.numel():
            pool = (torch.randperm(hi - lo, generator=gen, device=device) + lo).to(torch.int32)
            pos = 0
        out.append(pool[pos:pos + batch].contiguous())
        pos += batch
    return out


def edge_counts(g, batches, degree, cv):
    deg = g.degrees()
    s = f = 0
    for b in batches:
        d = deg[b.long()]
        s += int(torch.clamp(d, max=degree).sum())
        f += int(d.sum()) if cv else 0
    return s, f


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation of the path (compiled reference sampler + row
# slicer where oracle/_ref exists, oracle port otherwise; the TensorFlow aggregate is restated by
# the OpenMP C port because TensorFlow is not installed).  Only bench.py, tests/ and smoke() may
# touch oracle/.
# ------------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, w, g_host, feats_host, seed):
        import ctypes as C
        from oracle import native
        self.C, self.native, self.w = C, native, w
        self.lib = native.oracle_lib()
        self.threads = os.cpu_count() or 1
        self.kind_native = "reference" if native.have_ref() else "port"
        cls = native.RefSampler if native.have_ref() else native.OracleSampler
        self.adj_w, self.adj_i, self.adj_p = g_host
        self.s = cls(self.adj_w, self.adj_i, self.adj_p, cv=w["mode"] != "ns")
        self.s.seed(seed)
        self.slice = native.ref_dense_slice if native.have_ref() else native.oracle_dense_slice
        self.feats = feats_host
        self.hist = np.zeros((len(self.adj_p) - 1, w["hidden"]), dtype=np.float32)
        self.adj_p32 = np.ascontiguousarray(self.adj_p, dtype=np.int32)      # row pointers never change

    def step(self, ids):
        nat, lib, w = self.native, self.lib, self.w
        H = w["hidden"]
        ip, fp = nat._ip, nat._fp
        self.s.start_batch(ids)
        self.s.expand(w["degree"])
        field, edg_s, edg_t, edg_w = (self.s.vec(k) for k in ("field", "edg_s", "edg_t", "edg_w"))
        n_out, n_in, nnz_s = len(ids), len(field), len(edg_s)
        x0 = self.slice(self.feats, field)                                   # history.dense_slice
        x = np.ascontiguousarray(x0[:, :H])
        rowptr = np.searchsorted(edg_s, np.arange(n_out + 1)).astype(np.int32)
        z = np.empty((n_out, H), dtype=np.float32)
        if w["mode"] == "ns":
            lib.orc_spmm_csr_omp(n_out, ip(rowptr), ip(edg_t), fp(edg_w), fp(x), H, fp(z), self.threads)
            nnz_f = 0
        else:
            # fused CV forward over the sampler's own rows (same arithmetic as fadj @ H[ffield])
            adj_p = self.adj_p32
            ids32 = np.ascontiguousarray(ids, dtype=np.int32)
            ci, cw = self.s._get_i, self.s._get_f
            pi, pw = nat._i32p(), nat._f32p()
            ci(self.s._h, nat.INT_VECS["adj_i"], self.C.byref(pi))           # zero-copy views of the
            cw(self.s._h, nat.FLOAT_VECS["adj_w"], self.C.byref(pw))         # scheduler's permuted CSR
            lib.orc_cv_forward_omp(n_out, ip(rowptr), ip(edg_t), fp(edg_w), fp(x), ip(field), fp(self.hist), H,
                                   ip(ids32), ip(adj_p), pi, pw, fp(z), self.threads)
            nnz_f = int((adj_p[ids32 + 1] - adj_p[ids32]).sum())
        dy = np.ones((n_out, H), dtype=np.float32)
        dx = np.zeros((n_in, H), dtype=np.float32)
        lib.orc_spmm_coo_t(nnz_s, ip(edg_s), ip(edg_t), fp(edg_w), fp(dy), H, fp(dx), n_in)
        if w["mode"] != "ns":
            lib.orc_scatter_rows(n_in, H, ip(field), fp(x), fp(self.hist))   # tf.scatter_update
        return nnz_s, nnz_f


def cpu_leg(w, g, feats, seed, batches_host, budget_s, warmup):
    g_host = (g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy())
    cpu = CpuPath(w, g_host, feats.cpu().numpy(), seed)
    for b in batches_host[:warmup]:
        cpu.step(b)
    edges, steps, t0 = 0, 0, time.perf_counter()
    for b in batches_host[warmup:]:
        s, f = cpu.step(b)
        edges += s + f
        steps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": edges / dt, "unit": "edges/s", "cores": cpu.threads, "kind": "port",
            "ms_per_step": 1e3 * dt / max(steps, 1), "steps": steps,
            "sample": "%d steps of the same workload (batch %d); sampler + dense_slice = %s, 1 thread as shipped "
                      "(gcn/history.cpp:77 omp pragma commented out); TF aggregate restated in C "
                      "(oracle/sgcn_oracle.c), OpenMP over %d threads" % (steps, w["batch"],
                      "compiled unmodified reference (oracle/_ref)" if cpu.kind_native == "reference"
                      else "oracle port", cpu.threads)}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    g, feats = build_inputs(w, args.seed, dev, args.scale)
    n = args.steps + args.warmup
    batches = [b.cpu().numpy() for b in make_batches(g.n, w["batch"], n, args.seed, dev)]
    g_host = (g.data.cpu().numpy(), g.indices.cpu().numpy(), g.indptr.cpu().numpy())
    cpu = CpuPath(w, g_host, feats.cpu().numpy(), args.seed)
    del g, feats
    for b in batches[:args.warmup]:
        cpu.step(b)
    t0 = time.perf_counter()
    s_tot = f_tot = 0
    for b in batches[args.warmup:]:
        s, f = cpu.step(b)
        s_tot += s; f_tot += f
    dt = time.perf_counter() - t0
    val = (s_tot + f_tot) / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "edges/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, w),
            "sampled_edges_per_s": s_tot / dt,
            "cpu_baseline": {"value": val, "unit": "edges/s", "cores": cpu.threads,
                             "kind": "reference" if cpu.kind_native == "reference" else "port",
                             "sample": "every step of this run; sampler + dense_slice are the compiled unmodified "
                                       "reference (1 thread, as shipped), the TensorFlow aggregate is the C port "
                                       "with OpenMP on %d threads" % cpu.threads},
            "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def pick_graph_passes(k, cap):
    """Passes per captured CUDA graph for K timed steps: ONE graph of K passes when K <= cap (the driver's
    20-step run is one graph launch), otherwise the largest divisor of K that is <= cap and >= cap / 2, or
    cap with the K mod cap left-over passes issued as plain stream launches by the same native driver."""
    k, cap = int(k), max(2, int(cap))
    if k <= cap:
        return k
    for s in range(cap, cap // 2 - 1, -1):
        if k % s == 0:
            return s
    return cap


def workload_config(args, w):
    """`config` of the JSON line: identical in both arms (the driver compares them)."""
    return {"workload": "%s: synthetic %s-shaped graph, %s+PP degree %d, batch %d/GPU, hidden %d, PP input %d-d"
                        % (args.workload, w["shape"], w["mode"].upper(), w["degree"], w["batch"], w["hidden"], w["feat"]),
            "scale": args.scale, "seed": args.seed,
            "l2": "inputs larger than L2 (adjacency + features + history > 2 GB, a fresh random batch every step); "
                  "no explicit flush"}


class Rig:
    """One workload set up on this rank: graph, features, the step object, its batches."""

    def __init__(self, args, name, w, world, rank, dev, n_batches, inputs=None):
        from stochastic_gcn_b200.step import HotPathStep
        self.w, self.name, self.world, self.rank, self.dev = w, name, world, rank, dev
        # inputs: (graph, features) of an earlier rig on the same graph shape / feature width (same seed: same data)
        self.g, self.feats = inputs if inputs is not None else build_inputs(w, args.seed, dev, args.scale)
        if world > 1:
            from stochastic_gcn_b200.sharding import ShardedHotPathStep
            self.step = ShardedHotPathStep(self.g, self.feats, w["hidden"], w["batch"], w["degree"], mode=w["mode"],
                                           seed=args.seed + rank, rank=rank, world=world, transport=args.transport,
                                           tables=args.tables)
            lo, hi = self.step.lo, self.step.hi
            if args.tables == "sharded":
                self.feats = None            # the step keeps this rank's rows only; drop the full matrix
                torch.cuda.empty_cache()
        else:
            self.step = HotPathStep(self.g, self.feats, w["hidden"], w["batch"], w["degree"], mode=w["mode"],
                                    seed=args.seed)
            lo, hi = 0, self.g.n
        self.step.train = args.train
        self.step.fuse_write_back = args.fuse_write_back
        gen = torch.Generator(device=dev).manual_seed(7)
        self.step.d_out.normal_(generator=gen)
        self.step.history.normal_(generator=gen)      # a warm history table (zero rows would skip reductions)
        self.batches = make_batches(self.g.n, w["batch"], n_batches, args.seed + rank, dev, lo, hi)

    def close(self):
        if hasattr(self.step, "close"):
            self.step.close()


def time_trains(rig, args, timed, warm, barrier, host_io):
    """K = len(timed) passes of the trains schedule as CUDA graph replays (+ a plain-launch remainder);
    returns (ms, passes per graph, launches per pass).  host_io: pinned ids in, every pass's rows out."""
    from stochastic_gcn_b200 import _lib
    step, dev = rig.step, rig.dev
    K = len(timed)
    S = pick_graph_passes(K, args.graph_passes)
    full = (K // S) * S
    before = _lib.launch_count()
    step.capture_trains(S, torch.stack(warm[:S]), host_io=host_io, first_train=args.first_train)
    launches = (_lib.launch_count() - before) / ((3.0 if host_io else 2.0) * S)   # eager warm-up + capture(s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    table = torch.stack(timed).contiguous()
    if not host_io:
        step.replay_trains(torch.stack(warm[:S]))              # first launch of the graph (upload) untimed
        barrier()
        e0.record()
        if args.eager_trains:
            step.run_trains(table, first_train=args.first_train)
        else:
            step.replay_trains(table[:full])
            if full < K:
                step.run_trains(table[full:].contiguous(), first_train=args.first_train)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), S, launches
    pinned = table.cpu().pin_memory()
    width = step.outs[0].shape[1]
    tail_rows = torch.empty((K - full, rig.w["batch"], width), dtype=torch.float32).pin_memory() if full < K else None
    pending = []

    def fetch(first, count, rows, done):
        if pending:
            pending.pop().synchronize()              # the caller reads every replay's rows, one replay behind
        pending.append(done)
    step.replay_trains(torch.stack(warm[:S]).cpu().pin_memory().repeat(2, 1), on_chunk=fetch)   # both graphs once
    pending.pop().synchronize()
    barrier()
    e0.record()
    step.replay_trains(pinned[:full], on_chunk=fetch)
    if full < K:
        step.run_trains(pinned[full:], out_host=tail_rows, first_train=args.first_train)
    pending.pop().synchronize()
    torch.cuda.current_stream(dev).synchronize()     # the caller has every pass's rows
    e1.record()
    barrier()
    return e0.elapsed_time(e1), S, launches


def time_first_dense_layer(rig, n_rows, reps=50):
    """act(MyLayerNorm(features[field] @ W)) for one pass's input field -- the fused tcgen05 kernel
    (gather as the A-operand load, TF32 x 3) beside gather kernel + cuBLAS fp32 GEMM + layer-norm kernel."""
    from stochastic_gcn_b200 import nn, ops
    dev, feats = rig.dev, rig.feats
    k = feats.shape[1]
    gen = torch.Generator(device=dev).manual_seed(5)
    w = torch.randn((k, 128), generator=gen, device=dev) / np.sqrt(k)
    idx = torch.randint(0, feats.shape[0], (n_rows,), generator=gen, device=dev, dtype=torch.int32)
    packed = ops.pack_dense_weights(w)
    out = torch.empty((n_rows, 128), device=dev)
    x0 = torch.empty((n_rows, k), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def fused():
        ops.gathered_dense(feats, idx, packed, k, epilogue="ln_relu", out=out)

    def unfused():
        ops.gather_rows(feats, idx, out=x0)
        nn.layer_norm_act(torch.mm(x0, w), None, None, 1e-9, True)
    res = {}
    for name, fn in (("fused_tcgen05_us", fused), ("gather_cublas_fp32_ln_us", unfused)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        res[name] = 1e3 * e0.elapsed_time(e1) / reps
    fused()
    want = nn.layer_norm_act(torch.mm(feats[idx.long()].double(), w.double()).float(), None, None, 1e-9, True)
    res["max_abs_diff_vs_float64_product"] = float((out - want).abs().max())
    res["rows"], res["K"], res["N"] = n_rows, k, 128
    res["flop_per_call"] = 2.0 * n_rows * k * 128
    res["what"] = ("act(LN(features[idx] @ W)): one kernel (csrc/gemm.cu, tcgen05.mma kind::tf32 x 3 products, TMEM "
                   "accumulator, gather = A-operand load) vs sgcn_gather_rows + torch.mm (cuBLAS fp32) + sgcn_ln_act_fwd; "
                   "back-to-back launches, CUDA events")
    return res


def time_train_step(rig, w, steps=20, warm=3):
    """The WHOLE training step around the hot path (gcn/vrgcn.py:71-84 run_one_step: sampler -> input rows -> dense
    -> aggregate -> dense -> loss -> backward -> Adam -> history write-back) as the Python host loop of
    stochastic_gcn_b200.nn / layers drives it: eager launches, one host read of the field size per step."""
    from stochastic_gcn_b200 import nn, ops
    from stochastic_gcn_b200.layers import DeviceAdj, FullNeighbours, VRAggregator
    from stochastic_gcn_b200.sampler import DeviceSampler
    dev, g, feats = rig.dev, rig.g, rig.feats
    B, hid, ncls, deg = w["batch"], w["hidden"], 41, w["degree"]
    model = nn.PPModel(feats.shape[1], hid, ncls, num_fc_layers=1, normalization="graphsage", cvd=False,
                       layer_norm=True, dropout=0.2, weight_decay=5e-4, seed=1, device=dev)
    opt = nn.Adam(model.parameters(), learning_rate=0.01)
    sampler = DeviceSampler(g.data, g.indices, g.indptr, L=1, cv=True)
    sampler.seed(1)
    hist = torch.zeros((g.n, hid), device=dev)
    gen = torch.Generator(device=dev).manual_seed(9)
    labels = torch.nn.functional.one_hot(torch.randint(0, ncls, (g.n,), generator=gen, device=dev), ncls).float()
    batches = make_batches(g.n, B, steps + warm, 77, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges = 0
    for it, ids in enumerate(batches):
        if it == warm:
            torch.cuda.synchronize(dev)
            e0.record()
        sampler.start_batch(ids)
        sampler.expand(deg, materialize_full=False)
        z = sampler.sizes()
        field = sampler.view("field")[:z.n_in]
        adj = DeviceAdj(sampler.view("rowptr_s"), sampler.view("edg_t"), sampler.view("edg_w"), B, z.n_in,
                        tgt=sampler.view("tgt"))
        full = FullNeighbours.in_place(ids, sampler.view("rowptr_f"), sampler.view("adj_p"), sampler.view("adj_i"),
                                       sampler.view("adj_w"))
        aggr = VRAggregator(adj, full, None, field, None, [hist], None, False, normalization="graphsage")
        logits = model.forward(ops.gather_rows(feats, field), aggr)
        loss = model.loss(logits, labels[ids.long()])
        opt.zero_grad()
        loss.backward()
        opt.step()
        aggr.write_back()
        if it >= warm:
            edges += z.nnz_s + z.nnz_f
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    sampler.close()
    return {"steps": steps, "ms_per_step": ms / steps, "value": edges / (ms * 1e-3), "unit": "edges/s",
            "final_loss": float(loss),
            "what": "whole training step on the headline workload: device sampler -> gather -> dropout + dense "
                    "(%d -> %d, cuBLAS) + layer norm + relu -> CV aggregate -> dropout + dense (%d -> %d) -> softmax "
                    "cross entropy -> backward -> Adam -> history write-back; eager Python host loop "
                    "(stochastic_gcn_b200.nn.PPModel), one host read of the field size per step"
                    % (feats.shape[1], hid, 2 * hid, ncls)}


def run_ours(args, w):
    import torch.distributed as dist
    from stochastic_gcn_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_times(vals):
        """(max over ranks, sum over ranks) of a list of floats"""
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world == 1:
            return t.tolist(), t.tolist()
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        return tmax.tolist(), tsum.tolist()

    K, W = args.steps, args.warmup
    S_max = min(max(K, 2), args.graph_passes)
    rig = Rig(args, args.workload, w, world, rank, dev, W + K + S_max)
    step, g, batches = rig.step, rig.g, rig.batches
    timed, warm = batches[W:W + K], batches[W + K:]
    s_edges, f_edges = edge_counts(g, timed, w["degree"], w["mode"] != "ns")

    sharded_tables = world > 1 and args.tables == "sharded"     # only the trains schedule runs on sharded tables
    # warm-up: the first pass is eager (sizes buffers, counts launches), then one-pass graph replays
    launches_per_step = 0
    if not sharded_tables:
        step.capture(batches[0])
        launches_per_step = step.launches_per_step
        for b in batches[1:W]:
            step.replay(b)
    torch.cuda.synchronize(dev)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    # ---- (1) one graph per step, steps strictly back to back (no cross-step overlap): reference point ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    if not sharded_tables:
        for b in timed:
            step.replay(b)
    ev1.record()
    barrier()
    ms_serial = ev0.elapsed_time(ev1)
    sizes_last = step.sizes() if not sharded_tables else None

    # ---- (2) timed region of record: exactly K passes ----
    nccl_multi = world > 1 and args.transport == "nccl"      # collective runs eagerly between one-pass graphs
    driver = "one-graph-per-step" if (args.no_pipeline or nccl_multi) else args.driver
    if driver == "trains" and w["batch"] > 4096:
        driver = "native"        # beyond the one-launch train sampler's bound: per-batch general sampler path
    ms, S = ms_serial, 1
    if driver == "trains":
        ms, S, launches_per_step = time_trains(rig, args, timed, warm, barrier, host_io=False)
        sizes_last = step.sizes()
    elif driver == "native":
        before = _lib.launch_count()
        step.run_native(torch.stack(batches[:W]))
        launches_per_step = (_lib.launch_count() - before) / W
        timed_table = torch.stack(timed)
        barrier()
        ev0.record()
        step.run_native(timed_table)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        sizes_last = step.sizes()
    elif driver == "graph":
        S = pick_graph_passes(K, 16) // 2 * 2 or 2
        step.capture_pipelined(batches[0], batches[1], steps_per_graph=S)
        launches_per_step -= 2                       # fused dX-init/zero and forward+backward, mark in write-back
        step.run_pipelined(batches[2:W + 2])
        timed_table = torch.stack(timed)             # [K, B] ids resident in HBM: one copy per chunk
        barrier()
        ev0.record()
        step.run_pipelined(timed_table)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        sizes_last = step.sizes()

    # ---- e2e: the same K passes through the host-buffer API (pinned ids in, aggregated rows out) ----
    pinned = [b.cpu().pin_memory() for b in timed]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if driver == "trains":
        ms_e2e, _, _ = time_trains(rig, args, timed, warm, barrier, host_io=True)
    elif driver == "native":
        ids_host = torch.stack(pinned).pin_memory()
        out_host = torch.empty((K, w["batch"], step.out.shape[1]), dtype=torch.float32).pin_memory()
        step.run_native(ids_host[:4], out_host=out_host[:4])
        barrier()
        e0.record()
        step.run_native(ids_host, out_host=out_host)
        torch.cuda.current_stream(dev).synchronize()          # the caller reads the rows
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
    elif driver == "graph":
        step.capture_pipelined(batches[0], batches[1], host_io=True, steps_per_graph=S)
        pending = []

        def fetch(first, count, st, done):
            if pending:
                pending.pop().synchronize()
            pending.append(done)
        pinned_table = torch.stack(pinned).pin_memory()
        step.run_pipelined(pinned_table[:S], on_chunk=fetch)
        pending.pop().synchronize()
        barrier()
        e0.record()
        step.run_pipelined(pinned_table, on_chunk=fetch)
        pending.pop().synchronize()
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
    else:
        if world == 1:
            step.capture_host()
        step.step_host(pinned[0])
        barrier()
        e0.record()
        for p in pinned:
            step.step_host(p)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- dominant kernel, timed alone on its launch stream over the same K batches ----
    kern = step.time_dominant_kernel(timed)

    (ms, ms_e2e, ms_serial, _, _), (_, _, _, s_edges, f_edges) = reduce_times(
        [ms, ms_e2e, ms_serial, float(s_edges), float(f_edges)])
    alg = step.algorithmic_bytes(sizes_last)
    detail = {"nodes": g.n, "stored_edges": g.nnz, "last_step_sizes": sizes_last,
              "parallelism": ("single GPU" if world == 1 else
                              "row-range shards x%d, history + feature tables SHARDED (remote rows read over NVLink), "
                              "write-back rows exchanged by peer stores" % world if sharded_tables else
                              "row-range shards x%d, history + feature replicas, write-back rows exchanged by %s"
                              % (world, args.transport))}
    extra = {}
    if world == 1 and not args.no_also and w["feat"] % 4 == 0 and w["hidden"] == 128 and w["mode"] == "cv":
        for key, fn in (("first_dense_layer", lambda: time_first_dense_layer(rig, sizes_last["n_in"])),
                        ("train_step", lambda: time_train_step(rig, w))):
            try:
                extra[key] = fn()
            except Exception as exc:      # an extra key must never cost the headline line
                extra[key] = {"error": repr(exc)[:300]}
    rig.close()
    kept = (rig.g, rig.feats, w["shape"], w["feat"]) if rig.feats is not None else None    # re-used below
    del rig, step, g, batches

    # ---- the other BASELINE configurations, device-resident leg only (extra keys; the headline stays configs[2]) ----
    also = dict(extra)
    for name in ([] if args.no_also else [n for n in ("reddit_cvd", "powerlaw_ns") if n != args.workload]):
        try:
            w2 = WORKLOADS[name]
            k2 = min(K, 64)
            same = kept is not None and (w2["shape"], w2["feat"]) == kept[2:]
            r2 = Rig(args, name, w2, world, rank, dev, W + k2 + min(k2, args.graph_passes),
                     inputs=kept[:2] if same else None)
            t2, wm2 = r2.batches[W:W + k2], r2.batches[W + k2:]
            if not sharded_tables:
                r2.step.capture(r2.batches[0])
            se, fe = edge_counts(r2.g, t2, w2["degree"], w2["mode"] != "ns")
            if nccl_multi:
                barrier(); ev0.record()
                for b in t2:
                    r2.step.replay(b)
                ev1.record(); barrier()
                m2 = ev0.elapsed_time(ev1)
            else:
                m2, _, _ = time_trains(r2, args, t2, wm2, barrier, host_io=False)
            z2 = r2.step.sizes()
            a2 = r2.step.algorithmic_bytes(z2)
            (m2,), _ = reduce_times([m2])
            _, (se, fe) = reduce_times([float(se), float(fe)])
            also[name] = {"config": workload_config(argparse.Namespace(workload=name, scale=args.scale, seed=args.seed), w2)["workload"],
                          "steps": k2, "ms_per_step": m2 / k2, "value": (se + fe) / (m2 * 1e-3), "unit": "edges/s",
                          "sampled_edges_per_s": se / (m2 * 1e-3), "nodes": r2.g.n, "stored_edges": r2.g.nnz,
                          "last_step_sizes": z2, "algorithmic_bytes_per_step": a2["total"],
                          "frac_of_hbm_peak": a2["total"] / (m2 / k2 * 1e-3) / 1e9 / load_peaks()[0]}
            r2.close()
            del r2
        except Exception as exc:      # an extra key must never cost the headline line
            also[name] = {"error": repr(exc)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    total_edges = s_edges + f_edges
    value = total_edges / (ms * 1e-3)
    what = {"trains": "CUDA graph(s) of %d passes of the trains schedule: trains of %d batches sampled by one launch a "
                      "train ahead, gather one pass ahead, full-neighbour means back to back%s" % (
                          S, args.train, " (history write-back in the tail of each mean's launch)"
                          if args.fuse_write_back and world == 1 and w["mode"] != "ns" else ""),
            "graph": "CUDA graphs of %d steps; batch k+1's sampler (1 CTA) runs beside batch k's aggregate" % S,
            "native": "plain stream launches from C++ on three streams, two batches of sampler lookahead",
            "one-graph-per-step": "one CUDA graph per step, back to back"}[driver]
    line = {"metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, w), "workload_detail": detail,
            "sampled_edges_per_s": s_edges / (ms * 1e-3),
            "schedule": {"driver": driver, "passes_per_graph": S, "what": what,
                         "ms_per_step_one_graph_back_to_back": ms_serial / K},
            "step_hbm": {"algorithmic_bytes_per_step": alg["total"],
                         "achieved_gbs": alg["total"] / (ms / K * 1e-3) / 1e9,
                         "frac_of_peak": alg["total"] / (ms / K * 1e-3) / 1e9 / peak,
                         "stages": {k: v for k, v in alg.items() if k != "total"}},
            "e2e": {"value": total_edges / (ms_e2e * 1e-3), "unit": "edges/s",
                    "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(w["batch"] * 4),
                    "d2h_bytes_per_step": int(w["batch"] * alg_width(w) * 4),
                    "how": "pinned host id table in (H2D copy of each train's ids inside the graphs), every pass's "
                           "aggregated rows out to pinned host memory (one D2H copy per pass inside the graphs); the "
                           "caller waits for each graph's rows one replay behind the launches"},
            "gpu_launches": int(round(launches_per_step * K)),
            "launches_per_step": float(launches_per_step),
            "clocks": clock_info}
    if also:
        line["also"] = also
    if kern is not None:
        traffic = load_traffic(kern["kernel"])
        line["roofline"] = {"bound": "hbm", "kernel": kern["kernel"], "achieved": kern["bytes"] / kern["sec"] / 1e9,
                            "peak": peak, "unit": "GB/s", "frac": kern["bytes"] / kern["sec"] / 1e9 / peak,
                            "traffic": traffic["bytes"] if traffic else None,
                            "frac_dram": (traffic["bytes"] / kern["sec"] / 1e9 / peak) if traffic else None,
                            "traffic_source": traffic["source"] if traffic else None, "peak_source": peak_src,
                            "us_per_launch": kern["sec"] * 1e6, "algorithmic_bytes_per_launch": kern["bytes"],
                            "how": kern["how"]}
    if world == 1 and not args.no_cpu:
        g_cpu, feats_cpu = kept[:2] if kept is not None else build_inputs(w, args.seed, dev, args.scale)
        n_cpu = int(args.cpu_seconds * 400) + 8         # the CPU path takes >= 5 ms per step at every workload
        host_batches = [b.cpu().numpy() for b in make_batches(g_cpu.n, w["batch"], n_cpu, args.seed + 99, dev)]
        line["cpu_baseline"] = cpu_leg(w, g_cpu, feats_cpu, args.seed, host_batches, args.cpu_seconds, 3)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def alg_width(w):
    return w["hidden"] * 2            # [self | neighbour] rows (graphsage normalisation: every bench workload)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else (NCCL banners, library
    chatter) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)            # NCCL prints its version banner on fd 1: keep stdout for the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reddit_cv", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic graph (tests only)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="time one graph per step, no sampler lookahead")
    ap.add_argument("--graph-passes", type=int, default=64,
                    help="upper bound on the passes captured per CUDA graph (K <= this: one graph of K passes)")
    ap.add_argument("--train", type=int, default=16, help="batches sampled per launch by the trains schedule (2..32)")
    ap.add_argument("--first-train", type=int, default=4,
                    help="length of the first train of a graph (short: smaller start-up bubble)")
    ap.add_argument("--fuse-write-back", action="store_true",
                    help="single GPU: the history write-back in the tail of the full-neighbour mean's launch instead "
                         "of a launch of its own on the chain (A/B, see DESIGN section 1)")
    ap.add_argument("--eager-trains", action="store_true",
                    help="trains driver as plain stream launches from C++ (no CUDA graphs): A/B of the graph overhead")
    ap.add_argument("--no-also", action="store_true", help="skip the extra keys for the other BASELINE configurations")
    ap.add_argument("--driver", default="trains", choices=["trains", "native", "graph"],
                    help="schedule of the timed region: trains (default; csrc/step.cu:sgcn_step_run_trains as CUDA "
                         "graphs), native (round-1 C++ driver, plain stream launches), graph (round-1 multi-step "
                         "torch-captured graphs)")
    ap.add_argument("--tables", default="replicated", choices=["replicated", "sharded"],
                    help="multi-GPU: history + feature tables replicated on every GPU (default) or row-sharded with "
                         "remote rows read over NVLink (SURVEY 8e (1)); sharded needs the trains driver")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU write-back exchange: NVLink peer stores (default) or NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
