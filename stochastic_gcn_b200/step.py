"""One pass of the hot path over one batch -- the device-resident restatement of the reference's
per-step sequence (gcn/train.py:190,207; gcn/vrgcn.py:39-84; gcn/models.py:160-166,186-194):

    sampler.expand                      scheduler.cpp:46-189      (device sampler, no host trip)
    input = dense_slice(features, fields[0])   vrgcn.py:39-47    (row gather)
    Z = aggregate(adj, fadj, history, X)       layers.py:223-362 (fused CV / CVD / plain forward)
    dX = adj^T dZ (+ self rows)                 models.py:187     (SpMM backward)
    history[fields[0]] = X                      models.py:160-166 (row scatter-store)

All buffers are sized from upper bounds (|field| <= B(1+degree)) and every data-dependent length
is read by the kernels from the sampler's device meta block, so the whole pass is a fixed launch
sequence that is captured into CUDA graphs and replayed per batch.

The dense layers between the gather and the aggregate (``X = relu(LN(input @ W))``) are not part
of the hot path: the aggregator input X is taken as the first ``hidden`` columns of
the gathered feature rows (CVD: h = columns [0,hidden), mu = columns [hidden, 2 hidden)), and the
upstream gradient dZ is a resident synthetic tensor, so the data dependencies gather -> aggregate
-> backward -> write-back are the real ones.

Launch topology (measured on B200, tools/graph_overhead.py + tools/timeline.py): inside a CUDA
graph a dependent kernel starts ~0.8 us after its predecessor on the same branch and ~3 us after
one on another branch.  The pass therefore keeps its dominant kernel and the write-back that must
follow it on ONE chain,

    main :  full_mean ─► history_update ─► (next step) full_mean ─► ...
    side :  gather ─► sampled aggregate ─► dX init ─► SpMM backward ─► zero(next step's output)
    samp :  sampler of the NEXT batch (one CTA, into the other buffer set)

and everything short rides on branches that are long finished when the chain needs them.
"""
import torch

from . import _lib, ops
from .sampler import DeviceSampler

MODES = ("ns", "cv", "cvd")


class HotPathStep:
    def __init__(self, graph, features, hidden, batch_size, degree, mode="cv", normalization="graphsage",
                 seed=1, history=None):
        if mode not in MODES:
            raise ValueError("mode must be one of %s" % (MODES,))
        need = hidden * (2 if mode == "cvd" else 1)
        if features.shape[1] < need:
            raise ValueError("features need at least %d columns" % need)
        self.dev = features.device
        self.mode, self.hidden, self.B, self.degree = mode, int(hidden), int(batch_size), int(degree)
        self.concat = normalization != "gcn"
        self.features = features
        self.n_nodes = graph.n
        self.sampler = DeviceSampler(graph.data, graph.indices, graph.indptr, L=1, cv=mode != "ns")
        self.sampler.seed(seed)
        for slot in (0, 1, 2):                       # every per-batch buffer set, sized once
            self.sampler.set_slot(slot)
            self.sampler.reserve(self.B, [self.degree])
        self.sampler.set_slot(0)
        self.history = history if history is not None else torch.zeros(
            (graph.n, self.hidden), dtype=torch.float32, device=self.dev)     # vrgcn.py:23-36: zero-init
        self.n_in_bound = min(self.B * (1 + self.degree), max(graph.n, self.B))
        f = features.shape[1]
        width = self.hidden * (2 if self.concat else 1)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.dev)
        self.ids2 = [torch.zeros(self.B, dtype=torch.int32, device=self.dev) for _ in range(2)]
        self.ids = self.ids2[0]
        self.dynamic_full = False    # warps of full_mean_kernel pull chunks from a device counter
        self.x0 = z(self.n_in_bound, f)                 # gathered input rows
        # aggregated rows, one buffer per sampler buffer set: the set not in use is zeroed one step ahead
        self.outs = [z(self.B, width), z(self.B, width)]
        self.outs_mu = [z(self.B, width), z(self.B, width)] if mode == "cvd" else [None, None]
        self.d_out = z(self.B, width)                   # upstream gradient (synthetic, resident)
        self.dx = z(self.n_in_bound, self.hidden)
        self.graph = None
        self.graph_host = None
        self.launches_per_step = None
        self._views = {}
        self._last_sampler_slot = None      # set by the native driver (its sampler sets rotate over three)
        self._pinned_out = None
        self._pipe = None           # graphs of the cross-step pipelined driver
        self._pipeline_on = False
        self._pipe_done = None
        self._last_slot = 0
        self.train = 16                # batches sampled per launch by the trains schedule (run_trains)
        self.fuse_write_back = False   # trains schedule, single GPU: write-back in the tail of the full-neighbour mean
                                       # (measured equal-to-slower than a launch of its own: DESIGN section 1)
        self._trains = None            # captured graphs of the trains schedule
        self._last_x0 = self._last_dx = None
        self._s_b = torch.cuda.Stream(device=self.dev)      # side branch of the pass
        self._s_samp = torch.cuda.Stream(device=self.dev)   # sampler branch of the pipelined graphs
        self._s_chain = torch.cuda.Stream(device=self.dev)  # main chain of the pipelined graphs
        self._publish_stream = None                         # sharded steps: branch of the early write-back push

    @property
    def out(self):
        """aggregated rows of the most recent pass"""
        return self.outs[self._last_slot]

    @property
    def out_mu(self):
        return self.outs_mu[self._last_slot]

    # -- pieces ----------------------------------------------------------------------------------
    def _sample(self, slot=0, ids=None):
        s = self.sampler
        s.set_slot(slot)
        s.start_batch(self.ids2[slot] if ids is None else ids)
        s.expand(self.degree, materialize_full=False)
        return self._ensure_views(slot)

    def _ensure_views(self, slot):
        """zero-copy views of the sampler's buffer set `slot` (the addresses never change once reserved)"""
        if self._views.get(slot) is None:
            s = self.sampler
            s.set_slot(slot)
            names = ("field", "rowptr_s", "rowptr_f", "edg_t", "tgt", "edg_w", "scales", "meta")
            v = {k: s.view(k) for k in names}
            v["adj_p"], v["adj_i"], v["adj_w"] = s.view("adj_p"), s.view("adj_i"), s.view("adj_w")
            v["n_out_dev"], v["n_in_dev"], v["work"] = v["meta"][0:1], v["meta"][1:2], v["meta"][6:7]
            self._views[slot] = v
        return self._views[slot]

    def _out_views(self, slot):
        H = self.hidden
        out, out_mu = self.outs[slot], self.outs_mu[slot]
        nb = out[:, H:] if self.concat else out
        slf = out[:, :H] if self.concat else None
        nb_mu = slf_mu = None
        if self.mode == "cvd":
            nb_mu = out_mu[:, H:] if self.concat else out_mu
            slf_mu = out_mu[:, :H] if self.concat else None
        return nb, slf, nb_mu, slf_mu

    def _zero_out(self, slot):
        """Zero the halves of outs[slot] that the two aggregate kernels accumulate into."""
        if self.mode == "ns":
            return
        nb, _, nb_mu, _ = self._out_views(slot)
        ops.copy_rows_pad(None, 0, nb)
        if nb_mu is not None:
            ops.copy_rows_pad(None, 0, nb_mu)

    def _rest(self, slot, main, zero_next=False, after=None, fork=None, side=None, carry=None):
        """Everything after the sampler for the batch held by buffer set `slot`.

        outs[slot] must already be zero.  zero_next: also zero outs[1-slot] (for the next step) at the
        end of the side branch.  after(side_stream): optional extra work enqueued on the side branch
        once the aggregate is complete (e.g. the D2H copy of the step's result).  carry: a list used as a
        mailbox between consecutive steps of one capture -- the event that ends `after` is left there
        and awaited by the NEXT step's side branch (before it re-zeroes the buffer `after` reads) instead
        of by this step's main chain; the caller waits for whatever is left at the end."""
        v = self._ensure_views(slot)
        pipelined = self._pipeline_on        # the sampler's device guard counts consumer passes
        H, B = self.hidden, self.B
        cv = self.mode != "ns"
        nb, slf, nb_mu, slf_mu = self._out_views(slot)
        side = side or self._s_b
        ev_start = torch.cuda.Event()
        ev_start.record(main)

        # main chain first (capture order = the order the driver feeds its hardware queues): the
        # full-neighbour history mean (dominant); the write-back follows below
        if cv:
            ops.full_history_mean(v["field"], v["rowptr_f"], B, v["adj_p"], v["adj_i"], v["adj_w"],
                                  self.history, nb_mu if self.mode == "cvd" else nb,
                                  nb if self.mode == "cvd" else None, n_out_dev=v["n_out_dev"],
                                  work_counter=v["work"] if self.dynamic_full else None)
        ev_full = torch.cuda.Event()
        ev_full.record(main)
        if fork is not None:
            fork(ev_start)                       # e.g. the next batch's sampler branch

        # side branch: dX init (+ next step's output zeroing), feature-row gather, then the sampled part
        # of the aggregate fused with its backward -- three graph nodes
        with torch.cuda.stream(side):
            side.wait_event(ev_start)
            if carry:                            # the previous step's deferred `after` work (D2H of its rows)
                side.wait_event(carry.pop())     # must precede this branch's zeroing of that buffer
            d_nb = self.d_out[:, H:] if self.concat else self.d_out
            d_self = self.d_out[:, :H] if self.concat else None
            nxt = self._out_views(1 - slot) if (zero_next and cv) else None
            # dX init + next step's output zeroing + this step's feature-row gather: ONE graph node
            ops.gather_pad_pair(self.features, v["field"], self.x0, d_self, B if self.concat else 0, self.dx,
                                None, 0, nxt[0] if nxt else None, n_dev=v["n_in_dev"],
                                n0_dev=v["n_out_dev"] if self.concat else None)
            if nxt and nxt[2] is not None:
                ops.copy_rows_pad(None, 0, nxt[2])
            x = self.x0[:, :H]
            new_hist = None
            ev_pub = None
            if cv and self._publish_stream is not None:
                # the rows that will be written back exist now: a sharded step publishes them to its peers
                # on a branch of its own, beside the sampled aggregate (the peers need the payload by the
                # end of THEIR full-neighbour mean; the local write-back joins this branch)
                ev_g = torch.cuda.Event()
                ev_g.record(side)
                with torch.cuda.stream(self._publish_stream):
                    self._publish_stream.wait_event(ev_g)
                    self._publish_write_back(v, x if self.mode == "cv" else self.x0[:, H:2 * H])
                    ev_pub = torch.cuda.Event()
                    ev_pub.record(self._publish_stream)
            if cv:      # sharded + peer transport: the push rides on the sampled launch below (host-side only)
                self._attach_write_back_push(v, x if self.mode == "cv" else self.x0[:, H:2 * H])
            if self.mode == "ns":
                ops.spmm_csr(v["rowptr_s"], v["edg_t"], v["edg_w"], x, B, out=nb, n_out_dev=v["n_out_dev"])
                if self.concat:
                    ops.copy_rows_pad(x, B, slf, n_dev=v["n_out_dev"])
                ops.spmm_csr_bwd(v["rowptr_s"], v["edg_t"], v["edg_w"], d_nb, self.dx, B, n_out_dev=v["n_out_dev"])
            elif self.mode == "cv":
                ops.cv_sampled_fwd_bwd(v["rowptr_s"], v["edg_t"], v["edg_w"], v["tgt"], B, x, self.history, nb,
                                       d_nb, self.dx, self_out=slf, n_out_dev=v["n_out_dev"], accumulate=True)
                new_hist = x
            else:
                mu = self.x0[:, H:2 * H]
                ops.cvd_sampled_fwd_bwd(v["rowptr_s"], v["edg_t"], v["edg_w"], v["tgt"], v["scales"], B, x, mu,
                                        self.history, nb, nb_mu, d_nb, self.dx, self_h=slf, self_mu=slf_mu,
                                        n_out_dev=v["n_out_dev"], accumulate=True)
                new_hist = mu
            ev_fwd = torch.cuda.Event()          # last read of history on this branch
            ev_fwd.record(side)

        main.wait_event(ev_fwd)                  # every forward read of history precedes the write-back
        if ev_pub is not None:
            main.wait_event(ev_pub)
        marked = False
        if new_hist is not None:                 # (models.py:186-194)
            marked = self._write_back(v, new_hist, self._pipe_done if pipelined else None)
        if pipelined and not marked:             # adjacency rows of this batch are no longer read
            self.sampler.mark_consumed(main)
        if after is not None:
            with torch.cuda.stream(side):
                side.wait_event(ev_full)
                after(side)
                ev_side = torch.cuda.Event()
                ev_side.record(side)
            if carry is not None:                # off the critical path: the next step's side branch waits
                carry.append(ev_side)
            else:
                main.wait_event(ev_side)

    def _publish_write_back(self, v, new_hist):
        """Hook, on the side branch right after the gather: nothing to do on one GPU."""

    def _attach_write_back_push(self, v, new_hist):
        """Hook, right before the sampled-aggregate launch: nothing to do on one GPU."""

    def _write_back(self, v, new_hist, done_counter=None):
        """tf.scatter_update(history, fields[0], new_history); the sharded subclass exchanges instead.
        Returns True when it also bumped the pipelining guard's consumer counter (done_counter)."""
        ops.history_update(self.history, v["field"], new_hist, n_dev=v["n_in_dev"], done_counter=done_counter)
        return done_counter is not None

    def _pass(self):
        """One whole pass on the current stream: zero -> sampler -> rest (buffer set 0)."""
        main = torch.cuda.current_stream(self.dev)
        if getattr(self.sampler, "_stream", None) is None or self.sampler._stream.cuda_stream != main.cuda_stream:
            self.sampler.use_stream(main)
        self._zero_out(0)
        self._sample(0)
        self._last_slot, self._last_sampler_slot = 0, None
        self._last_x0 = self._last_dx = None
        self._rest(0, main)

    # -- drivers ---------------------------------------------------------------------------------
    def run(self, ids):
        """Eager pass (one launch per kernel).  ids: CUDA int32 [B] of distinct node ids."""
        self.ids.copy_(ids, non_blocking=True)
        before = _lib.launch_count()
        self._pass()
        self.launches_per_step = _lib.launch_count() - before
        return self.out

    def capture(self, warmup_ids):
        """Capture the pass into a CUDA graph (after one eager warm-up pass sized the buffers)."""
        self.run(warmup_ids)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.dev)
        self.sampler.use_stream(side)      # outside the capture: set_stream synchronises the old stream
        with torch.cuda.graph(g, stream=side):
            self._pass()
        self.graph = g
        self._capture_stream = side
        return g

    def replay(self, ids):
        self.ids.copy_(ids, non_blocking=True)
        self._last_slot, self._last_sampler_slot = 0, None
        self.graph.replay()
        return self.out

    def capture_host(self):
        """Second graph for the host-buffer API: H2D of the pinned ids, the pass, D2H of the result
        -- the copies are memcpy nodes of the same graph, so a step is ONE launch + ONE sync."""
        width = self.outs[0].shape[1]
        self._pin_ids = torch.zeros(self.B, dtype=torch.int32).pin_memory()
        self._pinned_out = torch.empty((self.B, width), dtype=torch.float32).pin_memory()
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        side = getattr(self, "_capture_stream", None) or torch.cuda.Stream(device=self.dev)
        self.sampler.use_stream(side)
        with torch.cuda.graph(g, stream=side):
            self.ids.copy_(self._pin_ids, non_blocking=True)
            self._pass()
            self._pinned_out.copy_(self.outs[0], non_blocking=True)
        self.graph_host = g
        self._capture_stream = side
        return g

    def step_host(self, ids_pinned):
        """End-to-end call with HOST buffers: int32 ids in (host memory), aggregated rows out (pinned
        host memory).  Per call: a 2 KB host copy into the staging buffer, one graph launch (H2D +
        pass + D2H), one stream synchronise."""
        self._last_slot, self._last_sampler_slot = 0, None
        if self.graph_host is not None:
            self._pin_ids.copy_(ids_pinned)
            self.graph_host.replay()
            torch.cuda.current_stream(self.dev).synchronize()
            return self._pinned_out
        if self._pinned_out is None:
            self._pinned_out = torch.empty(self.outs[0].shape, dtype=torch.float32).pin_memory()
        self.ids.copy_(ids_pinned, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._pass()
        self._pinned_out.copy_(self.outs[0], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self._pinned_out

    # -- cross-step pipelining --------------------------------------------------------------------
    def capture_pipelined(self, warm0, warm1, host_io=False, steps_per_graph=8):
        """Graphs of S consecutive steps.  Inside step k three branches run side by side: the main
        chain (full-neighbour mean -> write-back) and the side branch of pass k, both reading buffer
        set k&1, and the sampler of batch k+1 filling the other set.  The sampler is one
        latency-bound CTA (a 256-thread, 48-register variant that fits beside two resident
        full_mean CTAs): next to the aggregate that streams history rows it leaves the critical
        path.  Sampler order, RNG stream, history reads and write-backs stay exactly sequential; the
        in-place row permutation is guarded on the device against the one possible race (a node
        shared by consecutive batches, see sgcn_sampler_pipeline).

        host_io=True adds, as memcpy nodes, the H2D copy of the chunk's ids (from pinned staging)
        and the D2H copy of every step's aggregated rows (to pinned memory)."""
        S = int(steps_per_graph)
        if S < 2 or S % 2:
            raise ValueError("steps_per_graph must be even and >= 2")
        main = torch.cuda.current_stream(self.dev)
        for p, ids in ((0, warm0), (1, warm1)):       # eager warm-up of both buffer sets
            self.ids2[p].copy_(ids)
            self._eager_step(p, sample=True)
        torch.cuda.synchronize(self.dev)
        self._enable_pipeline_guard()
        B, width = self.B, self.outs[0].shape[1]
        # tab[c][k] = ids of the batch that step k of a parity-c chunk samples AHEAD (batch k+1 of the
        # chunk; row S-1 is the first batch of the next chunk)
        tab = [torch.zeros((S, B), dtype=torch.int32, device=self.dev) for _ in range(2)]
        pin_tab = pin_out = None
        if host_io:
            pin_tab = [torch.zeros((S, B), dtype=torch.int32).pin_memory() for _ in range(2)]
            pin_out = [torch.empty((S, B, width), dtype=torch.float32).pin_memory() for _ in range(2)]
        s_samp = self._s_samp
        self.sampler.use_stream(s_samp)               # outside any capture: set_stream synchronises
        side = torch.cuda.Stream(device=self.dev)

        pools = [[torch.cuda.Stream(device=self.dev) for _ in range(S)] for _ in range(2)]
        self._stream_pools = pools

        def set_prev(slot, ids_row):
            """host bookkeeping only: the batch held by `slot` lives at ids_row when the graph replays"""
            self.sampler.set_slot(slot)
            self.sampler.start_batch(ids_row)

        def capture_chunk(c, closed):
            g = torch.cuda.CUDAGraph()
            set_prev(0, tab[1 - c][S - 1])            # slot 0 was sampled from the previous chunk's last row
            chain = self._s_chain
            with torch.cuda.graph(g, stream=side):
                # the capture-origin stream only forks and joins: on this driver the first kernel of
                # the ORIGIN stream starts tens of microseconds after those of forked streams
                ev_in = torch.cuda.Event()
                ev_in.record(side)
                with torch.cuda.stream(chain):
                    chain.wait_event(ev_in)
                    if host_io:
                        tab[c].copy_(pin_tab[c], non_blocking=True)
                    carry = [] if host_io else None
                    for k in range(S):
                        slot_r, slot_s = k & 1, 1 - (k & 1)
                        ev_box = []

                        def fork(ev_start, slot_s=slot_s, k=k, ev_box=ev_box):
                            st = pools[0][k]                   # a stream of its own per step: the driver
                            self.sampler.use_stream(st, sync=False)   # feeds its queues in capture order
                            with torch.cuda.stream(st):
                                st.wait_event(ev_start)
                                self._sample(slot_s, tab[c][k])
                                ev_s = torch.cuda.Event()
                                ev_s.record(st)
                                ev_box.append(ev_s)
                        after = None
                        if host_io:
                            dst, src = pin_out[c][k], self.outs[slot_r]
                            after = lambda st, dst=dst, src=src: dst.copy_(src, non_blocking=True)
                        self._rest(slot_r, chain, zero_next=True, after=after,
                                   fork=None if (closed and k == S - 1) else fork, side=pools[1][k], carry=carry)
                        if ev_box:
                            chain.wait_event(ev_box[0])
                    if carry:
                        chain.wait_event(carry.pop())      # the last step's rows are on the host at graph end
                    ev_out = torch.cuda.Event()
                    ev_out.record(chain)
                side.wait_event(ev_out)
            return g

        first = torch.cuda.CUDAGraph()
        with torch.cuda.graph(first, stream=s_samp):
            if host_io:
                tab[1][S - 1].copy_(pin_tab[1][S - 1], non_blocking=True)
            self._zero_out(0)
            self._sample(0, tab[1][S - 1])
        self._pipe = {"S": S, "first": first, "tab": tab, "pin_tab": pin_tab, "pin_out": pin_out,
                      "host_io": host_io,
                      "open": [capture_chunk(0, False), capture_chunk(1, False)],
                      "closed": [capture_chunk(0, True), capture_chunk(1, True)]}
        return self._pipe

    def _eager_step(self, slot, sample):
        main = torch.cuda.current_stream(self.dev)
        if getattr(self.sampler, "_stream", None) is None or self.sampler._stream.cuda_stream != main.cuda_stream:
            self.sampler.use_stream(main)
        if sample:
            self._zero_out(slot)
            self._sample(slot, self.ids2[slot])
        self._last_slot, self._last_sampler_slot = slot, None
        self._last_x0 = self._last_dx = None
        self._rest(slot, main, zero_next=True)

    def run_pipelined(self, batches, on_chunk=None):
        """Run len(batches) consecutive passes with one-batch sampler lookahead on the current stream.
        Per chunk of S steps: the ids copy and ONE graph launch.  ``batches``: a contiguous int32
        [n, B] id table (one copy per chunk) or a list of [B] id tensors (one copy per step); CUDA, or
        host memory when captured with host_io=True.  ``on_chunk(first_step, count, step,
        done_event)`` is called after a chunk has been enqueued (host_io: once ``done_event`` has
        completed, ``step._pipe["pin_out"][c][:count]`` holds the rows of those steps, c = chunk
        parity; that buffer is rewritten by chunk c+2, so wait for the event before launching it --
        waiting one chunk behind keeps the GPU busy while the host consumes results)."""
        pipe = self._pipe
        # a contiguous [n, B] id table moves a chunk's ids with ONE copy (a list costs one per step)
        table = batches if isinstance(batches, torch.Tensor) and batches.dim() == 2 else None
        S, n = pipe["S"], (table.shape[0] if table is not None else len(batches))
        if n == 0:
            return self.out
        host_io = pipe["host_io"]
        stage = pipe["pin_tab"] if host_io else pipe["tab"]
        full, rem = divmod(n, S)
        done = pipe.setdefault("done", [None, None])
        if host_io and done[1] is not None:
            done[1].synchronize()                               # an earlier call's DMA may still read pin_tab[1]
        stage[1][S - 1].copy_(batches[0], non_blocking=not host_io)
        pipe["first"].replay()                                  # zero outs[0]; sample batch 0 into set 0
        if host_io:                                             # its H2D reads pin_tab[1][S-1]: chunk 1 restages it
            done[1] = torch.cuda.Event()
            done[1].record()
        for c in range(full):
            par = c & 1
            base = c * S
            closed = rem == 0 and c == full - 1
            ahead = batches[base + 1: base + S + (0 if closed else 1)]
            dst = stage[par]
            if host_io and pipe.get("done", [None, None])[par] is not None:
                pipe["done"][par].synchronize()                 # chunk c-2 has consumed this staging buffer
            if table is not None:
                if len(ahead):
                    dst[:len(ahead)].copy_(ahead, non_blocking=not host_io)
            else:
                for k, ids in enumerate(ahead):
                    dst[k].copy_(ids, non_blocking=not host_io)
            pipe["closed" if closed else "open"][par].replay()
            self._last_slot, self._last_sampler_slot = (S - 1) & 1, None
            ev = torch.cuda.Event()
            ev.record()
            pipe.setdefault("done", [None, None])[par] = ev
            if on_chunk is not None:
                on_chunk(base, S, self, ev)
        if rem:                                                 # tail shorter than a chunk: eager passes
            base = full * S
            for j in range(rem):
                slot = j & 1
                if j > 0:
                    self.ids2[slot].copy_(batches[base + j], non_blocking=True)
                self._eager_step(slot, sample=j > 0)            # batch `base` was already sampled ahead
                if on_chunk is not None:
                    ev = torch.cuda.Event()
                    ev.record()
                    on_chunk(base + j, 1, self, ev)
        return self.out

    # -- native driver ----------------------------------------------------------------------------
    def _enable_pipeline_guard(self):
        if not self._pipeline_on:
            self.sampler.pipeline(True)
            self._pipeline_on = True
            self._pipe_done = self.sampler.view("pipe")[1:2]    # the guard's consumer counter

    def _step_desc(self):
        """sgcn_step_desc of this step (the sharded subclass adds the exchange fields)."""
        d = _lib.StepDesc()
        d.mode = MODES.index(self.mode)
        d.concat = int(self.concat)
        d.batch, d.degree, d.hidden, d.feat_dim = self.B, self.degree, self.hidden, self.features.shape[1]
        d.x0_rows = self.x0.shape[0]
        d.world, d.rank, d.wb_bound = 1, 0, 0
        d.features, d.ld_feat = self.features.data_ptr(), self.features.stride(0)
        d.history, d.ld_hist = self.history.data_ptr(), self.history.stride(0)
        d.x0, d.ld_x0 = self.x0.data_ptr(), self.x0.stride(0)
        for i in (0, 1):
            d.out[i] = self.outs[i].data_ptr()
            d.out_mu[i] = self.outs_mu[i].data_ptr() if self.outs_mu[i] is not None else None
        d.ld_out = self.outs[0].stride(0)
        d.d_out, d.ld_dout = self.d_out.data_ptr(), self.d_out.stride(0)
        d.dx, d.ld_dx = self.dx.data_ptr(), self.dx.stride(0)
        if getattr(self, "_alt_bufs", None) is None:   # trains schedule: three x0 copies, two dx copies
            self._alt_bufs = (torch.zeros_like(self.x0), torch.zeros_like(self.x0), torch.zeros_like(self.dx))
        d.x0_alt[0], d.x0_alt[1] = self._alt_bufs[0].data_ptr(), self._alt_bufs[1].data_ptr()
        d.dx_alt = self._alt_bufs[2].data_ptr()
        d.train, d.fuse_write_back = int(self.train), int(bool(self.fuse_write_back))
        return d

    def run_native(self, batches, out_host=None):
        """Run len(batches) consecutive passes through the native step driver (csrc/step.cu): plain
        stream launches from C++ on three streams, the sampler of batch k+1 beside the aggregate of
        batch k -- one ctypes call for the whole run.  ``batches``: an int32 [n, B] tensor or a list
        of [B] tensors, on the GPU or in (pinned) host memory.  ``out_host``: optional pinned float32
        [n, B, width] tensor that receives every pass's aggregated rows."""
        import ctypes as C
        lib = _lib.load()
        self._native_handle()
        table = batches if isinstance(batches, torch.Tensor) else torch.stack(list(batches))
        table = table.to(torch.int32).contiguous()
        n = int(table.shape[0])
        if n == 0:
            return self.out
        if table.shape[1] != self.B:
            raise ValueError("every batch must hold exactly %d ids" % self.B)
        on_host = not table.is_cuda
        if on_host and not table.is_pinned():
            table = table.pin_memory()
        if out_host is not None and not (out_host.is_pinned() and out_host.is_contiguous()
                                         and tuple(out_host.shape) == (n, self.B, self.outs[0].shape[1])):
            raise ValueError("out_host must be a pinned contiguous [n, B, width] float32 tensor")
        self._native_keep = (table, out_host)          # borrowed by the driver until the run completes
        _lib.check(lib.sgcn_step_run(self._native_h, _lib.ptr(table), int(on_host), n,
                                     _lib.ptr(out_host) if out_host is not None else None, _lib.stream_ptr()))
        self._last_slot = (n - 1) & 1                  # output buffers alternate ...
        self._last_sampler_slot = (n - 1) % 3          # ... sampler buffer sets rotate over three
        self._last_x0 = self._last_dx = None
        self.sampler._stream = None                    # the driver left the sampler on its own stream
        return self.out

    # -- native handle of the C++ step drivers ---------------------------------------------------------------
    def _native_handle(self):
        import ctypes as C
        if getattr(self, "_native_h", None) is None:
            self._enable_pipeline_guard()
            desc = self._step_desc()
            h = C.c_void_p()
            _lib.check(_lib.load().sgcn_step_create(C.byref(h), self.sampler._h, C.byref(desc)))
            self._native_h, self._native_desc = h, desc
        return self._native_h

    # -- trains schedule (csrc/step.cu:sgcn_step_run_trains): the schedule bench.py times ------------------
    def _check_io(self, table, out_host):
        n = int(table.shape[0])
        if table.dim() != 2 or table.shape[1] != self.B or table.dtype != torch.int32 or not table.is_contiguous():
            raise ValueError("ids must be a contiguous int32 [n, %d] table" % self.B)
        if not table.is_cuda and not table.is_pinned():
            raise ValueError("host id tables must be pinned")
        if out_host is not None and not (out_host.is_pinned() and out_host.is_contiguous()
                                         and tuple(out_host.shape) == (n, self.B, self.outs[0].shape[1])):
            raise ValueError("out_host must be a pinned contiguous [n, B, width] float32 tensor")
        return n

    def _trains_done(self, n, first_train):
        """host bookkeeping after n passes of the trains schedule: where the last pass left its results"""
        T = self.train
        T0 = min(first_train, T) if first_train > 0 else T
        k = n - 1
        c = 0 if k < T0 else 1 + (k - T0) // T
        base = 0 if c == 0 else T0 + (c - 1) * T
        self._last_slot = k & 1
        self._last_sampler_slot = (c & 1) * T + (k - base)
        x0s = (self.x0,) + tuple(self._alt_bufs[:2])
        self._last_x0, self._last_dx = x0s[k % 3], (self.dx, self._alt_bufs[2])[k & 1]
        self.sampler._stream = None

    def run_trains(self, table, out_host=None, first_train=0):
        """n consecutive passes through sgcn_step_run_trains on the current stream: the sampler runs a TRAIN of
        `self.train` batches per launch, one train ahead of the passes; gather / dX init / zeroing one pass
        ahead; full-neighbour means back to back (the write-back rides in their tails when self.fuse_write_back).
        ``table``: contiguous int32 [n, B] ids on the GPU or in PINNED host memory; ``out_host``: optional pinned
        float32 [n, B, width] receiving every pass's aggregated rows; ``first_train``: length of the first
        train (0 = self.train; a short one shortens the start-up bubble).  Returns the last pass's rows."""
        n = self._check_io(table, out_host)
        if n == 0:
            return self.out
        h = self._native_handle()
        self._trains_keep = (table, out_host)          # borrowed by the driver until the run completes
        _lib.check(_lib.load().sgcn_step_run_trains(h, _lib.ptr(table), int(not table.is_cuda), n,
                                                    _lib.ptr(out_host) if out_host is not None else None,
                                                    int(first_train), _lib.stream_ptr()))
        self._trains_done(n, first_train)
        return self.out

    def capture_trains(self, n, warm_table, host_io=False, first_train=4):
        """Capture n passes of the trains schedule as CUDA graph(s) over fixed id tables.
        host_io=False: ONE graph over a device [n, B] id table (``replay_trains`` copies the ids in).
        host_io=True: TWO graphs, each bound to its own pinned staging set (ids in: one H2D copy per train;
        every pass's rows out: one D2H copy per pass), alternated by ``replay_trains`` so that the host fills /
        drains one set while the GPU runs the other.  ``warm_table``: [n, B] device ids for the eager warm-up."""
        n = int(n)
        if tuple(warm_table.shape) != (n, self.B):
            raise ValueError("warm_table must be [n, batch]")
        tab = warm_table.to(torch.int32).contiguous().clone()
        self.run_trains(tab, first_train=first_train)             # sizes everything, warms the allocators
        torch.cuda.synchronize(self.dev)
        cap = torch.cuda.Stream(device=self.dev)
        width = self.outs[0].shape[1]
        if not host_io:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap):
                self.run_trains(tab, first_train=first_train)
            self._trains = {"n": n, "tab": tab, "graph": g, "host_io": False, "first_train": first_train}
            return g
        pin_tab = [torch.zeros((n, self.B), dtype=torch.int32).pin_memory() for _ in range(2)]
        pin_out = [torch.empty((n, self.B, width), dtype=torch.float32).pin_memory() for _ in range(2)]
        graphs = []
        for p in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap):
                self.run_trains(pin_tab[p], out_host=pin_out[p], first_train=first_train)
            graphs.append(g)
        self._trains = {"n": n, "pin_tab": pin_tab, "pin_out": pin_out, "graphs": graphs, "host_io": True,
                        "done": [None, None], "first_train": first_train}
        return graphs

    def replay_trains(self, table, on_chunk=None):
        """len(table) passes as replays of the captured graph(s); len(table) must be a multiple of the captured n.
        Captured with host_io: ``table`` is a HOST [m, B] id table and ``on_chunk(first_pass, count, rows,
        done_event)`` is called after each replay has been enqueued -- once ``done_event`` has completed,
        ``rows`` ([n, B, width], pinned) holds those passes' rows; that staging set is refilled two replays
        later, so consume it one replay behind the launches."""
        t = self._trains
        n = t["n"]
        m = int(table.shape[0])
        if m % n:
            raise ValueError("the number of passes must be a multiple of the captured %d" % n)
        for c in range(m // n):
            if not t["host_io"]:
                t["tab"].copy_(table[c * n:(c + 1) * n], non_blocking=True)
                t["graph"].replay()
                continue
            p = c & 1
            if t["done"][p] is not None:
                t["done"][p].synchronize()                       # replay c-2 has left this staging set
            t["pin_tab"][p].copy_(table[c * n:(c + 1) * n])      # host -> pinned host
            t["graphs"][p].replay()
            ev = torch.cuda.Event()
            ev.record()
            t["done"][p] = ev
            if on_chunk is not None:
                on_chunk(c * n, n, t["pin_out"][p], ev)
        self._trains_done(n, t["first_train"])
        return self.out

    def check_flags(self):
        """Raise if a device-side wait of the fused write-back timed out (synchronises)."""
        import ctypes as C
        if getattr(self, "_native_h", None) is None:
            return
        bad = C.c_int32()
        _lib.check(_lib.load().sgcn_step_status(self._native_h, C.byref(bad)))
        if bad.value:
            raise _lib.SgcnError(_lib.SGCN_EDATA, "a device-side wait of the fused write-back timed out")

    @property
    def last_x0(self):
        """gathered input rows of the most recent pass (the trains schedule rotates three copies)"""
        return self._last_x0 if self._last_x0 is not None else self.x0

    @property
    def last_dx(self):
        return self._last_dx if self._last_dx is not None else self.dx

    def time_dominant_kernel(self, batches):
        """Device time of the dominant kernel -- the edge-balanced full-neighbour history mean for
        CV/CVD, the feature-row gather for NS -- launched back to back on ONE stream over the inputs
        of len(batches) DIFFERENT batches (so the cache state is the workload's own, not a replay of
        one batch), CUDA events around the train of launches; returns the average per launch and the
        algorithmic bytes of exactly those launches."""
        dev, B, H = self.dev, self.B, self.hidden
        v = self._ensure_views(0)
        deg = (v["adj_p"][1:] - v["adj_p"][:-1])
        scratch = torch.zeros_like(self.outs[0])
        total_bytes, calls = 0, []
        if self.mode == "ns":
            for b in batches:
                self.run(b)
                z = self.sizes()
                field = v["field"][:z["n_in"]].clone()
                total_bytes += self.algorithmic_bytes(z)["gather"]
                calls.append(field)
            launch = lambda f: ops.gather_rows(self.features, f, out=self.x0)
            name = "move_rows_vec4_kernel<0> (feature-row gather)"
        else:
            nb = scratch[:, H:] if self.concat else scratch
            for b in batches:
                d = deg[b.long()]
                rowptr_f = torch.zeros(B + 1, dtype=torch.int32, device=dev)
                rowptr_f[1:] = torch.cumsum(d, 0)
                nnz_f = int(rowptr_f[-1])
                total_bytes += 8 * nnz_f + 4 * (B + 1) + 4 * H * nnz_f
                calls.append((b.contiguous(), rowptr_f))
            launch = lambda c: ops.full_history_mean(c[0], c[1], B, v["adj_p"], v["adj_i"], v["adj_w"],
                                                     self.history, nb)
            name = "full_mean_kernel"
        for c in calls[:3]:
            launch(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for c in calls:
            launch(c)
        e1.record()
        torch.cuda.synchronize(dev)
        n = max(len(calls), 1)
        return {"kernel": name, "sec": e0.elapsed_time(e1) * 1e-3 / n, "bytes": total_bytes / n, "launches": n,
                "how": "CUDA events around %d back-to-back launches on one stream, each on a different batch of "
                       "the timed region (inputs > L2 in aggregate; no flush)" % n}

    def sizes(self):
        """(n_out, n_in, nnz_s, nnz_f) of the last pass (synchronises)."""
        slot = self._last_sampler_slot if self._last_sampler_slot is not None else self._last_slot
        m = self._ensure_views(slot)["meta"].cpu().tolist()
        if m[5]:
            raise _lib.SgcnError(_lib.SGCN_EDATA, "sampler status %d" % m[5])
        return {"n_out": m[0], "n_in": m[1], "nnz_s": m[2], "nnz_f": m[3]}

    def algorithmic_bytes(self, z=None):
        """SURVEY.md 8d per-step algorithmic HBM bytes, split per stage, for the measured sizes."""
        z = z or self.sizes()
        D, F = self.hidden, self.features.shape[1]
        n_out, n_in, s, f = z["n_out"], z["n_in"], z["nnz_s"], z["nnz_f"]
        cv = self.mode != "ns"
        b = {}
        b["sampler"] = 8 * n_out + 2 * 2 * 8 * s + (16 if cv else 12) * s + 4 * (n_in + 2 * n_out)
        b["gather"] = 2 * 4 * F * n_in + 4 * n_in
        fwd = 8 * s + 4 * (n_out + 1) + 4 * D * s + 4 * D * n_out
        if self.concat:
            fwd += 2 * 4 * D * n_out
        if cv:
            fwd += 4 * D * s + 4 * s
        if self.mode == "cvd":
            fwd += 4 * D * s + 4 * n_out + 4 * D * n_out * (2 if self.concat else 1)
        b["aggregate_sampled"] = fwd
        b["aggregate_full"] = (8 * f + 4 * (n_out + 1) + 4 * D * f) if cv else 0
        b["backward"] = 8 * s + 4 * D * n_out * (2 if self.concat else 1) + 2 * 4 * D * s + 4 * D * n_in
        b["writeback"] = (4 * n_in + 8 * D * n_in) if cv else 0
        b["total"] = sum(b.values())
        return b
