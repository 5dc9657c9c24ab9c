"""Row-range sharding of the hot path across the GPUs of one node.

One process per GPU (torchrun).  Rank r owns the contiguous node range [lo, hi): it holds only
those rows of the adjacency (its sampler permutes only its own rows, so shards never conflict) and
draws its batches from that range, which keeps ``expand`` local for the 2-layer + preprocessing
models the reference trains (L = 1 after PP, gcn/train.py:86).  What a batch needs from OTHER
shards are history rows and input-feature rows of its neighbours.

B200-first layout decision: with 180 GB per GPU the history table (119 MB at Reddit shape, 1 GB
at 2M nodes) and the PP feature matrix (1.1 GB / 4 GB) are REPLICATED, and the replicas are kept
coherent by exchanging only what changes -- each rank's per-step write-back rows (<= B(1+d) rows,
0.8 MB) -- instead of fetching the ~150k distinct neighbour rows (75 MB) a Reddit-shaped batch
would need from remote shards every step.  Two transports for that exchange:

  "nccl"  pack -> ``all_gather_into_tensor`` -> apply                     (baseline)
  "peer"  pack straight into every peer's NVLink-mapped receive slot + flag, then wait + apply;
          no collective launch, the whole step stays one CUDA graph       (default)

Parity definition: rank r is bit-identical (sampled indices) to a reference ``Scheduler`` built on
the same global CSR, seeded ``seed + r`` and fed rank r's batches; history semantics are those of
R reference processes sharing one table with synchronous steps (every rank reads the pre-step
table; write-backs are applied in rank order, the highest rank winning a contended row).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr
from .graphs import CSRGraph
from .step import HotPathStep

HEADER_INTS = 4


def row_range(n, rank, world):
    """Contiguous node range of `rank`: [n*rank//world, n*(rank+1)//world)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def restrict_rows(graph, lo, hi):
    """The global CSR with every row outside [lo, hi) emptied (column ids stay global): what rank
    [lo, hi) stores.  Row pointers keep all N+1 entries so node ids need no translation."""
    indptr = graph.indptr.long()
    a, b = int(indptr[lo]), int(indptr[hi])
    local = (indptr.clamp(min=a, max=b) - a).to(torch.int32)
    return CSRGraph(graph.data[a:b].contiguous(), graph.indices[a:b].contiguous(), local.contiguous(), graph.n)


def payload_layout(n_bound, d):
    """Byte offsets of one write-back payload (mirrors csrc/exchange.cu): (ids, rows, total)."""
    ids = HEADER_INTS * 4
    rows = ids + ((n_bound * 4 + 15) & ~15)
    total = (rows + n_bound * d * 4 + 255) & ~255
    return ids, rows, total


def unpack_payload(buf, n_bound, d):
    """(ids, rows) views of one payload held in a uint8 array (host-side helper for tests / debugging)."""
    buf = np.asarray(buf, dtype=np.uint8)
    ids_off, rows_off, _ = payload_layout(n_bound, d)
    count = int(buf[:4].view(np.int32)[0])
    ids = buf[ids_off:ids_off + 4 * count].view(np.int32)
    rows = buf[rows_off:rows_off + 4 * count * d].view(np.float32).reshape(count, d)
    return ids, rows


def merge_payloads(history, payloads, n_bound, d):
    """Host restatement of the apply rule (rank order, highest rank wins): used by the CPU tests."""
    for buf in payloads:
        ids, rows = unpack_payload(buf, n_bound, d)
        history[ids] = rows
    return history


class _DevBytes:
    def __init__(self, addr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "|u1", "data": (int(addr), False),
                                         "version": 2, "strides": None}


RING = 8      # receive areas of the ring form (include/sgcn_b200.h: sgcn_wb_push_ring)


class PeerExchange:
    """cudaIpc plumbing of the peer transport: one exported allocation per rank holding
    [flags int32[64] | epoch, timeout, ... int32[64] | RING receive areas of `world` slots each].
    The two-area (even / odd) protocol of the one-pass drivers uses areas 0 and 1, the ring protocol of the
    trains schedule all RING of them (never both in one run: they share the flag array)."""

    def __init__(self, rank, world, slot_bytes, device):
        import torch.distributed as dist
        lib = _lib.load()
        self.lib, self.rank, self.world, self.slot = lib, rank, world, int(slot_bytes)
        self.off_ctl, self.off_even = 256, 512
        self.off_ring_flags = 128                   # ring-form flags: ints 32 .. 47 of the flag block
        self.off_reads, self.off_applied = 320, 384  # sharded tables: "reads done" / "applied" flags (control block)
        self.ring, self.ring_stride = RING, world * self.slot
        self.off_odd = self.off_even + self.ring_stride
        total = self.off_even + self.ring * self.ring_stride
        base = C.c_void_p()
        check(lib.sgcn_ipc_alloc(C.byref(base), total, 1))
        self.base, self.total = base.value, total
        handle = (C.c_ubyte * 64)()
        check(lib.sgcn_ipc_export(C.c_void_p(self.base), handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        self.peer_base = []
        for r in range(world):
            if r == rank:
                self.peer_base.append(self.base)
            else:
                p = C.c_void_p()
                h = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                check(lib.sgcn_ipc_open(h, C.byref(p)))
                self.peer_base.append(p.value)
        arr = lambda vals: (C.c_void_p * world)(*[C.c_void_p(v) for v in vals])
        self.dst_even = arr([b + self.off_even + rank * self.slot for b in self.peer_base])
        self.dst_odd = arr([b + self.off_odd + rank * self.slot for b in self.peer_base])
        self.peer_flags = arr(self.peer_base)
        self.flags = C.c_void_p(self.base)
        self.epoch = C.c_void_p(self.base + self.off_ctl)
        self.timeout = C.c_void_p(self.base + self.off_ctl + 4)
        self.block_counter = C.c_void_p(self.base + self.off_ctl + 8)   # scratch of the fused pack + signal
        self.apply_epoch = C.c_void_p(self.base + self.off_ctl + 12)    # ring form: epochs applied so far
        self.apply_stash = C.c_void_p(self.base + self.off_ctl + 16)    # ring form: claim -> copy hand-over
        self.push_epoch = C.c_void_p(self.base + self.off_ctl + 20)     # ring form: epochs pushed so far
        self.recv_even = C.c_void_p(self.base + self.off_even)
        self.recv_odd = C.c_void_p(self.base + self.off_odd)
        self._ctl = torch.as_tensor(_DevBytes(self.base, 512), device=device).view(torch.int32)
        dist.barrier()

    def control(self):
        """(flags[world], epoch, timeout_flag) read back from the device (synchronises)."""
        c = self._ctl.cpu()
        return c[:self.world].tolist(), int(c[64]), int(c[65])

    def close(self):
        for r, b in enumerate(self.peer_base):
            if r != self.rank and b:
                self.lib.sgcn_ipc_close(C.c_void_p(b))
        if self.base:
            self.lib.sgcn_ipc_free(C.c_void_p(self.base))
            self.base = 0


class IpcShards:
    """One row-sharded [N, width] fp32 table: this rank's rows [rank * rows, ...) in a cudaIpc-exported
    allocation, every other rank's shard mapped into this process over NVLink."""

    def __init__(self, rank, world, rows, width, device, init=None):
        import torch.distributed as dist
        lib = _lib.load()
        self.lib, self.rank, self.world, self.rows, self.width = lib, rank, world, int(rows), int(width)
        nbytes = self.rows * self.width * 4
        base = C.c_void_p()
        check(lib.sgcn_ipc_alloc(C.byref(base), nbytes, 1))
        self.base = base.value
        self.local = torch.as_tensor(_DevBytes(self.base, nbytes), device=device).view(torch.float32).view(
            self.rows, self.width)
        if init is not None:
            self.local[:init.shape[0]].copy_(init)
        torch.cuda.synchronize(device)
        handle = (C.c_ubyte * 64)()
        check(lib.sgcn_ipc_export(C.c_void_p(self.base), handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        self.peer_base = []
        for r in range(world):
            if r == rank:
                self.peer_base.append(self.base)
            else:
                p = C.c_void_p()
                check(lib.sgcn_ipc_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p)))
                self.peer_base.append(p.value)
        dist.barrier()

    def gather_full(self, n):
        """the whole table [n, width] assembled on this rank (tests / checkpoints): peer shards read over NVLink"""
        parts = []
        for r in range(self.world):
            nbytes = self.rows * self.width * 4
            t = torch.as_tensor(_DevBytes(self.peer_base[r], nbytes), device=self.local.device).view(torch.float32)
            parts.append(t.view(self.rows, self.width))
        return torch.cat(parts)[:n].clone()

    def close(self):
        for r, b in enumerate(self.peer_base):
            if r != self.rank and b:
                self.lib.sgcn_ipc_close(C.c_void_p(b))
        if self.base:
            self.local = None
            self.lib.sgcn_ipc_free(C.c_void_p(self.base))
            self.base = 0


def shard_rows(n, world):
    """rows per shard of the sharded-table layout: ceil(n / world) rounded up to 8 (row i lives on rank
    i // rows; the last shard may be short)"""
    return ((n + world - 1) // world + 7) // 8 * 8


class ShardedHotPathStep(HotPathStep):
    """HotPathStep over one row-range shard, with the cross-GPU history write-back exchange.

    tables="replicated" (default): every rank holds the whole history table and PP feature matrix; only each
    pass's write-back rows cross NVLink.  tables="sharded" ("boundary fetch"): rank r holds rows
    [r * S, (r + 1) * S) of both tables (S = shard_rows(N, world)) and reads every other row its batches touch
    straight out of the owner's HBM over NVLink -- memory per GPU falls with the number of GPUs, each pass moves
    its (R-1)/R of distinct neighbour rows across the links.  Only the trains schedule (run_trains /
    capture_trains / replay_trains) runs on sharded tables."""

    def __init__(self, graph, features, hidden, batch_size, degree, mode="cv", normalization="graphsage",
                 seed=1, rank=0, world=1, transport="peer", tables="replicated"):
        if transport not in ("nccl", "peer"):
            raise ValueError("transport must be 'nccl' or 'peer'")
        if tables not in ("replicated", "sharded"):
            raise ValueError("tables must be 'replicated' or 'sharded'")
        if tables == "sharded" and (transport != "peer" or world < 2):
            raise ValueError("sharded tables need the peer transport and at least two ranks")
        self.rank, self.world, self.transport, self.tables = int(rank), int(world), transport, tables
        self._hist_shards = self._feat_shards = None
        history = None
        if tables == "sharded":
            S = self.shard_rows = shard_rows(graph.n, world)
            self.lo, self.hi = min(graph.n, rank * S), min(graph.n, (rank + 1) * S)
            dev = features.device
            # `features` may be the whole [N, F] matrix (this rank keeps its rows) or already the local rows
            mine = features[self.lo:self.hi] if features.shape[0] == graph.n else features
            self._feat_shards = IpcShards(rank, world, S, features.shape[1], dev, init=mine)
            features = self._feat_shards.local
            if mode != "ns":
                self._hist_shards = IpcShards(rank, world, S, hidden, dev)
                history = self._hist_shards.local
        else:
            self.lo, self.hi = row_range(graph.n, rank, world)
        local = restrict_rows(graph, self.lo, self.hi)
        super().__init__(local, features, hidden, batch_size, degree, mode=mode, normalization=normalization,
                         seed=seed, history=history)
        self._exchange = None
        if self.mode == "ns":
            return     # plain neighbour sampling keeps no history: shards are independent
        # peer transport: the push is carried by the sampled-aggregate launch (sgcn_wb_push_attach) unless
        # SGCN_FUSE_PUSH=0, in which case -- as for NCCL's pack -- it runs on a branch of its own
        self._fuse_push = transport == "peer" and os.environ.get("SGCN_FUSE_PUSH", "1") != "0"
        self._publish_stream = None if self._fuse_push else torch.cuda.Stream(device=self.dev)
        lib = _lib.load()
        nb = self.sampler_n_in_bound()
        self.slot_bytes = int(lib.sgcn_wb_payload_bytes(nb, self.hidden))
        self.wb_bound = nb
        self.owner = torch.full((graph.n,), -1, dtype=torch.int32, device=self.dev)
        if transport == "nccl":
            self.send = torch.zeros(self.slot_bytes, dtype=torch.uint8, device=self.dev)
            self.recv = torch.zeros(self.slot_bytes * world, dtype=torch.uint8, device=self.dev)
            self._send_ptr = (C.c_void_p * 1)(C.c_void_p(self.send.data_ptr()))
        else:
            self._exchange = PeerExchange(rank, world, self.slot_bytes, self.dev)

    def sampler_n_in_bound(self):
        return min(self.B * (1 + self.degree), max(self.n_nodes, self.B))

    # hooks called by HotPathStep._rest ---------------------------------------------------------
    def _publish_write_back(self, v, new_hist):
        """Side branch, right after the gather: the rows this rank will write back already exist, so
        they cross NVLink (or are packed for NCCL) while the full-neighbour mean is still running."""
        lib, D = _lib.load(), self.hidden
        ld = new_hist.stride(0)
        if self.transport == "peer":
            x = self._exchange
            check(lib.sgcn_wb_push(ptr(v["field"]), ptr(v["n_in_dev"]), self.wb_bound, ptr(new_hist), ld, D,
                                   x.dst_even, x.dst_odd, self.world, x.peer_flags, self.rank, x.epoch,
                                   x.block_counter, stream_ptr()))
        else:
            check(lib.sgcn_wb_pack(ptr(v["field"]), ptr(v["n_in_dev"]), self.wb_bound, ptr(new_hist), ld, D,
                                   self._send_ptr, 1, 0, stream_ptr()))

    def _attach_write_back_push(self, v, new_hist):
        if not self._fuse_push:
            return
        x = self._exchange
        check(_lib.load().sgcn_wb_push_attach(ptr(v["field"]), ptr(v["n_in_dev"]), self.wb_bound, ptr(new_hist),
                                              new_hist.stride(0), self.hidden, x.dst_even, x.dst_odd, self.world,
                                              x.peer_flags, self.rank, x.epoch, x.block_counter))

    def _write_back(self, v, new_hist, done_counter=None):
        """Main chain, after every forward read of history: merge all ranks' payloads."""
        if self.transport == "peer":
            x = self._exchange
            check(_lib.load().sgcn_wb_wait_apply(ptr(self.history), self.history.stride(0), self.hidden,
                                                 x.recv_even, x.recv_odd, self.slot_bytes, self.world,
                                                 self.wb_bound, ptr(self.owner), x.flags, x.epoch, x.timeout,
                                                 ptr(done_counter), stream_ptr()))
            return done_counter is not None
        return False

    def _step_desc(self):
        d = super()._step_desc()
        if self._exchange is not None:
            x = self._exchange
            d.world, d.rank, d.wb_bound, d.slot_bytes = self.world, self.rank, self.wb_bound, self.slot_bytes
            for i in range(self.world):
                d.dst_even[i], d.dst_odd[i], d.peer_flags[i] = x.dst_even[i], x.dst_odd[i], x.peer_flags[i]
            d.recv_even, d.recv_odd, d.flags = x.recv_even.value, x.recv_odd.value, x.flags.value
            d.epoch, d.timeout_flag, d.block_counter = x.epoch.value, x.timeout.value, x.block_counter.value
            d.owner = self.owner.data_ptr()
            if os.environ.get("SGCN_WB_RING", "1") != "0" or self.tables == "sharded":   # trains schedule: ring form
                d.ring, d.ring_stride = x.ring, x.ring_stride
                d.push_epoch, d.apply_epoch, d.apply_stash = x.push_epoch.value, x.apply_epoch.value, x.apply_stash.value
                d.ring_flags, d.ring_recv = x.base + x.off_ring_flags, x.base + x.off_even
                for i in range(self.world):
                    d.ring_dst[i] = x.peer_base[i] + x.off_even + self.rank * x.slot
                    d.ring_peer_flags[i] = x.peer_base[i] + x.off_ring_flags
                if self.tables == "sharded":
                    d.reads_flags, d.applied_flags = x.base + x.off_reads, x.base + x.off_applied
                    d.shard_counter = x.base + x.off_ctl + 24
                    for i in range(self.world):
                        d.reads_peer_flags[i] = x.peer_base[i] + x.off_reads
                        d.applied_peer_flags[i] = x.peer_base[i] + x.off_applied
        elif self.mode != "ns" and self.world > 1:
            raise RuntimeError("the native step driver needs the peer transport for multi-GPU runs")
        if self.tables == "sharded":
            d.world, d.rank, d.shard_rows = self.world, self.rank, self.shard_rows
            for i in range(self.world):
                d.feat_shards[i] = self._feat_shards.peer_base[i]
                d.hist_shards[i] = (self._hist_shards or self._feat_shards).peer_base[i]
        return d

    def _sharded_only_trains(self, what):
        if self.tables == "sharded":
            raise RuntimeError("%s addresses the tables directly; on sharded tables use run_trains / capture_trains / "
                               "replay_trains" % what)

    def _pass(self):
        self._sharded_only_trains("the one-pass drivers")
        return super()._pass()

    def _eager_step(self, slot, sample):
        self._sharded_only_trains("the pipelined driver")
        return super()._eager_step(slot, sample)

    def run_native(self, *a, **k):
        self._sharded_only_trains("run_native")
        return super().run_native(*a, **k)

    def time_dominant_kernel(self, batches):
        if self.tables == "sharded":
            return None
        return super().time_dominant_kernel(batches)

    def full_history(self):
        """the whole [N, hidden] history table on this rank (sharded tables: assembled over NVLink)"""
        if self._hist_shards is not None:
            return self._hist_shards.gather_full(self.n_nodes)
        return self.history

    def _no_nccl_inside_graphs(self, what):
        """The NCCL transport runs its collective and the merge EAGERLY after each one-pass graph
        (_finish_exchange); a multi-pass driver would never apply a write-back."""
        if self.transport == "nccl" and self.mode != "ns":
            raise RuntimeError("%s needs transport='peer': with transport='nccl' the history write-back exchange runs "
                               "between one-pass replays (run / replay / step_host)" % what)

    def capture_pipelined(self, *a, **k):
        self._no_nccl_inside_graphs("capture_pipelined")
        return super().capture_pipelined(*a, **k)

    def run_pipelined(self, *a, **k):
        self._no_nccl_inside_graphs("run_pipelined")
        return super().run_pipelined(*a, **k)

    def _finish_exchange(self):
        """NCCL transport: the collective and the merge run eagerly after the (captured) pass."""
        if self.transport == "nccl" and self.mode != "ns":
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.recv, self.send)
            check(_lib.load().sgcn_wb_apply(ptr(self.history), self.history.stride(0), self.hidden, ptr(self.recv),
                                            self.slot_bytes, self.world, self.wb_bound, ptr(self.owner),
                                            stream_ptr()))

    def run(self, ids):
        out = super().run(ids)
        self._finish_exchange()
        return out

    def replay(self, ids):
        out = super().replay(ids)
        self._finish_exchange()
        return out

    def step_host(self, ids_pinned):
        if self._pinned_out is None:
            self._pinned_out = torch.empty(self.out.shape, dtype=torch.float32, pin_memory=True)
        self.ids.copy_(ids_pinned, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._pass()
        self._finish_exchange()
        self._pinned_out.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self._pinned_out

    def check_exchange(self):
        """Raise if a peer never arrived (bounded spin in wb_wait_kernel timed out)."""
        if self._exchange is not None:
            flags, epoch, timeout = self._exchange.control()
            if timeout:
                raise _lib.SgcnError(_lib.SGCN_EDATA, "peer exchange timed out waiting for rank %d (epoch %d, flags %s)"
                                     % (timeout - 1, epoch, flags))

    def close(self):
        if self._exchange is not None:
            self._exchange.close()
            self._exchange = None
        if self._hist_shards is not None or self._feat_shards is not None:
            torch.cuda.synchronize(self.dev)
            self.history = self.features = None      # views of the shards that are about to be freed
        for sh in (self._hist_shards, self._feat_shards):
            if sh is not None:
                sh.close()
        self._hist_shards = self._feat_shards = None
