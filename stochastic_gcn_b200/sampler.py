"""DeviceSampler: Python handle on the HBM-resident neighbour sampler of libsgcn_b200.so.

Replaces the reference's ``Scheduler`` C++ class (gcn/scheduler.h:6-28) as bound by
gcn/_scheduler.pyx:10-19.  The sampled sub-adjacency never leaves the GPU unless ``host()`` is
called; ``view()`` returns zero-copy torch tensors over the sampler-owned device buffers.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

_NAMES = {
    "field": _lib.VEC_FIELD, "ffield": _lib.VEC_FFIELD, "edg_s": _lib.VEC_EDG_S, "edg_t": _lib.VEC_EDG_T,
    "fedg_s": _lib.VEC_FEDG_S, "fedg_t": _lib.VEC_FEDG_T, "adj_i": _lib.VEC_ADJ_I, "adj_p": _lib.VEC_ADJ_P,
    "rowptr_s": _lib.VEC_ROWPTR_S, "rowptr_f": _lib.VEC_ROWPTR_F, "tgt": _lib.VEC_TGT, "meta": _lib.VEC_META,
    "pipe": _lib.VEC_PIPE,
    "scales": _lib.VEC_SCALES, "edg_w": _lib.VEC_EDG_W, "medg_w": _lib.VEC_MEDG_W, "fedg_w": _lib.VEC_FEDG_W,
    "adj_w": _lib.VEC_ADJ_W, "importance": _lib.VEC_IMPORTANCE,
}


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw device pointer."""

    def __init__(self, addr, n, typestr, owner):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(addr), False),
                                         "version": 2, "strides": None}
        self._owner = owner   # keeps the sampler alive while the view exists


class LevelSizes:
    __slots__ = ("n_out", "n_in", "nnz_s", "nnz_f", "n_ff", "status")

    def __init__(self, v):
        self.n_out, self.n_in, self.nnz_s, self.nnz_f, self.n_ff, self.status = [int(x) for x in v]

    def __repr__(self):
        return "LevelSizes(n_out=%d, n_in=%d, nnz_s=%d, nnz_f=%d, n_ff=%d)" % (
            self.n_out, self.n_in, self.nnz_s, self.nnz_f, self.n_ff)


class DeviceSampler:
    """``Scheduler(adj_w, adj_i, adj_p, num_data, num_edges, L, cv, is)`` on the GPU.

    ``adj_w / adj_i / adj_p`` may be NumPy arrays (copied host->device, like the reference's deep
    copy at gcn/scheduler.cpp:14-16) or CUDA torch tensors (copied device->device).
    ``adj_p`` may have N or N+1 entries; the last one is taken to be num_edges.
    """

    def __init__(self, adj_w, adj_i, adj_p, num_data=None, L=1, cv=False, importance=False, device=None):
        lib = _lib.load()
        self._lib = lib
        on_dev = isinstance(adj_w, torch.Tensor) and adj_w.is_cuda
        if on_dev:
            device = adj_w.device.index if device is None else device
            adj_w = adj_w.contiguous().to(torch.float32)
            adj_i = adj_i.contiguous().to(torch.int32)
            adj_p = adj_p.contiguous().to(torch.int32)
            n_edges = int(adj_i.numel())
            n_p = int(adj_p.numel())
        else:
            device = torch.cuda.current_device() if device is None else device
            adj_w = np.ascontiguousarray(adj_w, dtype=np.float32)
            adj_i = np.ascontiguousarray(adj_i, dtype=np.int32)
            adj_p = np.ascontiguousarray(adj_p, dtype=np.int32)
            n_edges = int(adj_i.shape[0])
            n_p = int(adj_p.shape[0])
        if num_data is None:
            num_data = n_p - 1
        if not (n_p == num_data or n_p == num_data + 1):
            raise ValueError("adj_p must have num_data or num_data+1 entries")
        self.num_data, self.num_edges, self.device = int(num_data), n_edges, int(device)
        self.cv, self.importance, self.L = bool(cv), bool(importance), int(L)
        h = C.c_void_p()
        fn = lib.sgcn_sampler_create_device if on_dev else lib.sgcn_sampler_create
        if on_dev:
            torch.cuda.current_stream(self.device).synchronize()
        check(fn(C.byref(h), ptr(adj_w), ptr(adj_i), ptr(adj_p), self.num_data, n_edges, self.L,
                 int(self.cv), int(self.importance), self.device))
        self._h = h
        # run on torch's current stream so that sampler kernels and torch / aggregate kernels are
        # ordered without explicit events (the C default is a private non-blocking stream)
        with torch.cuda.device(self.device):
            self.use_stream(torch.cuda.current_stream())

    # -- Scheduler API ---------------------------------------------------------------------------
    def seed(self, seed):
        check(self._lib.sgcn_sampler_seed(self._h, int(seed)))

    def reserve(self, max_batch, degrees, materialize_full=False):
        d = np.ascontiguousarray(degrees, dtype=np.int32)
        check(self._lib.sgcn_sampler_reserve(self._h, int(max_batch), ptr(d), len(d), int(materialize_full)))

    def start_batch(self, ids):
        if isinstance(ids, torch.Tensor) and ids.is_cuda:
            ids = ids.contiguous().to(torch.int32)
            self._keep = ids
            check(self._lib.sgcn_sampler_start_batch_device(self._h, int(ids.numel()), ptr(ids)))
        else:
            ids = np.ascontiguousarray(ids.cpu().numpy() if isinstance(ids, torch.Tensor) else ids, dtype=np.int32)
            check(self._lib.sgcn_sampler_start_batch(self._h, int(ids.shape[0]), ptr(ids)))

    def expand(self, degree, materialize_full=False):
        check(self._lib.sgcn_sampler_expand(self._h, int(degree), int(materialize_full)))

    def reserve_sets(self, n_sets, batch, degree):
        """Size buffer sets 0 .. n_sets-1 for trains of batches (``expand_train``)."""
        check(self._lib.sgcn_sampler_reserve_sets(self._h, int(n_sets), int(batch), int(degree)))
        self._train_batch = int(batch)

    def expand_train(self, table, first_set=0, prev=None, stream=None):
        """Sample the n batches of ``table`` (CUDA int32 [n, batch]) with ONE launch: batch j lands in buffer
        set (first_set + j) % n_sets exactly as n sequential start_batch + expand calls would leave it.
        ``prev``: ids (CUDA int32, any shape) of earlier batches whose consumer passes may still run."""
        if not (table.is_cuda and table.dtype == torch.int32 and table.is_contiguous() and table.dim() == 2
                and table.shape[1] == self._train_batch):
            raise ValueError("table must be a contiguous CUDA int32 [n, %d] tensor" % self._train_batch)
        self._keep_train = (table, prev)
        check(self._lib.sgcn_sampler_expand_train(self._h, ptr(table), int(table.shape[0]), int(first_set),
                                                  ptr(prev) if prev is not None else None,
                                                  int(prev.numel()) if prev is not None else 0,
                                                  _lib.stream_ptr(stream)))

    def set_slot(self, slot):
        """Select which of the two per-batch buffer sets start_batch / expand / view / sizes address."""
        check(self._lib.sgcn_sampler_set_slot(self._h, int(slot)))

    def pipeline(self, enable=True):
        """Arm the device-side guard that lets batch i+1 be sampled while batch i is still consumed."""
        check(self._lib.sgcn_sampler_pipeline(self._h, int(bool(enable))))

    def mark_consumed(self, stream=None):
        """Tell the guard that one consumer pass has finished reading the adjacency rows."""
        check(self._lib.sgcn_sampler_mark_consumed(self._h, _lib.stream_ptr(stream)))

    def use_stream(self, stream, sync=True):
        """Run expand() on a torch stream (e.g. the current one, for CUDA-graph capture).  sync=False
        skips the synchronisation of the previous stream (needed while a capture is in progress)."""
        fn = self._lib.sgcn_sampler_set_stream if sync else self._lib.sgcn_sampler_set_stream_async
        check(fn(self._h, C.c_void_p(stream.cuda_stream)))
        self._stream = stream

    # -- results ---------------------------------------------------------------------------------
    def sizes(self, level=-1):
        out = (C.c_int32 * 6)()
        check(self._lib.sgcn_sampler_sizes(self._h, int(level), out))
        return LevelSizes(out)

    def view(self, name, level=-1, count=None):
        """Zero-copy torch view of a sampler-owned device vector (valid until the next start_batch)."""
        which = _NAMES[name]
        p, n = C.c_void_p(), C.c_int64()
        check(self._lib.sgcn_sampler_vec(self._h, int(level), which, C.byref(p), C.byref(n)))
        n = int(n.value) if count is None else int(count)
        dtype = torch.float32 if which in _lib.FLOAT_VECS else torch.int32
        if n == 0 or not p.value:
            return torch.empty(0, dtype=dtype, device="cuda:%d" % self.device)
        arr = _DevArray(p.value, n, "<f4" if which in _lib.FLOAT_VECS else "<i4", self)
        return torch.as_tensor(arr, device="cuda:%d" % self.device)

    def host(self, name, count, level=-1):
        """Copy the first `count` elements of a vector to a fresh NumPy array (synchronises)."""
        which = _NAMES[name]
        out = np.empty(int(count), dtype=np.float32 if which in _lib.FLOAT_VECS else np.int32)
        check(self._lib.sgcn_sampler_copy_vec(self._h, int(level), which, ptr(out), int(count)))
        return out

    def snapshot(self, level=-1):
        """All public vectors of one level in the reference's layout (host copies)."""
        z = self.sizes(level)
        s = {"field": self.host("field", z.n_in, level), "edg_s": self.host("edg_s", z.nnz_s, level),
             "edg_t": self.host("edg_t", z.nnz_s, level), "edg_w": self.host("edg_w", z.nnz_s, level),
             "scales": self.host("scales", 0 if self.importance else z.n_out, level)}
        if self.cv and not self.importance:
            s["medg_w"] = self.host("medg_w", z.nnz_s, level)
            full = z.n_ff > 0 or z.nnz_f == 0
            nf = z.nnz_f if full else 0
            s["ffield"] = self.host("ffield", z.n_ff, level)
            s["fedg_s"] = self.host("fedg_s", nf, level)
            s["fedg_t"] = self.host("fedg_t", nf, level)
            s["fedg_w"] = self.host("fedg_w", nf, level)
        else:
            for k, dt in (("medg_w", np.float32), ("ffield", np.int32), ("fedg_s", np.int32),
                          ("fedg_t", np.int32), ("fedg_w", np.float32)):
                s[k] = np.zeros(0, dt)
        return s

    # -- checkpointing (absent in the reference) -------------------------------------------------
    def get_rng(self):
        st = np.empty(624, dtype=np.uint32)
        pos = C.c_int32()
        check(self._lib.sgcn_sampler_get_rng(self._h, ptr(st), C.byref(pos)))
        return st, int(pos.value)

    def set_rng(self, state, pos):
        st = np.ascontiguousarray(state, dtype=np.uint32)
        check(self._lib.sgcn_sampler_set_rng(self._h, ptr(st), int(pos)))

    def get_state(self):
        """Everything a resumed run needs to continue the reference's sampling sequence bit for bit: the engine
        (``mt_state`` / ``mt_pos``) and the stored adjacency rows, which the uniform sampler permutes in place and
        never restores (gcn/scheduler.cpp:144-145; row pointers never change).  Synchronises."""
        st, pos = self.get_rng()
        return {"mt_state": st, "mt_pos": pos, "adj_i": self.host("adj_i", self.num_edges),
                "adj_w": self.host("adj_w", self.num_edges)}

    def set_state(self, state):
        """Restore ``get_state()`` (e.g. from io.load_checkpoint(...)["sampler_state"]) into this sampler, which must
        have been built over the same graph."""
        adj_i = np.ascontiguousarray(state["adj_i"], dtype=np.int32)
        adj_w = np.ascontiguousarray(state["adj_w"], dtype=np.float32)
        if adj_i.shape[0] != self.num_edges or adj_w.shape[0] != self.num_edges:
            raise ValueError("sampler state of a different graph: %d stored entries, this sampler has %d"
                             % (adj_i.shape[0], self.num_edges))
        torch.cuda.synchronize()
        self.view("adj_i", count=self.num_edges).copy_(torch.from_numpy(adj_i))
        self.view("adj_w", count=self.num_edges).copy_(torch.from_numpy(adj_w))
        torch.cuda.synchronize()
        self.set_rng(state["mt_state"], state["mt_pos"])

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.sgcn_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
