"""``PyScheduler`` -- drop-in for the reference's Cython class (gcn/_scheduler.pyx:28-151) over the
device sampler.

Same constructor, ``shuffle`` / ``batch`` / ``minibatch`` / ``get_feed_dict`` / ``get_t`` and the
same feed-dict layout (COO triples ``(int32[ne,2], float32[ne], shape)``, int32 fields, float32
scales, ``labels[fields[-1]]``; index 0 = input-side layer after the reversal at
_scheduler.pyx:121-126), bit-exact with the reference for identical seeds.  ``batch`` copies the
device results to fresh NumPy arrays exactly as the reference memcpy's its vectors
(_scheduler.pyx:21-25,69-111); training code that stays on the GPU uses ``batch_device`` instead,
which returns zero-copy views and never synchronises.
"""
import numpy as np
import torch

from .sampler import DeviceSampler


class DeviceLevel:
    """One expand() result, resident in HBM.  Arrays are capacity-sized views; the exact lengths
    live in ``meta`` (device int32[8]: n_out, n_in, nnz_s, nnz_f, n_ff, status)."""

    __slots__ = ("field", "rowptr_s", "rowptr_f", "edg_s", "edg_t", "tgt", "edg_w", "medg_w", "scales",
                 "meta", "n_out_bound", "n_in_bound", "s_bound")


class DeviceBatch:
    """levels[0] is the input-side layer (same order as the reference's reversed lists)."""

    def __init__(self, levels, batch_ids):
        self.levels = levels
        self.batch_ids = batch_ids


class PyScheduler:
    def __init__(self, adj, labels, L, degrees, placeholders, seed, data=None, cv=False, importance=False,
                 device=None):
        # Scheduler(&ad[0], &ai[0], &ap[0], labels.shape[0], adj.data.shape[0], L, cv, importance)
        self.c_sch = DeviceSampler(adj.data, adj.indices, adj.indptr, num_data=labels.shape[0], L=L, cv=cv,
                                   importance=importance, device=device)
        self.c_sch.seed(seed)
        self.labels = labels
        self.data = data
        self.degrees = degrees
        self.L = L
        self.start = 0
        self.placeholders = placeholders
        self.t = 0
        self.cv = bool(cv)

    def shuffle(self):
        np.random.shuffle(self.data)      # the reference uses the global NumPy RNG (_scheduler.pyx:51)
        self.start, self.t = 0, 0

    # -- reference-format (host) path ------------------------------------------------------------
    # placeholder key -> (argument of get_feed_dict, only with control variates); one entry per hop
    _PER_HOP = (("adj", "adjs", False), ("scales", "scales", False), ("madj", "madjs", True),
                ("fadj", "fadjs", True), ("ffields", "ffields", True))

    @staticmethod
    def _coo(rows, cols, vals, n_rows, n_cols):
        """(int32 [ne, 2] index pairs, float32 values, dense shape): the triple TF's sparse placeholders take"""
        return np.stack((rows, cols), axis=1).astype(np.int32, copy=False), vals, (n_rows, n_cols)

    def _hop_host(self, degree, n_rows):
        """one expand() copied to fresh host arrays (as _scheduler.pyx:69-111 memcpy's the C++ vectors)"""
        self.c_sch.expand(int(degree), materialize_full=self.cv)
        s = self.c_sch.snapshot()
        hop = {"fields": s["field"], "scales": s["scales"],
               "adjs": self._coo(s["edg_s"], s["edg_t"], s["edg_w"], n_rows, s["field"].shape[0])}
        if self.cv:
            idx, _, shape = hop["adjs"]
            hop["ffields"] = s["ffield"]
            hop["madjs"] = (idx.copy(), s["medg_w"], np.array(shape))
            hop["fadjs"] = self._coo(s["fedg_s"], s["fedg_t"], s["fedg_w"], n_rows, s["ffield"].shape[0])
        return hop

    def batch(self, data):
        data = np.ascontiguousarray(data, dtype=np.int32)
        self.c_sch.start_batch(data)
        hops, n_rows = [], data.shape[0]
        for degree in reversed(list(self.degrees[:self.L])):          # output side first, as the sampler walks
            hops.append(self._hop_host(degree, n_rows))
            n_rows = hops[-1]["fields"].shape[0]
        hops.reverse()                                                # index 0 = input-side layer
        lists = {k: [h[k] for h in hops if k in h] for k in ("fields", "ffields", "adjs", "madjs", "fadjs", "scales")}
        lists["fields"].append(data)
        return self.get_feed_dict(**lists)

    def minibatch(self, batch_size):
        n = self.data.shape[0]
        if self.start >= n:
            return None
        ids = self.data[self.start:self.start + batch_size]
        self.start = min(n, self.start + batch_size)
        return self.batch(ids)

    def get_feed_dict(self, fields, ffields, adjs, madjs, fadjs, scales):
        """placeholder -> value, the layout gcn/models.py feeds to sess.run (keys as _scheduler.pyx:136-148)"""
        ph, given = self.placeholders, locals()
        feed = {ph["labels"]: self.labels[fields[-1]]}
        feed.update(zip(ph["fields"], fields[:self.L + 1]))
        for key, arg, cv_only in self._PER_HOP:
            if not cv_only or self.cv:
                feed.update(zip(ph[key], given[arg]))
        return feed

    def get_t(self):
        return self.t

    # -- device-resident path --------------------------------------------------------------------
    def batch_device(self, data):
        """Sample on the GPU and return views of the results; no host round-trip, no sync.

        ``data`` may be a CUDA int32 tensor (stays on device) or host ids (one small H2D copy).
        """
        sch = self.c_sch
        sch.start_batch(data)
        levels = []
        for l in range(self.L):
            sch.expand(int(self.degrees[self.L - l - 1]), materialize_full=False)
            lv = DeviceLevel()
            for name in ("field", "rowptr_s", "rowptr_f", "edg_s", "edg_t", "tgt", "edg_w", "scales", "meta"):
                setattr(lv, name, sch.view(name))
            lv.medg_w = sch.view("medg_w") if self.cv else None
            lv.n_out_bound = lv.rowptr_s.numel() - 1
            lv.n_in_bound = lv.field.numel()
            lv.s_bound = lv.edg_s.numel()
            levels.append(lv)
        levels.reverse()
        ids = data if isinstance(data, torch.Tensor) else None
        return DeviceBatch(levels, ids)
