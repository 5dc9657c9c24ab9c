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
        self.start = 0
        self.t = 0

    # -- reference-format (host) path ------------------------------------------------------------
    def batch(self, data):
        data = np.ascontiguousarray(data, dtype=np.int32)
        fields, ffields, adjs, madjs, fadjs, scales = [data], [], [], [], [], []
        sch = self.c_sch
        sch.start_batch(data)
        for l in range(self.L):
            sch.expand(int(self.degrees[self.L - l - 1]), materialize_full=self.cv)
            s = sch.snapshot()
            fields.append(s["field"])
            scales.append(s["scales"])
            ne = s["edg_s"].shape[0]
            edg_i = np.zeros((ne, 2), dtype=np.int32)
            edg_i[:, 0] = s["edg_s"]
            edg_i[:, 1] = s["edg_t"]
            shape = (fields[-2].shape[0], fields[-1].shape[0])
            adjs.append((edg_i, s["edg_w"], shape))
            if self.cv:
                ffields.append(s["ffield"])
                ne2 = s["fedg_s"].shape[0]
                fedg_i = np.zeros((ne2, 2), dtype=np.int32)
                fedg_i[:, 0] = s["fedg_s"]
                fedg_i[:, 1] = s["fedg_t"]
                fshape = (fields[-2].shape[0], ffields[-1].shape[0])
                madjs.append((np.copy(edg_i), s["medg_w"], np.copy(shape)))
                fadjs.append((fedg_i, s["fedg_w"], fshape))
        for lst in (fields, ffields, adjs, madjs, fadjs, scales):
            lst.reverse()
        return self.get_feed_dict(fields, ffields, adjs, madjs, fadjs, scales)

    def minibatch(self, batch_size):
        if self.start == self.data.shape[0]:
            return None
        end = min(self.data.shape[0], self.start + batch_size)
        batch = self.data[self.start:end]
        self.start = end
        return self.batch(batch)

    def get_feed_dict(self, fields, ffields, adjs, madjs, fadjs, scales):
        ph = self.placeholders
        labels = self.labels[fields[-1]]
        feed_dict = {ph['adj'][i]: adjs[i] for i in range(self.L)}
        feed_dict.update({ph['scales'][i]: scales[i] for i in range(len(scales))})
        if self.cv:
            feed_dict.update({ph['madj'][i]: madjs[i] for i in range(len(madjs))})
            feed_dict.update({ph['fadj'][i]: fadjs[i] for i in range(len(fadjs))})
            feed_dict.update({ph['ffields'][i]: ffields[i] for i in range(len(ffields))})
        feed_dict[ph['labels']] = labels
        for i in range(self.L + 1):
            feed_dict[ph['fields'][i]] = fields[i]
        return feed_dict

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
