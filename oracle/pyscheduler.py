"""TEST INFRASTRUCTURE ONLY -- PyScheduler (gcn/_scheduler.pyx:28-151) restated over a CPU sampler.

``OraclePyScheduler(adj, labels, L, degrees, placeholders, seed, data, cv, importance,
backend="oracle"|"ref")`` drives either the C restatement or the compiled reference Scheduler
through the same L x expand loop, list reversal and feed-dict assembly as the Cython class, so
that the product's ``stochastic_gcn_b200.scheduler.PyScheduler`` can be compared key by key.
"""
import numpy as np

from . import native


def default_placeholders(L):
    """String placeholders exactly as gcn/test_scheduler.py:24-32 builds them."""
    return {
        'adj': ['adj_{}'.format(i) for i in range(L)],
        'madj': ['madj_{}'.format(i) for i in range(L)],
        'fadj': ['fadj_{}'.format(i) for i in range(L)],
        'fields': ['fields_{}'.format(i) for i in range(L + 1)],
        'ffields': ['ffields_{}'.format(i) for i in range(L + 1)],
        'scales': ['scales_{}'.format(i) for i in range(L)],
        'labels': 'labels',
    }


class OraclePyScheduler:
    def __init__(self, adj, labels, L, degrees, placeholders, seed, data=None, cv=False,
                 importance=False, backend="oracle"):
        cls = native.OracleSampler if backend == "oracle" else native.RefSampler
        self.c_sch = cls(adj.data, adj.indices, adj.indptr, cv=cv, importance=importance)
        self.c_sch.seed(seed)                                  # _scheduler.pyx:41
        self.labels, self.data, self.degrees = labels, data, degrees
        self.L, self.start, self.placeholders = L, 0, placeholders
        self.cv = bool(cv)

    def shuffle(self):                                         # _scheduler.pyx:50-53
        np.random.shuffle(self.data)
        self.start = 0

    def batch(self, data):                                     # _scheduler.pyx:55-127
        data = np.asarray(data, dtype=np.int32)
        fields, ffields, adjs, madjs, fadjs, scales = [data], [], [], [], [], []
        sch = self.c_sch
        sch.start_batch(data)
        for l in range(self.L):
            sch.expand(int(self.degrees[self.L - l - 1]))
            s = sch.snapshot()
            fields.append(s["field"])
            scales.append(s["scales"])
            shape = (fields[-2].shape[0], fields[-1].shape[0])
            edg_i = np.stack([s["edg_s"], s["edg_t"]], axis=1).astype(np.int32).reshape(-1, 2)
            adjs.append((edg_i, s["edg_w"], shape))
            if self.cv:
                ffields.append(s["ffield"])
                fedg_i = np.stack([s["fedg_s"], s["fedg_t"]], axis=1).astype(np.int32).reshape(-1, 2)
                fshape = (fields[-2].shape[0], s["ffield"].shape[0])
                madjs.append((edg_i.copy(), s["medg_w"], np.array(shape)))
                fadjs.append((fedg_i, s["fedg_w"], fshape))
        for lst in (fields, ffields, adjs, madjs, fadjs, scales):
            lst.reverse()
        return self.get_feed_dict(fields, ffields, adjs, madjs, fadjs, scales)

    def minibatch(self, batch_size):                           # _scheduler.pyx:129-135
        if self.start == self.data.shape[0]:
            return None
        end = min(self.data.shape[0], self.start + batch_size)
        batch = self.data[self.start:end]
        self.start = end
        return self.batch(batch)

    def get_feed_dict(self, fields, ffields, adjs, madjs, fadjs, scales):   # _scheduler.pyx:137-148
        ph = self.placeholders
        fd = {ph['adj'][i]: adjs[i] for i in range(self.L)}
        fd.update({ph['scales'][i]: scales[i] for i in range(len(scales))})
        if self.cv:
            fd.update({ph['madj'][i]: madjs[i] for i in range(len(madjs))})
            fd.update({ph['fadj'][i]: fadjs[i] for i in range(len(fadjs))})
            fd.update({ph['ffields'][i]: ffields[i] for i in range(len(ffields))})
        fd[ph['labels']] = self.labels[fields[-1]]
        for i in range(self.L + 1):
            fd[ph['fields'][i]] = fields[i]
        return fd
