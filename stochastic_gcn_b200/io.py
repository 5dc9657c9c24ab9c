"""On-disk formats.

* The reference's pre-processed dataset caches -- ``data/<name>_<normalization>.npz`` written by
  ``load_gcn_data`` (gcn/utils.py:172-181, read back at 38-49) with SPARSE feature matrices, and
  ``<prefix>[_deg<k>].npz`` written by ``load_graphsage_data`` (utils.py:325-333, read back at 201-213)
  with DENSE ones -- same key names, dtypes and the same 10-tuple ``load_data`` returns, so a cache the
  reference produced loads here and vice versa.  Pure NumPy / SciPy: no device involved.
* Checkpoints.  The reference saves ``vars + history_vars`` with ``tf.train.Saver`` into
  ``tmp/<name>.ckpt`` (gcn/models.py:204-220); that container is TensorFlow's own and is not readable
  without TensorFlow.  ``save_checkpoint`` keeps the same CONTENT (trainable variables in layer order,
  history tables) in one ``.npz`` and adds what the reference loses on restart: Adam slots + step, the
  sampler's ``mt19937`` state and its in-place row permutation (the reference's ``Scheduler`` cannot
  be resumed bit-exactly; this one can).
"""
import numpy as np
import scipy.sparse as sp

_CSR_PARTS = ("data", "indices", "indptr", "shape")


def _csr_keys(name):
    return tuple("%s_%s" % (name, p) for p in _CSR_PARTS)


def _put_csr(out, name, m):
    m = sp.csr_matrix(m)
    for key, val in zip(_csr_keys(name), (m.data, m.indices, m.indptr, m.shape)):
        out[key] = val


def _get_csr(data, name):
    d, i, p, shape = (data[k] for k in _csr_keys(name))
    return sp.csr_matrix((d, i, p), shape=tuple(int(x) for x in shape))


def save_cache(path, num_data, train_adj, full_adj, feats, train_feats, test_feats, labels, train_data,
               val_data, test_data):
    """Write a dataset cache with the reference's key set.  Sparse ``feats`` (SciPy) give the
    ``load_gcn_data`` schema (utils.py:172-181), dense ones the ``load_graphsage_data`` schema (325-333)."""
    out = {"num_data": num_data, "labels": labels, "train_data": train_data, "val_data": val_data,
           "test_data": test_data}
    _put_csr(out, "train_adj", train_adj)
    _put_csr(out, "full_adj", full_adj)
    if sp.issparse(feats):
        for name, m in (("feats", feats), ("train_feats", train_feats), ("test_feats", test_feats)):
            _put_csr(out, name, m)
    else:
        out.update(feats=np.asarray(feats), train_feats=np.asarray(train_feats), test_feats=np.asarray(test_feats))
    with open(path, "wb") as f:
        np.savez(f, **out)


def load_cache(path):
    """Read either cache schema; returns the reference's tuple
    ``(num_data, train_adj, full_adj, feats, train_feats, test_feats, labels, train_data, val_data, test_data)``
    (gcn/utils.py:183,335)."""
    data = np.load(path)
    if "feats" in data.files:                                  # load_graphsage_data schema
        feats, train_feats, test_feats = data["feats"], data["train_feats"], data["test_feats"]
    else:                                                      # load_gcn_data schema
        feats, train_feats, test_feats = (_get_csr(data, n) for n in ("feats", "train_feats", "test_feats"))
    return (data["num_data"], _get_csr(data, "train_adj"), _get_csr(data, "full_adj"), feats, train_feats,
            test_feats, data["labels"], data["train_data"], data["val_data"], data["test_data"])


# ---- checkpoints ------------------------------------------------------------------------------------
def save_checkpoint(path, variables, history, optimizer=None, sampler_state=None, host_state=None):
    """variables: list of arrays (trainable variables in layer order); history: list of [N, D] tables;
    optimizer: optional dict {"t": int, "m": [...], "v": [...]}; sampler_state: optional dict
    {"mt_state": uint32[624], "mt_pos": int, "adj_i": int32[E], "adj_w": float32[E]} -- the engine and
    the permuted adjacency rows (row pointers never change; DeviceSampler.get_state()); host_state: optional dict
    {"dropout_seed": int, "dropout_offset": int, "numpy_rng": np.random.get_state(), "epoch_data": int32[n],
    "epoch_start": int} -- the Philox counter of nn.DropoutState, the global NumPy RNG PyScheduler.shuffle draws
    from, and the shuffled id list with its cursor.  With all of them a resumed run repeats an uninterrupted one
    bit for bit (tests/test_resume_gpu.py)."""
    out = {"n_vars": len(variables), "n_history": len(history)}
    for k, v in enumerate(variables):
        out["var_%d" % k] = np.asarray(v, dtype=np.float32)
    for k, h in enumerate(history):
        out["history_%d" % k] = np.asarray(h, dtype=np.float32)
    if optimizer is not None:
        out["adam_t"] = int(optimizer["t"])
        for k, (m, v) in enumerate(zip(optimizer["m"], optimizer["v"])):
            out["adam_m_%d" % k] = np.asarray(m, dtype=np.float32)
            out["adam_v_%d" % k] = np.asarray(v, dtype=np.float32)
    if sampler_state is not None:
        out["mt_state"] = np.asarray(sampler_state["mt_state"], dtype=np.uint32)
        out["mt_pos"] = int(sampler_state["mt_pos"])
        out["adj_i"] = np.asarray(sampler_state["adj_i"], dtype=np.int32)
        out["adj_w"] = np.asarray(sampler_state["adj_w"], dtype=np.float32)
    if host_state is not None:
        out["dropout_seed"], out["dropout_offset"] = int(host_state["dropout_seed"]), int(host_state["dropout_offset"])
        if host_state.get("numpy_rng") is not None:
            kind, keys, pos, has_gauss, cached = host_state["numpy_rng"]
            out["np_rng_keys"], out["np_rng_pos"] = np.asarray(keys, dtype=np.uint32), int(pos)
            out["np_rng_gauss"] = np.asarray([has_gauss, cached], dtype=np.float64)
        if host_state.get("epoch_data") is not None:
            out["epoch_data"] = np.asarray(host_state["epoch_data"], dtype=np.int32)
            out["epoch_start"] = int(host_state.get("epoch_start", 0))
    with open(path, "wb") as f:
        np.savez(f, **out)


def load_checkpoint(path, load_history=True):
    """-> dict(variables, history, optimizer | None, sampler_state | None, host_state | None).  load_history=False mirrors
    ``Model.load(sess, load_history=False)`` (gcn/models.py:211-220): the tables are left out."""
    data = np.load(path)
    res = {"variables": [data["var_%d" % k] for k in range(int(data["n_vars"]))],
           "history": [data["history_%d" % k] for k in range(int(data["n_history"]))] if load_history else [],
           "optimizer": None, "sampler_state": None, "host_state": None}
    if "adam_t" in data.files:
        n = len(res["variables"])
        res["optimizer"] = {"t": int(data["adam_t"]), "m": [data["adam_m_%d" % k] for k in range(n)],
                            "v": [data["adam_v_%d" % k] for k in range(n)]}
    if "mt_state" in data.files:
        res["sampler_state"] = {"mt_state": data["mt_state"], "mt_pos": int(data["mt_pos"]), "adj_i": data["adj_i"],
                                "adj_w": data["adj_w"]}
    if "dropout_seed" in data.files:
        hs = {"dropout_seed": int(data["dropout_seed"]), "dropout_offset": int(data["dropout_offset"]),
              "numpy_rng": None, "epoch_data": None, "epoch_start": 0}
        if "np_rng_keys" in data.files:
            g = data["np_rng_gauss"]
            hs["numpy_rng"] = ("MT19937", data["np_rng_keys"], int(data["np_rng_pos"]), int(g[0]), float(g[1]))
        if "epoch_data" in data.files:
            hs["epoch_data"], hs["epoch_start"] = data["epoch_data"], int(data["epoch_start"])
        res["host_state"] = hs
    return res
