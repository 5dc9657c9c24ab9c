"""CPU: dataset-cache and checkpoint formats (stochastic_gcn_b200/io.py).  The cache written here must
be readable by the reference's own read-back lines (gcn/utils.py:38-49 and 201-213, restated verbatim
below: same keys, same constructors), and a cache laid out by the reference's write lines must load here."""
import numpy as np
import scipy.sparse as sp

from stochastic_gcn_b200 import io
from tests.graphs_small import random_graph


def _dataset(sparse_feats):
    n, f = 60, 9
    train_adj, full_adj = random_graph(n, 4, 1), random_graph(n, 6, 2)
    rng = np.random.RandomState(0)
    feats = rng.randn(n, f).astype(np.float32)
    if sparse_feats:
        feats = sp.csr_matrix(feats * (rng.rand(n, f) < 0.3))
        train_feats, test_feats = train_adj.dot(feats), full_adj.dot(feats)
    else:
        train_feats, test_feats = train_adj.dot(feats), full_adj.dot(feats)
    labels = np.eye(3, dtype=np.float32)[rng.randint(0, 3, n)]
    ids = rng.permutation(n).astype(np.int32)
    return n, train_adj, full_adj, feats, train_feats, test_feats, labels, ids[:30], ids[30:40], ids[40:]


def _reference_read_gcn(npz_file):
    """gcn/utils.py:38-49"""
    data = np.load(npz_file)
    train_adj = sp.csr_matrix((data['train_adj_data'], data['train_adj_indices'], data['train_adj_indptr']), shape=data['train_adj_shape'])
    full_adj = sp.csr_matrix((data['full_adj_data'], data['full_adj_indices'], data['full_adj_indptr']), shape=data['full_adj_shape'])
    feats = sp.csr_matrix((data['feats_data'], data['feats_indices'], data['feats_indptr']), shape=data['feats_shape'])
    train_feats = sp.csr_matrix((data['train_feats_data'], data['train_feats_indices'], data['train_feats_indptr']), shape=data['train_feats_shape'])
    test_feats = sp.csr_matrix((data['test_feats_data'], data['test_feats_indices'], data['test_feats_indptr']), shape=data['test_feats_shape'])
    return (data['num_data'], train_adj, full_adj, feats, train_feats, test_feats, data['labels'], data['train_data'],
            data['val_data'], data['test_data'])


def _reference_read_graphsage(npz_file):
    """gcn/utils.py:201-213"""
    data = np.load(npz_file)
    train_adj = sp.csr_matrix((data['train_adj_data'], data['train_adj_indices'], data['train_adj_indptr']), shape=data['train_adj_shape'])
    full_adj = sp.csr_matrix((data['full_adj_data'], data['full_adj_indices'], data['full_adj_indptr']), shape=data['full_adj_shape'])
    return (data['num_data'], train_adj, full_adj, data['feats'], data['train_feats'], data['test_feats'], data['labels'],
            data['train_data'], data['val_data'], data['test_data'])


def _same(a, b):
    if sp.issparse(a) or sp.issparse(b):
        a, b = sp.csr_matrix(a), sp.csr_matrix(b)
        return a.shape == b.shape and a.dtype == b.dtype and (a != b).nnz == 0
    return np.array_equal(np.asarray(a), np.asarray(b)) and np.asarray(a).dtype == np.asarray(b).dtype


def test_cache_round_trip_both_schemas(tmp_path):
    for sparse_feats, ref_read in ((True, _reference_read_gcn), (False, _reference_read_graphsage)):
        ds = _dataset(sparse_feats)
        path = str(tmp_path / ("cache_%d.npz" % sparse_feats))
        io.save_cache(path, *ds)
        for got in (io.load_cache(path), ref_read(path)):          # our reader and the reference's read-back
            assert len(got) == 10
            for a, b in zip(got, ds):
                assert _same(a, b)


def test_cache_written_with_the_reference_key_set_loads(tmp_path):
    n, train_adj, full_adj, feats, train_feats, test_feats, labels, tr, va, te = _dataset(False)
    path = str(tmp_path / "ref_written.npz")
    with open(path, "wb") as fwrite:                                   # gcn/utils.py:325-333, verbatim key set
        np.savez(fwrite, num_data=n,
                 train_adj_data=train_adj.data, train_adj_indices=train_adj.indices, train_adj_indptr=train_adj.indptr, train_adj_shape=train_adj.shape,
                 full_adj_data=full_adj.data, full_adj_indices=full_adj.indices, full_adj_indptr=full_adj.indptr, full_adj_shape=full_adj.shape,
                 feats=feats, train_feats=train_feats, test_feats=test_feats,
                 labels=labels, train_data=tr, val_data=va, test_data=te)
    got = io.load_cache(path)
    assert int(got[0]) == n and _same(got[1], train_adj) and _same(got[2], full_adj) and _same(got[4], train_feats)


def test_checkpoint_round_trip(tmp_path):
    rng = np.random.RandomState(3)
    variables = [rng.randn(12, 5).astype(np.float32), rng.randn(5).astype(np.float32)]
    history = [rng.randn(40, 5).astype(np.float32)]
    opt = {"t": 7, "m": [v * 0.1 for v in variables], "v": [v * v for v in variables]}
    samp = {"mt_state": rng.randint(0, 2**32, 624, dtype=np.uint64).astype(np.uint32), "mt_pos": 311,
            "adj_i": rng.randint(0, 40, 90).astype(np.int32), "adj_w": rng.rand(90).astype(np.float32)}
    path = str(tmp_path / "model.npz")
    io.save_checkpoint(path, variables, history, opt, samp)
    ck = io.load_checkpoint(path)
    assert all(np.array_equal(a, b) for a, b in zip(ck["variables"], variables))
    assert np.array_equal(ck["history"][0], history[0])
    assert ck["optimizer"]["t"] == 7 and np.array_equal(ck["optimizer"]["v"][1], opt["v"][1])
    assert np.array_equal(ck["sampler_state"]["mt_state"], samp["mt_state"]) and ck["sampler_state"]["mt_pos"] == 311
    assert io.load_checkpoint(path, load_history=False)["history"] == []            # models.py:211-220
    io.save_checkpoint(path, variables, history)
    ck = io.load_checkpoint(path)
    assert ck["optimizer"] is None and ck["sampler_state"] is None


def test_checkpoint_keeps_the_host_side_random_state(tmp_path):
    """dropout Philox counter, the global NumPy RNG PyScheduler.shuffle draws from, the shuffled epoch and its cursor"""
    rng = np.random.RandomState(5)
    variables, history = [rng.randn(3, 2).astype(np.float32)], []
    np.random.seed(123)
    np.random.rand(7)
    data = np.arange(50, dtype=np.int32)
    np.random.shuffle(data)
    host = {"dropout_seed": 9, "dropout_offset": 123456789012, "numpy_rng": np.random.get_state(),
            "epoch_data": data, "epoch_start": 32}
    want_next = np.random.rand(5)
    path = str(tmp_path / "model.npz")
    io.save_checkpoint(path, variables, history, host_state=host)
    ck = io.load_checkpoint(path)["host_state"]
    assert ck["dropout_seed"] == 9 and ck["dropout_offset"] == 123456789012 and ck["epoch_start"] == 32
    assert np.array_equal(ck["epoch_data"], data)
    np.random.seed(0)
    np.random.set_state(ck["numpy_rng"])
    assert np.array_equal(np.random.rand(5), want_next)
    io.save_checkpoint(path, variables, history)
    assert io.load_checkpoint(path)["host_state"] is None
