"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference aggregators.

The reference evaluates these with TensorFlow-1.x ops that are not in /root/reference
(tf.sparse_tensor_dense_matmul, tf.gather, tf.scatter_update, tf.concat); TensorFlow is not
installed here, and the reference holds no test or golden vector at this boundary, so this part
of the oracle is **parity unpinned**: it follows the reference's formulas line by line
(citations below) but cannot be checked against reference outputs.

Each function takes the feed-dict pieces exactly as ``PyScheduler.batch`` produces them
(COO triples ``(idx[ne,2], val[ne], shape)``, int32 fields, float32 scales) and a ``dtype``:
``np.float64`` gives the high-precision reference value the 1e-4-relative parity bound is
measured against; ``np.float32`` mirrors the TF CPU kernel's fp32 storage-order accumulation
(via oracle/sgcn_oracle.c) and is what the CPU baseline times.
"""
import numpy as np

from . import native


_SCIPY_ABOVE = 50_000


def _coo_matmul(adj, x, dtype):
    """dot(adj, x, sparse=True), gcn/layers.py:31-37."""
    idx, val, shape = adj
    idx = np.asarray(idx).reshape(-1, 2)
    if dtype == np.float32:
        return native.spmm_coo(idx, val, shape, x)
    if len(val) > _SCIPY_ABOVE:       # full-size cases: same float64 sums through SciPy's CSR product
        from scipy.sparse import coo_matrix
        a = coo_matrix((np.asarray(val, np.float64), (idx[:, 0], idx[:, 1])), shape=(int(shape[0]), int(shape[1])))
        return np.asarray(a.tocsr() @ x.astype(np.float64))
    y = np.zeros((int(shape[0]), x.shape[1]), dtype=np.float64)
    np.add.at(y, idx[:, 0], np.asarray(val, np.float64)[:, None] * x.astype(np.float64)[idx[:, 1]])
    return y


def _coo_matmul_t(adj, dy, dtype):
    """gradient of dot(adj, x) w.r.t. x (TF autodiff, gcn/models.py:187): adj^T dy."""
    idx, val, shape = adj
    idx = np.asarray(idx).reshape(-1, 2)
    if dtype == np.float32:
        return native.spmm_coo_t(idx, val, shape, dy)
    dx = np.zeros((int(shape[1]), dy.shape[1]), dtype=np.float64)
    np.add.at(dx, idx[:, 1], np.asarray(val, np.float64)[:, None] * dy.astype(np.float64)[idx[:, 0]])
    return dx


def plain_forward(adj, inputs, graphsage, dtype=np.float64):
    """PlainAggregator._call, non-tuple branch, gcn/layers.py:249-257."""
    n_out = int(adj[2][0])
    a_self = inputs[:n_out].astype(dtype)
    a_nb = _coo_matmul(adj, inputs, dtype)
    return np.concatenate((a_self, a_nb), axis=1) if graphsage else a_nb


def plain_backward(adj, d_out, n_in, graphsage, dtype=np.float64):
    """d(inputs) for plain_forward: adj^T d(a_neighbour) (+ d(a_self) on the first n_out rows)."""
    n_out = int(adj[2][0])
    d_out = d_out.astype(dtype)
    if graphsage:
        dim = d_out.shape[1] // 2
        d_self, d_nb = d_out[:, :dim], d_out[:, dim:]
    else:
        d_self, d_nb = None, d_out
    dx = _coo_matmul_t((adj[0], adj[1], (n_out, n_in)), d_nb, dtype)
    if d_self is not None:
        dx[:n_out] += d_self
    return dx


def cv_forward(adj, fadj, ifield, ffield, history, inputs, graphsage, dtype=np.float64):
    """VRAggregator._call, CV branch, gcn/layers.py:350-362.

    a_neighbour = adj@inputs - adj@history[ifield] + fadj@history[ffield]; new_history=[inputs].
    """
    n_out = int(adj[2][0])
    a_self = inputs[:n_out].astype(dtype)
    cur = _coo_matmul(adj, inputs, dtype)
    old = _coo_matmul(adj, history[ifield], dtype)
    mean = _coo_matmul(fadj, history[ffield], dtype)
    a_nb = cur - old + mean
    out = np.concatenate((a_self, a_nb), axis=1) if graphsage else a_nb
    return out, [inputs]


def cvd_forward(adj, fadj, ifield, ffield, history, scale, h, mu, graphsage, dtype=np.float64):
    """VRAggregator._call, CVD branch, gcn/layers.py:298-319.

    mu_nb = adj@(mu - history[ifield]) + fadj@history[ffield]
    h_nb  = (adj@(h - mu)) * scale[:,None] + mu_nb ; new_history=[mu]
    """
    n_out = int(adj[2][0])
    h_self, mu_self = h[:n_out].astype(dtype), mu[:n_out].astype(dtype)
    mu_small = history[ifield].astype(dtype)
    mu_large = history[ffield].astype(dtype)
    z = h.astype(dtype) - mu.astype(dtype)
    delta_mu = mu.astype(dtype) - mu_small
    mu_mean = _coo_matmul(fadj, mu_large, dtype)
    mu_nb = _coo_matmul(adj, delta_mu, dtype) + mu_mean
    h_nb = _coo_matmul(adj, z, dtype) * np.asarray(scale, dtype)[:, None] + mu_nb
    if graphsage:
        return (np.concatenate((h_self, h_nb), axis=1), np.concatenate((mu_self, mu_nb), axis=1)), [mu]
    return (h_nb, mu_nb), [mu]


def cvd_backward_h(adj, scale, d_h_out, n_in, graphsage, dtype=np.float64):
    """d(h) for cvd_forward.  mu is stop_gradient-ed where it is produced (gcn/layers.py:412) and
    history is non-trainable, so h is the only differentiable input: d(h) = adj^T (d(h_nb)*scale)
    (+ d(h_self) on the first n_out rows)."""
    n_out = int(adj[2][0])
    d = d_h_out.astype(dtype)
    if graphsage:
        dim = d.shape[1] // 2
        d_self, d_nb = d[:, :dim], d[:, dim:]
    else:
        d_self, d_nb = None, d
    dx = _coo_matmul_t((adj[0], adj[1], (n_out, n_in)), d_nb * np.asarray(scale, dtype)[:, None], dtype)
    if d_self is not None:
        dx[:n_out] += d_self
    return dx


def _sq(adj):
    """tf.square(SparseTensor): element-wise square of the stored values (gcn/layers.py:242,337,338)."""
    return (adj[0], np.asarray(adj[1], np.float64) ** 2, adj[2])


def plain_forward_det(adj, mu, var, graphsage, dtype=np.float64):
    """PlainAggregator._call, (mu, var) branch, gcn/layers.py:238-247."""
    n_out = int(adj[2][0])
    mu_nb = _coo_matmul(adj, mu, dtype)
    var_nb = _coo_matmul(_sq(adj), var, np.float64).astype(dtype)
    if graphsage:
        return (np.concatenate((mu[:n_out].astype(dtype), mu_nb), axis=1),
                np.concatenate((var[:n_out].astype(dtype), var_nb), axis=1))
    return mu_nb, var_nb


def det_forward(adj, fadj, madj, ifield, ffield, mu_history, var_history, mu, var, graphsage):
    """VRAggregator._call, det-dropout branch (inputs a tuple, cvd False), gcn/layers.py:320-349, float64.

    Returns ((mu_out, var_out), new_history=(mu, var), pre) where pre is the value under the relu."""
    n_out = int(adj[2][0])
    f8 = np.float64
    mu, var = mu.astype(f8), var.astype(f8)
    delta_mu = mu - mu_history[ifield].astype(f8)
    mu_bar = mu_history[ffield].astype(f8)
    sigma = np.sqrt(var)
    sigma_bar = np.sqrt(var_history[ifield].astype(f8))
    delta_sigma = sigma - sigma_bar
    var_bar = var_history[ffield].astype(f8)
    msigma = delta_sigma * sigma_bar
    mu_nb = _coo_matmul(adj, delta_mu, f8) + _coo_matmul(fadj, mu_bar, f8)
    pre = (_coo_matmul(_sq(adj), delta_sigma ** 2, f8) + _coo_matmul(_sq(fadj), var_bar, f8)
           + 2 * _coo_matmul(madj, msigma, f8))
    var_nb = np.maximum(pre, 0.0) + 1e-10
    if graphsage:
        out = (np.concatenate((mu[:n_out], mu_nb), axis=1), np.concatenate((var[:n_out], var_nb), axis=1))
    else:
        out = (mu_nb, var_nb)
    return out, (mu, var), pre


def det_backward_var(adj, madj, ifield, var_history, var, pre, d_var_out, graphsage):
    """d(var) for det_forward under TF autodiff (history is not trainable): with G = d(var_nb) * [pre > 0],
    d(delta_sigma) = 2 delta_sigma * (adj^2)^T G + 2 sigma_bar * madj^T G,  d(var) = d(delta_sigma) * 0.5 / sigma."""
    n_out, n_in = int(adj[2][0]), var.shape[0]
    f8 = np.float64
    d = d_var_out.astype(f8)
    if graphsage:
        dim = d.shape[1] // 2
        d_self, d_nb = d[:, :dim], d[:, dim:]
    else:
        d_self, d_nb = None, d
    G = d_nb * (pre > 0)
    sigma = np.sqrt(var.astype(f8))
    sigma_bar = np.sqrt(var_history[ifield].astype(f8))
    shape = (n_out, n_in)
    d_ds = 2 * (sigma - sigma_bar) * _coo_matmul_t((adj[0], _sq(adj)[1], shape), G, f8) \
        + 2 * sigma_bar * _coo_matmul_t((madj[0], madj[1], shape), G, f8)
    dv = d_ds * 0.5 / sigma
    if d_self is not None:
        dv[:n_out] += d_self
    return dv


def history_update(history, ifield, new_rows):
    """tf.scatter_update(history, fields[l], new_history), gcn/models.py:160-166 (in place)."""
    history[np.asarray(ifield)] = new_rows
    return history


def preprocess_features(adj_csr, feats, graphsage, dtype=np.float64):
    """Model input of the PP models: the reference computes ``train_adj.dot(feats)`` with SciPy
    (gcn/utils.py:168-169,321-322) and stacks ``[feats | adj.dot(feats)]`` (graphsage) or uses
    ``adj.dot(feats)`` alone (gcn) -- gcn/models.py:230-239.  adj_csr: scipy.sparse.csr_matrix."""
    nbr = adj_csr.astype(dtype).dot(np.asarray(feats, dtype=dtype))
    if graphsage:
        return np.hstack((np.asarray(feats, dtype=dtype), nbr))
    return nbr
