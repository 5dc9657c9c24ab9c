"""Small deterministic graphs shared by the tests and the golden-vector generator."""
import numpy as np
from scipy.sparse import csr_matrix, diags


def tree11():
    """The 11-node tree of the reference's gcn/test_scheduler.py:10-20, row-normalised, CSR float32."""
    edges = np.array([(0, 1, 1), (0, 2, 1), (0, 3, 1), (1, 4, 1), (1, 5, 1), (1, 6, 1),
                      (2, 7, 1), (2, 8, 1), (2, 9, 1), (3, 10, 1)])
    adj = csr_matrix((edges[:, 2], (edges[:, 0], edges[:, 1])), shape=(11, 11), dtype=np.float32)
    adj = adj + adj.transpose()
    deg = np.array(adj.sum(axis=0)).flatten()
    adj = diags(1.0 / deg, 0).dot(adj)
    adj = csr_matrix(adj, dtype=np.float32)
    adj.indices = adj.indices.astype(np.int32)
    adj.indptr = adj.indptr.astype(np.int32)
    return adj


def random_graph(n, avg_deg, seed, normalise="row"):
    """Directed random graph with Pareto-distributed out-degrees (some rows empty), no duplicate
    entries inside a row, float32 weights: row-normalised (graphsage, gcn/utils.py:299-309) or
    symmetric D^-1/2 (A) D^-1/2-like random positive weights."""
    rng = np.random.RandomState(seed)
    deg = np.minimum((rng.pareto(1.5, size=n) * avg_deg * 0.5).astype(np.int64), n - 1)
    deg[rng.rand(n) < 0.05] = 0
    indptr = np.zeros(n + 1, dtype=np.int32)
    indptr[1:] = np.cumsum(deg)
    indices = np.zeros(indptr[-1], dtype=np.int32)
    for i in range(n):
        indices[indptr[i]:indptr[i + 1]] = np.sort(rng.choice(n, size=deg[i], replace=False))
    if normalise == "row":
        data = np.repeat(1.0 / np.maximum(deg, 1), deg).astype(np.float32)
    else:
        data = (rng.rand(indptr[-1]) + 0.1).astype(np.float32)
    return csr_matrix((data, indices, indptr), shape=(n, n), dtype=np.float32)
