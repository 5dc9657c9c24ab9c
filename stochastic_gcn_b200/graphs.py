"""Synthetic graphs of the shapes BASELINE.json names (no datasets / network in this environment).

Generated with torch so that the same seeded graph can be produced on the GPU (bench, GPU tests)
or on the CPU.  Normalisation follows the reference's loaders: row-normalised ``D^-1 A`` without
self loops for the GraphSAGE-format datasets (Reddit / PPI, gcn/utils.py:299-309) and
``D^-1/2 (A+I) D^-1/2`` for the GCN-format ones (Cora / PubMed, gcn/utils.py:127-136).
"""
import torch


class CSRGraph:
    """CSR adjacency as torch tensors: data float32 [E], indices int32 [E], indptr int32 [N+1]."""

    def __init__(self, data, indices, indptr, n):
        self.data, self.indices, self.indptr, self.n = data, indices, indptr, int(n)

    @property
    def nnz(self):
        return int(self.indices.numel())

    @property
    def shape(self):
        return (self.n, self.n)

    def row_ids(self):
        deg = (self.indptr[1:] - self.indptr[:-1]).long()
        return torch.repeat_interleave(torch.arange(self.n, device=self.indptr.device), deg)

    def degrees(self):
        return (self.indptr[1:] - self.indptr[:-1]).long()

    def to_scipy(self):
        from scipy.sparse import csr_matrix
        return csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()),
                          shape=(self.n, self.n))

    def to(self, device):
        return CSRGraph(self.data.to(device), self.indices.to(device), self.indptr.to(device), self.n)


def _csr_from_sorted_keys(keys, n, device):
    rows = torch.div(keys, n, rounding_mode="floor")
    cols = (keys - rows * n).to(torch.int32)
    deg = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(deg, 0)
    return rows, cols, deg, indptr.to(torch.int32)


def powerlaw_graph(n, nnz_target, seed=0, device="cuda", exponent=2.1, max_degree=None, normalization="graphsage",
                   chunk=1 << 24, exact=False):
    """Undirected Chung-Lu power-law graph, symmetrised, duplicate- and self-loop-free.

    Endpoint i is drawn with probability proportional to w_i ~ Pareto(exponent-1) (capped so the
    expected maximum degree stays near ``max_degree``); ``nnz_target`` is the number of stored
    entries (2 x undirected edges) aimed for before duplicate removal.  ``exact``: top up with further
    draws of the same distribution until at least ``nnz_target`` DISTINCT entries are stored, then drop
    random undirected edges down to exactly ``nnz_target`` (rounded down to an even number) -- the
    named shapes are quoted by their stored size (E = nnz AFTER de-duplication).
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    u = torch.rand(n, generator=gen, device=dev, dtype=torch.float64).clamp_(min=1e-12)
    w = u.pow(-1.0 / (exponent - 1.0))
    if max_degree is not None:
        cap = float(max_degree) * float(w.sum()) / float(nnz_target)
        w = w.clamp_(max=cap)
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1]
    m = nnz_target // 2
    parts = []
    for start in range(0, m, chunk):
        k = min(chunk, m - start)
        a = torch.searchsorted(cdf, torch.rand(k, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=n - 1)
        b = torch.searchsorted(cdf, torch.rand(k, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=n - 1)
        keep = a != b
        a, b = a[keep], b[keep]
        parts.append(torch.cat((a * n + b, b * n + a)))
    keys = torch.unique(torch.cat(parts)) if parts else torch.zeros(0, dtype=torch.int64, device=dev)
    del parts
    target = (nnz_target // 2) * 2
    if exact and n * (n - 1) >= 2 * target:
        for _ in range(16):
            short = target - keys.numel()
            if short <= 0:
                break
            k = short // 2 + short // 6 + 1024              # duplicates again: draw ~1.3x what is missing
            a = torch.searchsorted(cdf, torch.rand(k, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=n - 1)
            b = torch.searchsorted(cdf, torch.rand(k, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=n - 1)
            keep = a != b
            a, b = a[keep], b[keep]
            keys = torch.unique(torch.cat((keys, a * n + b, b * n + a)))
        surplus = (keys.numel() - target) // 2
        if surplus > 0:                                       # drop whole undirected edges, both directions
            r = torch.div(keys, n, rounding_mode="floor")
            upper = keys[r < keys - r * n]
            pick = torch.randperm(upper.numel(), generator=gen, device=dev)[:surplus]
            drop = upper[pick]
            dr = torch.div(drop, n, rounding_mode="floor")
            mirror = (drop - dr * n) * n + dr
            keep = torch.ones(keys.numel(), dtype=torch.bool, device=dev)
            keep[torch.searchsorted(keys, torch.cat((drop, mirror)))] = False
            keys = keys[keep]
            del r, upper, pick, drop, dr, mirror, keep
    rows, cols, deg, indptr = _csr_from_sorted_keys(keys, n, dev)
    if normalization == "graphsage":
        data = (1.0 / deg.clamp(min=1).to(torch.float32))[rows]
    else:
        raise ValueError("use gcn_normalized_graph for the D^-1/2 (A+I) D^-1/2 form")
    return CSRGraph(data.contiguous(), cols.contiguous(), indptr, n)


def gcn_normalized_graph(n, n_undirected_edges, seed=0, device="cuda"):
    """Uniform random undirected graph with self loops, D^-1/2 (A+I) D^-1/2 (gcn/utils.py:127-136):
    the Cora / PubMed shape."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    a = torch.randint(0, n, (n_undirected_edges,), generator=gen, device=dev)
    b = torch.randint(0, n, (n_undirected_edges,), generator=gen, device=dev)
    loops = torch.arange(n, device=dev)
    keys = torch.unique(torch.cat((a * n + b, b * n + a, loops * n + loops)))
    rows, cols, deg, indptr = _csr_from_sorted_keys(keys, n, dev)
    dinv = deg.to(torch.float32).pow(-0.5)
    data = dinv[rows] * dinv[cols.long()]
    return CSRGraph(data.contiguous(), cols.contiguous(), indptr, n)


SHAPES = {
    # name: (N, stored nnz target, builder kwargs)                              BASELINE.json configs
    "reddit": dict(n=232_965, nnz=114_600_000, max_degree=22_000, feat=602, hidden=128, classes=41),
    "powerlaw2m": dict(n=2_000_000, nnz=100_000_000, max_degree=50_000, feat=256, hidden=128, classes=16),
    "pubmed": dict(n=19_717, edges=44_338, feat=500, hidden=32, classes=3),
    "cora": dict(n=2_708, edges=5_429, feat=1433, hidden=32, classes=7),
}


def make_shape(name, seed=0, device="cuda", scale=1.0):
    """Build one of the named synthetic shapes (``scale`` < 1 shrinks N and nnz proportionally)."""
    s = SHAPES[name]
    n = max(16, int(s["n"] * scale))
    if "nnz" in s:
        return powerlaw_graph(n, int(s["nnz"] * scale), seed=seed, device=device, max_degree=s["max_degree"],
                              exact=True)
    return gcn_normalized_graph(n, int(s["edges"] * scale), seed=seed, device=device)
