"""Aggregation layers -- the reference's ``PlainAggregator`` / ``VRAggregator`` protocol
(gcn/layers.py:214-257, 282-362) over the sm_100a kernels of libsgcn_b200.so.

Same constructor arguments and call protocol as the reference layers: ``layer(inputs)`` returns
the aggregated activations (a tensor, or a ``(h, mu)`` pair for CVD) and leaves ``new_history``
on the layer for the caller to write back (gcn/models.py:160-166 -> ``write_back``).  Inputs are
CUDA float32 tensors; the sparse operands are ``DeviceAdj`` / ``FullNeighbours`` descriptors built
either straight from the device sampler (no host round-trip) or from the reference-format feed
dict triples.  Gradients flow exactly where the reference's do: through ``adj @ inputs`` (CV) and
``(adj @ (h - mu)) * scale`` (CVD); history is not trainable.
"""
import numpy as np
import torch

from . import ops


class DeviceAdj:
    """Row-sorted sampled adjacency (CSR) in HBM: rows = output field, cols = index into the input field."""

    def __init__(self, rowptr, cols, vals, n_out, n_in, tgt=None, n_out_dev=None):
        self.rowptr, self.cols, self.vals, self.tgt = rowptr, cols, vals, tgt
        self.n_out, self.n_in, self.n_out_dev = int(n_out), int(n_in), n_out_dev

    @staticmethod
    def from_level(lv, exact_n_out=None):
        """From a ``scheduler.DeviceLevel`` (device sampler output)."""
        n_out = lv.n_out_bound if exact_n_out is None else exact_n_out
        return DeviceAdj(lv.rowptr_s, lv.edg_t, lv.edg_w, n_out, lv.n_in_bound, tgt=lv.tgt,
                         n_out_dev=lv.meta[0:1])

    @staticmethod
    def from_coo(triple, device="cuda", ifield=None):
        """From the reference's feed-dict triple ``(idx[ne,2], val[ne], shape)``; rows must be sorted
        ascending, which ``Scheduler::expand`` guarantees (edges are emitted row by row)."""
        idx, val, shape = triple
        idx = np.asarray(idx, dtype=np.int32).reshape(-1, 2)
        rows = idx[:, 0]
        if rows.size and np.any(np.diff(rows) < 0):
            raise ValueError("adjacency rows must be sorted ascending (use ops.spmm_coo for general COO)")
        n_out, n_in = int(shape[0]), int(shape[1])
        rowptr = np.searchsorted(rows, np.arange(n_out + 1), side="left").astype(np.int32)
        dev = torch.device(device)
        cols = torch.from_numpy(np.ascontiguousarray(idx[:, 1])).to(dev)
        tgt = None
        if ifield is not None:
            ifield_t = ifield if isinstance(ifield, torch.Tensor) else torch.from_numpy(
                np.ascontiguousarray(ifield, dtype=np.int32)).to(dev)
            tgt = ifield_t[cols.long()].contiguous()
        return DeviceAdj(torch.from_numpy(rowptr).to(dev), cols,
                         torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32)).to(dev), n_out, n_in, tgt=tgt)


class FullNeighbours:
    """The full-neighbour adjacency of the output field (``fadj`` + ``ffield`` of the reference).

    ``in_place``: rows are read straight out of the sampler's CSR (nodes[r] = global id of output
    row r) -- nothing is materialised.  ``from_coo``: the reference-format triple + ffield."""

    def __init__(self):
        self.nodes = self.rowptr_f = self.adj_p = self.adj_i = self.adj_w = None
        self.cols = self.vals = self.ffield = None

    @staticmethod
    def in_place(nodes, rowptr_f, adj_p, adj_i, adj_w):
        f = FullNeighbours()
        f.nodes, f.rowptr_f, f.adj_p, f.adj_i, f.adj_w = nodes, rowptr_f, adj_p, adj_i, adj_w
        return f

    @staticmethod
    def from_sampler(sampler, lv):
        return FullNeighbours.in_place(lv.field, lv.rowptr_f, sampler.view("adj_p"), sampler.view("adj_i"),
                                       sampler.view("adj_w"))

    @staticmethod
    def from_coo(triple, ffield, device="cuda"):
        idx, val, shape = triple
        idx = np.asarray(idx, dtype=np.int32).reshape(-1, 2)
        rows = idx[:, 0]
        if rows.size and np.any(np.diff(rows) < 0):
            raise ValueError("fadj rows must be sorted ascending")
        dev = torch.device(device)
        f = FullNeighbours()
        f.rowptr_f = torch.from_numpy(np.searchsorted(rows, np.arange(int(shape[0]) + 1), side="left")
                                      .astype(np.int32)).to(dev)
        f.cols = torch.from_numpy(np.ascontiguousarray(idx[:, 1])).to(dev)
        f.vals = torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32)).to(dev)
        f.ffield = ffield if isinstance(ffield, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(ffield, dtype=np.int32)).to(dev)
        return f

    def add_history_mean(self, n_out, n_out_dev, hist, y0, y1=None, square=False):
        """y0 (+y1) += fadj @ gather(history, ffield)   (gcn/layers.py:305,309,354,357);
        square: tf.square(fadj) @ gather(var_history, ffield)   (gcn/layers.py:338)"""
        if self.nodes is not None:
            ops.full_history_mean(self.nodes, self.rowptr_f, n_out, self.adj_p, self.adj_i, self.adj_w, hist,
                                  y0, y1, n_out_dev=n_out_dev, square=square)
        else:
            ops.spmm_csr(self.rowptr_f, self.cols, self.vals, hist, n_out, out=y0, row_map=self.ffield,
                         accumulate=True, n_out_dev=n_out_dev, square=square)
            if y1 is not None:
                ops.spmm_csr(self.rowptr_f, self.cols, self.vals, hist, n_out, out=y1, row_map=self.ffield,
                             accumulate=True, n_out_dev=n_out_dev, square=square)


def _split(d_out, dim, concat):
    """(d_self, d_neighbour) views of the gradient of a [self | neighbour] output."""
    if concat:
        return d_out[:, :dim], d_out[:, dim:]
    return None, d_out


def _grad_init(adj, d_self, n_rows, dim, device):
    """dx = [d_self ; 0]: the gradient of the self half of a concat output, zero elsewhere."""
    dx = torch.empty((n_rows, dim), dtype=torch.float32, device=device)
    if d_self is not None:
        ops.copy_rows_pad(d_self, adj.n_out, dx, n_dev=adj.n_out_dev)
    else:
        ops.copy_rows_pad(None, 0, dx)
    return dx


def _input_grad(adj, d_self, d_nb, n_rows, dim, rscale=None, square=False):
    """dx = adj^T (d_nb * rscale) (+ d_self on the first n_out rows)  -- the SpMM backward."""
    dx = _grad_init(adj, d_self, n_rows, dim, d_nb.device)
    ops.spmm_csr_bwd(adj.rowptr, adj.cols, adj.vals, d_nb, dx, adj.n_out, rscale=rscale, n_out_dev=adj.n_out_dev,
                     square=square)
    return dx


class _PlainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, adj, concat):
        dim = x.shape[1]
        out = torch.empty((adj.n_out, dim * (2 if concat else 1)), dtype=torch.float32, device=x.device)
        nb = out[:, dim:] if concat else out
        ops.spmm_csr(adj.rowptr, adj.cols, adj.vals, x, adj.n_out, out=nb, n_out_dev=adj.n_out_dev)
        if concat:
            ops.copy_rows_pad(x, adj.n_out, out[:, :dim], n_dev=adj.n_out_dev)
        ctx.adj, ctx.concat, ctx.shape = adj, concat, tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        d_out = d_out.contiguous()
        d_self, d_nb = _split(d_out, ctx.shape[1], ctx.concat)
        return _input_grad(ctx.adj, d_self, d_nb, ctx.shape[0], ctx.shape[1]), None, None


class _CVFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, adj, full, hist, concat):
        dim = x.shape[1]
        out = torch.empty((adj.n_out, dim * (2 if concat else 1)), dtype=torch.float32, device=x.device)
        nb = out[:, dim:] if concat else out
        ops.cv_sampled_fwd(adj.rowptr, adj.cols, adj.vals, adj.tgt, adj.n_out, x, hist, nb,
                           self_out=out[:, :dim] if concat else None, n_out_dev=adj.n_out_dev)
        full.add_history_mean(adj.n_out, adj.n_out_dev, hist, nb)
        ctx.adj, ctx.concat, ctx.shape = adj, concat, tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        d_out = d_out.contiguous()
        d_self, d_nb = _split(d_out, ctx.shape[1], ctx.concat)
        return _input_grad(ctx.adj, d_self, d_nb, ctx.shape[0], ctx.shape[1]), None, None, None, None


class _CVDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, mu, adj, full, hist, scale, concat):
        dim = h.shape[1]
        width = dim * (2 if concat else 1)
        out_h = torch.empty((adj.n_out, width), dtype=torch.float32, device=h.device)
        out_mu = torch.empty((adj.n_out, width), dtype=torch.float32, device=h.device)
        nb_h = out_h[:, dim:] if concat else out_h
        nb_mu = out_mu[:, dim:] if concat else out_mu
        ops.cvd_sampled_fwd(adj.rowptr, adj.cols, adj.vals, adj.tgt, scale, adj.n_out, h, mu, hist, nb_h, nb_mu,
                            self_h=out_h[:, :dim] if concat else None,
                            self_mu=out_mu[:, :dim] if concat else None, n_out_dev=adj.n_out_dev)
        full.add_history_mean(adj.n_out, adj.n_out_dev, hist, nb_mu, nb_h)
        ctx.adj, ctx.concat, ctx.shape, ctx.scale = adj, concat, tuple(h.shape), scale
        return out_h, out_mu

    @staticmethod
    def backward(ctx, d_h, d_mu):
        adj, dim = ctx.adj, ctx.shape[1]
        d_h = d_h.contiguous()
        dh_self, dh_nb = _split(d_h, dim, ctx.concat)
        grad_h = grad_mu = None
        if ctx.needs_input_grad[0]:
            grad_h = _input_grad(adj, dh_self, dh_nb, ctx.shape[0], dim, rscale=ctx.scale)
        if ctx.needs_input_grad[1]:
            # the reference never needs this (mu is stop_gradient-ed, gcn/layers.py:412); kept exact
            # for callers that do differentiate mu:  d mu = adj^T (d mu_nb + d h_nb (1 - scale)) + self
            d_mu = d_mu.contiguous()
            dm_self, dm_nb = _split(d_mu, dim, ctx.concat)
            mix = (dm_nb + dh_nb * (1.0 - ctx.scale[:adj.n_out, None])).contiguous()
            grad_mu = _input_grad(adj, dm_self, mix, ctx.shape[0], dim)
        return grad_h, grad_mu, None, None, None, None, None


class _PlainVarFn(torch.autograd.Function):
    """var stream of PlainAggregator's (mu, var) branch: tf.square(adj) @ var  (gcn/layers.py:238-247)."""

    @staticmethod
    def forward(ctx, var, adj, concat):
        dim = var.shape[1]
        out = torch.empty((adj.n_out, dim * (2 if concat else 1)), dtype=torch.float32, device=var.device)
        nb = out[:, dim:] if concat else out
        ops.spmm_csr(adj.rowptr, adj.cols, adj.vals, var, adj.n_out, out=nb, n_out_dev=adj.n_out_dev, square=True)
        if concat:
            ops.copy_rows_pad(var, adj.n_out, out[:, :dim], n_dev=adj.n_out_dev)
        ctx.adj, ctx.concat, ctx.shape = adj, concat, tuple(var.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        d_out = d_out.contiguous()
        d_self, d_nb = _split(d_out, ctx.shape[1], ctx.concat)
        return _input_grad(ctx.adj, d_self, d_nb, ctx.shape[0], ctx.shape[1], square=True), None, None


class _DetVarFn(torch.autograd.Function):
    """var stream of VRAggregator's det-dropout branch (gcn/layers.py:331-341):
    relu(adj^2 @ dsigma^2 + fadj^2 @ var_history[ffield] + 2 madj @ (dsigma * sigma_bar)) + 1e-10."""

    @staticmethod
    def forward(ctx, var, adj, mvals, full, hvar, concat):
        dim = var.shape[1]
        out = torch.empty((adj.n_out, dim * (2 if concat else 1)), dtype=torch.float32, device=var.device)
        nb = out[:, dim:] if concat else out
        ops.copy_rows_pad(None, 0, nb)                       # zero: the history term adds with REDs
        full.add_history_mean(adj.n_out, adj.n_out_dev, hvar, nb, square=True)
        pre = torch.empty((adj.n_out, dim), dtype=torch.float32, device=var.device)
        ops.det_sampled_fwd(adj.rowptr, adj.cols, adj.vals, mvals, adj.tgt, adj.n_out, var, hvar, nb, pre=pre,
                            self_out=out[:, :dim] if concat else None, n_out_dev=adj.n_out_dev, accumulate=True)
        ctx.adj, ctx.mvals, ctx.hvar, ctx.concat, ctx.shape = adj, mvals, hvar, concat, tuple(var.shape)
        ctx.save_for_backward(var, pre)
        return out

    @staticmethod
    def backward(ctx, d_out):
        var, pre = ctx.saved_tensors
        adj, dim = ctx.adj, ctx.shape[1]
        d_out = d_out.contiguous()
        d_self, d_nb = _split(d_out, dim, ctx.concat)
        dvar = _grad_init(adj, d_self, ctx.shape[0], dim, d_out.device)
        ops.det_sampled_bwd(adj.rowptr, adj.cols, adj.vals, ctx.mvals, adj.tgt, adj.n_out, var, ctx.hvar, d_nb, pre,
                            dvar, n_out_dev=adj.n_out_dev)
        return dvar, None, None, None, None, None


class Layer:
    """Call protocol of gcn/layers.py:40-84 (``layer(inputs)`` -> ``_call``), without TF scoping."""

    def __init__(self, name=None, normalization="gcn"):
        self.name = name or self.__class__.__name__.lower()
        self.normalization = normalization     # FLAGS.normalization: 'gcn' or 'graphsage'

    def __call__(self, inputs):
        return self._call(inputs)

    @property
    def _concat(self):
        return self.normalization != "gcn"


class GatherAggregator(Layer):
    """tf.gather(inputs, field)  (gcn/layers.py:205-211)"""

    def __init__(self, field, **kwargs):
        super().__init__(**kwargs)
        self.field = field

    def _call(self, inputs):
        return _GatherFn.apply(inputs, self.field)


class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, field):
        ctx.field, ctx.shape = field, tuple(x.shape)
        return ops.gather_rows(x, field)

    @staticmethod
    def backward(ctx, d_out):
        # gradient of a gather is a scatter-ADD (indices may repeat): route through the COO kernel
        n = ctx.field.numel()
        idx2 = torch.stack((torch.arange(n, dtype=torch.int32, device=d_out.device), ctx.field), dim=1).contiguous()
        ones = torch.ones(n, dtype=torch.float32, device=d_out.device)
        dx = ops.spmm_coo(idx2, ones, d_out.contiguous(), ctx.shape[0], transpose=True)
        return dx, None


class PlainAggregator(Layer):
    """H -> Z = A H  (gcn/layers.py:214-257, non-tuple branch)."""

    def __init__(self, adj, **kwargs):
        super().__init__(**kwargs)
        self.adj = adj

    def _call(self, inputs):
        if isinstance(inputs, tuple):                      # det-dropout (mu, var), gcn/layers.py:238-247
            mu, var = inputs
            return (_PlainFn.apply(mu, self.adj, self._concat), _PlainVarFn.apply(var, self.adj, self._concat))
        return _PlainFn.apply(inputs, self.adj, self._concat)


class VRAggregator(Layer):
    """Control-variate aggregator (gcn/layers.py:282-362), CV and CVD branches.

    ``fadj``/``ffield`` are folded into one ``FullNeighbours`` descriptor.  ``madj`` is only used by
    the det-dropout branch (inputs = ``(mu, var)`` with ``cvd`` False, gcn/layers.py:320-349): the
    weights of the sampler's ``medg_w`` as a CUDA float32 tensor aligned with ``adj.vals`` (or a
    ``DeviceAdj`` / reference triple whose values are taken); there ``history`` holds TWO tables
    (mean, variance).  Otherwise ``history`` is a list with one [N, dim] CUDA tensor.  ``ifield`` is
    kept for the write-back (``adj.tgt`` already holds ``ifield[cols]``)."""

    def __init__(self, adj, fadj, madj, ifield, ffield, history, scale, cvd, **kwargs):
        super().__init__(**kwargs)
        self.adj, self.fadj, self.madj = adj, fadj, madj
        self.ifield, self.ffield = ifield, ffield
        self.history, self.scale, self.cvd = history, scale, cvd
        self.new_history = None
        if adj.tgt is None:
            adj.tgt = ifield[adj.cols.long()].contiguous()

    def _call(self, inputs):
        hist = self.history[0]
        if self.cvd:
            h, mu = inputs
            out = _CVDFn.apply(h, mu, self.adj, self.fadj, hist, self.scale, self._concat)
            self.new_history = [mu]
            return out
        if isinstance(inputs, tuple):                      # det-dropout, gcn/layers.py:320-349
            mu, var = inputs
            mu_hist, var_hist = self.history
            out_mu = _CVFn.apply(mu, self.adj, self.fadj, mu_hist, self._concat)
            out_var = _DetVarFn.apply(var, self.adj, self._madj_vals(), self.fadj, var_hist, self._concat)
            self.new_history = (mu, var)
            return out_mu, out_var
        out = _CVFn.apply(inputs, self.adj, self.fadj, hist, self._concat)
        self.new_history = [inputs]
        return out

    def _madj_vals(self):
        m = self.madj
        if isinstance(m, torch.Tensor):
            return m
        if isinstance(m, DeviceAdj):
            return m.vals
        if m is None:
            raise ValueError("the det-dropout branch needs madj (the sampler's medg_w)")
        return torch.from_numpy(np.ascontiguousarray(m[1], dtype=np.float32)).to(self.adj.vals.device)

    def write_back(self, n_in_dev=None):
        """tf.scatter_update(history, fields[l], new_history) after the step (gcn/models.py:160-166,186-194)."""
        for hist, new in zip(self.history, self.new_history):
            ops.history_update(hist, self.ifield, new.detach(), n_dev=n_in_dev)
