"""Functional wrappers: torch CUDA tensors in, libsgcn_b200 kernels underneath.

Each function validates dtype / device / layout, then passes raw device pointers, row strides and
torch's current stream to the C ABI (include/sgcn_b200.h).  Nothing here computes on the host.
"""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _f32(t, name, dims=2):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == dims):
        raise TypeError("%s must be a CUDA float32 tensor with %d dims" % (name, dims))
    if dims == 2 and t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError("%s must be row-major (unit stride along columns)" % name)
    if dims == 1 and t.numel() > 1 and t.stride(0) != 1:
        raise ValueError("%s must be contiguous" % name)
    return t


def _i32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int32):
        raise TypeError("%s must be a CUDA int32 tensor" % name)
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _ld(t):
    """row stride in elements (a single-row tensor may report any stride(0))"""
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gather_rows(src, idx, out=None, n_dev=None):
    """out[i, :] = src[idx[i], :]  -- history.dense_slice / tf.gather (gcn/history.cpp:74-88)."""
    _f32(src, "src"); _i32(idx, "idx")
    n, c = idx.numel(), src.shape[1]
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=src.device)
    _f32(out, "out")
    check(_lib.load().sgcn_gather_rows(ptr(src), _ld(src), ptr(idx), n, ptr(n_dev), c, ptr(out), _ld(out),
                                       stream_ptr()))
    return out


def pack_dense_weights(w):
    """Split W [K, 128] into tf32 hi / lo parts in the shared-memory image of every K-chunk (sgcn_gemm_pack_w):
    the B operand of gathered_dense; repack after every update of W."""
    _f32(w, "w")
    k, n = w.shape
    lib = _lib.load()
    count = lib.sgcn_gemm_packed_floats(k, n)
    if count < 0:
        raise ValueError("the fused dense layer is built for 128 output columns")
    packed = torch.empty(count, dtype=torch.float32, device=w.device)
    check(lib.sgcn_gemm_pack_w(ptr(w), _ld(w), k, n, ptr(packed), stream_ptr()))
    return packed


def gathered_dense(src, idx, w_packed, k, epilogue="ln_relu", eps=1e-9, out=None, pre=None, stats=None, n_dev=None):
    """out[i, :] = act(LN(src[idx[i], :k] @ W)) in one kernel on the tcgen05 tensor cores (TF32 x 3 split, fp32
    result to ~1e-6): the first dense layer of the pre-processed models (gcn/layers.py:100-138) with the
    feature-row gather (history.dense_slice, gcn/train.py:190) as its A-operand load.  ``idx`` None: rows
    0 .. len(out)-1.  ``epilogue``: "none" | "ln_relu" | "ln".  ``pre`` / ``stats``: optional outputs for the
    backward (raw product; {mean, rstd} per row)."""
    _f32(src, "src")
    if idx is not None:
        _i32(idx, "idx")
    n = idx.numel() if idx is not None else (out.shape[0] if out is not None else src.shape[0])
    if out is None:
        out = torch.empty((n, 128), dtype=torch.float32, device=src.device)
    _f32(out, "out")
    code = {"none": 0, "ln_relu": 1, "ln": 2}[epilogue]
    check(_lib.load().sgcn_gather_gemm_tf32x3(ptr(src), _ld(src), ptr(idx), n, ptr(n_dev), int(k), ptr(w_packed), 128,
                                              ptr(out), _ld(out), ptr(pre), _ld(pre) if pre is not None else 0,
                                              ptr(stats), code, float(eps), stream_ptr()))
    return out


def history_update(hist, idx, rows, n_dev=None, done_counter=None):
    """hist[idx[i], :] = rows[i, :]  -- tf.scatter_update (gcn/models.py:160-166)."""
    _f32(hist, "hist"); _i32(idx, "idx"); _f32(rows, "rows")
    n = idx.numel()
    if rows.shape[0] < n or rows.shape[1] != hist.shape[1]:
        raise ValueError("rows must be [>=len(idx), hist.shape[1]]")
    check(_lib.load().sgcn_history_update(ptr(hist), _ld(hist), ptr(idx), n, ptr(n_dev), ptr(rows), _ld(rows),
                                          hist.shape[1], ptr(done_counter), stream_ptr()))
    return hist


def copy_rows_pad(src, n, out, n_dev=None):
    """out[i] = src[i] for i < n, 0 for n <= i < out.shape[0]."""
    _f32(out, "out")
    if n > 0:
        _f32(src, "src")
    check(_lib.load().sgcn_copy_rows_pad(ptr(src) if n > 0 else None, _ld(src) if n > 0 else 0, n, ptr(n_dev),
                                         out.shape[0], out.shape[1], ptr(out), _ld(out), stream_ptr()))
    return out


def csr_slice(a_d, a_i, a_p, r):
    """history.slice on device (gcn/_history.pyx:25-51): returns (idx2[nnz,2], val[nnz], indptr[n+1])."""
    _f32(a_d, "a_d", 1); _i32(a_i, "a_i"); _i32(a_p, "a_p"); _i32(r, "r")
    n = r.numel()
    lib = _lib.load()
    o_p = torch.empty(n + 1, dtype=torch.int32, device=r.device)
    check(lib.sgcn_csr_slice_indptr(ptr(a_p), ptr(r), n, ptr(o_p), stream_ptr()))
    nnz = int(o_p[n].item())
    idx2 = torch.empty((nnz, 2), dtype=torch.int32, device=r.device)
    val = torch.empty(nnz, dtype=torch.float32, device=r.device)
    if nnz:
        check(lib.sgcn_csr_slice(ptr(a_d), ptr(a_i), ptr(a_p), ptr(r), n, ptr(o_p), ptr(val), ptr(idx2),
                                 stream_ptr()))
    return idx2, val, o_p


def spmm_csr(rowptr, cols, vals, x, n_out, out=None, row_map=None, accumulate=False, n_out_dev=None,
             square=False):
    """out[r] (+)= sum_e vals[e] * x[row_map[cols[e]] or cols[e]]  (gcn/layers.py:31-37, CSR form);
    square=True uses vals[e]^2 (tf.square(adj), gcn/layers.py:242)."""
    _i32(rowptr, "rowptr"); _f32(x, "x")
    d = x.shape[1]
    if out is None:
        out = torch.empty((n_out, d), dtype=torch.float32, device=x.device)
    _f32(out, "out")
    fn = _lib.load().sgcn_spmm_csr_sq if square else _lib.load().sgcn_spmm_csr
    check(fn(ptr(rowptr), ptr(cols), ptr(vals), ptr(row_map), n_out, ptr(n_out_dev),
             ptr(x), _ld(x), d, ptr(out), _ld(out), 1 if accumulate else 0, stream_ptr()))
    return out


def spmm_coo(idx2, vals, x, n_rows, transpose=False, out=None):
    """out[rows[e]] += vals[e] * x[cols[e]] for the reference's COO triples (atomics)."""
    _i32(idx2, "idx2"); _f32(vals, "vals", 1); _f32(x, "x")
    d = x.shape[1]
    if out is None:
        out = torch.zeros((n_rows, d), dtype=torch.float32, device=x.device)
    _f32(out, "out")
    check(_lib.load().sgcn_spmm_coo(ptr(idx2), ptr(vals), vals.numel(), ptr(x), _ld(x), d, ptr(out),
                                    _ld(out), 1 if transpose else 0, stream_ptr()))
    return out


def spmm_csr_bwd(rowptr, cols, vals, dy, dx, n_out, rscale=None, n_out_dev=None, square=False):
    """dx[cols[e]] += vals[e] * rscale[r] * dy[r]   (gradient of the sampled SpMM; square: vals[e]^2)."""
    _i32(rowptr, "rowptr"); _f32(dy, "dy"); _f32(dx, "dx")
    d = dy.shape[1]
    if dx.shape[1] != d:
        raise ValueError("dx and dy must have the same width")
    fn = _lib.load().sgcn_spmm_csr_bwd_sq if square else _lib.load().sgcn_spmm_csr_bwd
    check(fn(ptr(rowptr), ptr(cols), ptr(vals), ptr(rscale), n_out, ptr(n_out_dev),
             ptr(dy), _ld(dy), d, ptr(dx), _ld(dx), stream_ptr()))
    return dx


def full_history_mean(nodes, rowptr_f, n_out, adj_p, adj_i, adj_w, hist, y0, y1=None, n_out_dev=None,
                      work_counter=None, square=False):
    """y0[r] (+= and y1[r] +=) sum over the stored row of nodes[r] of adj_w * hist[adj_i]
    (square: adj_w^2, tf.square(fadj) of gcn/layers.py:338)."""
    _i32(nodes, "nodes"); _i32(rowptr_f, "rowptr_f"); _f32(hist, "hist"); _f32(y0, "y0")
    d = hist.shape[1]
    if y0.shape[1] != d or (y1 is not None and y1.shape[1] != d):
        raise ValueError("outputs must have hist's width")
    fn = _lib.load().sgcn_full_history_mean_sq if square else _lib.load().sgcn_full_history_mean
    check(fn(ptr(nodes), ptr(rowptr_f), n_out, ptr(n_out_dev), ptr(adj_p),
             ptr(adj_i), ptr(adj_w), ptr(hist), _ld(hist), d, ptr(y0), _ld(y0),
             ptr(y1), _ld(y1) if y1 is not None else 0, ptr(work_counter), stream_ptr()))
    return y0


def cv_sampled_fwd(rowptr, cols, vals, tgt, n_out, x, hist, y, self_out=None, n_out_dev=None, accumulate=False):
    """y[r] = sum_e vals[e] (x[cols[e]] - hist[tgt[e]]); self_out[r] = x[r]  (gcn/layers.py:350-362)."""
    _f32(x, "x"); _f32(hist, "hist"); _f32(y, "y")
    d = x.shape[1]
    check(_lib.load().sgcn_cv_sampled_fwd(ptr(rowptr), ptr(cols), ptr(vals), ptr(tgt), n_out, ptr(n_out_dev),
                                          ptr(x), _ld(x), ptr(hist), _ld(hist), d, ptr(y), _ld(y),
                                          ptr(self_out), _ld(self_out) if self_out is not None else 0,
                                          1 if accumulate else 0, stream_ptr()))
    return y


def cvd_sampled_fwd(rowptr, cols, vals, tgt, scale, n_out, h, mu, hist, yh, ymu, self_h=None, self_mu=None,
                    n_out_dev=None, accumulate=False):
    """CVD sampled part (gcn/layers.py:298-319); see include/sgcn_b200.h:sgcn_cvd_sampled_fwd."""
    _f32(h, "h"); _f32(mu, "mu"); _f32(hist, "hist"); _f32(yh, "yh"); _f32(ymu, "ymu")
    d = h.shape[1]
    check(_lib.load().sgcn_cvd_sampled_fwd(
        ptr(rowptr), ptr(cols), ptr(vals), ptr(tgt), ptr(scale), n_out, ptr(n_out_dev), ptr(h), _ld(h),
        ptr(mu), _ld(mu), ptr(hist), _ld(hist), d, ptr(yh), _ld(yh), ptr(ymu), _ld(ymu),
        ptr(self_h), _ld(self_h) if self_h is not None else 0,
        ptr(self_mu), _ld(self_mu) if self_mu is not None else 0, 1 if accumulate else 0, stream_ptr()))
    return yh, ymu


def det_sampled_fwd(rowptr, cols, vals, mvals, tgt, n_out, var, hvar, y, pre=None, self_out=None, n_out_dev=None,
                    accumulate=True):
    """variance stream of the det-dropout VRAggregator (gcn/layers.py:331-341); see
    include/sgcn_b200.h:sgcn_det_sampled_fwd."""
    _f32(var, "var"); _f32(hvar, "hvar"); _f32(y, "y"); _f32(vals, "vals", 1); _f32(mvals, "mvals", 1)
    d = var.shape[1]
    check(_lib.load().sgcn_det_sampled_fwd(
        ptr(rowptr), ptr(cols), ptr(vals), ptr(mvals), ptr(tgt), n_out, ptr(n_out_dev), ptr(var), _ld(var),
        ptr(hvar), _ld(hvar), d, ptr(y), _ld(y), ptr(pre), _ld(pre) if pre is not None else 0,
        ptr(self_out), _ld(self_out) if self_out is not None else 0, 1 if accumulate else 0, stream_ptr()))
    return y


def det_sampled_bwd(rowptr, cols, vals, mvals, tgt, n_out, var, hvar, dy, pre, dvar, n_out_dev=None):
    """dvar += gradient of det_sampled_fwd w.r.t. var (dvar pre-initialised)."""
    _f32(var, "var"); _f32(hvar, "hvar"); _f32(dy, "dy"); _f32(pre, "pre"); _f32(dvar, "dvar")
    d = var.shape[1]
    check(_lib.load().sgcn_det_sampled_bwd(
        ptr(rowptr), ptr(cols), ptr(vals), ptr(mvals), ptr(tgt), n_out, ptr(n_out_dev), ptr(var), _ld(var),
        ptr(hvar), _ld(hvar), d, ptr(dy), _ld(dy), ptr(pre), _ld(pre), ptr(dvar), _ld(dvar), stream_ptr()))
    return dvar


def copy_rows_pad_pair(src0, n0, out0, src1, n1, out1, n0_dev=None, n1_dev=None):
    """Two copy_rows_pad jobs in one launch (out1 may be None)."""
    _f32(out0, "out0")
    a = (ptr(src0) if n0 > 0 else None, _ld(src0) if n0 > 0 else 0, n0, ptr(n0_dev), out0.shape[0], out0.shape[1],
         ptr(out0), _ld(out0))
    if out1 is None:
        b = (None, 0, 0, None, 0, 0, None, 0)
    else:
        _f32(out1, "out1")
        b = (ptr(src1) if n1 > 0 else None, _ld(src1) if n1 > 0 else 0, n1, ptr(n1_dev), out1.shape[0],
             out1.shape[1], ptr(out1), _ld(out1))
    check(_lib.load().sgcn_copy_rows_pad_pair(*a, *b, stream_ptr()))


def gather_pad_pair(src, idx, out, src0, n0, out0, src1, n1, out1, n_dev=None, n0_dev=None, n1_dev=None):
    """gather_rows(src, idx) -> out fused with copy_rows_pad_pair (out1 may be None): one launch."""
    _f32(src, "src"); _i32(idx, "idx"); _f32(out, "out"); _f32(out0, "out0")
    a = (ptr(src0) if n0 > 0 else None, _ld(src0) if n0 > 0 else 0, n0, ptr(n0_dev), out0.shape[0], out0.shape[1],
         ptr(out0), _ld(out0))
    if out1 is None:
        b = (None, 0, 0, None, 0, 0, None, 0)
    else:
        _f32(out1, "out1")
        b = (ptr(src1) if n1 > 0 else None, _ld(src1) if n1 > 0 else 0, n1, ptr(n1_dev), out1.shape[0],
             out1.shape[1], ptr(out1), _ld(out1))
    check(_lib.load().sgcn_gather_pad_pair(ptr(src), _ld(src), ptr(idx), idx.numel(), ptr(n_dev), src.shape[1],
                                           ptr(out), _ld(out), *a, *b, stream_ptr()))
    return out


def cv_sampled_fwd_bwd(rowptr, cols, vals, tgt, n_out, x, hist, y, dy, dx, self_out=None, n_out_dev=None,
                       accumulate=False):
    """cv_sampled_fwd fused with the backward scatter dx[cols[e]] += vals[e] * dy[r] (dx pre-initialised)."""
    _f32(x, "x"); _f32(hist, "hist"); _f32(y, "y"); _f32(dy, "dy"); _f32(dx, "dx")
    d = x.shape[1]
    check(_lib.load().sgcn_cv_sampled_fwd_bwd(
        ptr(rowptr), ptr(cols), ptr(vals), ptr(tgt), n_out, ptr(n_out_dev), ptr(x), _ld(x), ptr(hist), _ld(hist),
        d, ptr(y), _ld(y), ptr(self_out), _ld(self_out) if self_out is not None else 0, 1 if accumulate else 0,
        ptr(dy), _ld(dy), ptr(dx), _ld(dx), stream_ptr()))
    return y


def cvd_sampled_fwd_bwd(rowptr, cols, vals, tgt, scale, n_out, h, mu, hist, yh, ymu, dy, dx, self_h=None,
                        self_mu=None, n_out_dev=None, accumulate=False):
    """cvd_sampled_fwd fused with dx[cols[e]] += vals[e] * scale[r] * dy[r] (dx pre-initialised)."""
    _f32(h, "h"); _f32(mu, "mu"); _f32(hist, "hist"); _f32(yh, "yh"); _f32(ymu, "ymu"); _f32(dy, "dy"); _f32(dx, "dx")
    d = h.shape[1]
    check(_lib.load().sgcn_cvd_sampled_fwd_bwd(
        ptr(rowptr), ptr(cols), ptr(vals), ptr(tgt), ptr(scale), n_out, ptr(n_out_dev), ptr(h), _ld(h),
        ptr(mu), _ld(mu), ptr(hist), _ld(hist), d, ptr(yh), _ld(yh), ptr(ymu), _ld(ymu),
        ptr(self_h), _ld(self_h) if self_h is not None else 0,
        ptr(self_mu), _ld(self_mu) if self_mu is not None else 0, 1 if accumulate else 0,
        ptr(dy), _ld(dy), ptr(dx), _ld(dx), stream_ptr()))
    return yh, ymu


def csr_spmm(indptr, indices, data, x, out=None, tile_cols=0):
    """out = A @ x for a whole CSR matrix on the device (the reference's ``adj.dot(feats)``,
    gcn/utils.py:168-169,321-322).  x / out may be column views of wider row-major matrices."""
    _i32(indptr, "indptr"); _i32(indices, "indices"); _f32(data, "data", 1); _f32(x, "x")
    n = indptr.numel() - 1
    if out is None:
        out = torch.empty((n, x.shape[1]), dtype=torch.float32, device=x.device)
    _f32(out, "out")
    if out.shape[0] != n or out.shape[1] != x.shape[1]:
        raise ValueError("out must be [n_rows, x.shape[1]]")
    check(_lib.load().sgcn_csr_spmm(ptr(indptr), ptr(indices), ptr(data), n, ptr(x), _ld(x), x.shape[1],
                                    ptr(out), _ld(out), int(tile_cols), stream_ptr()))
    return out


def preprocess_features(indptr, indices, data, feats, normalization="graphsage", tile_cols=0):
    """Model input of the PP ('preprocess') models: ``[feats | A @ feats]`` for graphsage normalisation,
    ``A @ feats`` for gcn (gcn/models.py:230-239 with FLAGS.pp_nbr; gcn/utils.py:168-169,321-322)."""
    _f32(feats, "feats")
    n, f = feats.shape
    if normalization == "gcn":
        return csr_spmm(indptr, indices, data, feats, tile_cols=tile_cols)
    out = torch.empty((n, 2 * f), dtype=torch.float32, device=feats.device)
    out[:, :f].copy_(feats)
    csr_spmm(indptr, indices, data, feats, out=out[:, f:], tile_cols=tile_cols)
    return out
