"""Dense layers, loss and optimiser around the aggregate -- the reference's ``Dense``,
``AugmentedDropoutDense``, ``Dropout``, ``MyLayerNorm[2]`` (gcn/layers.py:87-138,365-433), the loss
(gcn/models.py:68-83), ``tf.train.AdamOptimizer`` (models.py:50-51) and the layer stacking of the
pre-processed ("PP") models (models.py:256-337) -- over the row-wise sm_100a kernels of
libsgcn_b200.so (csrc/dense.cu).  The matrix products ``X @ W`` are plain library GEMMs (cuBLAS via
torch.mm; sparse inputs go through the library's own CSR product).  Same layer names, constructor
arguments and call protocol as the reference, so a model is assembled the way ``GCN._build`` does.

Not reproducible from the reference: TensorFlow's random streams (dropout masks, Glorot draws) --
weights are inputs (``glorot`` below draws them from a NumPy generator) and dropout masks come from
the library's Philox generator or are injected.
"""
import math

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr
from .layers import Layer


# ---- functions --------------------------------------------------------------------------------------
class _LnActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, offset, eps, relu):
        x = x.contiguous()
        n, d = x.shape
        y = torch.empty_like(x)
        stats = torch.empty((n, 2), dtype=torch.float32, device=x.device)
        check(_lib.load().sgcn_ln_act_fwd(ptr(x), x.stride(0), n, None, d, ptr(scale), ptr(offset), float(eps),
                                          1 if relu else 0, ptr(y), y.stride(0), ptr(stats), stream_ptr()))
        ctx.save_for_backward(x, y, stats, scale if scale is not None else torch.empty(0, device=x.device))
        ctx.relu, ctx.has_scale, ctx.has_offset = relu, scale is not None, offset is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, stats, scale = ctx.saved_tensors
        dy = dy.contiguous()
        n, d = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        ds = torch.zeros(d, dtype=torch.float32, device=x.device) if ctx.has_scale and ctx.needs_input_grad[1] else None
        do = torch.zeros(d, dtype=torch.float32, device=x.device) if ctx.has_offset and ctx.needs_input_grad[2] else None
        check(_lib.load().sgcn_ln_act_bwd(ptr(x), x.stride(0), ptr(y), y.stride(0), ptr(dy), dy.stride(0), n, None, d,
                                          ptr(scale) if ctx.has_scale else None, ptr(stats), 1 if ctx.relu else 0,
                                          ptr(dx), dx.stride(0) if dx is not None else 0, ptr(ds), ptr(do),
                                          stream_ptr()))
        return dx, ds, do, None, None


def layer_norm_act(x, scale=None, offset=None, eps=1e-9, relu=True):
    """act(MyLayerNorm2(x, offset, scale)) (gcn/layers.py:87-97): row moments, batch_normalization form."""
    return _LnActFn.apply(x, scale, offset, eps, relu)


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, keep_prob, mask, seed, offset):
        x = x.contiguous()
        n, d = x.shape
        y = torch.empty_like(x)
        used = torch.empty((n, d), dtype=torch.uint8, device=x.device)
        if mask is not None:
            mask = mask.to(device=x.device, dtype=torch.uint8).contiguous()
        check(_lib.load().sgcn_dropout(ptr(x), x.stride(0), n, None, d, float(keep_prob), int(seed), int(offset),
                                       ptr(mask), ptr(used), ptr(y), y.stride(0), stream_ptr()))
        ctx.keep, ctx.used = keep_prob, used
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        n, d = dy.shape
        dx = torch.empty_like(dy)
        check(_lib.load().sgcn_dropout(ptr(dy), dy.stride(0), n, None, d, float(ctx.keep), 0, 0, ptr(ctx.used), None,
                                       ptr(dx), dx.stride(0), stream_ptr()))
        return dx, None, None, None, None


class DropoutState:
    """Seed + running counter of the library's Philox stream (one counter value per 4 elements)."""

    def __init__(self, seed=1):
        self.seed, self.offset = int(seed), 0

    def take(self, numel):
        off = self.offset
        self.offset += (numel + 3) // 4
        return self.seed, off


_default_dropout_state = DropoutState(1)


def dropout(x, keep_prob, mask=None, state=None):
    """tf.nn.dropout(x, keep_prob).  mask: optional injected keep-mask ([n, D], nonzero = keep)."""
    if keep_prob >= 1.0 and mask is None:
        return x
    seed, off = (0, 0) if mask is not None else (state or _default_dropout_state).take(x.numel())
    return _DropoutFn.apply(x, keep_prob, mask, seed, off)


class _XentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, sigmoid):
        logits, labels = logits.contiguous(), labels.contiguous()
        n, c = logits.shape
        loss = torch.zeros((), dtype=torch.float32, device=logits.device)
        d = torch.empty_like(logits)
        check(_lib.load().sgcn_xent(ptr(logits), logits.stride(0), ptr(labels), labels.stride(0), n, c,
                                    1 if sigmoid else 0, ptr(loss), ptr(d), d.stride(0), stream_ptr()))
        ctx.save_for_backward(d)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g, None, None


def cross_entropy(logits, labels, multitask=False):
    """tf.reduce_mean(softmax|sigmoid_cross_entropy_with_logits(...))  (gcn/models.py:76-83)."""
    return _XentFn.apply(logits, labels, multitask)


def glorot(shape, rng):
    """Glorot-uniform [fan_in, fan_out] weights (tf.get_variable's default initialiser, gcn/inits.py:10-12),
    drawn from a NumPy RandomState so that the oracle and the GPU path share them."""
    lim = math.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


class _SparseMMFn(torch.autograd.Function):
    """dot(x, W, sparse=True) for a row-sorted COO/CSR x (gcn/layers.py:31-37 on the sparse feature rows
    history.slice returns): y = x @ W ; dW = x^T dy.  x carries no gradient (it is data)."""

    @staticmethod
    def forward(ctx, w, rowptr, cols, vals, n_rows):
        ctx.save_for_backward(rowptr, cols, vals)
        ctx.n_rows, ctx.shape = n_rows, tuple(w.shape)
        return ops.spmm_csr(rowptr, cols, vals, w.contiguous(), n_rows)

    @staticmethod
    def backward(ctx, dy):
        rowptr, cols, vals = ctx.saved_tensors
        dw = torch.zeros(ctx.shape, dtype=torch.float32, device=dy.device)
        ops.spmm_csr_bwd(rowptr, cols, vals, dy.contiguous(), dw, ctx.n_rows)
        return dw, None, None, None, None


class SparseRows:
    """Sparse feature rows in CSR on the device (what ``history.slice`` + ``tf.sparse_reorder`` feed)."""

    def __init__(self, rowptr, cols, vals, n_rows, n_cols):
        self.rowptr, self.cols, self.vals, self.n_rows, self.n_cols = rowptr, cols, vals, int(n_rows), int(n_cols)

    @staticmethod
    def from_slice(idx2, val, indptr, n_cols):
        """from ``ops.csr_slice`` (device)"""
        return SparseRows(indptr, idx2[:, 1].contiguous(), val, indptr.numel() - 1, n_cols)


def dot(x, w, sparse=False):
    """gcn/layers.py:31-37"""
    if sparse or isinstance(x, SparseRows):
        return _SparseMMFn.apply(w, x.rowptr, x.cols, x.vals, x.n_rows)
    return torch.mm(x, w)


class _GatheredDenseFn(torch.autograd.Function):
    """act(MyLayerNorm(features[idx] @ W)) in ONE kernel on the tcgen05 tensor cores (csrc/gemm.cu: the row gather
    is the GEMM's A-operand load, TF32 x 3 split, layer norm + activation out of tensor memory).  Backward: the
    row-wise layer-norm backward kernel, then dW = X^T dPre as a library GEMM over the gathered rows."""

    @staticmethod
    def forward(ctx, w, features, idx, norm, relu):
        from . import ops
        n, k = idx.numel(), w.shape[0]
        packed = ops.pack_dense_weights(w.detach())
        y = torch.empty((n, w.shape[1]), dtype=torch.float32, device=w.device)
        pre = torch.empty_like(y) if norm else None
        stats = torch.empty((n, 2), dtype=torch.float32, device=w.device) if norm else None
        ops.gathered_dense(features, idx, packed, k, epilogue=("ln_relu" if relu else "ln") if norm else "none",
                           out=y, pre=pre, stats=stats)
        if not norm and relu:
            y = torch.relu_(y)
        ctx.save_for_backward(features, idx, y, pre if norm else y, stats if norm else torch.empty(0, device=w.device))
        ctx.norm, ctx.relu, ctx.k = norm, relu, k
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        features, idx, y, pre, stats = ctx.saved_tensors
        dy = dy.contiguous()
        n, d = y.shape
        if ctx.norm:
            d_pre = torch.empty_like(dy)
            check(_lib.load().sgcn_ln_act_bwd(ptr(pre), pre.stride(0), ptr(y), y.stride(0), ptr(dy), dy.stride(0), n, None,
                                              d, None, ptr(stats), 1 if ctx.relu else 0, ptr(d_pre), d_pre.stride(0),
                                              None, None, stream_ptr()))
        else:
            d_pre = dy * (y > 0) if ctx.relu else dy
        x = ops.gather_rows(features, idx)[:, :ctx.k]
        return torch.mm(x.t(), d_pre), None, None, None, None


# ---- layers -------------------------------------------------------------------------------------------
class Parameter:
    """A trainable tensor with Adam slots (kept outside torch.optim: the update is the library's kernel)."""

    def __init__(self, value, weight_decay=0.0):
        self.data = value.detach().clone().requires_grad_(True)
        self.m = torch.zeros_like(self.data)
        self.v = torch.zeros_like(self.data)
        self.weight_decay = float(weight_decay)


class Dense(Layer):
    """gcn/layers.py:100-138: act(MyLayerNorm(x @ W)); no bias; norm has fixed unit scale / zero offset."""

    def __init__(self, input_dim, output_dim, placeholders=None, sparse_inputs=False, act="relu", norm=True,
                 bias=False, featureless=False, weights=None, rng=None, device="cuda", **kwargs):
        super().__init__(**kwargs)
        if bias or featureless:
            raise NotImplementedError("the reference never builds Dense with bias / featureless")
        self.act, self.norm, self.sparse_inputs = act, norm, sparse_inputs
        w = weights if weights is not None else glorot((input_dim, output_dim), rng or np.random.RandomState(0))
        self.vars = {"weights": Parameter(torch.as_tensor(w, dtype=torch.float32, device=device))}

    def _call(self, inputs):
        out = dot(inputs, self.vars["weights"].data, sparse=self.sparse_inputs)
        relu = self.act == "relu"
        if self.norm:
            return layer_norm_act(out, None, None, 1e-9, relu)
        return torch.relu(out) if relu else out

    def forward_gathered(self, features, idx):
        """``self(features[idx])`` without materialising the gathered rows: one tensor-core kernel with the gather
        as its A-operand load (128 output columns, 16-byte aligned feature rows)."""
        if self.sparse_inputs or self.vars["weights"].data.shape[1] != 128:
            raise ValueError("the fused layer needs dense inputs and 128 output columns")
        return _GatheredDenseFn.apply(self.vars["weights"].data, features, idx, self.norm, self.act == "relu")


class AugmentedDropoutDense(Layer):
    """gcn/layers.py:365-412: the two-stream (h, mu) dense layer of CVD: dropout on h only, shared W,
    MyLayerNorm2 with learned offset / scale on both, relu, stop_gradient(mu)."""

    def __init__(self, keep_prob, input_dim, output_dim, sparse_inputs=False, act="relu", norm=True, weights=None,
                 rng=None, device="cuda", dropout_state=None, **kwargs):
        super().__init__(**kwargs)
        self.keep_prob, self.act, self.norm, self.sparse_inputs = keep_prob, act, norm, sparse_inputs
        self.dropout_state, self.mask = dropout_state, None
        w = weights if weights is not None else glorot((input_dim, output_dim), rng or np.random.RandomState(0))
        self.vars = {"weights": Parameter(torch.as_tensor(w, dtype=torch.float32, device=device))}
        if norm:
            self.vars["offset"] = Parameter(torch.zeros(output_dim, dtype=torch.float32, device=device))
            self.vars["scale"] = Parameter(torch.ones(output_dim, dtype=torch.float32, device=device))

    def _call(self, inputs):
        x, mu = inputs if isinstance(inputs, tuple) else (inputs, inputs)
        w = self.vars["weights"].data
        if isinstance(x, SparseRows):          # sparse_dropout: drop stored values (layers.py:23-28)
            keep = self.keep_prob
            vals = dropout(x.vals[:, None], keep, None if self.mask is None else self.mask.reshape(-1, 1),
                           self.dropout_state)[:, 0]
            xs = SparseRows(x.rowptr, x.cols, vals.contiguous(), x.n_rows, x.n_cols)
            hx = _SparseMMFn.apply(w, xs.rowptr, xs.cols, xs.vals, xs.n_rows)
            hm = _SparseMMFn.apply(w, mu.rowptr, mu.cols, mu.vals, mu.n_rows)
        else:
            hx = torch.mm(dropout(x, self.keep_prob, self.mask, self.dropout_state), w)
            hm = torch.mm(mu, w)
        relu = self.act == "relu"
        if self.norm:
            off, sc = self.vars["offset"].data, self.vars["scale"].data
            hx = layer_norm_act(hx, sc, off, 1e-9, relu)
            hm = layer_norm_act(hm, sc, off, 1e-9, relu)
        elif relu:
            hx, hm = torch.relu(hx), torch.relu(hm)
        return hx, hm.detach()                 # tf.stop_gradient(mu), layers.py:412


class Dropout(Layer):
    """gcn/layers.py:415-433 (the det-dropout sampling branch draws Gaussian noise and is not built)."""

    def __init__(self, keep_prob, cvd, dropout_state=None, **kwargs):
        super().__init__(**kwargs)
        self.keep_prob, self.cvd, self.dropout_state, self.mask = keep_prob, cvd, dropout_state, None

    def _call(self, inputs):
        if self.cvd and isinstance(inputs, tuple):
            return dropout(inputs[0], self.keep_prob, self.mask, self.dropout_state)
        if isinstance(inputs, tuple):
            raise NotImplementedError("Dropout on a det-dropout (mu, var) pair samples with tf.random_normal")
        if isinstance(inputs, SparseRows):
            vals = dropout(inputs.vals[:, None], self.keep_prob,
                           None if self.mask is None else self.mask.reshape(-1, 1), self.dropout_state)[:, 0]
            return SparseRows(inputs.rowptr, inputs.cols, vals.contiguous(), inputs.n_rows, inputs.n_cols)
        return dropout(inputs, self.keep_prob, self.mask, self.dropout_state)


# ---- data-parallel gradient exchange --------------------------------------------------------------------
def allreduce_gradients(params, group=None):
    """the dense weights are replicated on every rank of a row-range sharded run; their
    gradients (<= 1204x128 + 256x128 + 128x41 floats ~ 0.8 MB at Reddit shape) are AVERAGED over the ranks
    -- every rank's loss is a mean over its own equally sized batch, so the average is the gradient of the
    mean over the global batch -- in ONE flattened all-reduce (NCCL on GPUs; any torch.distributed backend).
    A parameter that got no gradient on some rank contributes zeros there.  No-op without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    params = list(params)
    grads = [p.data.grad if p.data.grad is not None else torch.zeros_like(p.data) for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        p.data.grad = flat[off:off + n].view_as(g).clone()
        off += n


# ---- optimiser ------------------------------------------------------------------------------------------
class Adam:
    """tf.train.AdamOptimizer(learning_rate, beta1, beta2) (gcn/models.py:50-51): epsilon = 1e-8 outside
    the square root, bias correction folded into the step size.  ``weight_decay * l2_loss(var)`` of the
    first layer (models.py:68-74) enters through each Parameter's ``weight_decay``."""

    def __init__(self, params, learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.params, self.lr, self.b1, self.b2, self.eps, self.t = list(params), learning_rate, beta1, beta2, epsilon, 0

    def zero_grad(self):
        for p in self.params:
            p.data.grad = None

    def step(self, group=None):
        """One update; in a multi-process run the gradients are averaged over ``group`` first."""
        allreduce_gradients(self.params, group)
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        lib = _lib.load()
        for p in self.params:
            g = p.data.grad
            if g is None:
                continue
            g = g.contiguous()
            with torch.no_grad():
                check(lib.sgcn_adam_step(ptr(p.data), ptr(g), ptr(p.m), ptr(p.v), p.data.numel(), lr_t, self.b1,
                                         self.b2, self.eps, p.weight_decay, stream_ptr()))


# ---- the pre-processed two-layer model (what every BASELINE config trains) ---------------------------
class PPModel:
    """``GCN(L=2, preprocess=True)`` of gcn/models.py:223-337 after pre-processing (one aggregator left):

        input rows of fields[0] -> num_fc_layers x dense -> aggregator -> num_fc_layers x dense -> logits

    with the reference's choice of layer per position: ``AugmentedDropoutDense`` before the aggregator when
    cvd, else ``Dropout`` + ``Dense``; after the last aggregator always ``Dropout`` + ``Dense`` (l+1 == L),
    the last one without norm / activation.  ``weight_decay`` applies to the first layer's variables."""

    def __init__(self, input_dim, hidden, n_classes, num_fc_layers=1, normalization="graphsage", cvd=False,
                 layer_norm=True, dropout=0.0, weight_decay=5e-4, sparse_inputs=False, multitask=False, seed=0,
                 device="cuda"):
        rng = np.random.RandomState(seed)
        keep = 1.0 - dropout
        self.cvd, self.multitask, self.normalization = cvd, multitask, normalization
        self.drop_state = DropoutState(seed + 1)
        dim_s = 1 if normalization == "gcn" else 2
        self.pre, self.post = [], []
        for l in range(num_fc_layers):
            d_in = input_dim if l == 0 else hidden
            sp_in = sparse_inputs and l == 0
            if cvd:
                self.pre.append(AugmentedDropoutDense(keep, d_in, hidden, sparse_inputs=sp_in, norm=layer_norm, rng=rng,
                                                      device=device, dropout_state=self.drop_state, name="dense%d" % l))
            else:
                self.pre.append(Dropout(keep, cvd, dropout_state=self.drop_state))
                self.pre.append(Dense(d_in, hidden, sparse_inputs=sp_in, norm=layer_norm, rng=rng, device=device,
                                      name="dense%d" % l))
        for l2 in range(num_fc_layers):
            last = l2 + 1 == num_fc_layers
            d_in = hidden * dim_s if l2 == 0 else hidden
            self.post.append(Dropout(keep, cvd, dropout_state=self.drop_state))
            self.post.append(Dense(d_in, n_classes if last else hidden, act=None if last else "relu",
                                   norm=False if last else layer_norm, rng=rng, device=device,
                                   name="dense%d" % (num_fc_layers + l2)))
        first = next(l for l in self.pre if getattr(l, "vars", None))
        for p in first.vars.values():
            p.weight_decay = weight_decay
        self.weight_decay = weight_decay
        self.first = first

    def parameters(self):
        return [p for layer in self.pre + self.post for p in getattr(layer, "vars", {}).values()]

    def forward(self, inputs, aggregator):
        h = (inputs, inputs) if self.cvd and not isinstance(inputs, tuple) else inputs
        for layer in self.pre:
            h = layer(h)
        h = aggregator(h)
        for layer in self.post:
            h = layer(h)
        return h

    def loss(self, logits, labels):
        """cross entropy only: the ``weight_decay * l2_loss`` term is applied inside the Adam kernel
        (its value is reported by ``l2_term``)."""
        return cross_entropy(logits, labels, self.multitask)

    def l2_term(self):
        return sum(self.weight_decay * 0.5 * float((p.data.detach() ** 2).sum()) for p in self.first.vars.values())
