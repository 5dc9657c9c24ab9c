"""TEST INFRASTRUCTURE ONLY -- float64 restatement (torch CPU autograd) of the dense layers, loss and
optimiser the reference builds from TensorFlow-1.x ops (parity unpinned: TensorFlow is absent and
the reference holds no test at this boundary; formulas follow the cited lines).

  layer_norm_act   MyLayerNorm / MyLayerNorm2 + act           gcn/layers.py:87-97,130-138,404-412
  dropout          tf.nn.dropout with an injected keep mask   gcn/layers.py:396,415-433
  cross_entropy    softmax / sigmoid xent, reduce_mean        gcn/models.py:76-83
  adam_step        tf.train.AdamOptimizer                     gcn/models.py:50-51
  PPReference      GCN(L=2, preprocess=True) layer stack      gcn/models.py:256-337, layers.py:282-362
"""
import math

import numpy as np
import torch

F8 = torch.float64


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=F8)


def layer_norm_act(x, scale=None, offset=None, eps=1e-9, relu=True):
    mean = x.mean(dim=1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=1, keepdim=True)          # tf.nn.moments: biased
    y = (x - mean) * torch.rsqrt(var + eps)                      # tf.nn.batch_normalization
    if scale is not None:
        y = y * scale
    if offset is not None:
        y = y + offset
    return torch.relu(y) if relu else y


def dropout(x, keep_prob, mask):
    """mask: 0/1 keep mask of x's shape (None = keep everything, no scaling when keep_prob == 1)"""
    if mask is None:
        return x
    return x * t64(mask) / keep_prob


def cross_entropy(logits, labels, multitask=False):
    if multitask:
        z, t = logits, labels
        return (torch.clamp(z, min=0) - z * t + torch.log1p(torch.exp(-z.abs()))).mean()
    return -(labels * torch.log_softmax(logits, dim=1)).sum(dim=1).mean()


def adam_step(p, g, m, v, t, lr=0.01, b1=0.9, b2=0.999, eps=1e-8, weight_decay=0.0):
    """NumPy float64, in place; t = step number starting at 1."""
    g = g + weight_decay * p
    m[...] = b1 * m + (1 - b1) * g
    v[...] = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    p -= lr_t * m / (np.sqrt(v) + eps)
    return p


def dense_adj(triple):
    idx, val, shape = triple
    a = torch.zeros((int(shape[0]), int(shape[1])), dtype=F8)
    idx = np.asarray(idx).reshape(-1, 2)
    a.index_put_((torch.from_numpy(idx[:, 0]).long(), torch.from_numpy(idx[:, 1]).long()), t64(val), accumulate=True)
    return a


class PPReference:
    """The pre-processed two-layer model in float64 with dense adjacency matrices (small graphs only).
    weights: list of NumPy arrays in layer order: for every dense layer W (and, for AugmentedDropoutDense with
    norm, offset and scale after it)."""

    def __init__(self, weights, num_fc_layers, graphsage, cvd, layer_norm, weight_decay, multitask=False):
        self.w = [t64(w).clone().requires_grad_(True) for w in weights]
        self.nfc, self.graphsage, self.cvd, self.ln, self.wd, self.multitask = (num_fc_layers, graphsage, cvd,
                                                                                 layer_norm, weight_decay, multitask)

    def forward(self, x, adj, fadj, ifield, ffield, history, scale, keep_prob=1.0, masks=None):
        """x: [n_in, F] input rows (dense).  masks: list of keep masks, one per dropout site in order."""
        masks = list(masks) if masks is not None else None
        nxt = (lambda: masks.pop(0)) if masks is not None else (lambda: None)
        A, Af = dense_adj(adj), (dense_adj(fadj) if fadj is not None else None)
        n_out = A.shape[0]
        H = t64(history) if history is not None else None
        wi = 0
        x = t64(x)
        h, mu = x, x
        first_vars = None
        for l in range(self.nfc):
            if self.cvd:
                W = self.w[wi]; wi += 1
                hx = dropout(h, keep_prob, nxt()) @ W
                hm = mu @ W
                if self.ln:
                    off, sc = self.w[wi], self.w[wi + 1]; wi += 2
                    hx, hm = layer_norm_act(hx, sc, off), layer_norm_act(hm, sc, off)
                    if first_vars is None:
                        first_vars = [W, off, sc]
                else:
                    hx, hm = torch.relu(hx), torch.relu(hm)
                    if first_vars is None:
                        first_vars = [W]
                h, mu = hx, hm.detach()
            else:
                W = self.w[wi]; wi += 1
                if first_vars is None:
                    first_vars = [W]
                h = dropout(h, keep_prob, nxt()) @ W
                h = layer_norm_act(h) if self.ln else torch.relu(h)
        # aggregator (gcn/layers.py:223-257, 298-319, 350-362)
        if H is None:
            nb = A @ h
            out = torch.cat((h[:n_out], nb), 1) if self.graphsage else nb
            self.new_history = None
        elif self.cvd:
            mu_nb = A @ (mu - H[ifield]) + Af @ H[ffield]
            h_nb = (A @ (h - mu)) * t64(scale)[:, None] + mu_nb
            out = torch.cat((h[:n_out], h_nb), 1) if self.graphsage else h_nb      # Dropout keeps h only
            self.new_history = mu.detach().numpy()
        else:
            nb = A @ h - A @ H[ifield] + Af @ H[ffield]
            out = torch.cat((h[:n_out], nb), 1) if self.graphsage else nb
            self.new_history = h.detach().numpy()
        h = out
        for l2 in range(self.nfc):
            last = l2 + 1 == self.nfc
            W = self.w[wi]; wi += 1
            h = dropout(h, keep_prob, nxt()) @ W
            if not last:
                h = layer_norm_act(h) if self.ln else torch.relu(h)
        self.first_vars = first_vars
        return h

    def loss(self, logits, labels):
        l2 = sum(self.wd * 0.5 * (v ** 2).sum() for v in self.first_vars)      # tf.nn.l2_loss = sum(v^2)/2
        return cross_entropy(logits, t64(labels), self.multitask) + l2
