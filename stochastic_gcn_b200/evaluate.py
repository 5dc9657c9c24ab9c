"""Evaluation path: the reference's ``evaluate`` / ``Test`` (gcn/train.py:133-160,
320-341) over the device sampler and the library's kernels.

Per batch of ``test_batch_size`` ids: ``eval_sch.batch`` (full adjacency, ``test_degree``, ``test_cv``)
-> input rows -> the model with dropout off (the ``dropout`` placeholder defaults to 0) -> loss, accuracy,
prediction -> ``test_op`` = the history write-back alone (gcn/models.py:186-194).  Batch results are
weighted by batch size (train.py:149-151); F1 comes from all predictions at the end (``calc_f1``,
gcn/utils.py:521-529).  With ``test_cv`` the reference runs ``Test()`` ``num_layers + 1`` times and, each
time, evaluates the REMAINING nodes too, which is what warms the test model's own history tables up
(train.py:320-341); ``Evaluator.test`` does the same.
"""
import contextlib

import numpy as np
import torch

from . import nn, ops
from .layers import DeviceAdj, FullNeighbours, PlainAggregator, VRAggregator


def _f1(tp, fp, fn):
    den = 2.0 * tp + fp + fn
    return np.where(den > 0, 2.0 * tp / np.maximum(den, 1e-300), 0.0)


def calc_f1(y_pred, y_true, multitask):
    """(micro, macro) F1 as ``sklearn.metrics.f1_score`` computes them for the reference
    (gcn/utils.py:521-529): multitask -> multilabel indicators thresholded at 0.5, macro over ALL label
    columns; otherwise argmax classes, macro over the classes present in y_true or y_pred."""
    y_pred, y_true = np.asarray(y_pred), np.asarray(y_true)
    if multitask:
        p = y_pred > 0.5
        t = y_true > 0.5
        tp = (p & t).sum(0).astype(np.float64)
        fp = (p & ~t).sum(0).astype(np.float64)
        fn = (~p & t).sum(0).astype(np.float64)
        micro = float(_f1(tp.sum(), fp.sum(), fn.sum()))
        return micro, float(_f1(tp, fp, fn).mean())
    t = np.argmax(y_true, axis=1)
    p = np.argmax(y_pred, axis=1)
    classes = np.union1d(t, p)
    tp = np.array([np.sum((p == c) & (t == c)) for c in classes], dtype=np.float64)
    fp = np.array([np.sum((p == c) & (t != c)) for c in classes], dtype=np.float64)
    fn = np.array([np.sum((p != c) & (t == c)) for c in classes], dtype=np.float64)
    return float(_f1(tp.sum(), fp.sum(), fn.sum())), float(_f1(tp, fp, fn).mean())


@contextlib.contextmanager
def dropout_off(model):
    """keep_prob = 1 on every dropout site for the duration (the reference feeds dropout = 0 at test time)."""
    sites = [l for l in model.pre + model.post if hasattr(l, "keep_prob")]
    saved = [l.keep_prob for l in sites]
    for l in sites:
        l.keep_prob = 1.0
    try:
        yield
    finally:
        for l, k in zip(sites, saved):
            l.keep_prob = k


class Evaluator:
    """``evaluate(data)`` of gcn/train.py:133-160 for a pre-processed two-layer model.

    sampler: ``DeviceSampler`` over the FULL adjacency (the reference's ``eval_sch``); features: the
    [N, F'] device matrix of model inputs (``test_feats`` stacked as gcn/models.py:235-239 does);
    labels: [N, C] device tensor; history: list of [N, hidden] tables of the TEST model (its own, separate
    from the training model's; empty for plain neighbour sampling)."""

    def __init__(self, model, sampler, features, labels, history, degree, batch_size=1000, cv=False, cvd=False,
                 multitask=False):
        self.model, self.sampler, self.features, self.labels = model, sampler, features, labels
        self.history, self.degree, self.batch_size = list(history), int(degree), int(batch_size)
        self.cv, self.cvd, self.multitask = bool(cv), bool(cvd), bool(multitask)

    def _aggregator(self, n_out):
        s = self.sampler
        z = s.sizes()
        field = s.view("field")[:z.n_in]
        adj = DeviceAdj(s.view("rowptr_s"), s.view("edg_t"), s.view("edg_w"), n_out, z.n_in, tgt=s.view("tgt"))
        norm = self.model.normalization
        if not self.cv:
            return PlainAggregator(adj, normalization=norm), field
        full = FullNeighbours.in_place(field[:n_out], s.view("rowptr_f"), s.view("adj_p"), s.view("adj_i"),
                                       s.view("adj_w"))
        scale = s.view("scales") if self.cvd else None
        return VRAggregator(adj, full, None, field, None, self.history, scale, self.cvd, normalization=norm), field

    def run_batch(self, ids):
        """One ``test_model.run_one_step``: (loss, accuracy, predictions [n, C] on the host)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        n = int(ids.shape[0])
        self.sampler.start_batch(ids)
        self.sampler.expand(self.degree, materialize_full=False)
        aggr, field = self._aggregator(n)
        labels = self.labels[torch.from_numpy(ids.astype(np.int64)).to(self.labels.device)]
        with torch.no_grad(), dropout_off(self.model):
            logits = self.model.forward(ops.gather_rows(self.features, field), aggr)
            loss = float(self.model.loss(logits, labels)) + self.model.l2_term()       # models.py:68-83
            if self.multitask:
                acc = float(((logits > 0) == (labels > 0.5)).float().mean())           # models.py:85-95
                pred = torch.sigmoid(logits)
            else:
                acc = float((logits.argmax(1) == labels.argmax(1)).float().mean())
                pred = torch.softmax(logits, dim=1)
            if self.cv:
                aggr.write_back()                                                       # test_op
        return loss, acc, pred.cpu().numpy()

    def evaluate(self, data):
        """-> (loss, accuracy, micro F1, macro F1), batch results weighted by batch size."""
        data = np.ascontiguousarray(data, dtype=np.int32)
        total_loss = total_acc = 0.0
        preds, labs = [], []
        n = len(data)
        for start in range(0, n, self.batch_size):
            batch = data[start:start + self.batch_size]
            los, acc, prd = self.run_batch(batch)
            total_loss += los * len(batch)
            total_acc += acc * len(batch)
            preds.append(prd)
            labs.append(self.labels[torch.from_numpy(batch.astype(np.int64)).to(self.labels.device)].cpu().numpy())
        micro, macro = calc_f1(np.vstack(preds), np.vstack(labs), self.multitask)
        return total_loss / n, total_acc / n, micro, macro

    def test(self, test_data, num_data, num_layers=2):
        """``Test()`` repeated as gcn/train.py:339-341 does: with cv, ``num_layers + 1`` passes, each followed
        by an evaluation of every other node (the history warm-up); returns the last pass's result."""
        test_data = np.ascontiguousarray(test_data, dtype=np.int32)
        remaining = np.setdiff1d(np.arange(num_data, dtype=np.int32), test_data).astype(np.int32)
        result = None
        for _ in range(num_layers + 1 if self.cv else 1):
            result = self.evaluate(test_data)
            if self.cv and len(remaining):
                self.evaluate(remaining)
        return result
