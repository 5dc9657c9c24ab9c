"""GPU: the evaluation path (gcn/train.py:133-160,320-341) -- device sampler over the full adjacency,
dropout off, loss / accuracy / predictions per batch, history write-back as the only side effect --
against the float64 model fed by the CPU oracle sampler with the same seed, batch after batch."""
import numpy as np
import pytest
import torch

from oracle import aggregators as agg
from oracle import dense as od
from oracle import native
from tests.graphs_small import random_graph

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("cv", [True, False])
def test_evaluate_matches_float64_model(cv):
    from stochastic_gcn_b200 import nn
    from stochastic_gcn_b200.evaluate import Evaluator, calc_f1
    from stochastic_gcn_b200.sampler import DeviceSampler
    n, f, hid, ncls, bs, degree, seed = 300, 24, 16, 4, 64, 2, 5
    g = random_graph(n, 12, 31)
    rng = np.random.RandomState(2)
    feats = rng.randn(n, f).astype(np.float32)
    labels = np.eye(ncls, dtype=np.float32)[rng.randint(0, ncls, n)]
    hist0 = (rng.randn(n, hid) * 0.3).astype(np.float32)
    data = rng.permutation(n)[:150].astype(np.int32)                 # 64 + 64 + 22: a ragged last batch

    model = nn.PPModel(f, hid, ncls, num_fc_layers=1, normalization="graphsage", cvd=False, layer_norm=True,
                       dropout=0.5, weight_decay=5e-4, seed=4)         # dropout 0.5 must be OFF at test time
    weights = [p.data.detach().cpu().numpy().copy() for p in model.parameters()]
    sampler = DeviceSampler(g.data, g.indices, g.indptr, L=1, cv=cv)
    sampler.seed(seed)
    hist_t = dev(hist0)
    ev = Evaluator(model, sampler, dev(feats), dev(labels), [hist_t] if cv else [], degree, batch_size=bs, cv=cv)
    got = ev.evaluate(data)

    osamp = native.OracleSampler(g.data, g.indices, g.indptr, cv=cv)
    osamp.seed(seed)
    hist = hist0.astype(np.float64).copy()
    tot_loss = tot_acc = 0.0
    preds = []
    for start in range(0, len(data), bs):
        ids = data[start:start + bs]
        osamp.start_batch(ids)
        osamp.expand(degree)
        z = osamp.snapshot()
        adj = (np.stack([z["edg_s"], z["edg_t"]], 1).astype(np.int32), z["edg_w"], (len(ids), len(z["field"])))
        fadj = None
        if cv:
            fadj = (np.stack([z["fedg_s"], z["fedg_t"]], 1).astype(np.int32), z["fedg_w"], (len(ids), len(z["ffield"])))
        ref = od.PPReference(weights, 1, True, False, True, 5e-4)
        logits = ref.forward(feats[z["field"]], adj, fadj, z["field"], z.get("ffield"), hist if cv else None,
                             z["scales"])
        loss = float(ref.loss(logits, labels[ids]).detach())
        lg = logits.detach().numpy()
        acc = float((lg.argmax(1) == labels[ids].argmax(1)).mean())
        e = np.exp(lg - lg.max(1, keepdims=True))
        preds.append(e / e.sum(1, keepdims=True))
        tot_loss += loss * len(ids)
        tot_acc += acc * len(ids)
        if cv:
            hist = agg.history_update(hist, z["field"], ref.new_history)
    want_loss, want_acc = tot_loss / len(data), tot_acc / len(data)
    micro, macro = calc_f1(np.vstack(preds), labels[data], False)
    assert abs(got[0] - want_loss) <= 1e-4 * abs(want_loss)
    assert abs(got[1] - want_acc) <= 1.0 / len(data) + 1e-9       # an argmax tie may flip one prediction
    assert abs(got[2] - micro) <= 2.0 / len(data) and abs(got[3] - macro) <= 0.02
    if cv:
        err = np.abs(hist_t.cpu().numpy().astype(np.float64) - hist).max() / np.abs(hist).max()
        assert err <= 1e-5, "history after the evaluation's write-backs: %.3e" % err
    # dropout sites are restored after the evaluation
    assert all(l.keep_prob == 0.5 for l in model.pre + model.post if hasattr(l, "keep_prob"))
