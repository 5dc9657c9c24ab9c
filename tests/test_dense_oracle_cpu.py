"""CPU: the float64 dense-layer restatement (oracle/dense.py) against hand-computed values of the
TensorFlow formulas it cites (tf.nn.moments + batch_normalization, softmax / sigmoid cross entropy,
AdamOptimizer), so that the GPU parity tests compare against a checked checker."""
import numpy as np
import torch

from oracle import dense as od


def test_layer_norm_is_moments_then_batch_normalization():
    rng = np.random.RandomState(0)
    x = rng.randn(5, 7)
    sc, off = rng.rand(7) + 0.5, rng.randn(7)
    mean = x.mean(1, keepdims=True)
    var = x.var(1, keepdims=True)                      # biased, as tf.nn.moments
    want = np.maximum((x - mean) / np.sqrt(var + 1e-9) * sc + off, 0)
    got = od.layer_norm_act(od.t64(x), od.t64(sc), od.t64(off), 1e-9, True).numpy()
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_cross_entropy_values():
    z = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]])
    t = np.array([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0]])
    want = np.mean([np.log(np.exp(z[0]).sum()) - 3.0, np.log(3.0)])
    assert abs(float(od.cross_entropy(od.t64(z), od.t64(t))) - want) < 1e-12
    sig = np.mean(np.maximum(z, 0) - z * t + np.log1p(np.exp(-np.abs(z))))
    assert abs(float(od.cross_entropy(od.t64(z), od.t64(t), True)) - sig) < 1e-12


def test_adam_first_steps_closed_form():
    p, m, v = np.array([1.0, -2.0]), np.zeros(2), np.zeros(2)
    g = np.array([0.5, -0.25])
    od.adam_step(p, g, m, v, 1, lr=0.01)
    # t = 1: m = 0.1 g, v = 0.001 g^2, lr_t = lr sqrt(0.001) / 0.1  ->  step = lr * g / (|g| + eps sqrt(1000)) ~ lr sign(g)
    assert np.allclose(p, [1.0 - 0.01, -2.0 + 0.01], atol=1e-8)
    assert np.allclose(m, 0.1 * g) and np.allclose(v, 0.001 * g * g)


def test_reference_model_gradients_by_finite_differences():
    from tests.graphs_small import random_graph
    from tests.test_dense_gpu import sample
    n, f, hid, ncls = 120, 6, 5, 3
    g = random_graph(n, 8, 2)
    rng = np.random.RandomState(1)
    feats, labels = rng.randn(n, f), np.eye(ncls)[rng.randint(0, ncls, n)]
    ids = rng.choice(n, 20, replace=False).astype(np.int32)
    z, adj, fadj = sample(g, ids, 1)
    hist = rng.randn(n, hid)
    # CV (not CVD: there stop_gradient(mu) makes the reference's gradient differ from the true derivative on purpose)
    ws = [rng.randn(f, hid) * 0.3, rng.randn(2 * hid, ncls) * 0.3]

    def loss_of(ws_):
        ref = od.PPReference(ws_, 1, True, False, True, 5e-4)
        lg = ref.forward(feats[z["field"]], adj, fadj, z["field"], z["ffield"], hist, z["scales"])
        return ref, ref.loss(lg, labels[ids])

    ref, l0 = loss_of(ws)
    l0.backward()
    for wi, idx in ((0, (2, 3)), (0, (5, 0)), (1, (4, 1))):
        wp, wm = [w.copy() for w in ws], [w.copy() for w in ws]
        wp[wi][idx] += 1e-5
        wm[wi][idx] -= 1e-5
        fd = (float(loss_of(wp)[1].detach()) - float(loss_of(wm)[1].detach())) / 2e-5      # central difference
        assert abs(fd - float(ref.w[wi].grad[idx])) < 1e-6 * max(1.0, abs(fd))
