"""CPU: the F1 restatement of the evaluation path against scikit-learn, the library the reference calls
(gcn/utils.py:521-529)."""
import numpy as np
import pytest

sk = pytest.importorskip("sklearn.metrics")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_calc_f1_matches_sklearn_multiclass(seed):
    from stochastic_gcn_b200.evaluate import calc_f1
    rng = np.random.RandomState(seed)
    n, c = 500, 7
    y_true = np.eye(c)[rng.randint(0, c - 2, n)]            # two classes never occur in y_true
    y_pred = rng.rand(n, c) + 0.8 * y_true
    y_pred[:, c - 1] = -1.0                                   # ... and one never predicted either
    micro, macro = calc_f1(y_pred, y_true, False)
    t, p = y_true.argmax(1), y_pred.argmax(1)
    assert abs(micro - sk.f1_score(t, p, average="micro")) < 1e-12
    assert abs(macro - sk.f1_score(t, p, average="macro")) < 1e-12


@pytest.mark.parametrize("seed", [0, 1])
def test_calc_f1_matches_sklearn_multitask(seed):
    from stochastic_gcn_b200.evaluate import calc_f1
    rng = np.random.RandomState(seed)
    n, c = 400, 12
    y_true = (rng.rand(n, c) < 0.2).astype(np.float64)
    y_true[:, 3] = 0                                          # a label that never occurs
    y_pred = np.clip(0.6 * y_true + 0.5 * rng.rand(n, c), 0, 1)
    micro, macro = calc_f1(y_pred.copy(), y_true, True)
    yp = (y_pred > 0.5).astype(int)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert abs(micro - sk.f1_score(y_true.astype(int), yp, average="micro")) < 1e-12
        assert abs(macro - sk.f1_score(y_true.astype(int), yp, average="macro")) < 1e-12
