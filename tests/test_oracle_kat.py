"""Known-answer tests for the oracle, restating the control computations the reference's own tests use
(reference tests/testcrbm.py:182-198 bottom-up; :202-220 top-down; :340,363 all-zero hidden -> 0.25;
:283-287 one base per sampled position). Tolerance is the reference's: rtol=atol=1e-5. CPU-only."""
import numpy as np

from oracle import crbm_oracle as O
from tests._golden import onehot

RS = np.random.RandomState(7)
TOL = dict(rtol=1e-5, atol=1e-5)


def _model(K=10, M=5):
    W = RS.randn(K, 1, 4, M)
    b = RS.randn(1, K) - 2.0
    c = 0.1 * RS.randn(1, 4)
    return W, b, c


def _control_bottom_up(D, W, b, flip):
    Wc = W[:, :, ::-1, ::-1] if flip else W
    N, L = D.shape[0], D.shape[3]
    K, M = W.shape[0], W.shape[3]
    out = np.zeros((N, K, 1, L - M + 1))
    for n in range(N):
        for k in range(K):
            for j in range(L - M + 1):
                out[n, k, 0, j] = (Wc[k, 0] * D[n, 0, :, j:j + M]).sum() + b[0, k]
    return out


def _control_top_down(W, c, h, hp=None):
    N, K, _, Lh = h.shape
    M = W.shape[3]
    out = np.zeros((N, 1, 4, Lh + M - 1)) + c[0][None, None, :, None]
    for n in range(N):
        for k in range(K):
            for j in range(Lh):
                out[n, 0, :, j:j + M] += W[k, 0] * h[n, k, 0, j]
                if hp is not None:
                    out[n, 0, :, j:j + M] += W[k, 0, ::-1, ::-1] * hp[n, k, 0, j]
    return out


def test_bottom_up_both_orientations():
    W, b, c = _model()
    D = onehot(RS.randint(0, 4, size=(11, 60)), np.float64)
    for flip in (False, True):
        ctrl = _control_bottom_up(D, W, b, flip)
        x = O.bottom_up_activity(D, W, b, flip)
        assert x.shape == (11, 10, 1, 56)
        np.testing.assert_allclose(x, ctrl, **TOL)
        np.testing.assert_allclose(O.bottom_up_probability(x, 1), 1.0 / (1.0 + np.exp(-ctrl)), **TOL)


def test_gather_form_equals_dense_form():
    """One-hot input: x = b_k + sum_m W[k, base(j+m), m]; rc strand: W[k, 3-base(j+m), M-1-m]."""
    W, b, c = _model(K=6, M=7)
    codes = RS.randint(0, 4, size=(3, 40))
    D = onehot(codes, np.float64)
    xf = O.bottom_up_activity(D, W, b, False)
    xr = O.bottom_up_activity(D, W, b, True)
    M = 7
    for n in range(3):
        for j in range(40 - M + 1):
            gf = b[0] + sum(W[:, 0, codes[n, j + m], m] for m in range(M))
            gr = b[0] + sum(W[:, 0, 3 - codes[n, j + m], M - 1 - m] for m in range(M))
            np.testing.assert_allclose(xf[n, :, 0, j], gf, rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(xr[n, :, 0, j], gr, rtol=1e-12, atol=1e-12)


def test_top_down_zeros_ones_random():
    W, b, c = _model()
    Lh = 196
    zeros = np.zeros((11, 10, 1, Lh))
    p = O.top_down_probability(O.top_down_activity(zeros, zeros, W, np.zeros((1, 4))))
    assert p.shape == (11, 1, 4, 200)
    np.testing.assert_allclose(p, 0.25, **TOL)
    ones = np.ones((3, 10, 1, Lh))
    rnd = RS.binomial(1, 0.1, size=(3, 10, 1, Lh)).astype(np.float64)
    rnd2 = RS.binomial(1, 0.1, size=(3, 10, 1, Lh)).astype(np.float64)
    for h, hp in ((ones, None), (ones, ones), (rnd, None), (rnd, rnd2)):
        ctrl = _control_top_down(W, c, h, hp)
        a = O.top_down_activity(h, hp, W, c)
        np.testing.assert_allclose(a, ctrl, **TOL)
        e = np.exp(ctrl)
        np.testing.assert_allclose(O.top_down_probability(a), e / e.sum(axis=2, keepdims=True), **TOL)


def test_visible_sample_is_one_hot_and_hidden_sample_at_most_one_per_group():
    W, b, c = _model()
    D = onehot(RS.randint(0, 4, size=(5, 64)), np.float64)
    P = 4
    p = O.bottom_up_probability(O.bottom_up_activity(D, W, b + 2.0, False), P)
    h = O.bottom_up_sample(p, P, RS.rand(p.size // P))
    assert set(np.unique(h)) <= {0.0, 1.0}
    assert h.reshape(5, 10, 1, -1, P).sum(axis=4).max() <= 1.0
    pv = O.top_down_probability(O.top_down_activity(h, None, W, c))
    v = O.top_down_sample(pv, RS.rand(pv.shape[0] * pv.shape[3]))
    assert v.shape == pv.shape
    np.testing.assert_array_equal(v.sum(axis=2), 1.0)


def test_pool_normaliser_is_P_plus_sum_exp():
    x = RS.randn(2, 3, 1, 12)
    p = O.bottom_up_probability(x, 3)
    g = x.reshape(2, 3, 1, 4, 3)
    np.testing.assert_allclose(p.reshape(2, 3, 1, 4, 3), np.exp(g) / (3.0 + np.exp(g).sum(-1, keepdims=True)), rtol=1e-12)


def test_sparsity_gradient_matches_finite_differences():
    W, b, c = _model(K=3, M=4)
    D = onehot(RS.randint(0, 4, size=(4, 23)), np.float64)
    P, rho = 2, 0.01

    def R(Wx, bx):
        p = O.bottom_up_probability(O.bottom_up_activity(D, Wx, bx, False), P).mean(axis=(0, 2, 3))
        return -np.mean(rho * np.log(p) + (1 - rho) * np.log(1 - p))
    gW, gb = O.sparsity_gradient(D, W, b, P, rho)
    eps = 1e-6
    for idx in [(0, 0, 1, 2), (2, 0, 3, 0), (1, 0, 0, 3)]:
        Wp, Wm = W.copy(), W.copy()
        Wp[idx] += eps
        Wm[idx] -= eps
        np.testing.assert_allclose(gW[idx], (R(Wp, b) - R(Wm, b)) / (2 * eps), rtol=1e-5, atol=1e-9)
    for k in range(3):
        bp, bm = b.copy(), b.copy()
        bp[0, k] += eps
        bm[0, k] -= eps
        np.testing.assert_allclose(gb[0, k], (R(W, bp) - R(W, bm)) / (2 * eps), rtol=1e-5, atol=1e-9)


def test_fit_helpers():
    assert O.truncate_for_pooling(1000, 20, 5) == 999
    assert O.truncate_for_pooling(200, 15, 1) == 200
    assert O.iterate_batch_indices(45, 20) == [[0, 20], [20, 40], [40, 45]]
