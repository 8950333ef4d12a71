"""Row-dense tcgen05 scan kernel (csrc/scan_rd.cuh: motif = TMEM lane, pooled softmax in registers, rows staged in shared memory and
stored by the TMA engine) against the NumPy oracle: `motifHitProbs` with the pooling factors the kernel is built for, both
getHitProbs branches (reference secomo/convRBM.py:507-514), ragged shapes (hidden length not a multiple of 8 / of the 4-block step,
K not a multiple of 128, more than one motif chunk, fewer sequences than CTAs and many more), and — forced by CRBM_SCAN_KERNEL=rows
in a subprocess — the un-pooled scan suite of test_gpu_scan_tc.py on the same kernel."""
import os
import sys

import numpy as np
import pytest

from oracle import crbm_oracle as O

pytestmark = pytest.mark.gpu

TOL = dict(rtol=1e-5, atol=1e-5)          # the reference's own tolerance (tests/testcrbm.py:191-198)


def _model(K, M, P, ds, rs, wscale=0.5):
    import crbm_b200
    m = crbm_b200.CRBM(K, M, doublestranded=ds, pooling=P, seed=1)
    W = (wscale * rs.randn(K, 1, 4, M)).astype(np.float32)
    b = (-2 + rs.randn(1, K)).astype(np.float32)
    m.motifs.set_value(W)
    m.bias.set_value(b)
    return m, W.astype(np.float64), b.astype(np.float64)


@pytest.mark.parametrize('K,M,P,L,N', [(128, 20, 5, 999, 5), (100, 15, 2, 200, 23), (33, 9, 4, 76, 37), (16, 25, 10, 524, 3),
                                       (260, 32, 5, 156, 7), (7, 1, 10, 130, 9), (40, 17, 5, 21, 301), (130, 12, 2, 267, 11),
                                       (5, 8, 4, 1031, 2)])
def test_pooled_hit_probs_match_oracle(K, M, P, L, N):
    import crbm_b200
    assert (L - M + 1) % P == 0
    rs = np.random.RandomState(K * 1000 + M + P)
    codes = rs.randint(0, 4, size=(N, L))
    D = crbm_b200.codesToOneHot(codes)
    D64 = D.astype(np.float64)
    for ds in (True, False):
        m, W, b = _model(K, M, P, ds, rs)
        want = O.hit_probs(D64, W, b, P, ds)
        got = m.motifHitProbs(D)
        assert got.shape == (N, K, 1, L - M + 1) and got.dtype == np.float32
        np.testing.assert_allclose(got, want, **TOL)


def test_pooled_large_weights_stay_finite():
    """Activities of +-100: the exponentials are clamped, the probabilities stay in [0, 1] and match the oracle."""
    import crbm_b200
    rs = np.random.RandomState(12)
    K, M, P, L, N = 64, 12, 5, 411, 6
    D = crbm_b200.codesToOneHot(rs.randint(0, 4, size=(N, L)))
    m, W, b = _model(K, M, P, True, rs, wscale=4.0)
    got = m.motifHitProbs(D)
    assert np.isfinite(got).all() and got.min() >= 0.0 and got.max() <= 1.0 + 1e-6
    np.testing.assert_allclose(got, O.hit_probs(D.astype(np.float64), W, b, P, True), rtol=1e-4, atol=1e-5)


def test_pooled_many_sequences_same_window_property():
    """Full-size property (no oracle): every sequence is the same, so every output row block must be bit-identical whichever CTA,
    step and staging slot produced it; the first one is checked against the oracle."""
    import crbm_b200
    rs = np.random.RandomState(19)
    K, M, P, L = 128, 20, 5, 999
    base = rs.randint(0, 4, size=(1, L))
    N = 1500
    m, W, b = _model(K, M, P, True, rs)
    seqs = crbm_b200.PackedSequences.from_onehot(crbm_b200.codesToOneHot(np.repeat(base, N, axis=0)))
    got = m.motifHitProbs(seqs)
    assert np.array_equal(got, np.broadcast_to(got[:1], got.shape))
    np.testing.assert_allclose(got[:1], O.hit_probs(crbm_b200.codesToOneHot(base).astype(np.float64), W, b, P, True), **TOL)


@pytest.mark.parametrize('kernel', ['rows', 'pos', 'rows-bulk'])
def test_unpooled_scan_suite_on_either_kernel(kernel):
    """The un-pooled scan picks the row-dense or the position-major kernel by the model shape (train_tc_impl.cuh: rowdense_wanted);
    CRBM_SCAN_KERNEL=rows | pos (read once per process: subprocess) forces one of them for every shape, and the scan suite of
    test_gpu_scan_tc.py (activities and probabilities, all strand modes) must pass on both."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # 'rows-bulk': the row-dense kernel without the tensor map (per-row bulk copies, the path taken when cuTensorMapEncodeTiled is
    # not available)
    env = dict(os.environ, CRBM_SCAN_KERNEL=kernel.split('-')[0])
    if kernel.endswith('-bulk'):
        env['CRBM_RD_STORE'] = 'bulk'
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu', os.path.join(root, 'tests', 'test_gpu_scan_tc.py')],
                       capture_output=True, text=True, timeout=900, cwd=root, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
