"""tcgen05 / TMEM scan kernel (csrc/scan_tc.cu, `crbm_v2h_dense` with pooling 1) against the NumPy oracle:
activities and probabilities for the three strand modes, motif banks wider than one MMA (K > 256), every k-step count
(motif lengths 1..33), ragged tails (hidden positions not a multiple of 128, K not a multiple of 16), argmax exact."""
import numpy as np
import pytest

from oracle import crbm_oracle as O

pytestmark = pytest.mark.gpu

TOL = dict(rtol=1e-5, atol=1e-5)          # the reference's own tolerance (tests/testcrbm.py:191-198)


def _model(K, M, ds, rs, wscale=0.5):
    import crbm_b200
    m = crbm_b200.CRBM(K, M, doublestranded=ds, seed=1)
    W = (wscale * rs.randn(K, 1, 4, M)).astype(np.float32)
    b = (-2 + rs.randn(1, K)).astype(np.float32)
    m.motifs.set_value(W)
    m.bias.set_value(b)
    return m, W.astype(np.float64), b.astype(np.float64)


@pytest.mark.parametrize('K,M,L,N', [(100, 15, 200, 23), (33, 9, 77, 37), (16, 25, 500, 5), (512, 25, 300, 3),
                                     (128, 20, 999, 4), (2, 5, 40, 3), (300, 33, 120, 6), (7, 1, 130, 9), (40, 17, 17, 300)])
def test_activity_and_probability_match_oracle(K, M, L, N):
    import crbm_b200
    rs = np.random.RandomState(K * 1000 + M)
    codes = rs.randint(0, 4, size=(N, L))
    D = crbm_b200.codesToOneHot(codes)
    D64 = D.astype(np.float64)
    for ds in (True, False):
        m, W, b = _model(K, M, ds, rs)
        for flip in (False, True):
            np.testing.assert_allclose(m._bottomUpActivity(D, flip), O.bottom_up_activity(D64, W, b, flip), **TOL)
        want = O.hit_probs(D64, W, b, 1, ds)
        got = m.motifHitProbs(D)
        assert got.shape == (N, K, 1, L - M + 1) and got.dtype == np.float32
        np.testing.assert_allclose(got, want, **TOL)


def test_argmax_hit_positions_exact_and_large_weights():
    """Well-separated maxima: the hit position of every (sequence, motif) equals the oracle's; weights of magnitude ~8
    (activities of +-100) stay within the relative tolerance."""
    import crbm_b200
    rs = np.random.RandomState(5)
    K, M, L, N = 64, 12, 400, 16
    codes = rs.randint(0, 4, size=(N, L))
    D = crbm_b200.codesToOneHot(codes)
    m, W, b = _model(K, M, True, rs, wscale=4.0)
    act = m._bottomUpActivity(D, False)
    want = O.bottom_up_activity(D.astype(np.float64), W, b, False)
    np.testing.assert_allclose(act, want, rtol=1e-5, atol=1e-4)
    srt = np.sort(want[:, :, 0, :], axis=2)
    clear = (srt[:, :, -1] - srt[:, :, -2]) > 1e-3
    assert clear.mean() > 0.9
    np.testing.assert_array_equal(act[:, :, 0, :].argmax(axis=2)[clear], want[:, :, 0, :].argmax(axis=2)[clear])


def test_many_tiles_round_trip_against_unpacked_windows():
    """Full-size property (no oracle): at the bench shape the activity of hidden position j only depends on the M bases
    under it, so two different sequences sharing a window give bit-identical outputs wherever they sit in a tile."""
    import crbm_b200
    rs = np.random.RandomState(9)
    K, M, L = 100, 15, 200
    base = rs.randint(0, 4, size=(1, L))
    N = 4000                                           # 4000 x 186 positions = 5813 tiles over all CTAs
    codes = np.repeat(base, N, axis=0)
    m, W, b = _model(K, M, True, rs)
    seqs = crbm_b200.PackedSequences.from_codes(codes) if hasattr(crbm_b200.PackedSequences, 'from_codes') else \
        crbm_b200.PackedSequences.from_onehot(crbm_b200.codesToOneHot(codes))
    got = m.motifHitProbs(seqs)
    assert np.array_equal(got, np.broadcast_to(got[:1], got.shape))
    want = O.hit_probs(crbm_b200.codesToOneHot(base).astype(np.float64), W, b, 1, True)
    np.testing.assert_allclose(got[:1], want, **TOL)
