"""Parity at the motif-bank shapes of BASELINE.json's configs 3-5 (small batches, so the NumPy oracle finishes in seconds):
C3 = 128 motifs x 20 bp, pooling 5, 999 bp (1 kb truncated as fit() does); C4 = 512 motifs x 25 bp, CD-5, 500 bp;
C5 = the 512 x 25 scan in both `doublestranded` branches of getHitProbs (convRBM.py:507-514), dense and with the
callers' reductions fused (secomo/utils.py:116,155,245-247,306). Everything goes through the C ABI.

Tolerances: probabilities / free energies / parameters after a step at the reference's own rtol = atol = 1e-5;
every gradient term (data, model, sparsity) at 1e-5 relative to the largest magnitude of that term ("gradients
... 1e-5 relative to max-abs") — the statistics are sums of fp32 probabilities split into two 11-bit terms for the
tensor cores (relative 2^-22 per term), accumulated in fp32; sampled states and argmax positions exact."""
import ctypes

import numpy as np
import pytest

from tests._golden import onehot

pytestmark = pytest.mark.gpu

TOL = dict(rtol=1e-5, atol=1e-5)


@pytest.fixture(scope='module')
def crbm():
    import crbm_b200
    return crbm_b200


@pytest.fixture(scope='module')
def O():
    from oracle import crbm_oracle
    return crbm_oracle


def make_model(crbm, rs, K, M, P, ds, bs, cd_k, wscale=0.3, bias=-4.0):
    W = (wscale * rs.randn(K, 1, 4, M)).astype(np.float32)
    b = (bias + 0.5 * rs.randn(1, K)).astype(np.float32)
    c = (0.1 * rs.randn(1, 4)).astype(np.float32)
    m = crbm.CRBM(K, M, doublestranded=ds, pooling=P, batchsize=bs, cd_k=cd_k, seed=3)
    m.motifs.set_value(W); m.bias.set_value(b); m.c.set_value(c)
    return m, W, b, c


def sparse_chain(rs, bs, K, P, density=0.02):
    fh = (rs.rand(bs, K, 1, 200) < density).astype(np.float64)
    g = fh.reshape(bs, K, 1, 200 // P, P)
    g[...] = g * (np.cumsum(g, axis=4) == 1)            # at most one active unit per pooling group
    return fh


def rel_to_max(got, want, tol=1e-5):
    want = np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(np.asarray(got, dtype=np.float64) - want).max() / scale
    assert err <= tol, "max error %.3g of the largest magnitude %.3g" % (err, scale)


def gradient_terms(m, crbm, D, uniforms):
    """The un-normalised statistics of one step (crbm_train_accumulate) turned into the reference's gradient terms."""
    import torch
    from crbm_b200 import _lib
    lib = _lib.load()
    seqs = crbm.PackedSequences.from_onehot(D, device=m._device)
    conv = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(m._device)
    u = {k: conv(v) for k, v in uniforms.items()}
    a = m._fill_args(seqs.words.data_ptr(), seqs.row_words, seqs.N, seqs.L, u['v'], u['h'], u.get('hp'))
    _lib.check(lib.crbm_train_accumulate(ctypes.byref(a), _lib.stream_handle()), 'train_accumulate')
    torch.cuda.synchronize()
    st = m._train_state['stats'].cpu().numpy().astype(np.float64)
    lay = m._train_state['layout']
    K, M = m.num_motifs, m.motif_length
    nW = K * 4 * M
    S = 2.0 if m.doublestranded else 1.0
    nhd, nhm = st[lay.n_hidden_data], st[lay.n_hidden_model]
    pk = st[lay.h_fwd:lay.h_fwd + K] / nhd
    gk = -(m.rho / pk - (1 - m.rho) / (1 - pk)) / (K * nhd)
    out = {
        'G_W_data': st[lay.vh_data:lay.vh_data + nW].reshape(K, 1, 4, M) / (S * nhd),
        'G_W_model': st[lay.vh_model:lay.vh_model + nW].reshape(K, 1, 4, M) / (S * nhm),
        'reg_W': st[lay.vh_q:lay.vh_q + nW].reshape(K, 1, 4, M) * gk[:, None, None, None],
        'G_b_data': st[lay.h_data:lay.h_data + K][None, :] / (S * nhd),
        'G_b_model': st[lay.h_model:lay.h_model + K][None, :] / (S * nhm),
        'reg_b': (st[lay.h_q:lay.h_q + K] * gk)[None, :],
    }
    vd, vm = st[lay.v_data:lay.v_data + 4], st[lay.v_model:lay.v_model + 4]
    out['G_c_data'] = ((vd + vd[::-1]) / st[lay.n_vis_data])[None, :]
    out['G_c_model'] = ((vm + vm[::-1]) / st[lay.n_vis_model])[None, :]
    return out


SHAPES = [
    # K,   M,  P, L,   ds,    cd_k, N, bs
    (128, 20, 5, 999, True, 1, 4, 4),        # C3
    (512, 25, 1, 500, True, 5, 4, 4),        # C4: 32 m-tiles, 16 k-tiles (h2v split 4), separate P / Q launches, CD-5
    (512, 25, 1, 500, True, 2, 3, 12),       # C4 bank, short minibatch against a larger persistent chain
    (128, 20, 5, 999, False, 2, 3, 3),       # C3 bank, single stranded, CD-2
    (100, 15, 1, 200, True, 1, 20, 20),      # C2
    (24, 50, 1, 300, True, 2, 5, 5),         # motif_length 50: beyond the tuned kernels -> shape-generic kernels (csrc/generic.cuh)
    (12, 40, 20, 259, False, 1, 4, 6),       # pooling 20 (and motif_length 40), single stranded
]


@pytest.mark.parametrize('K,M,P,L,ds,cd_k,N,bs', SHAPES)
def test_gradient_terms_match_oracle(crbm, O, K, M, P, L, ds, cd_k, N, bs):
    """Positive-phase, negative-phase and sparsity gradient terms of one step, each against the oracle's."""
    rs = np.random.RandomState(K + 3 * M + P + cd_k)
    m, W, b, c = make_model(crbm, rs, K, M, P, ds, bs, cd_k)
    o = O.OracleCRBM(W, b, c, doublestranded=ds, batchsize=bs, pooling=P, cd_k=cd_k)
    fh, fhp = sparse_chain(rs, bs, K, P), sparse_chain(rs, bs, K, P)
    o.fantasy_h, o.fantasy_hp = fh.copy(), (fhp.copy() if ds else None)
    m.fantasy_h.set_value(fh.astype(np.float32))
    if ds:
        m.fantasy_h_prime.set_value(fhp.astype(np.float32))
    lv = 199 + M
    codes = rs.randint(0, 4, size=(N, L))
    us = {'v%d' % it: rs.rand(bs * lv).astype(np.float32) for it in range(cd_k)}
    us.update({'h%d' % it: rs.rand(bs * K * 200 // P).astype(np.float32) for it in range(cd_k)})
    us.update({'hp%d' % it: rs.rand(bs * K * 200 // P).astype(np.float32) for it in range(cd_k)})
    want = o.train_step(onehot(codes, np.float64), lambda tag, rows: us[tag].astype(np.float64))
    u = {'v': np.concatenate([us['v%d' % it] for it in range(cd_k)]),
         'h': np.concatenate([us['h%d' % it] for it in range(cd_k)])}
    if ds:
        u['hp'] = np.concatenate([us['hp%d' % it] for it in range(cd_k)])
    got = gradient_terms(m, crbm, onehot(codes), u)
    for name in ('G_W_data', 'G_b_data', 'G_c_data', 'reg_W', 'reg_b'):
        rel_to_max(got[name], want[name])
    mism = np.mean(m.fantasy_h.get_value() != o.fantasy_h)
    assert mism < 2e-5, mism
    if ds:
        assert np.mean(m.fantasy_h_prime.get_value() != o.fantasy_hp) < 2e-5
    v = crbm.PackedSequences(m._train_state['v_model'], bs, lv).to_onehot()
    if mism == 0 and np.array_equal(v, want['v_model']):
        for name in ('G_W_model', 'G_b_model', 'G_c_model'):
            rel_to_max(got[name], want[name])
    else:
        # a uniform within fp32 noise of a threshold flipped one sample: the model statistics legitimately differ by that unit
        assert np.mean(v != want['v_model']) < 1e-3


@pytest.mark.parametrize('K,M,P,L,ds,cd_k,N,bs', SHAPES[:4] + SHAPES[5:])
def test_train_steps_match_oracle_baseline_shapes(crbm, O, K, M, P, L, ds, cd_k, N, bs):
    """Two whole steps (parameters, velocities, persistent chain) at the C3 / C4 shapes."""
    rs = np.random.RandomState(7 * K + M + P + cd_k)
    m, W, b, c = make_model(crbm, rs, K, M, P, ds, bs, cd_k)
    o = O.OracleCRBM(W, b, c, doublestranded=ds, batchsize=bs, pooling=P, cd_k=cd_k)
    fh, fhp = sparse_chain(rs, bs, K, P), sparse_chain(rs, bs, K, P)
    o.fantasy_h, o.fantasy_hp = fh.copy(), (fhp.copy() if ds else None)
    m.fantasy_h.set_value(fh.astype(np.float32))
    if ds:
        m.fantasy_h_prime.set_value(fhp.astype(np.float32))
    lv = 199 + M
    for step in range(2):
        codes = rs.randint(0, 4, size=(N, L))
        us = {'v%d' % it: rs.rand(bs * lv).astype(np.float32) for it in range(cd_k)}
        us.update({'h%d' % it: rs.rand(bs * K * 200 // P).astype(np.float32) for it in range(cd_k)})
        us.update({'hp%d' % it: rs.rand(bs * K * 200 // P).astype(np.float32) for it in range(cd_k)})
        o.train_step(onehot(codes, np.float64), lambda tag, rows: us[tag].astype(np.float64))
        u = {'v': np.concatenate([us['v%d' % it] for it in range(cd_k)]),
             'h': np.concatenate([us['h%d' % it] for it in range(cd_k)])}
        if ds:
            u['hp'] = np.concatenate([us['hp%d' % it] for it in range(cd_k)])
        m._trainingFct(onehot(codes), uniforms=u)
        mism = np.mean(m.fantasy_h.get_value() != o.fantasy_h)
        assert mism < 2e-5, mism
        if mism > 0:
            pytest.skip("a threshold-adjacent uniform flipped a sample; chains legitimately diverge after that")
        np.testing.assert_allclose(m.motifs.get_value(), o.W, **TOL)
        np.testing.assert_allclose(m.bias.get_value(), o.b, **TOL)
        np.testing.assert_allclose(m.c.get_value(), o.c, **TOL)
        # velocities: lr * (data - model - lambda * reg) is a difference of terms of similar size, so the bound is relative to
        # the largest velocity of the tensor (each term is checked on its own in test_gradient_terms_match_oracle)
        rel_to_max(m.motif_velocity.get_value(), o.vW, 2e-5)
        rel_to_max(m.bias_velocity.get_value(), o.vb, 2e-5)
        rel_to_max(m.c_velocity.get_value(), o.vc, 2e-5)


# ---- C5: the 512 x 25 scan, both getHitProbs branches, dense and fused ------------------------------------------------------
@pytest.mark.parametrize('K,M,P,L,ds', [(512, 25, 1, 500, True), (512, 25, 1, 500, False), (512, 25, 4, 500, False),
                                        (128, 20, 5, 999, True), (100, 15, 1, 200, False),
                                        (24, 50, 1, 300, True), (12, 40, 20, 259, False), (10, 64, 8, 327, True)])
def test_scan_c5_both_strand_modes(crbm, O, K, M, P, L, ds):
    rs = np.random.RandomState(K + M + P + int(ds))
    N = 3
    m, W, b, c = make_model(crbm, rs, K, M, P, ds, N, 1, wscale=0.25, bias=-3.0)
    Lt = m._truncated_length(L)
    D = onehot(rs.randint(0, 4, size=(N, Lt)))
    W64, b64, c64, D64 = W.astype(np.float64), b.astype(np.float64), c.astype(np.float64), D.astype(np.float64)
    want = O.hit_probs(D64, W64, b64, P, ds)
    got = m.motifHitProbs(D)
    assert got.shape == (N, K, 1, Lt - M + 1) and got.dtype == np.float32
    np.testing.assert_allclose(got, want, **TOL)
    np.testing.assert_allclose(m.freeEnergy(D), O.free_energy(D64, W64, b64, c64, P, ds), **TOL)
    # the callers' reductions, fused into the scan (never materialising the tensor), against the ORACLE's tensor
    s = m.motifHitSummary(D)
    w = want[:, :, 0, :]
    np.testing.assert_allclose(s['max'], w.max(axis=2), **TOL)
    np.testing.assert_allclose(s['mean'], w.mean(axis=2), **TOL)
    srt = np.sort(w, axis=2)
    clear = (srt[:, :, -1] - srt[:, :, -2]) > 1e-5 * (1 + srt[:, :, -1])
    np.testing.assert_array_equal(s['argmax'][clear], w.argmax(axis=2)[clear])
    assert clear.mean() > 0.8
    # wherever the maximum is not clear-cut the reported position must still hold a value within tolerance of the maximum
    pos = s['argmax'].astype(np.int64)
    np.testing.assert_allclose(np.take_along_axis(w, pos[:, :, None], axis=2)[:, :, 0], w.max(axis=2), **TOL)


@pytest.mark.parametrize('K,M,P,L,ds,N', [(100, 15, 1, 200, True, 37), (16, 9, 4, 105, False, 11), (130, 25, 1, 330, True, 5)])
def test_position_mean_matches_oracle(crbm, O, K, M, P, L, ds, N):
    """Per-(motif, position) mean over sequences (secomo/utils.py:113-116), reduced on the device in chunks, against the oracle."""
    rs = np.random.RandomState(K + M + N)
    m, W, b, c = make_model(crbm, rs, K, M, P, ds, N, 1, wscale=0.3, bias=-2.5)
    Lt = m._truncated_length(L)
    D = onehot(rs.randint(0, 4, size=(N, Lt)))
    want = O.hit_probs(D.astype(np.float64), W.astype(np.float64), b.astype(np.float64), P, ds).mean(axis=(0, 2))
    got = m.motifHitPositionMean(D, max_chunk_bytes=4 * K * (Lt - M + 8) * 7)      # several chunks, a ragged last one
    assert got.shape == want.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, want, **TOL)
    np.testing.assert_allclose(m.motifHitPositionMean(D), want, **TOL)


def test_small_shapes_on_the_tcgen05_training_path():
    """Small un-pooled banks train on the mma.sync kernels by default (train_tc_impl.cuh: train_uses_tc); CRBM_TRAIN_KERNEL=tc
    forces the tcgen05 kernels, and the whole-step parity tests must pass on them too (read once per process: subprocess)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CRBM_TRAIN_KERNEL='tc')
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu', os.path.join(root, 'tests', 'test_gpu_parity.py'),
                        os.path.join(root, 'tests', 'test_gpu_api.py'), '-k', 'train or step or graph or phase or chain or fantasy'],
                       capture_output=True, text=True, timeout=900, cwd=root, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert ' passed' in r.stdout


def test_shape_generic_kernels_pass_the_parity_suite():
    """The shape-generic kernels (csrc/generic.cuh: any motif_length, any pooling factor) forced onto the small shapes of the
    parity suite (CRBM_GENERIC=1, read once per process: subprocess): fixtures, oracle, graph replay == eager, phases, loop-back."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CRBM_GENERIC='1')
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu', os.path.join(root, 'tests', 'test_gpu_parity.py'),
                        os.path.join(root, 'tests', 'test_gpu_api.py')],
                       capture_output=True, text=True, timeout=1200, cwd=root, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert ' passed' in r.stdout
