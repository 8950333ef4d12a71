"""The multithreaded torch-CPU restatement (bench.py's CPU arm) against the NumPy oracle. CPU-only."""
import numpy as np
import torch

from oracle import crbm_oracle as O
from oracle.crbm_torch_cpu import TorchCpuCRBM, random_onehot


def test_deterministic_parts_match_numpy_oracle():
    rs = np.random.RandomState(0)
    K, M, P, N, L = 6, 5, 2, 4, 34
    W = (0.5 * rs.randn(K, 1, 4, M)).astype(np.float32)
    b = (-2 + rs.randn(1, K)).astype(np.float32)
    c = (0.1 * rs.randn(1, 4)).astype(np.float32)
    t = TorchCpuCRBM(K, M, pooling=P, batchsize=N, W=W, b=b, c=c)
    D = random_onehot(N, L, 1)
    Dn = D.numpy().astype(np.float64)
    W64, b64, c64 = W.astype(np.float64), b.astype(np.float64), c.astype(np.float64)
    for flip in (False, True):
        np.testing.assert_allclose(t.act(D, flip).numpy(), O.bottom_up_activity(Dn, W64, b64, flip), rtol=1e-5, atol=1e-5)
    p = t.prob(t.act(D))
    np.testing.assert_allclose(p.numpy(), O.hit_probs(Dn, W64, b64, P, True), rtol=1e-5, atol=1e-6)
    h = (torch.rand(N, K, 1, L - M + 1) < 0.1).float()
    hp = (torch.rand(N, K, 1, L - M + 1) < 0.1).float()
    np.testing.assert_allclose(t.v_act(h, hp).numpy(),
                               O.top_down_activity(h.numpy().astype(np.float64), hp.numpy().astype(np.float64), W64, c64),
                               rtol=1e-5, atol=1e-5)
    php = t.prob(t.act(D, True))
    gW, gb, gc = t.stats(p, php, D)
    oW, ob, oc = O.collect_update_statistics(p.numpy().astype(np.float64), php.numpy().astype(np.float64), Dn)
    np.testing.assert_allclose(gW.numpy(), oW, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gb.numpy(), ob, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gc.numpy(), oc, rtol=1e-5, atol=1e-6)
    rW, rb = t.sparsity_grad(D)
    sW, sb = O.sparsity_gradient(Dn, W64, b64, P, 0.01)
    np.testing.assert_allclose(rW.numpy(), sW, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(rb.numpy(), sb, rtol=1e-4, atol=1e-7)
    s = t.v_sample(t.v_act(h, hp))
    assert torch.all(s.sum(dim=2) == 1)
    t.train_step(D)     # runs
