"""Data-parallel design, checked on CPU with world_size-2 gloo (no GPU, no product kernels involved):

the statistics vector each rank produces (un-normalised sums, include/crbm_b200.h `crbm_stats_layout`) is linear in the
samples, so ONE sum all-reduce followed by the per-motif sparsity factor and the momentum update reproduces the
single-process reference step (convRBM.py:373-438) on the global minibatch with the union of the fantasy particles.
Each rank computes its sums with the NumPy oracle; the update restates crbm_train_apply in NumPy.
Also covers the host-side sharding arithmetic of CRBM._train_packed / _local_shard.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import crbm_oracle as O          # noqa: E402
from tests._golden import onehot             # noqa: E402

K, M, P, L, CDK = 6, 5, 2, 28, 2
N_GLOBAL, B_GLOBAL = 6, 4
HYP = dict(learning_rate=0.1, momentum=0.95, rho=0.01, lambda_rate=0.1)


def make_problem():
    rs = np.random.RandomState(42)
    W = 0.5 * rs.randn(K, 1, 4, M)
    b = -2.0 + 0.3 * rs.randn(1, K)
    c = 0.1 * rs.randn(1, 4)
    D = onehot(rs.randint(0, 4, size=(N_GLOBAL, L)), np.float64)
    fh = (rs.rand(B_GLOBAL, K, 1, 200) < 0.05).astype(np.float64)
    fhp = (rs.rand(B_GLOBAL, K, 1, 200) < 0.05).astype(np.float64)
    for arr in (fh, fhp):
        g = arr.reshape(B_GLOBAL, K, 1, 200 // P, P)
        g[...] = g * (np.cumsum(g, axis=4) == 1)
    lv = 199 + M
    us = {}
    for it in range(CDK):
        us['v%d' % it] = rs.rand(B_GLOBAL, lv)
        us['h%d' % it] = rs.rand(B_GLOBAL, K * 200 // P)
        us['hp%d' % it] = rs.rand(B_GLOBAL, K * 200 // P)
    return W, b, c, D, fh, fhp, us


def local_sums(W, b, c, D, fh, fhp, us):
    """Un-normalised statistics of one rank: the quantities crbm_train_accumulate writes."""
    Lh = D.shape[3] - M + 1
    ph = O.bottom_up_probability(O.bottom_up_activity(D, W, b, False), P)
    php = O.bottom_up_probability(O.bottom_up_activity(D, W, b, True), P)
    n_hd = D.shape[0] * Lh
    vh_data = (O.collect_vh(ph, D) + O.collect_vh(php, D)[:, :, ::-1, ::-1]) * n_hd
    grp = ph.reshape(D.shape[0], K, 1, Lh // P, P)
    Q = (grp * (1.0 - grp.sum(axis=4, keepdims=True))).reshape(ph.shape)
    vh_q = O.collect_vh(Q, D) * n_hd
    h, hp = fh, fhp
    for it in range(CDK):
        pv = O.top_down_probability(O.top_down_activity(h, hp, W, c))
        v = O.top_down_sample(pv, us['v%d' % it].reshape(-1))
        pm = O.bottom_up_probability(O.bottom_up_activity(v, W, b, False), P)
        pmp = O.bottom_up_probability(O.bottom_up_activity(v, W, b, True), P)
        h = O.bottom_up_sample(pm, P, us['h%d' % it].reshape(-1))
        hp = O.bottom_up_sample(pmp, P, us['hp%d' % it].reshape(-1))
    n_hm = v.shape[0] * 200
    vh_model = (O.collect_vh(pm, v) + O.collect_vh(pmp, v)[:, :, ::-1, ::-1]) * n_hm
    vec = np.concatenate([vh_data.ravel(), vh_model.ravel(), vh_q.ravel(),
                          (ph + php).sum(axis=(0, 2, 3)), (pm + pmp).sum(axis=(0, 2, 3)), Q.sum(axis=(0, 2, 3)),
                          ph.sum(axis=(0, 2, 3)), D.sum(axis=(0, 1, 3)), v.sum(axis=(0, 1, 3)),
                          [n_hd, n_hm, D.shape[0] * D.shape[3], v.shape[0] * v.shape[3]]])
    return vec, h, hp


def apply_update(W, b, c, vec):
    """NumPy restatement of crbm_train_apply (csrc/train.cu) from the all-reduced statistics vector."""
    nW = K * 4 * M
    o = 0
    vh_data = vec[o:o + nW].reshape(K, 1, 4, M); o += nW
    vh_model = vec[o:o + nW].reshape(K, 1, 4, M); o += nW
    vh_q = vec[o:o + nW].reshape(K, 1, 4, M); o += nW
    h_data, h_model, h_q, h_fwd = (vec[o + i * K:o + (i + 1) * K] for i in range(4)); o += 4 * K
    v_data, v_model = vec[o:o + 4], vec[o + 4:o + 8]; o += 8
    n_hd, n_hm, n_vd, n_vm = vec[o:o + 4]
    pk = h_fwd / n_hd
    gk = -(HYP['rho'] / pk - (1 - HYP['rho']) / (1 - pk)) / (K * n_hd)
    lr, lam = HYP['learning_rate'], HYP['lambda_rate']
    vW = lr * (vh_data / (2 * n_hd) - vh_model / (2 * n_hm) - lam * gk[:, None, None, None] * vh_q)
    vb = lr * (h_data / (2 * n_hd) - h_model / (2 * n_hm) - lam * gk * h_q)
    vc = lr * ((v_data + v_data[::-1]) / n_vd - (v_model + v_model[::-1]) / n_vm)
    return W + vW, b + vb[None, :], c + vc[None, :]


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    W, b, c, D, fh, fhp, us = make_problem()
    per = (N_GLOBAL + world - 1) // world
    lo, hi = min(N_GLOBAL, rank * per), min(N_GLOBAL, (rank + 1) * per)        # CRBM._train_packed's shard
    lb = B_GLOBAL // world
    sl = slice(rank * lb, (rank + 1) * lb)                                     # this rank's fantasy particles
    vec, h, hp = local_sums(W, b, c, D[lo:hi], fh[sl], fhp[sl], {k: u[sl] for k, u in us.items()})
    t = torch.from_numpy(vec)
    dist.all_reduce(t)
    Wn, bn, cn = apply_update(W, b, c, t.numpy())
    out[rank] = (Wn, bn, cn, h, hp)
    dist.destroy_process_group()


def _run(rank, world, port, q):
    out = {}
    _worker(rank, world, port, out)
    q.put((rank, out[rank]))


def test_allreduced_statistics_reproduce_single_process_step():
    W, b, c, D, fh, fhp, us = make_problem()
    o = O.OracleCRBM(W, b, c, doublestranded=True, batchsize=B_GLOBAL, pooling=P, cd_k=CDK, **HYP)
    o.fantasy_h, o.fantasy_hp = fh.copy(), fhp.copy()
    o.train_step(D, lambda tag, rows: us[tag].reshape(-1))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_run, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        Wn, bn, cn, h, hp = res[r]
        np.testing.assert_allclose(Wn, o.W, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(bn, o.b, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(cn, o.c, rtol=1e-10, atol=1e-12)
        lb = B_GLOBAL // 2
        np.testing.assert_array_equal(h, o.fantasy_h[r * lb:(r + 1) * lb])
        np.testing.assert_array_equal(hp, o.fantasy_hp[r * lb:(r + 1) * lb])


def test_local_shard_row_selection():
    """CRBM._local_shard: rank r owns rows [r*lb, (r+1)*lb) of every full global minibatch, batch-major."""
    import crbm_b200
    from crbm_b200.sequences import PackedSequences
    m = crbm_b200.CRBM(3, 4, batchsize=6, device='cpu')
    words = torch.arange(20 * 5, dtype=torch.int32).reshape(20, 5)       # 20 sequences, 3 full batches of 6
    seqs = PackedSequences(words, 20, 40)
    m._world, m._local_batch = 3, 2
    for r in range(3):
        m._rank = r
        shard, n_full = m._local_shard(seqs)
        assert n_full == 3 and shard.shape == (6, 5)
        want = torch.cat([words[bi * 6 + r * 2: bi * 6 + r * 2 + 2] for bi in range(3)])
        assert torch.equal(shard, want)
    with pytest.raises(Exception):
        os.environ.pop('RANK', None)
        crbm_b200.CRBM(3, 4, batchsize=0)


# ---- replicas start identical: rank 0's unseeded draw of W is broadcast (reference :127 draws from the global NumPy generator) ----
def _replica_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import crbm_b200
    np.random.seed(1000 + rank)                        # the ranks' generators differ, as they do for an unseeded caller
    m = crbm_b200.CRBM(5, 7, batchsize=4, device='cpu')
    first = m.motifs.get_value().copy()
    # a caller that sets different values per rank is brought back in line at the start of fit(): _sync_replicas
    m.motifs.set_value(np.full((5, 1, 4, 7), float(rank), dtype=np.float32))
    m._sync_replicas()
    q.put((rank, first, m.motifs.get_value(), m._local_batch))
    dist.destroy_process_group()


def test_replicas_share_rank0_parameters():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_replica_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (a, b, lb) for r, a, b, lb in (q.get(timeout=180) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(res[0][0], res[1][0])            # same W although the generators were seeded differently
    assert np.abs(res[0][0]).max() > 0
    np.testing.assert_array_equal(res[1][1], 0.0)                  # rank 1 took rank 0's values again
    assert res[0][2] == 2 and res[1][2] == 2


def test_epoch_metric_batches_are_dealt_out_once():
    """CRBM._epoch_metrics: rank r evaluates batches [r*nb//W, (r+1)*nb//W) — every batch exactly once, contiguous rows."""
    for nb in (1, 2, 5, 8, 13):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = (r * nb) // world, ((r + 1) * nb) // world
                seen += list(range(lo, hi))
            assert seen == list(range(nb))
