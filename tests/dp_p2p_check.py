"""Multi-GPU check of the fused peer-memory all-reduce + update (crbm_train_accumulate_p2p / crbm_train_apply_p2p).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_p2p_check.py

Every rank trains the same model twice on its own shard, once with the NCCL all-reduce between accumulate and apply
(CRBM_DP=nccl) and once with the one-shot peer-memory path (CRBM_DP=p2p), eagerly and by CUDA-graph replay. Checks:
replicas stay bit-identical across ranks in both modes; p2p == NCCL within fp32 summation-order tolerance after the
first step (identical samples) and stays finite / close in distribution afterwards; no peer time-out was flagged.
Run by tests/test_gpu_multi.py when the box has >= 2 GPUs.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import crbm_b200
    # DP_CHECK_SHAPE=tc: a bank that trains on the tcgen05 kernels (K * M >= 2048), pooled; default: the mma.sync path (C2's bank)
    tc = os.environ.get('DP_CHECK_SHAPE', '') == 'tc'
    K, M, B, L, NB, P = (128, 20, 32, 209, 40, 5) if tc else (100, 15, 32, 200, 40, 1)
    rs = np.random.RandomState(3)
    W = (0.5 * rs.randn(K, 1, 4, M)).astype(np.float32)
    data = crbm_b200.PackedSequences.random(NB * B, L, seed=50 + rank, device=dev)
    data.is_local_shard = True

    def run(mode, steps_eager, steps_graph):
        os.environ['CRBM_DP'] = mode
        m = crbm_b200.CRBM(K, M, epochs=1, batchsize=B * world, cd_k=1, pooling=P, seed=11, device=dev)
        m.motifs.set_value(W)
        snaps = []
        for i in range(steps_eager):
            m._train_packed(data, i * B, (i + 1) * B, local_shard=True)
            snaps.append(m.motifs.get_value())
        if steps_graph:
            m._train_full_batches(data, steps_eager, steps_graph, steps_per_graph=4)
            snaps.append(m.motifs.get_value())
        torch.cuda.synchronize()
        if mode == 'p2p':
            assert m._p2p, "peer-memory path was not set up"
            m._p2p_check()
        return m, snaps

    def same_on_all_ranks(x, what):
        t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        ts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        for r in range(world):
            assert torch.equal(ts[r], ts[0]), "%s differs between rank 0 and rank %d" % (what, r)

    m_n, s_n = run('nccl', 3, 9)
    m_p, s_p = run('p2p', 3, 9)
    for i, (a, b) in enumerate(zip(s_n, s_p)):
        assert np.isfinite(a).all() and np.isfinite(b).all()
        same_on_all_ranks(a, 'nccl W after snapshot %d' % i)
        same_on_all_ranks(b, 'p2p W after snapshot %d' % i)
    # first step: same samples in both modes, only the summation order over ranks differs
    np.testing.assert_allclose(s_p[0], s_n[0], rtol=1e-5, atol=2e-6)
    # later steps: sampling thresholds may flip on 1-ulp differences; the trajectories must stay statistically the same
    d = np.abs(s_p[-1] - s_n[-1])
    assert np.median(d) < 1e-4 and d.max() < 0.05, (float(np.median(d)), float(d.max()))
    same_on_all_ranks(m_p.bias.get_value(), 'p2p bias')
    same_on_all_ranks(m_p.c.get_value(), 'p2p c')

    # timing of the two modes, graph replay, a few hundred steps (informational)
    for mode in ('nccl', 'p2p'):
        os.environ['CRBM_DP'] = mode
        m = crbm_b200.CRBM(K, M, epochs=1, batchsize=B * world, cd_k=1, pooling=P, seed=11, device=dev)
        m._train_full_batches(data, 0, 16)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m._train_full_batches(data, 0, 32)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        if rank == 0:
            print("dp%d %s: %.1f us/step" % (world, mode, e0.elapsed_time(e1) * 1e3 / 320))
    # fit() under data parallelism without any seeding by the caller: every rank draws its own W (reference :127 uses the global
    # NumPy generator), rank 0's is broadcast, the ranks train on their shards of the same global minibatches, the epoch metrics
    # are all-reduced, and the replicas must end bit-identical
    os.environ['CRBM_DP'] = 'p2p'
    np.random.seed(1000 + rank)
    m = crbm_b200.CRBM(K, M, epochs=2, batchsize=8 * world, cd_k=1, pooling=P, device=dev)
    rs2 = np.random.RandomState(77)                       # the same global training set on every rank
    codes = rs2.randint(0, 4, size=(8 * world * 3 + 3, L))
    D = np.zeros((codes.shape[0], 1, 4, L), dtype=np.float32)
    for a in range(4):
        D[:, 0, a, :] = codes == a
    m.fit(D)
    for name in ('motifs', 'bias', 'c'):
        v = getattr(m, name).get_value()
        assert np.isfinite(v).all()
        same_on_all_ranks(v, 'fit() replica %s' % name)
    if rank == 0:
        print("DP_P2P_CHECK_OK world=%d shape=%s" % (world, 'tc' if tc else 'mma'))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
