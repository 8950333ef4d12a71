"""Eager training steps at a given shape on the tcgen05 path (debugging aid): python scratch/tt_debug.py K M P L N cd_k steps"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import crbm_b200
K, M, P, L, N, cd_k, steps = [int(x) for x in sys.argv[1:8]]
np.random.seed(0)
m = crbm_b200.CRBM(K, M, epochs=1, batchsize=N, pooling=P, cd_k=cd_k, seed=1)
m._use_graphs = False
Lt = m._truncated_length(L)
data = crbm_b200.PackedSequences.random(N * 4, Lt, seed=5, device=m._device)
t0 = time.time()
for i in range(steps):
    m._train_packed(data, (i % 4) * N, (i % 4 + 1) * N)
torch.cuda.synchronize()
print('ok', K, M, P, L, N, cd_k, steps, '%.3f s' % (time.time() - t0), float(m.motifs.get_value().std()))
