"""Scan reductions / free energy at a given shape (profiling aid): python scratch/rt_debug.py K M P L N"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import crbm_b200
from crbm_b200 import _lib
K, M, P, L, N = [int(x) for x in sys.argv[1:6]]
np.random.seed(0)
m = crbm_b200.CRBM(K, M, epochs=1, batchsize=16, pooling=P, cd_k=1, seed=1)
Lt = m._truncated_length(L)
seqs = crbm_b200.PackedSequences.random(N, Lt, seed=5, device=m._device)
lib = _lib.load()
mx = torch.empty((N, K), dtype=torch.float32, device=m._device)
am = torch.empty((N, K), dtype=torch.int32, device=m._device)
mean = torch.empty((N, K), dtype=torch.float32, device=m._device)
def red():
    _lib.check(lib.crbm_scan_reduce(_lib.ptr(seqs.words), seqs.row_words, N, Lt, _lib.ptr(m._t_W), _lib.ptr(m._t_bias), K, M, P,
                                    _lib.STRAND_FWD, _lib.OUT_PROBABILITY, _lib.ptr(mx), _lib.ptr(am), _lib.ptr(mean), _lib.stream_handle()), 'red')
for fn, name in ((red, 'reduce'), (lambda: m._free_energy_device(seqs, True, False, True), 'fe+count')):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(name, K, M, P, Lt, N, '%.3f ms' % ms, '%.3e nt/s' % (N * Lt / ms * 1e3))
