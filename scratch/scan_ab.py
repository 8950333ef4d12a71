"""Times crbm_v2h_dense_strided at a given shape (padded rows): python scratch/scan_ab.py K M P L [mode]; env selects the kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crbm_b200
from crbm_b200 import _lib
K, M, P, L = [int(x) for x in sys.argv[1:5]]
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
lib = _lib.load()
dev = torch.device('cuda', 0)
m = crbm_b200.CRBM(K, M, pooling=P, seed=1, device=dev)
Lh = L - M + 1
n = max(1024, min(200000, int(8e9 // (4 * K * Lh))))
seqs = crbm_b200.PackedSequences.random(n, L, seed=99, device=dev)
pitch = (Lh + 7) // 8 * 8
out = torch.empty((n, K, 1, pitch), dtype=torch.float32, device=dev)
s = _lib.stream_handle()
def call():
    _lib.check(lib.crbm_v2h_dense_strided(_lib.ptr(seqs.words), seqs.row_words, n, L, _lib.ptr(m._t_W), _lib.ptr(m._t_bias), K, M, P,
                                          mode, _lib.OUT_PROBABILITY, _lib.ptr(out), pitch, s), 'scan')
if len(sys.argv) > 6 and sys.argv[6] == 'reduce':
    mx = torch.empty((n, K), dtype=torch.float32, device=dev); am = torch.empty((n, K), dtype=torch.int32, device=dev); mean = torch.empty((n, K), dtype=torch.float32, device=dev)
    def call():
        _lib.check(lib.crbm_scan_reduce(_lib.ptr(seqs.words), seqs.row_words, n, L, _lib.ptr(m._t_W), _lib.ptr(m._t_bias), K, M, P,
                                        mode, _lib.OUT_PROBABILITY, _lib.ptr(mx), _lib.ptr(am), None if os.environ.get('NOMEAN') else _lib.ptr(mean), s), 'scan_reduce')
for _ in range(2): call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
alg = n * L / 4.0 + 4.0 * n * K * Lh
print(("REDUCE " if len(sys.argv) > 6 else "") + "K=%d M=%d P=%d L=%d mode=%d env=%s/%s: %.3f ms, %.3e nt/s, %.0f GB/s (%.1f %% of 6544)" % (
    K, M, P, L, mode, os.environ.get('CRBM_SCAN_KERNEL', '-'), os.environ.get('CRBM_POOL_SCAN', '-'), ms, n * L / ms * 1e3, alg / ms / 1e6, alg / ms / 1e6 / 65.443))
