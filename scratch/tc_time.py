import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import crbm_b200
from crbm_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
shapes = [(100, 15, 200, 0), (100, 15, 200, 2), (512, 25, 500, 0), (128, 20, 999, 0)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (K, M, Lt, mode) in shapes:
    Lh = Lt - M + 1
    n = max(1024, min(200000, int(8e9 // (4 * K * Lh))))
    np.random.seed(0)
    model = crbm_b200.CRBM(K, M, seed=1)
    seqs = crbm_b200.PackedSequences.random(n, Lt, seed=99, device=dev)
    s = _lib.stream_handle()
    for pitch in (Lh, (Lh + 7) // 8 * 8):
        out = torch.empty((n, K, 1, pitch), dtype=torch.float32, device=dev)

        def go():
            _lib.check(lib.crbm_v2h_dense_strided(_lib.ptr(seqs.words), seqs.row_words, n, Lt, _lib.ptr(model._t_W),
                                                  _lib.ptr(model._t_bias), K, M, 1, mode, _lib.OUT_PROBABILITY, _lib.ptr(out),
                                                  pitch, s), 'scan')
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        alg = n * (Lt / 4.0) + 4.0 * n * K * Lh
        print("K=%d M=%d L=%d mode=%d n=%d pitch=%d: %.3f ms, %.3e nt/s, %.0f GB/s (%.1f%% of 6544)"
              % (K, M, Lt, mode, n, pitch, ms, n * Lt / ms * 1e3, alg / ms / 1e6, alg / ms / 1e6 / 65.443))
        del out
