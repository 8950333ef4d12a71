"""Timeline of CTA 0 of the dense h2v kernel at C4 (CRBM_H2V_DEBUG=1)."""
import os, sys, ctypes
os.environ['CRBM_H2V_DEBUG'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import crbm_b200
from crbm_b200 import _lib
np.random.seed(0)
m = crbm_b200.CRBM(512, 25, epochs=1, batchsize=128, cd_k=5, seed=1)
data = crbm_b200.PackedSequences.random(4096, 500, seed=3, device=m._device)
for i in range(6):
    m._train_packed(data, 128 * i, 128 * i + 128)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 16)()
lib.crbm_debug_h2v_timeline.argtypes = [ctypes.c_void_p]
lib.crbm_debug_h2v_timeline(buf)
t = list(buf)
names = ['entry', 'prologue done', 'pdl_wait done', 'MMA: first full_a', 'MMA: first full_b', 'MMA: all issued', 'EPI: acc_full', 'EPI: words done',
         'EPI: finalize done', 'exit', 'MMA: it8 full_a', 'MMA: it8 full_b', 'EPI: w1 tmem loaded', 'EPI: w1 staged', 'EPI: w1 diag done', 'EPI: w0 diag done']
for i in sorted(range(16), key=lambda i: t[i]):
    print('%-22s %8.2f us' % (names[i], (t[i] - t[0]) / 1965.0))
