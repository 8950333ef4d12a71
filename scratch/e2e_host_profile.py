"""Where does the host time of the one-hot training call go? (run on a GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import crbm_b200
from crbm_b200 import _lib

K, M, B, L = 100, 15, 100, 200
np.random.seed(0)
m = crbm_b200.CRBM(K, M, epochs=1, batchsize=B, cd_k=1, seed=1)
rs = np.random.RandomState(7)
host = []
for i in range(16):
    codes = torch.from_numpy(rs.randint(0, 4, size=(B, L)))
    host.append(torch.nn.functional.one_hot(codes, 4).permute(0, 2, 1).reshape(B, 1, 4, L).float().contiguous().pin_memory())
sink = torch.empty((64, 4)).pin_memory()
c_dev = m._t_c.reshape(-1)
for i in range(10):
    m._trainingFct(host[i % 16])
torch.cuda.synchronize()
N = 2000

def timeit(name, fn, sync_every=64):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        fn(i)
        if i % sync_every == sync_every - 1:
            torch.cuda.current_stream().synchronize()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print('%-50s host loop %.1f us/step, incl. drain %.1f us/step' % (name, 1e6 * t_host / N, 1e6 * t_all / N))

timeit('trainingFct only', lambda i: m._trainingFct(host[i % 16]))
timeit('trainingFct + async readback', lambda i: (m._trainingFct(host[i % 16]), sink[i % 64].copy_(c_dev, non_blocking=True)))
timeit('async readback only', lambda i: sink[i % 64].copy_(c_dev, non_blocking=True))
eng = m._onehot_engine
a = eng['args']
s = torch.cuda.current_stream().cuda_stream
def direct(i):
    a.step = i
    eng['fn'](a, None, host[i % 16].data_ptr(), eng['c_staging'], eng['c_bad'], eng['c_side'], eng['c_done'], eng['c_ready'], s)
timeit('C call only (crbm_train_step_onehot)', direct)
lib = _lib.load()
def phases(i):
    a.step = i
    lib.crbm_train_phases(a, 15, s)
timeit('C call crbm_train_phases(all) only', phases)
def stepfn(i):
    a.step = i
    lib.crbm_train_step(a, s)
timeit('C call crbm_train_step only', stepfn)
timeit('python noop', lambda i: None)
timeit('host[i].data_ptr()', lambda i: host[i % 16].data_ptr())
timeit('current_stream', lambda i: torch.cuda.current_stream(m._device).cuda_stream)
