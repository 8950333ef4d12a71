"""Host time vs device time of the host-fed training call (crbm_train_step_hostfed), and the cost of short graph-replay runs."""
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
trace = m.traceVisibleBias(64)
for i in range(10):
    m._trainingFct(host[i % 16])
torch.cuda.synchronize()
N = 2048

def timeit(name, fn, sync_every=64):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        fn(i)
        if sync_every and i % sync_every == sync_every - 1:
            torch.cuda.current_stream().synchronize()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print('%-60s host loop %.1f us/step, incl. drain %.1f us/step' % (name, 1e6 * t_host / N, 1e6 * t_all / N), flush=True)

timeit('trainingFct (hostfed), sync every 64', lambda i: m._trainingFct(host[i % 16]))
timeit('trainingFct (hostfed), no sync', lambda i: m._trainingFct(host[i % 16]), 0)
eng = m._onehot_engine
a = eng['args']
s = torch.cuda.current_stream().cuda_stream
def direct(i):
    a.step = i
    eng['fn_hostfed'](a, None, host[i % 16].data_ptr(), eng['c_bad'], s)
timeit('C call only (crbm_train_step_hostfed), no sync', direct, 0)
timeit('C call only (crbm_train_step_hostfed), sync 64', direct, 64)
lib = _lib.load()
def stepfn(i):
    a.step = i
    lib.crbm_train_step(a, s)
timeit('C call crbm_train_step (resident packed rows), no sync', stepfn, 0)
# device time of the hostfed step in a graph-free back-to-back run measured by events
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(N):
    direct(i)
e1.record()
torch.cuda.synchronize()
print('hostfed C calls, device time by events: %.1f us/step' % (1e3 * e0.elapsed_time(e1) / N))

# short graph-replay runs: K = 20 steps with different graph sizes
data = crbm_b200.PackedSequences.random(100000, L, seed=3, device=m._device)
for spg in (1, 2, 4, 5, 10, 20):
    m._engine = None
    m._train_full_batches(data, 0, 3 * spg + 20, steps_per_graph=spg)
    m._train_full_batches(data, 0, 20, steps_per_graph=spg)
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        e0.record()
        m._train_full_batches(data, 7, 20, steps_per_graph=spg)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('20 steps, %2d steps per graph: %.1f us/step (best of 5)' % (spg, 1e3 * best / 20), flush=True)
# eager
best = 1e9
for rep in range(5):
    torch.cuda.synchronize()
    e0.record()
    for i in range(20):
        stepfn(i)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print('20 steps, eager C calls: %.1f us/step (best of 5)' % (1e3 * best / 20))
