"""Where do the ~80 us of fixed cost per timed region of a short run come from? Replays of a 20-step graph timed one by one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import crbm_b200
dev = torch.device('cuda', 0)
np.random.seed(0)
m = crbm_b200.CRBM(100, 15, epochs=1, batchsize=100, cd_k=1, seed=1, device=dev)
data = crbm_b200.PackedSequences.random(100000, 200, seed=2, device=dev)
m._train_full_batches(data, 0, 60, steps_per_graph=20)
eng = m._engine
g = m._graph_for(eng, 20)
torch.cuda.synchronize()
def ev(): return torch.cuda.Event(enable_timing=True)
for trial in range(3):
    e = [ev() for _ in range(5)]
    torch.cuda.synchronize()
    e[0].record(); g.replay(); e[1].record(); g.replay(); e[2].record(); g.replay(); e[3].record()
    torch.cuda.synchronize()
    print("after sync, 3 replays back to back: %.1f %.1f %.1f us" % tuple(e[i].elapsed_time(e[i + 1]) * 1e3 for i in range(3)))
for trial in range(3):
    e = [ev() for _ in range(5)]
    torch.cuda.synchronize()
    e[0].record(); eng['step_dev'].fill_(7); eng['batch_dev'].fill_(3); e[1].record(); g.replay(); e[2].record()
    torch.cuda.synchronize()
    print("after sync: two fills %.1f us, then replay %.1f us" % (e[0].elapsed_time(e[1]) * 1e3, e[1].elapsed_time(e[2]) * 1e3))
for trial in range(3):
    e = [ev() for _ in range(5)]
    torch.cuda.synchronize()
    torch.cuda._sleep(4000000)
    e[0].record(); g.replay(); e[1].record(); g.replay(); e[2].record()
    torch.cuda.synchronize()
    print("behind a 2 ms spin kernel: %.1f %.1f us" % (e[0].elapsed_time(e[1]) * 1e3, e[1].elapsed_time(e[2]) * 1e3))
g8 = m._graph_for(eng, 8)
for trial in range(2):
    e = [ev() for _ in range(12)]
    torch.cuda.synchronize()
    for i in range(10):
        e[i].record(); g8.replay()
    e[10].record()
    torch.cuda.synchronize()
    print("8-step graph x 10 back to back: " + " ".join("%.1f" % (e[i].elapsed_time(e[i + 1]) * 1e3) for i in range(10)))
