"""DRAM traffic of the C5-shape dense scan launch and of the C4 positive-phase launch, for profiles/r02_traffic.json:
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
       -k regex:'v2h_rowdense_tc_kernel|v2h_train_tc_kernel' --csv --log-file gpurun_out/traffic_c5_c4.csv python scratch/ncu_traffic_c5_c4.py
Shapes are those of bench.py's c5.dense record (4002 x 1000 bp, K=512, M=25) and c4 record (128 x 500 bp, positive phase)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import crbm_b200
from crbm_b200 import _lib

lib = _lib.load()
np.random.seed(0)
dev = torch.device('cuda', 0)
m = crbm_b200.CRBM(512, 25, epochs=1, batchsize=128, pooling=1, cd_k=5, seed=1, device=dev)
seqs = crbm_b200.PackedSequences.random(4002, 1000, seed=3, device=dev)
for _ in range(2):
    v = m._v2h_device(seqs, _lib.STRAND_FWD, _lib.OUT_PROBABILITY)
torch.cuda.synchronize()
del v
data = crbm_b200.PackedSequences.random(1024, 500, seed=4, device=dev)
full = m._fill_args(data.words.data_ptr(), data.row_words, 128, data.L)
for _ in range(2):
    _lib.check(lib.crbm_train_phases(full, _lib.PHASE_POSITIVE, _lib.stream_handle()), 'train_phases')
torch.cuda.synchronize()
print('done')
