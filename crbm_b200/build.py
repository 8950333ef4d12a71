"""Builds crbm_b200/libcrbm_b200.so (the C-ABI library of sm_100a kernels) in-tree with nvcc.

    python -m crbm_b200.build [--force]

One object per translation unit, compiled in parallel; `v2h_inst.cu` is compiled once per padded motif length
(-DCRBM_MW=3..33) because its fully unrolled kernels are the bulk of the compile time. Objects are cached under
crbm_b200/csrc/_build and only rebuilt when a source or header is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_build')
LIB = os.path.join(HERE, 'libcrbm_b200.so')
MWS = (3, 6, 9, 12, 15, 18, 21, 24, 27, 30, 33)
KSS = (1, 2, 3, 4, 5, 6, 7, 8, 9)
TT_POOLS = (1, 2, 4, 5, 10)       # pooling factors the tcgen05 training kernel is instantiated for
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hs.append(os.path.join(os.path.dirname(HERE), 'include', 'crbm_b200.h'))
    return max(os.path.getmtime(h) for h in hs)


def _jobs():
    jobs = []
    for src in ('api.cu', 'ingest.cu', 'ops.cu', 'train.cu', 'scan_tc.cu'):
        jobs.append((os.path.join(CSRC, src), os.path.join(OBJ, src[:-3] + '.o'), []))
    for mw in MWS:
        jobs.append((os.path.join(CSRC, 'v2h_inst.cu'), os.path.join(OBJ, 'v2h_inst_%d.o' % mw), ['-DCRBM_MW=%d' % mw]))
    for ks in KSS:
        jobs.append((os.path.join(CSRC, 'v2h_mma_inst.cu'), os.path.join(OBJ, 'v2h_mma_inst_%d.o' % ks), ['-DCRBM_KS=%d' % ks]))
    for pp in TT_POOLS:
        jobs.append((os.path.join(CSRC, 'train_tc_inst.cu'), os.path.join(OBJ, 'train_tc_inst_%d.o' % pp), ['-DCRBM_TT_P=%d' % pp]))
        jobs.append((os.path.join(CSRC, 'scan_rt_inst.cu'), os.path.join(OBJ, 'scan_rt_inst_%d.o' % pp), ['-DCRBM_TT_P=%d' % pp]))
    return jobs


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr = _newest_header()
    todo = []
    for src, obj, extra in _jobs():
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr):
            todo.append((src, obj, extra))

    def compile_one(job):
        src, obj, extra = job
        cmd = [nvcc] + NVCC_FLAGS + extra + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (' '.join(cmd), r.stdout, r.stderr))
        if verbose:
            print("built", os.path.basename(obj))
        return obj

    if todo:
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, todo))
    objs = [obj for _, obj, _ in _jobs()]
    if todo or not os.path.exists(LIB) or force:
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static',
                                                      '-Xcompiler', '-fPIC']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s\n%s" % (' '.join(cmd), r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose=True)
    print(path)
