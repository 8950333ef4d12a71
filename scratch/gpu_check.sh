#!/bin/bash
# On the GPU box: the GPU test-suite, then short training benches of C3 / C4 (and C2 forced onto the tcgen05 path).
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
for w in c3 c4; do timeout 100 python bench.py --workload $w --no-cpu-baseline --no-scan --no-e2e --only-headline > gpurun_out/chk_$w.json 2> gpurun_out/chk_$w.err; done
CRBM_TRAIN_KERNEL=tc timeout 100 python bench.py --workload c2 --no-cpu-baseline --no-scan --no-e2e --only-headline > gpurun_out/chk_c2tc.json 2> gpurun_out/chk_c2tc.err
