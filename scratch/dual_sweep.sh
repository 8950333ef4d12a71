#!/bin/bash
# On the GPU box: parity tests, then C2 step time for the two-role launch with several CTA splits.
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
B="python bench.py --workload c2 --no-cpu-baseline --no-scan --no-e2e --only-headline --steps 4000"
run() { name=$1; shift; env "$@" timeout 120 $B > gpurun_out/dual_$name.json 2> gpurun_out/dual_$name.err; python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/dual_$name.json') if l.startswith('{')][0])
    print('$name', round(d['ms_per_step']*1e3,2), 'us/step', d['step_kernels_us'])
except Exception as e:
    print('$name', 'failed', e)
P
}
run off CRBM_DUAL=0
run full CRBM_DUAL=1
run g21_21 CRBM_DUAL_GY_POS=21 CRBM_DUAL_GY_NEG=21
run g28_14 CRBM_DUAL_GY_POS=28 CRBM_DUAL_GY_NEG=14
run g14_28 CRBM_DUAL_GY_POS=14 CRBM_DUAL_GY_NEG=28
run g42_21 CRBM_DUAL_GY_POS=42 CRBM_DUAL_GY_NEG=21
run g21_42 CRBM_DUAL_GY_POS=21 CRBM_DUAL_GY_NEG=42
run g32_32 CRBM_DUAL_GY_POS=32 CRBM_DUAL_GY_NEG=32
