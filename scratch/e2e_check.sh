#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
B="python bench.py --workload c2 --no-cpu-baseline --no-scan --only-headline --steps 2000"
run() { name=$1; shift; env "$@" timeout 200 $B > gpurun_out/e2e_$name.json 2> gpurun_out/e2e_$name.err; python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/e2e_$name.json') if l.startswith('{')][0])
    print('$name', round(d['ms_per_step']*1e3,2), 'us/step; e2e', '%.3e'%d['e2e']['value'], 'blocking %.3e'%d['e2e']['blocking_read_value'], 'us/step e2e', round(20000/d['e2e']['value']*1e6,2))
except Exception as e:
    print('$name', 'failed', e); print(open('gpurun_out/e2e_$name.err').read()[-1500:])
P
}
run zc_dual CRBM_DUAL=1
run zc_nodual CRBM_DUAL=0
run copy CRBM_HOSTFED=copy
run g16_26 CRBM_DUAL_GY_POS=16 CRBM_DUAL_GY_NEG=26
run g17_25 CRBM_DUAL_GY_POS=17 CRBM_DUAL_GY_NEG=25
run g18_24 CRBM_DUAL_GY_POS=18 CRBM_DUAL_GY_NEG=24
run g15_27 CRBM_DUAL_GY_POS=15 CRBM_DUAL_GY_NEG=27
run g21_21 CRBM_DUAL_GY_POS=21 CRBM_DUAL_GY_NEG=21
python bench.py --gpus 1 --steps 20 --warmup 5 --only-headline --no-scan > gpurun_out/e2e_driver20.json 2> gpurun_out/e2e_driver20.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/e2e_driver20.json') if l.startswith('{')][0]); print('driver-like 20 steps:', d['ms_per_step']*1e3, d['e2e']['value'])"
