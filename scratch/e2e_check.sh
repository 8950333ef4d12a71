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
run copy CRBM_HOSTFED=copy
run zerocopy CRBM_HOSTFED=zerocopy
python bench.py --gpus 1 --steps 20 --warmup 5 --only-headline --no-scan > gpurun_out/e2e_driver20.json 2> gpurun_out/e2e_driver20.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/e2e_driver20.json') if l.startswith('{')][0]); print('driver-like 20 steps:', d['ms_per_step']*1e3, d['e2e']['value'])"
