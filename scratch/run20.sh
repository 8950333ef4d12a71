python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv --log-file gpurun_out/r2q_launches_c2.csv python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/ncu_c2.log 2>&1
python bench.py --workload c3 --steps 16 --warmup 3 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv --log-file gpurun_out/r2q_launches_c3.csv python bench.py --workload c3 --steps 16 --warmup 3 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/ncu_c3.log 2>&1
python bench.py --workload c4 --steps 16 --warmup 3 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 320 --csv --log-file gpurun_out/r2q_launches_c4.csv python bench.py --workload c4 --steps 16 --warmup 3 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/ncu_c4.log 2>&1
