timeout 900 python -m pytest tests/test_gpu_api.py -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --only-headline --no-cpu-baseline --no-scan > gpurun_out/r2m_e2e_$i.json 2> gpurun_out/r2m_e2e_$i.err; done
python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan > gpurun_out/r2m_e2e_20.json 2> gpurun_out/r2m_e2e_20.err
