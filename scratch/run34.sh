timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan > gpurun_out/r2w_20_$i.json 2> gpurun_out/r2w_20_$i.err
done
python bench.py --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2w_def.json 2> gpurun_out/r2w_def.err
