timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c3 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2r_c3.json 2> gpurun_out/r2r_c3.err
python bench.py --workload c4 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2r_c4.json 2> gpurun_out/r2r_c4.err
