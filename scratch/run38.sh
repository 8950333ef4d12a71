timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --only-headline --no-cpu-baseline --no-scan > gpurun_out/r3a_c2.json 2> gpurun_out/r3a_c2.err
python bench.py --workload c4 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r3a_c4.json 2> gpurun_out/r3a_c4.err
