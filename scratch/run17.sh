python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan > gpurun_out/r2n_20.json 2> gpurun_out/r2n_20.err
python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan > gpurun_out/r2n_20b.json 2> gpurun_out/r2n_20b.err
python bench.py --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2n_def.json 2> gpurun_out/r2n_def.err
