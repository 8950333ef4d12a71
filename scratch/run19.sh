python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2p_20.json 2> gpurun_out/r2p_20.err
python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2p_20b.json 2> gpurun_out/r2p_20b.err
python bench.py --steps 40 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2p_40.json 2> gpurun_out/r2p_40.err
