python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan > gpurun_out/r2y_20.json 2> gpurun_out/r2y_20.err
time python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_full20.json 2> gpurun_out/r2y_full20.err
