for i in 1 2; do
python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2v_20_$i.json 2> gpurun_out/r2v_20_$i.err
done
BENCH_PREFLUSH=0 python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2v_20_off.json 2> gpurun_out/r2v_20_off.err
python bench.py --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2v_def.json 2> gpurun_out/r2v_def.err
