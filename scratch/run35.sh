for r in 1 10 40; do
BENCH_REHEARSALS=$r python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2x_$r.json 2> gpurun_out/r2x_$r.err
done
