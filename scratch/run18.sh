for spg in 20 10 5 4; do
python bench.py --steps 20 --warmup 5 --steps-per-graph $spg --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2o_$spg.json 2> gpurun_out/r2o_$spg.err
done
python bench.py --steps 40 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2o_s40.json 2> gpurun_out/r2o_s40.err
python bench.py --steps 64 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2o_s64.json 2> gpurun_out/r2o_s64.err
python bench.py --steps 200 --warmup 5 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2o_s200.json 2> gpurun_out/r2o_s200.err
