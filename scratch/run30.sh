timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -x -q -k "train_steps or gradient" 2>&1 | tail -2
python bench.py --workload c4 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2t_c4.json 2> gpurun_out/r2t_c4.err
