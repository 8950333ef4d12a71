timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-headline > gpurun_out/r3d.json 2> gpurun_out/r3d.err
