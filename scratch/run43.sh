timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
time python bench.py --steps 20 --warmup 5 > gpurun_out/r3f.json 2> gpurun_out/r3f.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r3f_ref.json 2> gpurun_out/r3f_ref.err; tail -c 300 gpurun_out/r3f_ref.json
