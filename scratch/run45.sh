timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
time python bench.py --steps 20 --warmup 5 > gpurun_out/r4b.json 2> gpurun_out/r4b.err
tail -c 600 gpurun_out/r4b.err
