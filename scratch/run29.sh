time python bench.py > gpurun_out/r2s_full.json 2> gpurun_out/r2s_full.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
