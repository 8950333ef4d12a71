time python bench.py > gpurun_out/r2i_full.json 2> gpurun_out/r2i_full.err
