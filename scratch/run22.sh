CRBM_RT_DEBUG=1 timeout 120 python scratch/scan_ab.py 512 25 1 1000 0 reduce 2> gpurun_out/rt_dbg_c5.txt
CRBM_RT_DEBUG=1 timeout 120 python scratch/scan_ab.py 100 15 1 200 0 reduce 2> gpurun_out/rt_dbg_c2.txt
