set -x
python scratch/scan_ab.py 128 20 5 999 > gpurun_out/plain_rd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rowdense -s 2 -c 1 -f -o gpurun_out/r2_rd_c3b python scratch/scan_ab.py 128 20 5 999 > gpurun_out/ncu_rd.log 2>&1
tail -3 gpurun_out/ncu_rd.log
