timeout 900 python -m pytest tests/test_gpu_scan_rd.py -x -q 2>&1 | tail -5
for d in 0 3; do
  CRBM_RD_DEBUG=$d CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 128 20 1 999
  CRBM_RD_DEBUG=$d timeout 120 python scratch/scan_ab.py 128 20 5 999
done
CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 100 15 1 200
CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 512 25 1 1000
timeout 120 python scratch/scan_ab.py 100 15 2 200
timeout 120 python scratch/scan_ab.py 512 25 5 1004
