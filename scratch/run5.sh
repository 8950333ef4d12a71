set -x
timeout 900 python -m pytest tests/test_gpu_scan_rd.py -x -q 2>&1 | tail -5
for cfg in "100 15 2 200" "128 20 4 999"; do
  timeout 120 python scratch/scan_ab.py $cfg
  CRBM_RD_STORE=bulk timeout 120 python scratch/scan_ab.py $cfg
done
for cfg in "100 15 1 200" "512 25 1 1000" "128 20 1 999"; do
  CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py $cfg
done
CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 100 15 1 200 2
CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 512 25 1 1000 2
