set -x
timeout 900 python -m pytest tests/test_gpu_scan_rd.py -x -q 2>&1 | tail -5
for cfg in "128 20 5 999" "100 15 2 200" "128 20 10 999" "512 25 5 1004"; do
  timeout 120 python scratch/scan_ab.py $cfg
done
for cfg in "100 15 1 200" "512 25 1 1000" "128 20 1 999"; do
  CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py $cfg
done
