for cfg in "128 25 1 1000" "256 20 1 999" "512 15 1 200" "64 12 1 500" "384 25 1 1000" "200 25 1 1000"; do
  timeout 120 python scratch/scan_ab.py $cfg
  CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py $cfg
done
