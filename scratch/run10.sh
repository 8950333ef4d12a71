for cfg in "200 25 1 1000" "64 12 1 500" "384 25 1 1000" "300 15 1 200" "150 20 1 500"; do
  timeout 120 python scratch/scan_ab.py $cfg
  CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py $cfg
done
timeout 120 python scratch/scan_ab.py 100 15 1 200 0 reduce
timeout 120 python scratch/scan_ab.py 512 25 1 1000 0 reduce
timeout 120 python scratch/scan_ab.py 128 20 5 999 0 reduce
