for d in 0 1 2 3; do
  CRBM_RD_DEBUG=$d CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 128 20 1 999
  CRBM_RD_DEBUG=$d CRBM_RD_STORE=bulk CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 128 20 1 999
  CRBM_RD_DEBUG=$d timeout 120 python scratch/scan_ab.py 128 20 5 999
done
