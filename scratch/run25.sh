timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_baseline_shapes.py -x -q -k "fused_scan or scan_c5 or position_mean" 2>&1 | tail -3
NOMEAN=1 timeout 120 python scratch/scan_ab.py 100 15 1 200 0 reduce
NOMEAN=1 timeout 120 python scratch/scan_ab.py 512 25 1 1000 0 reduce
NOMEAN=1 timeout 120 python scratch/scan_ab.py 512 25 1 1000 2 reduce
