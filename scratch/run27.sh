timeout 900 python -m pytest tests/test_gpu_scan_rd.py tests/test_gpu_scan_tc.py -x -q 2>&1 | tail -2
timeout 120 python scratch/scan_ab.py 100 15 1 200 0 reduce
timeout 120 python scratch/scan_ab.py 512 25 1 1000 0 reduce
timeout 120 python scratch/scan_ab.py 128 20 1 999
timeout 120 python scratch/scan_ab.py 100 15 1 200
timeout 120 python scratch/scan_ab.py 512 25 1 1000
timeout 120 python scratch/scan_ab.py 128 20 5 999
timeout 120 python scratch/scan_ab.py 100 15 2 200
