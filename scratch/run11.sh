timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload c3 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2h_c3.json 2> gpurun_out/r2h_c3.err
python bench.py --workload c4 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2h_c4.json 2> gpurun_out/r2h_c4.err
timeout 120 python scratch/scan_ab.py 128 20 1 999
timeout 120 python scratch/scan_ab.py 100 15 1 200
timeout 120 python scratch/scan_ab.py 512 25 1 1000
CRBM_SCAN_KERNEL=pos timeout 120 python scratch/scan_ab.py 512 25 1 1000
CRBM_SCAN_KERNEL=pos timeout 120 python scratch/scan_ab.py 100 15 1 200
CRBM_SCAN_KERNEL=pos timeout 120 python scratch/scan_ab.py 200 25 1 1000
timeout 120 python scratch/scan_ab.py 128 20 5 999
timeout 120 python scratch/scan_ab.py 100 15 1 200 0 reduce
timeout 120 python scratch/scan_ab.py 512 25 1 1000 0 reduce
