timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload c3 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2g_c3.json 2> gpurun_out/r2g_c3.err
python bench.py --workload c4 --only-headline --no-cpu-baseline --no-scan --no-e2e > gpurun_out/r2g_c4.json 2> gpurun_out/r2g_c4.err
CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 128 20 1 999
CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 100 15 1 200
CRBM_SCAN_KERNEL=rows timeout 120 python scratch/scan_ab.py 512 25 1 1000
timeout 120 python scratch/scan_ab.py 512 25 1 1000
timeout 120 python scratch/scan_ab.py 100 15 1 200
timeout 120 python scratch/scan_ab.py 128 20 5 999
