timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_baseline_shapes.py -x -q -k "fused_scan or scan_c5 or position_mean" 2>&1 | tail -2
CRBM_PINNED_RESULT_MAX=0 timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_scan_tc.py -x -q -k "fused_scan or activity_and" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3e.json 2> gpurun_out/r3e.err
