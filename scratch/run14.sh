timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2k_dp2.json 2> gpurun_out/r2k_dp2.err
