timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z_dp2_20.json 2> gpurun_out/r2z_dp2_20.err
