DP_CHECK_SHAPE=tc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/dp_p2p_check.py > gpurun_out/r2l_p2p8_tc.log 2>&1
tail -3 gpurun_out/r2l_p2p8_tc.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2l_dp8.json 2> gpurun_out/r2l_dp8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 > gpurun_out/r2l_dp4.json 2> gpurun_out/r2l_dp4.err
