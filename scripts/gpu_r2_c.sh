#!/bin/bash
# round 2, call C (2 GPUs): NCCL window-sharded path against the single-GPU result, then the 2-GPU bench line (lv_parity, configs4)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/mgpu_tests.log 2>&1
echo "2-gpu tests exit $?"; tail -n 5 gpurun_out/mgpu_tests.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench exit $?"; tail -c 1800 gpurun_out/bench_2gpu.json; tail -n 8 gpurun_out/bench_2gpu.err
