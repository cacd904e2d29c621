#!/bin/bash
# round 2, call 39 (2 GPUs): multi-device tests on the final library (length hint on a multi-device context, search, NCCL)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_c39_multi_tests.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_c39_bench_n2.json 2> gpurun_out/r02_c39_bench_n2.err
tail -c 200 gpurun_out/r02_c39_bench_n2.json
