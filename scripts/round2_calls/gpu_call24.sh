#!/bin/bash
# round 2, call 24 (2 GPUs): multi-GPU tests and the N = 2 bench lines on the new search path; sanitizers over the round-2 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -2
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_c24_multi_tests.txt
cat gpurun_out/r02_c24_multi_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_c24_bench_n2.json 2> gpurun_out/r02_c24_bench_n2.err
tail -c 600 gpurun_out/r02_c24_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_c24_bench_ref_n2.json 2> gpurun_out/r02_c24_bench_ref_n2.err
tail -c 400 gpurun_out/r02_c24_bench_ref_n2.json
for wl in search_n32_h4096 search_n64_h4096; do
  timeout 300 python bench.py --inproc --gpus 2 --workload $wl --steps 10 --warmup 3 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/r02_c24_inproc.jsonl
done
tail -c 700 gpurun_out/r02_c24_inproc.jsonl
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_round2.py > gpurun_out/r02_c24_memcheck.log 2>&1; tail -3 gpurun_out/r02_c24_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python scripts/sanitize_round2.py > gpurun_out/r02_c24_racecheck.log 2>&1; tail -3 gpurun_out/r02_c24_racecheck.log
