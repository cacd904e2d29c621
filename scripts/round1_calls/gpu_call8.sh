#!/bin/bash
# 2-GPU call: weak-scaling bench lines (torchrun, one rank per GPU) + the exponential-search schedule on 1 GPU
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/c8_bench_2gpu.json 2> gpurun_out/c8_bench_2gpu.err
tail -c 1500 gpurun_out/c8_bench_2gpu.json
$TR bench.py --gpus 2 --workload search_n32_h4096 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c8_search_2gpu.json 2>> gpurun_out/c8_bench_2gpu.err
tail -c 700 gpurun_out/c8_search_2gpu.json
$TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/c8_reference_2gpu.json 2>> gpurun_out/c8_bench_2gpu.err
tail -c 300 gpurun_out/c8_reference_2gpu.json
export STEPS=50
{
echo "== exp default (first round k=16)"; bash scripts/quick_bench.sh exp_len1024
echo "== exp reference schedule (k=30 first)"; TA_EXP_FIRST_K=30 bash scripts/quick_bench.sh exp_len1024
} > gpurun_out/c8_variants.log 2>&1
cat gpurun_out/c8_variants.log
python -m pytest tests -m gpu -x -q -k "exp or dist" 2>&1 | tail -2
