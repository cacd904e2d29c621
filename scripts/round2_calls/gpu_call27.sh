#!/bin/bash
# round 2, call 27 (8 GPUs): end-of-round validation of the N = 8 paths on the final library -- multi-device tests, the
# driver's bench line at N = 8 and N = 4 (both arms), one in-library multi-device search
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_c27_multi_tests.txt
cat gpurun_out/r02_c27_multi_tests.txt
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_c27_bench_n$N.json 2> gpurun_out/r02_c27_bench_n$N.err
  tail -c 300 gpurun_out/r02_c27_bench_n$N.json; echo
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02_c27_bench_ref_n8.json 2> gpurun_out/r02_c27_bench_ref_n8.err
tail -c 300 gpurun_out/r02_c27_bench_ref_n8.json; echo
for wl in search_n32_h4096 lev_k8_len128; do
  timeout 300 python bench.py --inproc --gpus 8 --workload $wl --steps 10 --warmup 3 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/r02_c27_inproc.jsonl
done
tail -c 600 gpurun_out/r02_c27_inproc.jsonl
