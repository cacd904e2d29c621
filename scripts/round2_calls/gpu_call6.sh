#!/bin/bash
# round 2, call 6 (2 GPUs): multi-device context + NCCL tests, new bench.py at N=1 and N=2 (torchrun), in-process multi ctx,
# staging ceiling, diagonal-extension kernel v4 timings
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_c6_gpus.txt; nvidia-smi topo -m >> gpurun_out/r02_c6_gpus.txt 2>&1; nproc >> gpurun_out/r02_c6_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_c6_multi_tests.txt
cat gpurun_out/r02_c6_multi_tests.txt
tools/h2d_ceiling 256 10 > gpurun_out/r02_c6_h2d.jsonl 2>&1; cat gpurun_out/r02_c6_h2d.jsonl
for wl in lev_k16_len4096 exp_len1024; do
    echo "== $wl TA_FR=1" >> gpurun_out/r02_c6_bench.txt
    TA_FR=1 timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c6_bench.txt
done
cat gpurun_out/r02_c6_bench.txt
( time timeout 600 python bench.py --steps 100 --warmup 5 ) > gpurun_out/r02_c6_bench_n1.json 2> gpurun_out/r02_c6_bench_n1.err
tail -c 400 gpurun_out/r02_c6_bench_n1.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 ) > gpurun_out/r02_c6_bench_n2.json 2> gpurun_out/r02_c6_bench_n2.err
tail -c 400 gpurun_out/r02_c6_bench_n2.err
for wl in lev_k8_len128 rdamerau_k16_len512 search_n32_h4096; do
  for g in 1 2; do
    timeout 300 python bench.py --inproc --gpus $g --workload $wl --steps 10 --warmup 3 >> gpurun_out/r02_c6_inproc.jsonl 2>> gpurun_out/r02_c6_inproc.err
  done
done
cat gpurun_out/r02_c6_inproc.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config']['name'], d['n_gpus'], round(d['ms_per_step'],3), d['units_per_s'], d['e2e']['h2d_gbs'], d['parity_ok'], d['uses_nccl'], d['needle_broadcasts'])"
