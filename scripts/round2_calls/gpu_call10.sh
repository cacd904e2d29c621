#!/bin/bash
# round 2, call 10 (8 GPUs): staging ceiling of the box, multi-device context over 8 GPUs, torchrun bench at N=4 and N=8
mkdir -p gpurun_out
{ nvidia-smi -L; nvidia-smi topo -m; nproc; free -g | head -2; ls /sys/devices/system/node/ | grep node; } > gpurun_out/r02_8gpu_box.txt 2>&1
tools/h2d_ceiling 256 10 > gpurun_out/r02_8gpu_h2d.jsonl 2>&1; cat gpurun_out/r02_8gpu_h2d.jsonl
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_8gpu_multi_tests.txt
cat gpurun_out/r02_8gpu_multi_tests.txt
for n in 8 4; do
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 5 ) > gpurun_out/r02_8gpu_bench_n$n.json 2> gpurun_out/r02_8gpu_bench_n$n.err
tail -c 300 gpurun_out/r02_8gpu_bench_n$n.err
done
for wl in lev_k8_len128 rdamerau_k16_len512 search_n32_h4096; do
  for g in 8 4; do
    timeout 300 python bench.py --inproc --gpus $g --workload $wl --steps 10 --warmup 3 2>> gpurun_out/r02_8gpu_inproc.err | grep '^{' >> gpurun_out/r02_8gpu_inproc.jsonl
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_8gpu_inproc.jsonl'):
    d=json.loads(l); print(d['config']['name'], d['n_gpus'], round(d['ms_per_step'],3), '%.3g'%d['units_per_s'], round(d['e2e']['h2d_gbs'],1), d['parity_ok'], d['uses_nccl'], d['needle_broadcasts'])
for n in (8,4):
    try:
        d=json.loads(open('gpurun_out/r02_8gpu_bench_n%d.json'%n).read().strip().splitlines()[-1])
        print(n, 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['staging'])
    except Exception as ex: print('ERR', n, ex)
PY
