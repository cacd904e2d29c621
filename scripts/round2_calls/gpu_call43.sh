#!/bin/bash
# round 2, call 43 (2 GPUs): the bench line with the extended configs array under torchrun, both arms
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_c43_bench_n2.json 2> gpurun_out/r02_c43_bench_n2.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c43_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['bench_wall_s'], d['e2e']['ms_per_step'])
for c in d['configs']: print(c['name'], c['scaling'], round(c['ms_per_step'],4), c['parity_ok'])
PY
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_c43_bench_ref_n2.json 2> gpurun_out/r02_c43_bench_ref_n2.err ) 2>&1 | grep real
tail -c 200 gpurun_out/r02_c43_bench_ref_n2.json
