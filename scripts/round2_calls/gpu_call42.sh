#!/bin/bash
# round 2, call 42: default bench line with the three extra search workloads in `configs` (wall time), Hamming dev_len test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_entry or hamming" 2>&1 | tail -2
( time python bench.py > gpurun_out/r02_c42_bench_default.json 2> gpurun_out/r02_c42_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c42_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['bench_wall_s'])
for c in d['configs']: print(c['name'], round(c['ms_per_step'],4), round(c['frac_hbm'],4), c['kernel'][:30], c['parity_ok'])
PY
