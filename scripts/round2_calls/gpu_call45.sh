#!/bin/bash
# round 2, call 45: final validation of the library as committed -- full GPU suite, smoke(), sanitizers, default bench line, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_c45_tests.txt
cat gpurun_out/r02_c45_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_round2.py > gpurun_out/r02_c45_memcheck.log 2>&1; tail -2 gpurun_out/r02_c45_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_round2.py > gpurun_out/r02_c45_racecheck.log 2>&1; tail -2 gpurun_out/r02_c45_racecheck.log
python bench.py > gpurun_out/r02_c45_bench_default.json 2> gpurun_out/r02_c45_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c45_bench_ref.json 2> gpurun_out/r02_c45_bench_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c45_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['bench_wall_s'])
for c in d['configs']: print(c['name'], round(c['ms_per_step'],4), round(c['frac_hbm'],4), c['kernel'][:34], c['parity_ok'])
r=json.loads(open('gpurun_out/r02_c45_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['ms_per_step'])
PY
