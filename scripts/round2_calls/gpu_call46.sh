#!/bin/bash
# round 2, call 46: larger q-gram queue (16 entries per 4 KB): search tests incl. a match in every haystack, stress, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or Search" 2>&1 | tail -3
TA_SEARCH_FILTER=qgram timeout 200 python scripts/stress_search.py 45 21 2>&1 | tail -1
timeout 200 python bench.py --workload search_n32_h4096 --steps 30 --warmup 3 --no-cpu-baseline --no-configs 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['parity_ok'], d['e2e']['ms_per_step'])"
