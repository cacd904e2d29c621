#!/bin/bash
# usage: quick_bench.sh workload...   (prints one compact line per workload; env passes through)
for w in "$@"; do python bench.py --workload $w --no-cpu-baseline --no-e2e --steps ${STEPS:-50} | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"config\"][\"name\"], \"ms\", round(d[\"ms_per_step\"],4), \"units/s\", int(d[\"pairs_per_s\"]), \"GB/s\", round(d[\"roofline\"][\"achieved\"],1), \"frac\", round(d[\"roofline\"][\"frac\"],4), d[\"parity_ok\"])"; done
