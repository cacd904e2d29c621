#!/bin/bash
# every bench workload once (kernel-only, e2e, CPU baseline) -> gpurun_out/bench_all.jsonl
mkdir -p gpurun_out
: > gpurun_out/bench_all.jsonl
for w in lev_k8_len128 lev_k16_len128 lev_k8_len128_R lev_k16_len128_R lev_k8_ragged96_160 rdamerau_k16_len512 exp_len1024 search_n32_h4096 lev_k16_len4096 lev_k60_len1024 affine_k16_len128 hamming_len64 hamming_len4096; do
  python bench.py --workload $w --steps ${STEPS:-50} --warmup 5 2>/dev/null | tail -1 >> gpurun_out/bench_all.jsonl
done
TA_FORCE_BAND=1 python bench.py --workload lev_k8_len128 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | sed 's/"name": "lev_k8_len128"/"name": "lev_k8_len128 (general lev_band_kernel forced)"/' >> gpurun_out/bench_all.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/bench_all.jsonl"):
    d = json.loads(l)
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    print("%-48s %8.3f ms %12.0f units/s %9.1f GCUPS  %7.1f GB/s frac %.4f | e2e %10.0f units/s | cpu %9.0f units/s x%s | parity %s" % (
        d["config"]["name"], d["ms_per_step"], d["pairs_per_s"], d["value"], d["roofline"]["achieved"], d["roofline"]["frac"],
        e.get("pairs_per_s", 0), c.get("pairs_per_s", 0), c.get("cores"), d["parity_ok"]))
PY
