#!/usr/bin/env python3
"""profiles/ncu_counts.json: per workload, the dominant kernel's warp instructions and DRAM bytes per launch, read from
the `ncu --set full` captures (run on the CPU box with `ncu -i`).  bench.py reports them as roofline.traffic and
roofline.issue (instruction-issue roofline: warp instructions / (SMs x 4 schedulers x SM clock x launch time)).
usage: python scripts/ncu_counts.py   (edit CAPTURES below when a kernel changes)"""
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# workload -> (capture, units per launch in that capture)
CAPTURES = {
    "lev_k8_len128": ("gpurun_out/prof_final_lev_k8_len128.ncu-rep", 1_000_000),
    "lev_k16_len128": ("gpurun_out/prof_lev_k16_len128.ncu-rep", 1_000_000),
    "lev_k8_len128_R": ("gpurun_out/prof_lev_k8_len128_R.ncu-rep", 1_000_000),
    "rdamerau_k16_len512": ("gpurun_out/prof_rdamerau_k16_len512.ncu-rep", 1_000_000),
    "lev_k16_len4096": ("gpurun_out/prof_fr6_lev_k16_len4096.ncu-rep", 262_144),
    "exp_len1024": ("gpurun_out/prof_fr6_exp_len1024.ncu-rep", 1_000_000),
    "affine_k16_len128": ("gpurun_out/prof_diag16v3_affine_k16_len128.ncu-rep", 1_000_000),
    "search_n32_h4096": ("gpurun_out/prof_search_qgram_n32_h4096.ncu-rep", 100_000),
}


def main():
    out = {}
    for wl, (rep, units) in CAPTURES.items():
        path = os.path.join(ROOT, rep)
        if not os.path.exists(path):
            continue
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = rows[0]
        best = None
        for r in rows[2:]:
            t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
            if best is None or t > best[0]:
                best = (t, r)
        t, r = best

        def num(name):
            return float(r[hdr.index(name)].replace(",", ""))

        def to_bytes(name):
            unit = rows[1][hdr.index(name)]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            return num(name) * mult
        out[wl] = {"kernel": r[hdr.index("Kernel Name")].split("(")[0][:80], "units": units,
                   "warp_instructions": int(num("smsp__inst_executed.sum")),
                   "dram_bytes": int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")),
                   "time_under_ncu_us": t * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3,
                                             "ms": 1e3, "second": 1e6}.get(rows[1][hdr.index("gpu__time_duration.sum")], 1.0),
                   "source": os.path.basename(rep)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_counts.json"), "w"), indent=1)
    for k, v in out.items():
        print(k, v)


if __name__ == "__main__":
    main()
