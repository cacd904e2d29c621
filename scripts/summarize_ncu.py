#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into profiles/<name>.json + .md.

usage: python scripts/summarize_ncu.py gpurun_out/prof_X.ncu-rep profiles/r01_X [launches.csv]
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT:
                d[h] = "%s %s" % (r[i], units[i])
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]] = float(r[i])
                except ValueError:
                    pass
        d["warp_stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        kernels.append(d)
    launches = None
    if len(sys.argv) > 3:
        launches = []
        for r in csv.reader(open(sys.argv[3])):
            if len(r) > 5 and r[0].isdigit():
                launches.append({"kernel": r[4].split("(")[0], "block": r[7], "grid": r[8], "ns": float(r[-1])})
    json.dump({"source": rep, "kernels": kernels, "launches": launches}, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("# ncu summary: %s\n\n" % rep)
        for d in kernels:
            f.write("## %s\n\n| metric | value |\n|---|---|\n" % d["kernel"][:110])
            for k, v in d.items():
                if k not in ("kernel", "warp_stall_cycles_per_issue"):
                    f.write("| %s | %s |\n" % (k, v))
            f.write("\nwarp stall reasons (cycles per issued instruction, top 8): %s\n\n" %
                    ", ".join("%s %.2f" % kv for kv in d["warp_stall_cycles_per_issue"].items()))
        if launches:
            tot = sum(x["ns"] for x in launches)
            f.write("## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold, serialised)\n\n")
            agg = {}
            for x in launches:
                a = agg.setdefault(x["kernel"], [0, 0.0])
                a[0] += 1
                a[1] += x["ns"]
            f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
            for kname, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("| %s | %d | %.1f | %.1f%% |\n" % (kname[:90], c, ns / 1e3, 100 * ns / tot))
    print("wrote", out + ".json", out + ".md")


if __name__ == "__main__":
    main()
