#!/usr/bin/env python3
"""Instruction mix of the largest loop of a kernel (per DP column): scripts/sass_loop.py <obj> <name-substring> [cols]

Reads `cuobjdump -sass`, finds the longest backward-branch loop body and prints total / ALU-pipe / FMA-pipe / LSU
instruction counts per column -- the ALU pipe (LOP3, SHF, PRMT, SEL, IADD3, LEA ... at 16 lanes/clk/SMSP) is what
binds the bit-parallel kernels, so this is the number to look at before spending GPU time.
"""
import re
import subprocess
import sys
from collections import Counter

obj, sub = sys.argv[1], sys.argv[2]
cols = int(sys.argv[3]) if len(sys.argv) > 3 else 32
names = [l.split()[2] for l in subprocess.check_output(["cuobjdump", "-sass", obj], text=True).splitlines()
         if "Function :" in l and sub in l]
for fn in names:
    txt = subprocess.check_output(["cuobjdump", "-sass", "-fun", fn, obj], text=True)
    lines = [l for l in txt.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    adr = lambda l: int(re.match(r"\s+/\*([0-9a-f]{4})\*/", l).group(1), 16)
    best = None
    for l in lines:
        m = re.search(r"BRA (0x[0-9a-f]+)", l)
        if m:
            t, a = int(m.group(1), 16), adr(l)
            if t < a and (best is None or a - t > best[1] - best[0]) and a - t < 0x8000:
                best = (t, a)
    body = [l for l in lines if best[0] <= adr(l) <= best[1]]
    c = Counter()
    for l in body:
        t = l.split()
        op = t[1] if not t[1].startswith("@") else t[2]
        c[op.rstrip(";")] += 1
    base = Counter()
    for k, v in c.items():
        base[k.split(".")[0]] += v
    fma = base["IMAD"] + base["FFMA"]
    lsu = sum(base[k] for k in ("LDS", "STS", "LDG", "STG", "ATOMS", "LDGSTS"))
    ctl = sum(base[k] for k in ("BRA", "BSSY", "BSYNC", "NOP", "EXIT"))
    alu = len(body) - fma - lsu - ctl
    print("%s\n  loop %d instr = %.1f/col: alu %.1f fma %.1f lsu %.1f | %s" % (
        fn[-60:], len(body), len(body) / cols, alu / cols, fma / cols, lsu / cols,
        " ".join("%s=%d" % kv for kv in c.most_common(14))))
