"""List the loops (backward branches) of a kernel in `cuobjdump -sass` output with their instruction mix.

usage: cuobjdump -sass <binary> | python tools/sass_loops.py <kernel-name-substring> [min_body]
"""
import re
import sys
from collections import Counter

want = sys.argv[1]
min_body = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ins = []
on = False
for line in sys.stdin:
    if "Function :" in line:
        on = want in line
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt >= a or tgt not in addr_index:
        continue
    body = ins[addr_index[tgt]: i + 1]
    if len(body) < min_body:
        continue
    c = Counter()
    for _, bt in body:
        op = bt.split()[1] if bt.startswith("@") else bt.split()[0]
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LD", "ST")) and "." in op else "")
        c[op] += 1
    f64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr, fp64 {f64}, issue~{len(body) + f64}: " +
          ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
