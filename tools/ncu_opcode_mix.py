"""Executed-instruction mix of one kernel from an ncu report's source page:
    ncu -i report.ncu-rep --page source --csv --kernel-name regex:k_equation > src.csv
    python tools/ncu_opcode_mix.py src.csv [units]
Sums "Instructions Executed" (warp level) by SASS opcode and groups the opcodes by the pipe they issue on."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
names = rows[hdr]
si, ei = names.index("Source"), names.index("Instructions Executed")
by = collections.Counter()
samples = collections.Counter()
smp = names.index("# Samples")
for r in rows[hdr + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):
        break       # next launch of the report: the first one is enough
    if len(r) <= ei:
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
    if not m:
        continue
    op = m.group(1)
    by[op] += int(r[ei] or 0)
    samples[op] += int(r[smp] or 0)
total = sum(by.values())


def grp(op):
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
        return "multiply pipe: wide/high multiplies"
    if op.startswith("IMAD") or op.startswith("IMUL"):
        return "multiply pipe: other IMAD forms (moves, carries, adds)"
    if op.startswith("DFMA") or op.startswith("DADD") or op.startswith("DMUL"):
        return "fp64 pipe"
    if op.startswith(("LD", "ST", "ATOM", "RED")):
        return "memory"
    if op.startswith(("BRA", "CALL", "RET", "EXIT", "BSSY", "BSYNC", "WARPSYNC")):
        return "control"
    return "alu and the rest"


g = collections.Counter()
for op, n in by.items():
    g[grp(op)] += n
print(f"warp instructions executed: {total}" + (f"  ({total * 32 / units:.0f} thread instructions per unit)" if units else ""))
for k, n in g.most_common():
    print(f"  {k:55s} {n:14d}  {100 * n / total:5.1f} %")
print("top opcodes:")
for op, n in by.most_common(14):
    print(f"  {op:28s} {n:14d}  {100 * n / total:5.1f} %   stall samples {samples[op]}")
