#!/usr/bin/env python3
"""Executed 32x32->64 multiplies per unit of each kernel, from ncu source pages (csv) of this build:
    tools/ncu_executed.py <tag> <kind>:<kernel>:<units>:<source.csv> ...
writes profiles/<tag>_executed_mac32.json = {"per_unit": {kind: {kernel: IMAD.WIDE + IMAD.HI thread instructions per unit}}, ...}
and one profiles/<tag>_<kind>_<kernel>_opcode_mix.txt per input (the opcode histogram, grouped by issue pipe).
bench.py multiplies these by the units of a step for roofline.executed."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mix(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    names = rows[hdr]
    si, ei = names.index("Source"), names.index("Instructions Executed")
    by = collections.Counter()
    for r in rows[hdr + 1:]:
        if r and r[0] in ("Address", "Kernel Name"):
            break
        if len(r) <= ei:
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        if m:
            by[m.group(1)] += int(r[ei] or 0)
    return by


def group(op):
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
        return "multiply pipe: wide/high multiplies"
    if op.startswith("IMAD"):
        return "multiply pipe: other IMAD forms (moves, carries, adds)"
    if op[0] == "D" and op[:4] in ("DFMA", "DADD", "DMUL", "DSET"):
        return "fp64 pipe"
    if op.startswith(("LD", "ST", "ATOM", "RED")):
        return "memory"
    if op.startswith(("BRA", "CALL", "RET", "EXIT", "BSSY", "BSYNC", "WARPSYNC")):
        return "control"
    return "alu and the rest"


if __name__ == "__main__":
    tag = sys.argv[1]
    outdir = os.environ.get("JJS_PROFILE_OUT", os.path.join(ROOT, "profiles"))   # on the GPU box: a directory under gpurun_out/
    os.makedirs(outdir, exist_ok=True)
    out = {"per_unit": {}, "instructions_per_unit": {}, "non_multiply_imad_share": {}, "how": "ncu --set full --import-source on, source page, 'Instructions Executed' (warp level) x 32 / units"}
    for spec in sys.argv[2:]:
        kind, kernel, units, path = spec.split(":", 3)
        units = int(units)
        by = mix(path)
        total = sum(by.values())
        groups = collections.Counter()
        for op, v in by.items():
            groups[group(op)] += v
        wide = groups["multiply pipe: wide/high multiplies"]
        out["per_unit"].setdefault(kind, {})[kernel] = wide * 32.0 / units
        out["instructions_per_unit"].setdefault(kind, {})[kernel] = total * 32.0 / units
        out["non_multiply_imad_share"].setdefault(kind, {})[kernel] = groups["multiply pipe: other IMAD forms (moves, carries, adds)"] / max(1, total)
        with open(os.path.join(outdir, f"{tag}_{kind}_{kernel}_opcode_mix.txt"), "w") as f:
            f.write(f"warp instructions executed: {total}  ({total * 32.0 / units:.0f} thread instructions per unit, {units} units)\n")
            for g, v in groups.most_common():
                f.write(f"  {g:58s} {v:12d}  {100.0 * v / max(1, total):5.1f} %\n")
            f.write("top opcodes:\n")
            for op, v in by.most_common(16):
                f.write(f"  {op:28s} {v:14d}  {100.0 * v / max(1, total):5.1f} %\n")
    with open(os.path.join(outdir, f"{tag}_executed_mac32.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["per_unit"]))
