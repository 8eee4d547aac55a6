#!/usr/bin/env python3
"""Static SASS view of one kernel of a cubin / object / shared library: the kernel body and every subroutine it calls
(ptxas places non-inlined device functions inside the calling kernel), each with its opcode histogram and its cost on the
integer-multiply pipe under the measured issue model (IMAD.WIDE / IMAD.HI: 4 cycles per warp instruction per scheduler,
any other IMAD form: 2).  usage: tools/sass_fn.py <file> <kernel-name-substring> [--dump]"""
import collections
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, fns = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fns[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m and cur:
            fns[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return fns


def split_subroutines(instrs):
    """[(start_addr, [instr...])]: the kernel body first, then one block per call target."""
    targets = sorted({int(m.group(1), 16) for _, t in instrs for m in [re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", t)] if m})
    bounds = [instrs[0][0]] + targets + [instrs[-1][0] + 16]
    blocks = []
    for a, b in zip(bounds, bounds[1:]):
        blocks.append((a, [(x, t) for x, t in instrs if a <= x < b]))
    return blocks


def opcode(t):
    parts = t.split()
    if parts[0].startswith("@"):
        parts = parts[1:]
    return parts[0]


def summarize(block):
    h = collections.Counter(opcode(t) for _, t in block)
    wide = sum(v for k, v in h.items() if k.startswith("IMAD.WIDE") or k.startswith("IMAD.HI"))
    other = sum(v for k, v in h.items() if k.startswith("IMAD") and not (k.startswith("IMAD.WIDE") or k.startswith("IMAD.HI")))
    return h, wide, other


if __name__ == "__main__":
    path, name = sys.argv[1], sys.argv[2]
    fns = functions(path)
    for fn, instrs in fns.items():
        if name not in fn or not instrs:
            continue
        print("kernel", fn[:100], len(instrs), "instructions")
        for start, block in split_subroutines(instrs):
            h, wide, other = summarize(block)
            calls = collections.Counter(m.group(1) for _, t in block for m in [re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", t)] if m)
            print(f"  block @0x{start:x}: {len(block)} instr, IMAD.WIDE {wide}, other IMAD {other} (pipe cycles {4 * wide + 2 * other}), calls {dict(calls)}")
            print("     ", ", ".join(f"{k} {v}" for k, v in h.most_common(14)))
            if "--dump" in sys.argv:
                for x, t in block:
                    print(f"        {x:06x} {t}")
