"""Per-instruction stall samples of one ncu source page (csv): prints, for an address range, every instruction with its executed
count, samples and the dominant stall reasons; then a per-range summary.  usage: ncu_hot.py src.csv <lo_hex> <hi_hex> [--sum]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
names = rows[hdr]
idx = {n: i for i, n in enumerate(names)}
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
stalls = [n for n in names if n.startswith("stall_") and "Not Issued" not in n]
base = None
tot = {s: 0 for s in stalls}
nsamp = 0
for r in rows[hdr + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):
        break
    try:
        a = int(r[0], 16)
    except Exception:
        continue
    if base is None:
        base = a
    off = a - base
    if not (lo <= off < hi):
        continue
    s = int(r[idx["# Samples"]] or 0)
    nsamp += s
    parts = []
    for st in stalls:
        v = int(r[idx[st]] or 0)
        tot[st] += v
        if v and v >= 0.15 * max(1, s):
            parts.append(f"{st[6:]}={v}")
    if "--sum" not in sys.argv:
        print(f"{off:06x} {r[idx['Source']][:70]:70s} exec={r[idx['Instructions Executed']]:>9s} samp={s:5d} {' '.join(parts)}")
print("range samples", nsamp, {k[6:]: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
