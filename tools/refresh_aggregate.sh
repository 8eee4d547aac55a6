#!/bin/bash
# Refresh of the aggregate-path evidence after a change to k_aggregate / k_agg_coeffs only: ncu capture of the two kernels, their
# executed multiplies merged into profiles/<base>_executed_mac32.json (on the box and, from gpurun_out/, here), then the bench
# lines of the two workloads that use them.     usage: tools/refresh_aggregate.sh <base tag, e.g. r02g> <new tag>
set -u
base=$1; tag=$2; out=gpurun_out; mkdir -p $out
python bench.py --workload aggregate --log2n 17 --steps 1 --warmup 1 --no-cpu-baseline --no-strong > $out/${tag}_plain_agg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_aggregate|k_agg_coeffs' --launch-skip 2 --launch-count 2 \
    -o $out/${tag}_prof_agg -f python bench.py --workload aggregate --log2n 17 --steps 1 --warmup 1 --no-cpu-baseline --no-strong > $out/${tag}_ncu_agg.log 2>&1
ncu -i $out/${tag}_prof_agg.ncu-rep --page raw --csv > $out/${tag}_prof_agg.raw.csv 2>/dev/null
for k in k_aggregate k_agg_coeffs; do
  ncu -i $out/${tag}_prof_agg.ncu-rep --page source --csv --kernel-name regex:"$k" > $out/${tag}_src_agg_$k.csv 2>/dev/null
done
JJS_PROFILE_OUT=$out/${tag}_profiles python - "$base" "$tag" "$out" <<'PY'
import json, os, subprocess, sys
base, tag, out = sys.argv[1:4]
d = json.loads([l for l in open(f"{out}/{tag}_plain_agg.log") if l.startswith("{")][-1])
keys = (d["e2e"]["h2d_bytes_per_step"] - (1 << 17) * 100 - 4) // 32
specs = [f"aggregate:{k}:{n}:{out}/{tag}_src_agg_{k}.csv" for k, n in (("k_agg_coeffs", keys), ("k_aggregate", 1 << 17))]
subprocess.check_call([sys.executable, "tools/ncu_executed.py", tag] + specs)
new = json.load(open(f"{out}/{tag}_profiles/{tag}_executed_mac32.json"))
path = f"profiles/{base}_executed_mac32.json"
cur = json.load(open(path))
for sec in new:
    if isinstance(new[sec], dict) and "aggregate" in new[sec]:
        cur.setdefault(sec, {})["aggregate"] = new[sec]["aggregate"]
cur["aggregate_refreshed_by"] = tag
json.dump(cur, open(path, "w"), indent=1)
json.dump(cur, open(f"{out}/{tag}_merged_executed_mac32.json", "w"), indent=1)
print(cur["per_unit"]["aggregate"])
PY
rm -f $out/${tag}_src_*.csv $out/${tag}_prof_*.ncu-rep
for w in aggregate mixed4; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-strong > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err || echo "bench $w failed"
done
python - "$tag" <<'PY'
import json, sys
for w in ("aggregate", "mixed4"):
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench_{w}.json")); ex = d["roofline"].get("executed") or {}
    print(w, round(d["value"] / 1e6, 2), round(d["e2e"]["value"] / 1e6, 2), round(d["e2e_pageable"]["value"] / 1e6, 2), "exec", round(ex.get("frac", 0), 3), d["roofline"]["stage_ms_per_step"])
PY
