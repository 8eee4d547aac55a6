#!/bin/bash
# A/B harness: runs bench.py once per library variant under build/ (same C ABI, different compile-time options) and
# prints one summary line per variant.  usage: tools/ab_bench.sh [bench.py args] -- name1 name2 ...
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
for v in "$@"; do
  JJS_B200_LIB=$PWD/build/lib_$v.so python bench.py --no-cpu-baseline "${args[@]}" > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || { echo "$v FAILED"; tail -3 gpurun_out/ab_$v.err; continue; }
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.load(open(f"gpurun_out/ab_{v}.json"))
st = d["roofline"]["stage_ms_per_step"]
print(f"{v:12s} value {d['value']/1e6:7.3f} M/s  e2e {d['e2e']['value']/1e6:7.3f}  ms {d['ms_per_step']:7.2f}  " + " ".join(f"{k}={x:.2f}" for k, x in st.items()))
PY
done
