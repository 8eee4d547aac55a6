#!/bin/bash
# After tools/profile_round.sh <tag> has come back from the GPU box: condense the raw ncu pages into profiles/<tag>_ncu_summary.md and
# profiles/<tag>_ncu_traffic.json, and copy the opcode mixes, the executed-multiply counts and the launch list into profiles/.
#   usage: tools/collect_profile.sh <tag>
set -e
tag=$1; out=gpurun_out
keys=$(python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/{tag}_plain_agg.log") if l.startswith("{")][-1])
print((d["e2e"]["h2d_bytes_per_step"] - (1 << 17) * 100 - 4) // 32)
PY
)
units() { python - "$1" "$2" <<'PY'
import ast, sys
line = [l for l in open(sys.argv[1]) if "units" in l][-1]
print(ast.literal_eval(line[line.index("units") + 5:].strip())[sys.argv[2]])
PY
}
u0=$out/${tag}_units_0.log
python tools/ncu_summarize.py $tag $out/${tag}_prof_v0.raw.csv $out/${tag}_prof_v2.raw.csv $out/${tag}_prof_agg.raw.csv \
    k_decode=$(units $u0 k_decode) k_challenge=$(units $u0 k_challenge) k_equation=$(units $u0 k_equation) \
    k_equation_vargen=$(units $out/${tag}_units_2.log k_equation) k_rtest=7000 k_agg_coeffs=$keys k_aggregate=131072 > /dev/null
cp $out/${tag}_profiles/* profiles/
cp $out/${tag}_launches.csv profiles/
ls profiles | grep "^${tag}_"
