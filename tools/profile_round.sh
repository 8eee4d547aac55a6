#!/bin/bash
# Round profile on ONE B200 (run through gpurun): every bench workload, the ncu launch list of the default bench
# command, and one `ncu --set full` capture per kernel and variant (after the same command has exited 0 without ncu).
# Outputs land in gpurun_out/; here, tools/ncu_summarize.py and tools/ncu_executed.py turn the exported pages into
# profiles/<tag>_*.   usage: tools/profile_round.sh <tag>         e.g. r02
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_single.json 2> $out/${tag}_bench_single.err || echo "bench single failed"
for w in double vargen aggregate mixed4 mixed5; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-strong > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err || echo "bench $w failed"
done
python bench.py --impl reference --steps 5 --warmup 2 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err || echo "reference arm failed"
# launch list of the default bench workload (cold-cache, serialised launches: shares must agree with the stage timers)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > $out/${tag}_ncu_launches.log 2>&1
# full captures at 2^18 items per variant (ncu replays every kernel ~40 times); the second verify call is captured
for v in 0 1 2; do
  python tools/time_stages.py $v 18 > $out/${tag}_units_$v.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'k_decode|k_challenge|k_equation|k_rtest' --launch-skip 4 --launch-count 4 \
      -o $out/${tag}_prof_v$v -f python tools/time_stages.py $v 18 > $out/${tag}_ncu_v$v.log 2>&1
done
python bench.py --workload aggregate --log2n 17 --steps 1 --warmup 1 --no-cpu-baseline --no-strong > $out/${tag}_plain_agg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_aggregate|k_agg_coeffs' --launch-skip 2 --launch-count 2 \
    -o $out/${tag}_prof_agg -f python bench.py --workload aggregate --log2n 17 --steps 1 --warmup 1 --no-cpu-baseline --no-strong > $out/${tag}_ncu_agg.log 2>&1
for r in v0 v1 v2 agg; do
  [ -f $out/${tag}_prof_$r.ncu-rep ] || continue
  ncu -i $out/${tag}_prof_$r.ncu-rep --page raw --csv > $out/${tag}_prof_$r.raw.csv 2>/dev/null
  for k in k_decode k_challenge k_equation k_aggregate k_agg_coeffs; do
    ncu -i $out/${tag}_prof_$r.ncu-rep --page source --csv --kernel-name regex:"$k" > $out/${tag}_src_${r}_$k.csv 2>/dev/null
    [ -s $out/${tag}_src_${r}_$k.csv ] || rm -f $out/${tag}_src_${r}_$k.csv
  done
done
# executed-instruction mix per kernel and variant, computed here: the source pages and the reports are too big to travel (64 MiB)
JJS_PROFILE_OUT=$out/${tag}_profiles python - "$tag" "$out" <<'PY'
import ast, os, re, subprocess, sys
tag, out = sys.argv[1], sys.argv[2]
specs = []
for v, kind in ((0, "single"), (1, "double"), (2, "vargen")):
    try:
        line = [l for l in open(f"{out}/{tag}_units_{v}.log") if "units" in l][-1]
        units = ast.literal_eval(line[line.index("units") + 5:].strip())
    except Exception as e:
        print("no units for", kind, e)
        continue
    for k, n in units.items():
        path = f"{out}/{tag}_src_v{v}_{k}.csv"
        if os.path.exists(path):
            specs.append(f"{kind}:{k}:{n // 4 if False else n}:{path}")
# the captures ran at 2^18 items: time_stages printed the units of THAT run
# aggregate path (bench.py --workload aggregate --log2n 17): k_agg_coeffs runs one thread per signer key, k_aggregate one per item
try:
    import json
    d = json.loads([l for l in open(f"{out}/{tag}_plain_agg.log") if l.startswith("{")][-1])
    keys = (d["e2e"]["h2d_bytes_per_step"] - (1 << 17) * (64 + 32 + 4) - 4) // 32
    for k, n in (("k_agg_coeffs", keys), ("k_aggregate", 1 << 17)):
        path = f"{out}/{tag}_src_agg_{k}.csv"
        if os.path.exists(path):
            specs.append(f"aggregate:{k}:{n}:{path}")
except Exception as e:
    print("no aggregate units:", e)
subprocess.check_call([sys.executable, "tools/ncu_executed.py", tag] + specs)
PY
rm -f $out/${tag}_src_*.csv $out/${tag}_prof_*.ncu-rep
ls -la $out | grep $tag
