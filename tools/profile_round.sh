#!/bin/bash
# Round profile on ONE B200 (run through gpurun): every bench workload, the ncu launch list of the default bench
# command, and one `ncu --set full` capture per kernel (after the same command has exited 0 without ncu).
# Outputs land in gpurun_out/; tools/ncu_summarize.py turns the reports into profiles/<tag>_ncu_summary.md.
#   usage: tools/profile_round.sh <tag>         e.g. r01b
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
for w in single double vargen aggregate mixed4 mixed5; do
  python bench.py --workload $w --steps 5 --warmup 3 > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err || echo "bench $w failed"
done
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err || echo "reference arm failed"
# launch list of the default bench command (cold-cache, serialised launches: shares must agree with the stage timers)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
# full captures at 2^18 items (ncu replays every kernel ~40 times)
python bench.py --log2n 18 --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_plain18.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_decode|k_challenge|k_equation|k_rtest' --launch-skip 5 --launch-count 5 \
    -o $out/${tag}_prof_single -f python bench.py --log2n 18 --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_single.log 2>&1
python bench.py --workload aggregate --log2n 17 --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_plain_agg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_aggregate|k_agg_coeffs' --launch-skip 2 --launch-count 2 \
    -o $out/${tag}_prof_agg -f python bench.py --workload aggregate --log2n 17 --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_agg.log 2>&1
for r in single agg; do
  [ -f $out/${tag}_prof_$r.ncu-rep ] && ncu -i $out/${tag}_prof_$r.ncu-rep --page raw --csv > $out/${tag}_prof_$r.raw.csv 2>/dev/null
done
ls -la $out | grep $tag
