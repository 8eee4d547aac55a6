#!/bin/bash
# One GPU call of an A/B round: stage times of every build/lib_<name>.so for the variants given, then (optionally) the aggregate
# workload through bench.py for the same libraries, then the GPU suite on the in-tree library.
#   usage: tools/ab_round.sh <tag> <variants e.g. 0,1,2> <agg: 0|1> <tests: 0|1|expr> name1 name2 ...
tag=$1; variants=$2; agg=$3; tests=$4; shift 4
out=gpurun_out; mkdir -p $out
log=$out/${tag}_stages.log; : > $log
tools/ab_stages.sh $variants $log "$@"
cat $log
if [ "$agg" = "1" ]; then
  tools/ab_bench.sh --workload aggregate --steps 3 --warmup 3 --no-strong -- "$@" | tee $out/${tag}_agg.log
fi
if [ "$tests" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/${tag}_tests.log
elif [ "$tests" != "0" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$tests" 2>&1 | tail -15 | tee $out/${tag}_tests.log
fi
