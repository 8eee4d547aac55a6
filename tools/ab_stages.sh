#!/bin/bash
# A/B harness for kernel work: per-stage device times of every library variant build/lib_<name>.so given on the command line
# (same C ABI, different compile-time options), statuses checked against the constructed expectation.
#   usage: tools/ab_stages.sh <variants, e.g. 0,1,2> <log file> name1 name2 ...
variants=$1; log=$2; shift 2
for v in "$@"; do
  JJS_B200_LIB=$PWD/build/lib_$v.so python tools/time_stages.py $variants >> $log 2>&1 || echo "$v FAILED" >> $log
done
