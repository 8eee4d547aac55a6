#!/bin/bash
# Bench lines of every workload on the in-tree build (the committed profiles/<tag>_bench_*.json), plus smoke().
tag=${1:-rXX}; out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_single.json 2> $out/${tag}_bench_single.err || echo "bench single failed"
for w in double vargen aggregate mixed4 mixed5; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-strong > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err || echo "bench $w failed"
done
python bench.py --impl reference --steps 5 --warmup 2 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err || echo "reference arm failed"
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for w in ("single", "double", "vargen", "aggregate", "mixed4", "mixed5"):
    d = json.load(open(f"gpurun_out/{tag}_bench_{w}.json"))
    ex = d["roofline"].get("executed") or {}
    print(w, round(d["value"] / 1e6, 2), round(d["e2e"]["value"] / 1e6, 2), round(d["e2e_pageable"]["value"] / 1e6, 2), d["roofline"]["kernel"],
          "canon", round(d["roofline"]["frac"], 3), "exec", round(ex.get("frac", 0), 3), "step", round(ex.get("step", {}).get("frac", 0), 3), ex.get("source"))
PY
