set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "two or devices" 2>&1 | tail -3
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02h_bench_single_${n}gpu.json 2> gpurun_out/r02h_bench_single_${n}gpu.err || tail -5 gpurun_out/r02h_bench_single_${n}gpu.err
python - <<EOF
import json
d=json.load(open("gpurun_out/r02h_bench_single_${n}gpu.json"))
s=d["strong_2p24"]
print($n, "weak", d["value"], d["e2e"]["value"], "strong", s["value"], s["e2e"]["value"], s["e2e_pageable"]["value"])
EOF
done
