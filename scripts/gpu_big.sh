#!/bin/bash
# Large / wide configurations: C3 (k=128) and C5 (1M x 200k, 200M entries) — time and memory sanity.
mkdir -p gpurun_out
for CFG in C3 C5; do
  timeout 1500 python bench.py --config $CFG --steps 20 --warmup 3 --no-cpu-baseline --profile-iters 5 > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$CFG.json"))
    print("$CFG ms/iter %.4f value %.3e e2e_s %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["seconds"]), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()}, "iter roofline frac %.3f" % d["roofline"]["iteration"]["frac"])
except Exception as e:
    print("$CFG failed", e)
PY
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
