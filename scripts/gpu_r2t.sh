#!/bin/bash
# C5 and C3 bench lines at HEAD (tiled passes chosen automatically at C5).
mkdir -p gpurun_out
for CFG in C3 C5; do
  timeout 420 python bench.py --config $CFG --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r2t_bench_$CFG.json 2> gpurun_out/r2t_bench_$CFG.err
  echo "$CFG rc=$?"; tail -2 gpurun_out/r2t_bench_$CFG.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2t_bench_$CFG.json").read().strip().splitlines()[-1])
print("$CFG value %.3e ms/step %.4f e2e %.1f ms frac %.3f" % (d["value"], d["ms_per_step"], 1e3*d["e2e"]["seconds"], d["roofline"]["frac"]), d["roofline"]["kernel_ms_per_iter"])
PY
done
