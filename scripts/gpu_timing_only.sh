#!/bin/bash
mkdir -p gpurun_out
for V in "$@"; do
  ENSTOP_B200_VARIANT=$V timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --profile-iters 20 > gpurun_out/bench_t$V.json 2> gpurun_out/bench_t$V.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_t$V.json"))
print("variant=$V ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
done
