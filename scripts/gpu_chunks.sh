#!/bin/bash
# Sweep the work-item length and the gather path (texture vs LDG): usage gpu_chunks.sh CONFIG "T C" ...
mkdir -p gpurun_out
CFG=$1; shift
for TC in "$@"; do set -- $TC; T=$1; C=$2
  ENSTOP_B200_TEXTURE=$T ENSTOP_B200_CHUNK=$C timeout 300 python bench.py --config $CFG --steps 30 --warmup 3 --no-cpu-baseline --profile-iters 10 > gpurun_out/bench_${CFG}_t${T}c$C.json 2> gpurun_out/bench_${CFG}_t${T}c$C.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${CFG}_t${T}c$C.json"))
print("$CFG texture=$T chunk=$C ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
done
