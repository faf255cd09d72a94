#!/bin/bash
# Sweep the work-item length and the gather path (texture vs LDG) on C2; C1/C3 at defaults.
mkdir -p gpurun_out
for TC in "1 128" "1 256" "1 384" "1 512" "0 256"; do set -- $TC; T=$1; C=$2
  ENSTOP_B200_TEXTURE=$T ENSTOP_B200_CHUNK=$C timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --profile-iters 20 > gpurun_out/bench_t${T}c$C.json 2> gpurun_out/bench_t${T}c$C.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_t${T}c$C.json"))
print("texture=$T chunk=$C ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
done
for CFG in C1 C3; do
  timeout 300 python bench.py --config $CFG --steps 20 --warmup 3 --no-cpu-baseline --profile-iters 10 > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$CFG.json"))
print("$CFG ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
done
