#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for VC in "3 128" "3 192" "3 256" "3 384" "1 2048" "1 256"; do set -- $VC; V=$1; C=$2
  ENSTOP_B200_VARIANT=$V ENSTOP_B200_CHUNK=$C timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --profile-iters 20 > gpurun_out/bench_c$C.json 2> gpurun_out/bench_c$C.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c$C.json"))
print("variant=$V chunk=$C ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()}, "e2e_s %.3f" % d["e2e"]["seconds"])
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
