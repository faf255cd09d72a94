#!/bin/bash
# Experiment: row-pass kernel variants (bit0 = texture gather, bit1 = coalesced entries + shuffles).
mkdir -p gpurun_out
for V in 1 2 3; do
  echo "=== variant=$V"
  ENSTOP_B200_VARIANT=$V timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or split or repeatable or scale or width" 2>&1 | tail -2
  ENSTOP_B200_VARIANT=$V timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --profile-iters 20 > gpurun_out/bench_var$V.json 2> gpurun_out/bench_var$V.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_var$V.json"))
print("variant=$V ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
done
for V in 2 3; do
  ENSTOP_B200_VARIANT=$V timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline --profile-iters 5 > gpurun_out/bench_c3_var$V.json 2> gpurun_out/bench_c3_var$V.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c3_var$V.json"))
print("C3 variant=$V ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
  ENSTOP_B200_VARIANT=$V timeout 300 python bench.py --config C1 --steps 50 --warmup 3 --no-cpu-baseline --profile-iters 20 > gpurun_out/bench_c1_var$V.json 2> gpurun_out/bench_c1_var$V.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c1_var$V.json"))
print("C1 variant=$V ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
done
