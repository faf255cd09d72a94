#!/bin/bash
# Experiment: lane-group mappings of the row pass (ENSTOP_B200_LANES).  Parity subset + C2 bench each.
mkdir -p gpurun_out
for L in 0 4 2 1; do
  echo "=== lanes=$L"
  ENSTOP_B200_LANES=$L timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or split or repeatable or scale" 2>&1 | tail -2
  ENSTOP_B200_LANES=$L timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --profile-iters 20 > gpurun_out/bench_lanes$L.json 2> gpurun_out/bench_lanes$L.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_lanes$L.json"))
print("lanes=$L ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
done
for L in 0 16 8; do
  ENSTOP_B200_LANES=$L timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline --profile-iters 5 > gpurun_out/bench_c3_lanes$L.json 2> gpurun_out/bench_c3_lanes$L.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c3_lanes$L.json"))
print("C3 lanes=$L ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
done
