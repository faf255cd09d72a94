#!/bin/bash
# Round 2, second GPU pass: the tiled doc pass (plsa_tile.cuh) — parity suite with the tiled path
# forced on, then C2 with and without it.
mkdir -p gpurun_out
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_tiled.log
tail -25 gpurun_out/pytest_tiled.log
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
for T in 1 0; do
  ENSTOP_B200_TILED=$T timeout 300 python bench.py --config C2 --steps 50 --warmup 3 --no-cpu-baseline --profile-iters 10 --e2e-repeats 1 > gpurun_out/r2b_c2_tiled$T.json 2> gpurun_out/r2b_c2_tiled$T.err
  tail -3 gpurun_out/r2b_c2_tiled$T.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_c2_tiled$T.json"))
    print("tiled=$T ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()}, "ll", d["config"]["ll_first_last"], "e2e_s", d["e2e"]["seconds"])
except Exception as e:
    print("tiled=$T failed", e)
PY
done
