#!/bin/bash
# Round 2, eleventh GPU pass: tile kernel with the three-deep header pipeline (far L2 / near L1
# prefetch): tests, C2 and C5 A/B, host profile.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiled.py -m gpu -q 2>&1 | tail -5 | cut -c1-250
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4 | cut -c1-250
timeout 300 python scripts/host_profile.py 2>&1 | grep -E "steady|front|end to end" | head -8 | tee gpurun_out/r2k_host_profile.txt
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
run() { # tag, cfg, env...
  TAG=$1; CFG=$2; shift 2
  env "$@" timeout 1500 python bench.py --config $CFG --steps 30 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 5 --e2e-repeats 1 > gpurun_out/r2k_$TAG.json 2> gpurun_out/r2k_$TAG.err
  tail -2 gpurun_out/r2k_$TAG.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2k_$TAG.json").read().strip().splitlines()[-1])
    print("$TAG ms/iter %.4f value %.3e" % (d["ms_per_step"], d["value"]), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
except Exception as e:
    print("$TAG failed", e)
PY
}
run c2_off C2 ENSTOP_B200_TILED=0
run c2_doc C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0
run c2_doc_t176 C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0 ENSTOP_B200_TILE_KB=176
run c2_both C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1 ENSTOP_B200_TERM_TILE_MIN=24
run c5_auto C5
run c5_t176 C5 ENSTOP_B200_TILE_KB=176
