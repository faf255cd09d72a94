#!/bin/bash
# Round 2, thirteenth GPU pass: half-row mapping of the tile kernel (16 lanes per item) at 512 /
# 640 / 768 threads against the octet mapping; e2e with page-locked inputs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiled.py -m gpu -q 2>&1 | tail -5 | cut -c1-250
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4 | cut -c1-250
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
run() { # tag, cfg, env...
  TAG=$1; CFG=$2; shift 2
  env "$@" timeout 1500 python bench.py --config $CFG --steps 30 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 5 --e2e-repeats 1 > gpurun_out/r2m_$TAG.json 2> gpurun_out/r2m_$TAG.err
  tail -2 gpurun_out/r2m_$TAG.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2m_$TAG.json").read().strip().splitlines()[-1])
    print("$TAG ms/iter %.4f value %.3e e2e_s %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["seconds"]), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
except Exception as e:
    print("$TAG failed", e)
PY
}
run c2_octet C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0 ENSTOP_B200_LIB=$PWD/build/libplsa_octet.so
run c2_half512 C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0 ENSTOP_B200_LIB=$PWD/build/libplsa_half512.so
run c2_half640 C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0
run c2_half768 C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0 ENSTOP_B200_LIB=$PWD/build/libplsa_half768.so
run c2_both_half640 C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1 ENSTOP_B200_TERM_TILE_MIN=24
run c5_half640 C5
run c5_half768 C5 ENSTOP_B200_LIB=$PWD/build/libplsa_half768.so
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r2m_bench_pinned.json 2> gpurun_out/r2m_bench_pinned.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-c4 --pageable > gpurun_out/r2m_bench_pageable.json 2> gpurun_out/r2m_bench_pageable.err
python - <<'PY'
import json
for t in ("pinned", "pageable"):
    try:
        d=json.loads(open("gpurun_out/r2m_bench_%s.json" % t).read().strip().splitlines()[-1])
        print(t, "e2e ms", [round(1e3*x,2) for x in d["e2e"]["seconds_all_runs"]], "first", round(1e3*d["e2e"]["first_call_seconds"],1))
    except Exception as e:
        print(t, "failed", e)
PY
timeout 300 python scripts/time_set_factors.py 2>&1 | tail -5
