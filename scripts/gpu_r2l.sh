#!/bin/bash
# Round 2, twelfth GPU pass: prefetch depth of the tile kernel (mode 0: next batch into L2; mode 1:
# two batches ahead into L2), at C2 (doc side) and C5.
mkdir -p gpurun_out
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
run() { # tag, cfg, env...
  TAG=$1; CFG=$2; shift 2
  env "$@" timeout 1500 python bench.py --config $CFG --steps 30 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 5 --e2e-repeats 1 > gpurun_out/r2l_$TAG.json 2> gpurun_out/r2l_$TAG.err
  tail -2 gpurun_out/r2l_$TAG.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_$TAG.json").read().strip().splitlines()[-1])
    print("$TAG ms/iter %.4f value %.3e" % (d["ms_per_step"], d["value"]), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
except Exception as e:
    print("$TAG failed", e)
PY
}
run c2_doc_pf0 C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0
run c2_doc_pf1 C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0 ENSTOP_B200_LIB=$PWD/build/libplsa_pf1.so
run c2_doc_pf0b C2 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0
run c5_pf0 C5
run c5_pf1 C5 ENSTOP_B200_LIB=$PWD/build/libplsa_pf1.so
