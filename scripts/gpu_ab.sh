#!/bin/bash
# A/B of row-pass variants on one box: usage gpu_ab.sh CONFIG "LIB VEC CHUNK [TEX]" ...
# LIB: d = enstop_b200/libplsa_b200.so, anything else = build/libplsa_<LIB>.so
mkdir -p gpurun_out
export ENSTOP_B200_CORPUS_CACHE=${ENSTOP_B200_CORPUS_CACHE:-/dev/shm}   # generate each corpus once per box
CFG=$1; shift
for V in "$@"; do set -- $V; L=$1; VEC=$2; C=$3; TX=${4:-1}
  LIBP=""; [ "$L" != "d" ] && LIBP=$PWD/build/libplsa_$L.so
  TAG=${CFG}_${L}_v${VEC}_c${C}_t${TX}
  ENSTOP_B200_LIB=$LIBP ENSTOP_B200_TEXTURE=$TX ENSTOP_B200_VEC=$VEC ENSTOP_B200_CHUNK=$C timeout ${AB_TIMEOUT:-300} python bench.py --config $CFG --steps 50 --warmup 3 --no-cpu-baseline --profile-iters 10 --e2e-repeats 1 > gpurun_out/ab_$TAG.json 2> gpurun_out/ab_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$TAG.json"))
    print("$TAG ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()}, "ll", d["config"]["ll_first_last"])
except Exception as e:
    print("$TAG failed", e)
PY
done
