#!/bin/bash
mkdir -p gpurun_out
run() { # lib variant tag
  ENSTOP_B200_LIB=$1 ENSTOP_B200_VARIANT=$2 timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --profile-iters 20 > gpurun_out/bench_$3.json 2> gpurun_out/bench_$3.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$3.json"))
print("$3 ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
PY
}
run "" 1 base_v1
for U in 4 5 6; do
  ENSTOP_B200_LIB=$PWD/build/libplsa_u$U.so ENSTOP_B200_VARIANT=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or split or scale" 2>&1 | tail -1
  run $PWD/build/libplsa_u$U.so 1 u${U}_v1
  run $PWD/build/libplsa_u$U.so 5 u${U}_v5
  run $PWD/build/libplsa_u$U.so 0 u${U}_v0
done
