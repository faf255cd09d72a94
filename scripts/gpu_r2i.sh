#!/bin/bash
# Round 2, ninth GPU pass: tiled/ensemble tests at HEAD, host-side profile of PLSA.fit, f2 timing,
# bench lines for C1 / C3 / C5 (C5 also with the tiled passes: P(z|d) exceeds the L2 there).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_ensemble.py -m gpu -q 2>&1 | tail -8 | cut -c1-250
timeout 300 python scripts/host_profile.py 2>&1 | tail -25 | tee gpurun_out/r2i_host_profile.txt
timeout 300 python scripts/time_distances.py 2>&1 | tail -8 | tee gpurun_out/r2i_time_distances.txt
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
for CFG in C1 C3; do
  timeout 600 python bench.py --config $CFG --steps 50 --warmup 3 --no-c4 > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err
  tail -2 gpurun_out/bench_$CFG.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$CFG.json").read().strip().splitlines()[-1])
    print("$CFG value %.3e ms/step %.4f e2e %.3e frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()}, "cpu", d["cpu_baseline"] and d["cpu_baseline"]["ms_per_iter"])
except Exception as e:
    print("$CFG failed", e)
PY
done
run5() { TAG=$1; shift
  env "$@" timeout 1500 python bench.py --config C5 --steps 30 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 5 --e2e-repeats 1 > gpurun_out/bench_C5_$TAG.json 2> gpurun_out/bench_C5_$TAG.err
  tail -2 gpurun_out/bench_C5_$TAG.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_C5_$TAG.json").read().strip().splitlines()[-1])
    print("C5 $TAG value %.3e ms/step %.4f e2e_s %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["seconds"]), {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
except Exception as e:
    print("C5 $TAG failed", e)
PY
}
run5 default
run5 tiled_doc ENSTOP_B200_TILED=1
run5 tiled_both ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1
