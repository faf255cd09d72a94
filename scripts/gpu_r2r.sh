#!/bin/bash
# 1-GPU pass: all GPU tests, fit phases, init time line, default-size e2e.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-250 | tee gpurun_out/r2r_pytest_all.log
timeout 300 python scripts/fit_phases.py C2 pinned 2>&1 | tail -8 | tee gpurun_out/r2r_fit_phases_pinned.txt
timeout 300 python scripts/fit_phases.py C2 pageable 2>&1 | tail -8 | tee gpurun_out/r2r_fit_phases_pageable.txt
ENSTOP_B200_INIT_PROFILE=1 timeout 120 python scripts/time_init.py 2>&1 | tail -24 | tee gpurun_out/r2r_time_init.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2r_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.4f e2e %.3e (%s ms, first %.2f ms) frac %.3f c4 %s %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], [round(1e3*x,2) for x in d["e2e"]["seconds_all_runs"]], 1e3*d["e2e"]["first_call_seconds"], d["roofline"]["frac"], d.get("c4_wall_s"), d.get("c4_wall_s_all_runs")))
PY
