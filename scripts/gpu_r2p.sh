#!/bin/bash
# 1-GPU pass: presort-during-upload + sliced seeded init: tests, fit phases, e2e.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiled.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-250 | tee gpurun_out/r2p_pytest.log
DRAIN=0 timeout 300 python scripts/fit_phases.py C2 pinned 2>&1 | tail -8 | tee gpurun_out/r2p_fit_phases_pinned.txt
DRAIN=1 timeout 300 python scripts/fit_phases.py C2 pinned 2>&1 | tail -8 | tee gpurun_out/r2p_fit_phases_pinned_drain.txt
DRAIN=0 timeout 300 python scripts/fit_phases.py C2 pageable 2>&1 | tail -8 | tee gpurun_out/r2p_fit_phases_pageable.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-c4 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2p_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.4f e2e %.3e (%s ms, first %.2f ms) frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], [round(1e3*x,2) for x in d["e2e"]["seconds_all_runs"]], 1e3*d["e2e"]["first_call_seconds"], d["roofline"]["frac"]))
PY
