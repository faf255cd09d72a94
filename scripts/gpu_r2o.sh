#!/bin/bash
# 1-GPU pass at HEAD: all GPU tests, smoke(), default bench, reference arm, fit phases.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-250 | tee gpurun_out/r2o_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r2o_smoke.log
timeout 600 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2o_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.4f e2e %.3e (%.2f ms, first %.2f ms) frac %.3f c4 %s fit %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], 1e3*d["e2e"]["seconds"], 1e3*d["e2e"]["first_call_seconds"], d["roofline"]["frac"], d.get("c4_wall_s"), d.get("c4_fit_wall_s"), d.get("gpu_launches")))
print(d["roofline"]["kernel_ms_per_iter"], d["clocks"], d["cpu_baseline"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2o_bench_reference.json 2> gpurun_out/r2o_bench_reference.err
echo "reference rc=$?"; tail -1 gpurun_out/r2o_bench_reference.json | cut -c1-600
timeout 300 python scripts/fit_phases.py C2 pinned 2>&1 | tail -8 | tee gpurun_out/r2o_fit_phases_pinned.txt
timeout 300 python scripts/fit_phases.py C2 pageable 2>&1 | tail -8 | tee gpurun_out/r2o_fit_phases_pageable.txt
