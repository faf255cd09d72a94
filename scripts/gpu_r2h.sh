#!/bin/bash
# Round 2, eighth GPU pass: threshold debug, full GPU test suite at HEAD, default bench.
mkdir -p gpurun_out
timeout 600 python scripts/debug_thresh.py 2>&1 | tail -40 | tee gpurun_out/r2h_debug_thresh.txt
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_loglik or baseline" > gpurun_out/pytest_fused.log 2>&1
grep -E "where False|passed|failed" gpurun_out/pytest_fused.log | cut -c1-700 | head -6
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-250 | tee gpurun_out/pytest_all_r2h.log
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
tail -3 gpurun_out/bench_r2h.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2h.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.4f e2e %.3e (%.2f ms, first %.2f ms) frac %.3f c4 %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], 1e3*d["e2e"]["seconds"], 1e3*d["e2e"]["first_call_seconds"], d["roofline"]["frac"], d.get("c4_wall_s")))
print(d["roofline"]["kernel_ms_per_iter"]); print(d.get("parity")); print(d.get("parity_checks")); print(d["cpu_baseline"])
PY
