#!/bin/bash
# Round 2, tenth GPU pass: everything at HEAD — all GPU tests, sanitizer, distances timing,
# default bench, launch list and full ncu captures of the default C2 path.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-250 | tee gpurun_out/pytest_all_r2j.log
timeout 300 python scripts/time_distances.py 2>&1 | tail -8 | tee gpurun_out/r2j_time_distances.txt
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err
tail -3 gpurun_out/bench_r2j.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2j.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.4f e2e %.3e (%.2f ms, first %.2f ms) frac %.3f c4 %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], 1e3*d["e2e"]["seconds"], 1e3*d["e2e"]["first_call_seconds"], d["roofline"]["frac"], d.get("c4_wall_s")))
print(d["roofline"]["kernel_ms_per_iter"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r2j.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 2 \
    > gpurun_out/ncu_launches_r2j.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:row_pass -s 9 -c 3 \
    -o gpurun_out/prof_r2j python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 1 --e2e-repeats 1 \
    > gpurun_out/ncu_full_r2j.log 2>&1
tail -2 gpurun_out/ncu_full_r2j.log | cut -c1-200
bash scripts/gpu_sanitizer.sh 2>&1 | tail -20
