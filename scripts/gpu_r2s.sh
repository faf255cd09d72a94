#!/bin/bash
# Final 1-GPU pass at HEAD: all GPU tests, smoke(), default bench, reference arm, launch list and
# full ncu capture of the default C2 path.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-250 | tee gpurun_out/r2s_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r2s_smoke.log
timeout 600 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2s_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print("steps %d value %.3e ms/step %.4f e2e %.3e (%.2f ms, first %.2f ms) frac %.3f c4 %s launches %s" % (d["steps"], d["value"], d["ms_per_step"], d["e2e"]["value"], 1e3*d["e2e"]["seconds"], 1e3*d["e2e"]["first_call_seconds"], d["roofline"]["frac"], d.get("c4_wall_s"), d.get("gpu_launches")))
print(d["roofline"]["kernel_ms_per_iter"], d["clocks"])
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r2s_bench_steps20.json 2> gpurun_out/r2s_bench_steps20.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2s_bench_steps20.json").read().strip().splitlines()[-1])
print("steps %d e2e %s ms" % (d["steps"], [round(1e3*x,2) for x in d["e2e"]["seconds_all_runs"]]))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r2s.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 2 \
    > gpurun_out/ncu_launches_r2s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_pass -s 9 -c 3 \
    -o gpurun_out/prof_r2s python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 1 --e2e-repeats 1 \
    > gpurun_out/ncu_full_r2s.log 2>&1
tail -2 gpurun_out/ncu_full_r2s.log | cut -c1-200
