#!/bin/bash
# Round 2, seventh GPU pass: balanced dealing + overflow-safe quotient in the tile kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiled.py -m gpu -q 2>&1 | tail -60 | cut -c1-260 > gpurun_out/pytest_gpu_tiled.log
grep -E "^FAILED|passed|failed|AssertionError: \(" gpurun_out/pytest_gpu_tiled.log | head -30
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_loglik" > gpurun_out/pytest_fused.log 2>&1
grep -E "where False|passed|failed" gpurun_out/pytest_fused.log | cut -c1-900 | head -6
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -6 | cut -c1-200
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
run() { # tag, env...
  TAG=$1; shift
  env "$@" timeout 300 python bench.py --config C2 --steps 30 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 10 --e2e-repeats 1 > gpurun_out/r2g_$TAG.json 2> gpurun_out/r2g_$TAG.err
  tail -2 gpurun_out/r2g_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2g_$TAG.json"))
    print("$TAG ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
except Exception as e:
    print("$TAG failed", e)
PY
}
run doc ENSTOP_B200_TILED=1
run doc_t160 ENSTOP_B200_TILED=1 ENSTOP_B200_TILE_KB=160
run doc_t160_pfl1 ENSTOP_B200_TILED=1 ENSTOP_B200_TILE_KB=160 ENSTOP_B200_LIB=$PWD/build/libplsa_tpfl1.so
run doc_t128_pfl1 ENSTOP_B200_TILED=1 ENSTOP_B200_TILE_KB=128 ENSTOP_B200_LIB=$PWD/build/libplsa_tpfl1.so
run both_min24 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1 ENSTOP_B200_TERM_TILE_MIN=24
ENSTOP_B200_TILED=1 timeout 600 ncu --clock-control none -k regex:"tile_pass" -s 4 -c 2 --csv \
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio \
  --log-file gpurun_out/r2g_ncu_counters.csv \
  python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 1 --e2e-repeats 1 > gpurun_out/r2g_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2g_ncu_counters.csv")) if len(r)>10]
hdr=rows[0]
ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
out={}
for r in rows[1:]:
    out.setdefault((r[ii], r[ki][:40]), {})[r[mi].split("__")[-1][:44]]=r[vi]
for k,v in out.items(): print(k, v)
PY
