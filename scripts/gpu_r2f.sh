#!/bin/bash
# Round 2, sixth GPU pass: what bounds the tile kernel?  Ablation builds (no LDS / no
# accumulation / no entry stream), 768-thread CTAs, a full ncu capture; test details.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiled.py -m gpu -q 2>&1 | tail -60 | cut -c1-260 > gpurun_out/pytest_gpu_tiled.log
grep -E "^FAILED|passed|failed|AssertionError: \(" gpurun_out/pytest_gpu_tiled.log | head -30
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_loglik" 2>&1 | grep -E "array|passed|failed" | cut -c1-700 | head -8
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
run() { # tag, env...
  TAG=$1; shift
  env "$@" timeout 300 python bench.py --config C2 --steps 30 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 10 --e2e-repeats 1 > gpurun_out/r2f_$TAG.json 2> gpurun_out/r2f_$TAG.err
  tail -2 gpurun_out/r2f_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f_$TAG.json"))
    print("$TAG ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()})
except Exception as e:
    print("$TAG failed", e)
PY
}
run doc ENSTOP_B200_TILED=1
for V in abl1 abl2 abl3 t768; do
  run doc_$V ENSTOP_B200_TILED=1 ENSTOP_B200_LIB=$PWD/build/libplsa_$V.so
done
ENSTOP_B200_TILED=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 6 -c 1 \
    -o gpurun_out/prof_r2f_tile python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 1 --e2e-repeats 1 \
    > gpurun_out/ncu_full_r2f.log 2>&1
tail -2 gpurun_out/ncu_full_r2f.log | cut -c1-200
ncu -i gpurun_out/prof_r2f_tile.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    for k,v in zip(hdr,r):
        if any(t in k for t in ('issue_stalled','pipe_','gpu__time','warps_active','issue_active','inst_executed.sum','l1tex__data_pipe','throughput.avg.pct')) and 'per_issue' in k or 'pipe_fma' in k or 'pipe_alu' in k or 'pipe_lsu' in k or 'pipe_xu' in k or 'gpu__time_duration.sum' in k or k in ('smsp__issue_active.avg.pct_of_peak_sustained_active',):
            print(k, v)
"
