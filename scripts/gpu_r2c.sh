#!/bin/bash
# Round 2, third GPU pass: tiled doc + term pass — parity (forced on / default), C2 A/B with the
# per-kernel profile, ncu counters of the tile and tail kernels.
mkdir -p gpurun_out
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_tiled.log
tail -12 gpurun_out/pytest_tiled.log
ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_tiled_doc_only.log
tail -4 gpurun_out/pytest_tiled_doc_only.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_default.log
tail -4 gpurun_out/pytest_default.log
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
run() { # tag, env...
  TAG=$1; shift
  env "$@" timeout 300 python bench.py --config C2 --steps 50 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 10 --e2e-repeats 1 > gpurun_out/r2c_$TAG.json 2> gpurun_out/r2c_$TAG.err
  tail -2 gpurun_out/r2c_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_$TAG.json"))
    print("$TAG ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()}, "ll", d["config"]["ll_first_last"], "e2e_s", round(d["e2e"]["seconds"],5))
except Exception as e:
    print("$TAG failed", e)
PY
}
run off ENSTOP_B200_TILED=0
run doc ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0
run both ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1
run both_t100 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1 ENSTOP_B200_TILE_KB=100
run both_min24 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1 ENSTOP_B200_TERM_TILE_MIN=24
ENSTOP_B200_TILED=1 timeout 600 ncu --clock-control none -k regex:"tile_pass|row_pass" -s 12 -c 8 --csv \
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,launch__registers_per_thread,launch__grid_size \
  --log-file gpurun_out/r2c_ncu_counters.csv \
  python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 1 --e2e-repeats 1 > gpurun_out/r2c_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2c_ncu_counters.csv")) if len(r)>10]
hdr=rows[0]; 
try:
    ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
    out={}
    for r in rows[1:]:
        out.setdefault((r[ii], r[ki][:60]), {})[r[mi]]=r[vi]
    for k,v in out.items():
        print(k, {a.split("__")[-1][:38]: b for a,b in v.items()})
except Exception as e:
    print("ncu parse failed", e)
PY
