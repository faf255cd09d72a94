#!/bin/bash
# Round 2, fourth GPU pass: tiled doc + term pass with the lane-major entry blocks and deep entry
# prefetch; entry-stream L2 prefetch in the group-per-row kernel (A/B against build/libplsa_nopf.so).
mkdir -p gpurun_out
ENSTOP_B200_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_tiled.log
tail -25 gpurun_out/pytest_tiled.log | cut -c1-300
ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_tiled_doc_only.log
tail -4 gpurun_out/pytest_tiled_doc_only.log | cut -c1-300
ENSTOP_B200_TILED=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_untiled.log
tail -4 gpurun_out/pytest_untiled.log | cut -c1-300
export ENSTOP_B200_CORPUS_CACHE=/dev/shm
run() { # tag, env...
  TAG=$1; shift
  env "$@" timeout 300 python bench.py --config C2 --steps 50 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 10 --e2e-repeats 1 > gpurun_out/r2d_$TAG.json 2> gpurun_out/r2d_$TAG.err
  tail -2 gpurun_out/r2d_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_$TAG.json"))
    print("$TAG ms/iter %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms_per_iter"].items()}, "ll", d["config"]["ll_first_last"], "e2e_s", round(d["e2e"]["seconds"],5))
except Exception as e:
    print("$TAG failed", e)
PY
}
run off_nopf ENSTOP_B200_TILED=0 ENSTOP_B200_LIB=$PWD/build/libplsa_nopf.so
run off ENSTOP_B200_TILED=0
run doc ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=0
run both ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1
run both_min24 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1 ENSTOP_B200_TERM_TILE_MIN=24
run both_min6 ENSTOP_B200_TILED=1 ENSTOP_B200_TERM_TILED=1 ENSTOP_B200_TERM_TILE_MIN=6
ENSTOP_B200_TILED=1 timeout 600 ncu --clock-control none -k regex:"tile_pass|row_pass|fixup|normalise|colsum|compact" -s 20 -c 12 --csv \
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,launch__grid_size \
  --log-file gpurun_out/r2d_ncu_counters.csv \
  python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline --no-c4 --profile-iters 1 --e2e-repeats 1 > gpurun_out/r2d_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2d_ncu_counters.csv")) if len(r)>10]
hdr=rows[0]
try:
    ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
    out={}
    for r in rows[1:]:
        out.setdefault((r[ii], r[ki][:48]), {})[r[mi]]=r[vi]
    short={"gpu__time_duration.sum":"ns","smsp__inst_executed.sum":"inst","smsp__issue_active.avg.pct_of_peak_sustained_active":"issue%","sm__warps_active.avg.pct_of_peak_sustained_active":"warps%","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed":"lsu%","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum":"bankconf","lts__t_sectors_srcunit_tex_op_read.sum":"l2sect","l1tex__t_sector_hit_rate.pct":"l1hit%","dram__bytes_read.sum":"dramR","dram__bytes_write.sum":"dramW","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio":"longsb","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio":"shortsb","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio":"barrier","launch__grid_size":"grid"}
    for k,v in out.items():
        print(k, {short.get(a,a): b for a,b in v.items()})
except Exception as e:
    print("ncu parse failed", e)
PY
