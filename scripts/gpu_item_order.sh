#!/bin/bash
# First GPU measurement of the work-item launch orders (plsa_set_option "item_order"):
# C2 / C3 at the default chunk, order 0 vs 1, twice each (run-to-run spread), then the
# L1 / L2 counters of the term pass for both orders.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_item_order.sh'
mkdir -p gpurun_out
for CFG in C2 C3; do
  bash scripts/gpu_ab.sh $CFG "d 1 0 1 0" "d 1 0 1 1" "d 1 0 1 0" "d 1 0 1 1"
done
for ORD in 0 1; do
  ENSTOP_B200_ITEM_ORDER=$ORD timeout 600 ncu --clock-control none -k regex:row_pass -s 9 -c 2 --csv \
    --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum \
    --log-file gpurun_out/item_order_${ORD}_ncu.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-iters 1 --e2e-repeats 1 \
    > gpurun_out/item_order_${ORD}_ncu.log 2>&1
  grep -E "row_pass|Metric" gpurun_out/item_order_${ORD}_ncu.csv | cut -c1-240 | tail -12
done
