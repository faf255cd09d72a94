#!/bin/bash
# First GPU measurement of two changes made after round 1's GPU budget was spent:
#  * work-item launch order (plsa_set_option "item_order"): C2 / C3 at the default chunk,
#    order 0 vs 1, twice each (run-to-run spread), then the L1 / L2 counters of both orders;
#  * the flush-to-zero threshold build (scripts/build_variants.sh -> build/libplsa_ftz.so,
#    run build_variants.sh BEFORE gpurun): parity suite with that library, then the same A/B.
#   bash scripts/build_variants.sh && gpurun --timeout 2400 -- 'bash scripts/gpu_item_order.sh'
mkdir -p gpurun_out
for CFG in C2 C3; do
  bash scripts/gpu_ab.sh $CFG "d 1 0 1 0" "d 1 0 1 1" "d 1 0 1 2" "d 1 0 1 0" "d 1 0 1 1" "d 1 0 1 2"
done
bash scripts/gpu_ab.sh C2 "d 1 192 1 2" "d 1 320 1 2" "d 1 192 1 1" "d 1 320 1 1"   # item length x order
AB_TIMEOUT=1500 bash scripts/gpu_ab.sh C5 "d 1 0 1 0" "d 1 0 1 1" "d 1 0 1 2"   # order 2 is meant for this size
if [ -f build/libplsa_ftz.so ]; then
  ENSTOP_B200_LIB=$PWD/build/libplsa_ftz.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_ftz.log
  tail -3 gpurun_out/pytest_ftz.log
  for CFG in C2 C3; do
    bash scripts/gpu_ab.sh $CFG "ftz 1 0 1 0" "ftz 1 0 1 1" "d 1 0 1 0"
  done
fi
for L in ftz128_9 d128_9; do
  if [ -f build/libplsa_$L.so ]; then
    ENSTOP_B200_LIB=$PWD/build/libplsa_$L.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_$L.log
    tail -2 gpurun_out/pytest_$L.log
    bash scripts/gpu_ab.sh C2 "$L 1 0 1 0" "$L 1 0 1 1" "d 1 0 1 0"
  fi
done
for ORD in 0 1 2; do
  ENSTOP_B200_ITEM_ORDER=$ORD timeout 600 ncu --clock-control none -k regex:row_pass -s 9 -c 2 --csv \
    --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum \
    --log-file gpurun_out/item_order_${ORD}_ncu.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-iters 1 --e2e-repeats 1 \
    > gpurun_out/item_order_${ORD}_ncu.log 2>&1
  grep -E "row_pass|Metric" gpurun_out/item_order_${ORD}_ncu.csv | cut -c1-240 | tail -12
done
