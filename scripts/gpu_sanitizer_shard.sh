#!/bin/bash
# compute-sanitizer memcheck over a 2-GPU document-sharded fit (one-shot and two-shot exchange
# over peer memory, and ncclAllReduce).  Log -> gpurun_out/sanitizer/.
mkdir -p gpurun_out/sanitizer
LOG=gpurun_out/sanitizer/memcheck_sharded_2gpu.log
timeout 420 compute-sanitizer --tool memcheck --print-limit 400 --error-exitcode 9 \
  python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k "two_shot_exchange or one_rank_shard_path_matches" > $LOG 2>&1
echo "memcheck sharded rc=$?" | tee -a gpurun_out/sanitizer/summary_sharded.txt
grep -E "ERROR SUMMARY|passed|failed|Error|error" $LOG | tail -8 | cut -c1-300 | tee -a gpurun_out/sanitizer/summary_sharded.txt
