#!/bin/bash
# 2-GPU pass: sharded-fit tests (one-shot and two-shot exchange), then the sharded bench with
# either exchange forced.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_ensemble.py -m gpu -q 2>&1 | tail -15 | cut -c1-250 | tee gpurun_out/r2n_pytest_multi_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for CFG in C2 C5; do
for TS in 0 1; do
  ENSTOP_B200_TWO_SHOT=$TS timeout 900 $TR --master-port 2953$TS bench.py --gpus $N --mode shard --config $CFG --steps 30 --warmup 3 > gpurun_out/r2n_shard_${CFG}_n${N}_ts$TS.json 2> gpurun_out/r2n_shard_${CFG}_n${N}_ts$TS.err
  echo "shard $CFG two_shot=$TS rc=$?"; tail -2 gpurun_out/r2n_shard_${CFG}_n${N}_ts$TS.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n_shard_${CFG}_n${N}_ts$TS.json").read().strip().splitlines()[-1])
    print("shard $CFG N=$N two_shot=$TS ms/step %.4f value %.3e e2e_s %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["seconds"]), d["roofline"]["kernel_ms_per_iter_max_over_ranks"])
except Exception as e:
    print("parse failed", e)
PY
done
done
