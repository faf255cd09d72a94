#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): 2-GPU tests, then bench.py as the driver launches it
# (ensemble mode: sharded-fit parity checks + config 4), then the sharded mode.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -12
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_ensemble.py -m gpu -q 2>&1 | tail -15 | cut -c1-250 | tee gpurun_out/pytest_multi_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_ens_n$N.json 2> gpurun_out/bench_ens_n$N.err
echo "ensemble rc=$?"; tail -3 gpurun_out/bench_ens_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ens_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.3e e2e %.3e ms/step %.4f c4_wall_s %s gather %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("c4_wall_s"), d.get("ensemble_gather_ms")))
    for c in d.get("parity_checks", []): print("  ", c)
except Exception as e:
    print("parse failed", e)
PY
for CFG in C2; do
  timeout 900 $TR --master-port 29512 bench.py --gpus $N --mode shard --config $CFG --steps 30 --warmup 3 > gpurun_out/bench_shard_${CFG}_n$N.json 2> gpurun_out/bench_shard_${CFG}_n$N.err
  echo "shard $CFG rc=$?"; tail -2 gpurun_out/bench_shard_${CFG}_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_shard_${CFG}_n$N.json").read().strip().splitlines()[-1])
    print("shard $CFG N=$N ms/step %.4f value %.3e" % (d["ms_per_step"], d["value"]), d["roofline"]["kernel_ms_per_iter_max_over_ranks"])
except Exception as e:
    print("parse failed", e)
PY
done
