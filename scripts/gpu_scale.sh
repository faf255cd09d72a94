#!/bin/bash
# bench.py exactly as the driver launches it at N ranks (ensemble mode: parity checks, weak-scaling
# EM timing, config 4), plus the sharded mode at C2 / C5.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_ens_n$N.json 2> gpurun_out/bench_ens_n$N.err
echo "ensemble rc=$?"; tail -3 gpurun_out/bench_ens_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ens_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.3e e2e %.3e (%.2f ms) ms/step %.4f c4_wall_s %s (fit %s gather %s) gather_ms %s" % (d["value"], d["e2e"]["value"], 1e3*d["e2e"]["seconds"], d["ms_per_step"], d.get("c4_wall_s"), d.get("c4_fit_wall_s"), d.get("c4_gather_s"), d.get("ensemble_gather_ms")))
    for c in d.get("parity_checks", []): print("  ", {k: (round(v, 9) if isinstance(v, float) else v) for k, v in c.items()})
except Exception as e:
    print("parse failed", e)
PY
for CFG in C2 C5; do
  timeout 1500 $TR --master-port 29522 bench.py --gpus $N --mode shard --config $CFG --steps 30 --warmup 3 > gpurun_out/bench_shard_${CFG}_n$N.json 2> gpurun_out/bench_shard_${CFG}_n$N.err
  echo "shard $CFG rc=$?"; tail -2 gpurun_out/bench_shard_${CFG}_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_shard_${CFG}_n$N.json").read().strip().splitlines()[-1])
    print("shard $CFG N=$N ms/step %.4f value %.3e e2e_s %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["seconds"]), d["roofline"]["kernel_ms_per_iter_max_over_ranks"])
except Exception as e:
    print("parse failed", e)
PY
done
