#!/bin/bash
# bench.py exactly as the driver launches it at N ranks (ensemble mode only).
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_ens_n$N.json 2> gpurun_out/bench_ens_n$N.err
echo "ensemble rc=$?"; tail -3 gpurun_out/bench_ens_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ens_n$N.json").read().strip().splitlines()[-1])
print("N=$N value %.3e e2e %.3e (%.2f ms) ms/step %.4f c4_wall_s %s %s gather_ms %s" % (d["value"], d["e2e"]["value"], 1e3*d["e2e"]["seconds"], d["ms_per_step"], d.get("c4_wall_s"), d.get("c4_wall_s_all_runs"), d.get("ensemble_gather_ms")))
print(all(c["ok"] for c in d["parity_checks"]), len(d["parity_checks"]))
PY
