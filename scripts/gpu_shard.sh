#!/bin/bash
# Sharded single fit on N GPUs of one box (bench.py --mode shard), then single-GPU runs of the
# large / wide configs side by side on two of the GPUs.  usage: gpu_shard.sh N CONFIG STEPS
N=${1:-2}; CFG=${2:-C5}; STEPS=${3:-30}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29533 bench.py --gpus $N --mode shard --config $CFG --steps $STEPS --warmup 3 \
    --e2e-repeats 1 --profile-iters 5 > gpurun_out/bench_shard${N}_$CFG.json 2> gpurun_out/bench_shard${N}_$CFG.err
tail -3 gpurun_out/bench_shard${N}_$CFG.err | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout 1500 python bench.py --config $CFG --steps $STEPS --warmup 3 --no-cpu-baseline \
    --profile-iters 5 --e2e-repeats 1 > gpurun_out/bench_single_$CFG.json 2> gpurun_out/bench_single_$CFG.err &
CUDA_VISIBLE_DEVICES=1 timeout 900 python bench.py --config C3 --steps 50 --warmup 3 --no-cpu-baseline \
    --profile-iters 10 --e2e-repeats 2 > gpurun_out/bench_single_C3.json 2> gpurun_out/bench_single_C3.err
wait
python - <<PY
import json
for f in ("bench_shard${N}_$CFG", "bench_single_$CFG", "bench_single_C3"):
    try:
        line = [l for l in open("gpurun_out/%s.json" % f).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        r = d["roofline"]
        print(f, "ms/iter %.4f value %.3e e2e_s %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["seconds"]),
              r.get("kernel_ms_per_iter") or r.get("kernel_ms_per_iter_max_over_ranks"))
    except Exception as e:
        print(f, "failed", e)
PY
