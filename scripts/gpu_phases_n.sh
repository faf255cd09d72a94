#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/fit_phases_ranks.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -20 | tee gpurun_out/fit_phases_n$N.txt
