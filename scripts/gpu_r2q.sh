#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "presort" 2>&1 | tail -6 | cut -c1-250 | tee gpurun_out/r2q_pytest.log
ENSTOP_B200_TRACE=1 timeout 300 python scripts/fit_phases.py C2 pinned 2>&1 | tail -40 | tee gpurun_out/r2q_trace.txt
