#!/bin/bash
# One single-GPU box visit for the fused distributed transform: tests (world 1 and two processes), check/timing at world 1.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "fused" 2>&1 | tee gpurun_out/pytest_fused.log | tail -15
timeout 300 python tools/fused_check.py 20 24 28 30 2>&1 | tee gpurun_out/fused_w1.log | grep -E "FUSED|phases|rror" | tail -12
