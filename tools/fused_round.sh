#!/bin/bash
# One single-GPU box visit for the fused distributed transform at world = 1: tests, check/timing, four-step regression sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "fused or six_step or large_complex" 2>&1 | tee gpurun_out/pytest_fused.log | tail -15
timeout 300 python tools/fused_check.py 14 17 20 21 24 26 28 30 2>&1 | tee gpurun_out/fused_w1.log | tail -12
timeout 200 python tools/fused_check.py --passes 3 28 30 2>&1 | tee -a gpurun_out/fused_w1.log | tail -4
