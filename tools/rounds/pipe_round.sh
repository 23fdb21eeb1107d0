#!/bin/bash
# single-GPU visit for the pipelined two-pass kernel: tests, then the sweep with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "pipelined or large_complex or large_real" 2>&1 | tee gpurun_out/pytest_pipe.log | tail -15
echo "== PIPE=1"; CKFFT_B200_PIPE=1 timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2>&1 | grep -E "c2c|r2c" | tee gpurun_out/sweep_pipe1.log
echo "== PIPE=0"; CKFFT_B200_PIPE=0 timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2>&1 | grep -E "c2c" | tee gpurun_out/sweep_pipe0.log
