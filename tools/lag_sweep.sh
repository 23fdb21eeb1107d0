#!/bin/bash
for lag in 0 16 29 48 64 96; do echo "== LAG=$lag"; CKFFT_B200_PIPE_LAG=$lag CKFFT_B200_PIPE_RING_MB=128 timeout 120 python tools/gpu_check.py 65536 2>&1 | grep -E "c2c"; done
for lag in 57 96 128 200; do echo "== LAG=$lag"; CKFFT_B200_PIPE_LAG=$lag CKFFT_B200_PIPE_RING_MB=128 timeout 120 python tools/gpu_check.py 32768 2>&1 | grep -E "c2c"; done
for lag in 2 4 6 8; do echo "== LAG=$lag"; CKFFT_B200_PIPE_LAG=$lag CKFFT_B200_PIPE_RING_MB=128 timeout 120 python tools/gpu_check.py 1048576 2>&1 | grep -E "c2c"; done
