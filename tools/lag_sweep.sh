#!/bin/bash
for nbuf in 1 2; do for lag in 30 44 60 80; do echo "== N=65536 NBUF=$nbuf LAG=$lag"; CKFFT_B200_PIPE_NBUF=$nbuf CKFFT_B200_PIPE_LAG=$lag CKFFT_B200_PIPE_RING_MB=100 timeout 120 python tools/gpu_check.py 65536 2>&1 | grep -E "c2c"; done; done
for nbuf in 1 2; do for lag in 60 90 120; do echo "== N=32768 NBUF=$nbuf LAG=$lag"; CKFFT_B200_PIPE_NBUF=$nbuf CKFFT_B200_PIPE_LAG=$lag CKFFT_B200_PIPE_RING_MB=100 timeout 120 python tools/gpu_check.py 32768 2>&1 | grep -E "c2c"; done; done
for lag in 8 12 16 24; do echo "== N=262144 LAG=$lag"; CKFFT_B200_PIPE_LAG=$lag CKFFT_B200_PIPE_RING_MB=100 timeout 120 python tools/gpu_check.py 262144 2>&1 | grep -E "c2c"; done
for lag in 2 3 4 5; do echo "== N=1048576 LAG=$lag"; CKFFT_B200_PIPE_LAG=$lag CKFFT_B200_PIPE_RING_MB=100 timeout 120 python tools/gpu_check.py 1048576 2>&1 | grep -E "c2c"; done
