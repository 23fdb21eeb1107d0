#!/bin/bash
# GPU visit for the layout variants (planar, in place): their tests, the sweep with the extra kinds, and ncu captures of
# the real-transform kernels that sit lowest in the single-pass table.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 420 python -m pytest tests/test_layouts_gpu.py -x -q 2>&1 | tee gpurun_out/pytest_layouts.log | tail -15
echo "== sweep"; CKFFT_SWEEP_LAYOUTS=1 timeout 300 python tools/gpu_check.py 16 32 64 128 256 512 1024 2048 4096 8192 16384 32768 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep_layouts.log
echo "== ncu"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -o gpurun_out/prof_real_r1e \
    python tools/prof_many.py r2c:8192 r2c:16384 c2r:8192 c2c:4096 c2c:16384 --reps 1 > gpurun_out/ncu_real.log 2>&1; tail -3 gpurun_out/ncu_real.log
ls -la gpurun_out/
