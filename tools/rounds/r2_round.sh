#!/bin/bash
# Round-2 GPU visit: suite, smoke, both bench arms, two-pass sweep A/B, one ncu capture.  bash tools/rounds/r2_round.sh <tag>
mkdir -p gpurun_out
TAG=${1:-r2a}
SECONDS=0
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 --durations=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -25 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "--- bench (${SECONDS}s)"
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$? (${SECONDS}s)"; tail -c 6000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1; tail -c 900 gpurun_out/bench_ref_${TAG}.json
echo "--- two-pass sweep (${SECONDS}s)"
timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2097152 > gpurun_out/sweep2p_${TAG}.log 2>&1; cat gpurun_out/sweep2p_${TAG}.log
echo "--- NBUF=2"
CKFFT_B200_PIPE_NBUF=2 timeout 300 python tools/gpu_check.py 32768 65536 > gpurun_out/sweep2p_nbuf2_${TAG}.log 2>&1; grep c2c gpurun_out/sweep2p_nbuf2_${TAG}.log
echo "--- two kernels"
CKFFT_B200_PIPE=0 timeout 300 python tools/gpu_check.py 65536 1048576 > gpurun_out/sweep2p_nopipe_${TAG}.log 2>&1; grep c2c gpurun_out/sweep2p_nopipe_${TAG}.log
echo "--- pcie (${SECONDS}s)"
timeout 200 python tools/pcie_peak.py --gpus 1 | tee gpurun_out/pcie_${TAG}.json
timeout 200 python tools/pcie_peak.py --gpus 1 --pageable --mb 1024 | tee gpurun_out/pcie_pageable_${TAG}.json
echo "--- ncu (${SECONDS}s)"
for n in 65536 1048576; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 2 -c 1 -o gpurun_out/prof_pipe_${n}_${TAG} python tools/prof_one.py c2c $n > gpurun_out/ncu_pipe_${n}_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_pipe_${n}_${TAG}.log
done
echo "done after ${SECONDS}s"
