#!/bin/bash
# One GPU-box visit: smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r1d}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1; tail -c 1500 gpurun_out/bench_ref_${TAG}.json
nproc; lscpu | grep -E "Model name|Socket|Thread|Core" | head -5
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
tail -3 gpurun_out/launches_${TAG}.csv
ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 2 -o gpurun_out/prof_c2c1024_${TAG} python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
