#!/bin/bash
# One GPU-box visit for the whole tree: full GPU test-suite (timed), smoke, default bench.  Logs to gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r1e}
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -14 gpurun_out/pytest_gpu_${TAG}.log
SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke ${SECONDS}s"
SECONDS=0
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$? ${SECONDS}s"; tail -c 2500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
