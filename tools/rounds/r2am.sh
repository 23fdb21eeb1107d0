# round 2, visit am: non-temporal staging copies of the pageable path (host_copy.cpp): whole GPU suite on the new library, then A/B
mkdir -p gpurun_out; TAG=r2m; SECONDS=0
timeout 110 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "suite rc=$? after ${SECONDS}s"; tail -2 gpurun_out/pytest_gpu_${TAG}.log
timeout 45 python tools/pageable_probe.py 0 1 2 3 1 0 2>&1 | tee gpurun_out/pageable_probe_${TAG}.log
lscpu | grep -i "model name" | tee -a gpurun_out/pageable_probe_${TAG}.log
echo "done ${SECONDS}s"
