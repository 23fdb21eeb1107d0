mkdir -p gpurun_out; TAG=r2ab; SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python tools/gpu_check.py 4096 8192 16384 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep_${TAG}.log
echo "done ${SECONDS}s"
