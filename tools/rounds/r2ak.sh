# round 2, visit ak: the direct-reference parity cases (tests/test_reference_gpu.py), then the whole GPU suite
mkdir -p gpurun_out; TAG=r2k; SECONDS=0
timeout 200 python -m pytest tests/test_reference_gpu.py -m gpu -q --maxfail=5 > gpurun_out/pytest_reference_${TAG}.log 2>&1; echo "reference cases rc=$? after ${SECONDS}s"; tail -4 gpurun_out/pytest_reference_${TAG}.log
timeout 400 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "suite rc=$? after ${SECONDS}s"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
echo "done ${SECONDS}s"
