mkdir -p gpurun_out; TAG=r2j; SECONDS=0
timeout 900 python -m pytest tests/test_pageable_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_pageable_multi_${TAG}.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
echo "done ${SECONDS}s"
