#!/bin/bash
# run on the GPU box: GPU test-suite, log to gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -25
