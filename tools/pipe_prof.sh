#!/bin/bash
# ncu --set full with source correlation of the two-pass dataflow kernel at a few lengths
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_layouts_gpu.py -x -q 2>&1 | tail -4
for n in 65536 262144 1048576; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 2 -c 1 -o gpurun_out/prof_pipe_${n}_r1e \
      python tools/prof_one.py c2c $n 27 > gpurun_out/ncu_pipe_$n.log 2>&1; tail -2 gpurun_out/ncu_pipe_$n.log
done
ls -la gpurun_out/
