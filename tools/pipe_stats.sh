#!/bin/bash
# Cycle statistics of the dataflow kernel (development): needs the instrumented library
#   CKFFT_B200_BUILD_TAG=stats CKFFT_B200_NVCC_FLAGS="-DCKB_PIPE_STATS=1" python -m ckfft_b200.build
# bash tools/pipe_stats.sh [sizes...]   -> one "pipe_stats ..." line per launch on stderr
export CKFFT_B200_LIB=$PWD/ckfft_b200/lib/libckfft_b200_stats.so
SIZES=${@:-"32768 65536 262144 1048576"}
for n in $SIZES; do
  for nbuf in 1 2; do
    CKFFT_B200_PIPE_NBUF=$nbuf python tools/prof_one.py c2c $n 2>&1 | grep pipe_stats | tail -1
  done
done
