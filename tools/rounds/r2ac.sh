# round 2, visit ac: split tile prefetch in the dataflow kernel (512 / 1024-point tile plans)
mkdir -p gpurun_out; TAG=r2ac; SECONDS=0
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_layouts_gpu.py -m gpu -x -q -k "large or pipelined or 2_26 or strided" 2>&1 | tail -4
echo "--- sweep split=1 (${SECONDS}s)"
timeout 300 python tools/gpu_check.py 131072 262144 524288 1048576 2097152 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep_split1_${TAG}.log
echo "--- sweep split=0 (${SECONDS}s)"
CKFFT_B200_PIPE_SPLIT=0 timeout 300 python tools/gpu_check.py 131072 262144 524288 1048576 2097152 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep_split0_${TAG}.log
echo "done ${SECONDS}s"
