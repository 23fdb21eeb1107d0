mkdir -p gpurun_out; TAG=r2d; SECONDS=0
echo "--- default lib"
timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2097152 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep2p_default_${TAG}.log
echo "--- packed lib"
export CKFFT_B200_LIB=$PWD/ckfft_b200/lib/libckfft_b200_packed.so
timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2097152 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep2p_packed_${TAG}.log
echo "--- packed nbuf2"
CKFFT_B200_PIPE_NBUF=2 CKFFT_B200_PIPE_LAG=80 timeout 300 python tools/gpu_check.py 65536 2>&1 | grep -E "c2c"
CKFFT_B200_PIPE_NBUF=2 CKFFT_B200_PIPE_LAG=150 timeout 300 python tools/gpu_check.py 32768 2>&1 | grep -E "c2c"
echo "--- packed: multi-pass parity tests (${SECONDS}s)"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "large or pipelined or fused or six_step or 2_26" 2>&1 | tail -5
echo "--- packed ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 2 -c 1 -o gpurun_out/prof_pipe_65536_packed_${TAG} python tools/prof_one.py c2c 65536 > gpurun_out/ncu_pipe_packed_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_pipe_packed_${TAG}.log
echo "done ${SECONDS}s"
