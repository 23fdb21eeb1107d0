mkdir -p gpurun_out; TAG=r2e; SECONDS=0
echo "--- default (nbuf2 where it fits)"
timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep2p_${TAG}.log
for fl in 1 2 3; do echo "--- flags=$fl"; CKFFT_B200_PIPE_FLAGS=$fl timeout 300 python tools/gpu_check.py 32768 65536 262144 1048576 2>&1 | grep -E "c2c"; done
echo "--- c2r via old path"
CKFFT_B200_PIPE_REAL=0 timeout 300 python tools/gpu_check.py 65536 262144 1048576 2>&1 | grep -E "r2c|c2r"
echo "--- stats c2r/r2c (${SECONDS}s)"
export CKFFT_B200_LIB=$PWD/ckfft_b200/lib/libckfft_b200_stats.so
for k in c2r r2c; do for n in 131072 1048576; do python tools/prof_one.py $k $n 2>&1 | grep pipe_stats | tail -1; done; done
unset CKFFT_B200_LIB
echo "--- tests (${SECONDS}s)"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "large or pipelined or 2_26" 2>&1 | tail -4
echo "--- ncu stft / r2c4096 (${SECONDS}s)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 1 -o gpurun_out/prof_stft4096_${TAG} python bench.py --workload stft4096 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_stft_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_stft_${TAG}.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 1 -o gpurun_out/prof_r2c4096_${TAG} python bench.py --workload r2c4096 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_r2c4096_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_r2c4096_${TAG}.log
echo "done ${SECONDS}s"
