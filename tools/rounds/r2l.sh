timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "large or pipelined or 2_26" 2>&1 | tail -4
echo "--- default flags"
timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2>&1 | grep -E "c2c|r2c|c2r"
echo "--- + flag 4 (two-look-up inter-pass twiddles)"
for n in 32768 65536; do CKFFT_B200_PIPE_FLAGS=5 timeout 300 python tools/gpu_check.py $n 2>&1 | grep -E "c2c|r2c|c2r"; done
for n in 131072 262144 524288 1048576; do CKFFT_B200_PIPE_FLAGS=7 timeout 300 python tools/gpu_check.py $n 2>&1 | grep -E "c2c|r2c|c2r"; done
