mkdir -p gpurun_out; TAG=r2g; SECONDS=0
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "real_inverse_fused or large_real" 2>&1 | tail -15
echo "--- sweep (${SECONDS}s)"
timeout 300 python tools/gpu_check.py 65536 131072 262144 524288 1048576 2097152 2>&1 | grep -E "r2c|c2r" | tee gpurun_out/sweep_c2r_${TAG}.log
echo "done ${SECONDS}s"
