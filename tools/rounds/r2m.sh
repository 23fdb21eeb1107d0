timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_layouts_gpu.py -m gpu -q -k "large or pipelined or 2_26 or fused or six_step" 2>&1 | tail -4
timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2097152 4194304 2>&1 | grep -E "c2c|r2c|c2r"
