timeout 600 python tools/twist_probe.py 2>&1 | tail -16
timeout 300 python tools/gpu_check.py 65536 131072 262144 524288 1048576 2097152 2>&1 | grep -E "r2c|c2r"
