timeout 600 python tools/twist_probe.py
