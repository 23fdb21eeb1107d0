mkdir -p gpurun_out; TAG=r2f; SECONDS=0
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_parity_gpu.py -m gpu -q -k "multi or two_processes or flag_block" 2>&1 | tail -6
echo "--- bench --gpus 2 (${SECONDS}s)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_2gpu_${TAG}.json 2> gpurun_out/bench_2gpu_${TAG}.err; echo "rc=$? (${SECONDS}s)"
tail -c 3000 gpurun_out/bench_2gpu_${TAG}.err | tail -8
python - <<'PY'
import json
for line in open('gpurun_out/bench_2gpu_r2f.json'):
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline']['frac'])
    for k in ('e2e','e2e_pageable','e2e_multi','cpu_baseline'): print(k, json.dumps(d.get(k))[:900])
    s=d['secondary']; print('sweep', s['sweep']['frac']); print('real', s['real_large']); print(s['r2c4096']['frac'], s['c2r4096']['frac'], s['stft4096']['frac'])
    print('dist30', json.dumps(s.get('dist30')))
PY
echo "--- reference arm under torchrun (${SECONDS}s)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -c 700
echo "--- pcie 2 gpus (${SECONDS}s)"
timeout 300 python tools/pcie_peak.py --gpus 2 | tee gpurun_out/pcie_2gpu_${TAG}.json
echo "done ${SECONDS}s"
