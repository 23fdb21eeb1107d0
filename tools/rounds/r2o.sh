mkdir -p gpurun_out; TAG=r2o; SECONDS=0
nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader | head -8; nproc; python -c "import os; print(sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0)))"
for i in 0 1 2 3 4 5 6 7; do b=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $i | tr 'A-Z' 'a-z' | cut -c5-); echo "gpu $i node $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) cpus $(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)"; done
echo "--- pcie 8 gpus, no binding (${SECONDS}s)"
timeout 300 python tools/pcie_peak.py --gpus 8 --mb 1024 | tee gpurun_out/pcie_8gpu_${TAG}.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['gpu_numa_nodes'], d['concurrent_1']['sum_both'], {k:d['concurrent_8'][k] for k in ('sum_both','sum_h2d','sum_d2h')}, [r['both'] for r in d['concurrent_8']['per_gpu']])"
echo "--- pcie 8 gpus, NUMA-bound (${SECONDS}s)"
timeout 300 python tools/pcie_peak.py --gpus 8 --mb 1024 --numa | tee gpurun_out/pcie_8gpu_numa_${TAG}.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['gpu_numa_nodes'], d['concurrent_1']['sum_both'], {k:d['concurrent_8'][k] for k in ('sum_both','sum_h2d','sum_d2h')}, [r['both'] for r in d['concurrent_8']['per_gpu']])"
echo "--- bench --gpus 8 (${SECONDS}s)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/bench_8gpu_${TAG}.json 2> gpurun_out/bench_8gpu_${TAG}.err; echo "rc=$? (${SECONDS}s)"
tail -c 1500 gpurun_out/bench_8gpu_${TAG}.err | tail -5
python - <<'PY'
import json
for line in open('gpurun_out/bench_8gpu_r2o.json'):
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline']['frac'])
    for k in ('e2e','e2e_pageable','e2e_multi','cpu_baseline'): print(k, json.dumps(d.get(k))[:1100])
    s=d['secondary']; print('sweep', s['sweep']['frac']); print('real', s['real_large']); print(s['r2c4096']['frac'], s['c2r4096']['frac'], s['stft4096']['frac'])
    print('dist30', json.dumps(s.get('dist30')))
PY
echo "done ${SECONDS}s"
