N=$1; mkdir -p gpurun_out; SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_${N}gpu_r2s.json 2> gpurun_out/bench_${N}gpu_r2s.err; echo "rc=$? (${SECONDS}s)"
python - <<PY
import json
for line in open('gpurun_out/bench_${N}gpu_r2s.json'):
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['roofline']['peak'], 'multi', d['e2e_multi']['value'])
    print('dist30', json.dumps(d['secondary'].get('dist30')))
PY
