mkdir -p gpurun_out; SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'])
for k in ('e2e','e2e_multi'): print(k, json.dumps(d.get(k))[:900])"
echo "done ${SECONDS}s"
