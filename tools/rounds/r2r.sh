SECONDS=0
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'peak', d['e2e']['roofline']['peak'])
print('e2e_multi', json.dumps(d.get('e2e_multi'))[:400])"
echo "done ${SECONDS}s"
