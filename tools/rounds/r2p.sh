mkdir -p gpurun_out; TAG=r2p; SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -k "oracle or golden or host or chunk or staging or multi or harness or c_abi or thread" 2>&1 | tail -4
echo "--- bench e2e (${SECONDS}s)"
timeout 600 python bench.py --no-secondary --no-cpu-baseline --steps 50 --e2e-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('e2e','e2e_pageable','e2e_multi'): print(k, json.dumps(d.get(k))[:700])"
for mb in 16 64; do echo "chunk $mb MiB"; CKFFT_B200_CHUNK_MB=$mb timeout 600 python bench.py --no-secondary --no-cpu-baseline --steps 20 --e2e-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['e2e']['value'], d['e2e']['roofline']['peak'], d['e2e_multi']['value'])"; done
echo "done ${SECONDS}s"
