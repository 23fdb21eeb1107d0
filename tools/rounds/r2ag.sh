mkdir -p gpurun_out; TAG=r2ag; SECONDS=0
for t in 0 12 16; do
CKFFT_B200_HOST_THREADS=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['e2e_pageable']; print('threads $t: pageable', p['value'], 'driver', p['driver_staged']['value'], 'identical', p['bit_identical_to_pinned_path'], '| e2e pinned', d['e2e']['value'])" | tee -a gpurun_out/pageable_${TAG}.log
done
nproc
echo "done ${SECONDS}s"
