# round 2, visit af: library-side staging of pageable host arrays
mkdir -p gpurun_out; TAG=r2af; SECONDS=0
timeout 600 python -m pytest tests/test_pageable_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "pageable or host or classic or abi or error" 2>&1 | tail -4
echo "--- e2e (${SECONDS}s)"
for t in 2 4 8; do
CKFFT_B200_HOST_THREADS=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['e2e_pageable']; print('threads $t: pageable', p['value'], 'driver', p['driver_staged']['value'], 'pin', p['with_registration']['value'], 'identical', p['bit_identical_to_pinned_path'], '| e2e pinned', d['e2e']['value'])" | tee -a gpurun_out/pageable_${TAG}.log
done
for mb in 8 32; do
CKFFT_B200_PAGEABLE_CHUNK_MB=$mb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['e2e_pageable']; print('chunk $mb MiB: pageable', p['value'])" | tee -a gpurun_out/pageable_${TAG}.log
done
echo "done ${SECONDS}s"
