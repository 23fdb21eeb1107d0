mkdir -p gpurun_out; TAG=r2b; SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -12 gpurun_out/pytest_gpu_${TAG}.log
echo "--- stats"; bash tools/pipe_stats.sh 2>&1 | tee gpurun_out/pipe_stats_${TAG}.log
echo "--- sweep (${SECONDS}s)"
timeout 300 python tools/gpu_check.py 32768 65536 131072 262144 524288 1048576 2097152 > gpurun_out/sweep2p_${TAG}.log 2>&1; cat gpurun_out/sweep2p_${TAG}.log
echo "--- bench (${SECONDS}s)"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$? (${SECONDS}s)"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])
for k in ('e2e','e2e_pageable','e2e_multi'): print(k, d.get(k))
s=d['secondary']; print(s['sweep']['frac']); print(s['real_large']); print(s['r2c4096']['frac'], s['c2r4096']['frac'], s['stft4096']['frac'])
PY
tail -3 gpurun_out/bench_${TAG}.err
echo "done ${SECONDS}s"
