# round 2, visit x: new plan tables (multi-group split prefetch) -- GPU suite, size sweep, config 3 / STFT, full bench line
mkdir -p gpurun_out; TAG=r2x; SECONDS=0
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_${TAG}.log
echo "--- sweep (${SECONDS}s)"
timeout 300 python tools/gpu_check.py 16 32 64 128 256 512 1024 2048 4096 8192 16384 32768 65536 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep_${TAG}.log
echo "--- bench (${SECONDS}s)"
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${TAG}.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], 'e2e', d['e2e']['value'])
s=d['secondary']
for k in ('r2c4096','c2r4096','stft4096'): print(k, s[k]['ms_per_step'], s[k]['frac'])
print('sweep', list(zip(s['sweep']['n'], s['sweep']['frac'])))
print('real_large', s['real_large'])
PY
echo "done ${SECONDS}s"
