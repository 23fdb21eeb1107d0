# round 2, visit ai: final validation of the tree: GPU suite, smoke, both bench arms
mkdir -p gpurun_out; TAG=r2i; SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$? (${SECONDS}s)"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1; tail -c 300 gpurun_out/bench_ref_${TAG}.json
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${TAG}.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], 'e2e', d['e2e']['value'], 'pageable', d['e2e_pageable']['value'], d['e2e_pageable']['driver_staged']['value'], 'multi', d['e2e_multi']['value'], 'cpu', d['cpu_baseline']['value'])
PY
echo "done ${SECONDS}s"
