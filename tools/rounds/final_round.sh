#!/bin/bash
# Round-closing GPU visit: full suite, smoke, both bench arms, size sweep with layouts, config 3, ncu launch list + full captures.
# Everything written under gpurun_out/ stays small (gpurun copies back at most 64 MiB): the ncu reports are summarised ON the box
# (tools/ncu_summary.py, tools/ncu_hot.py) and only the report of the headline kernel travels.
mkdir -p gpurun_out
TAG=${1:-r2g}
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -9 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$? (${SECONDS}s)"; head -c 700 gpurun_out/bench_${TAG}.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1; tail -c 600 gpurun_out/bench_ref_${TAG}.json
for w in sweep r2c4096 c2r4096 stft4096; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_${w}_${TAG}.json 2> gpurun_out/bench_${w}_${TAG}.err; tail -c 300 gpurun_out/bench_${w}_${TAG}.json; echo; done
CKFFT_SWEEP_LAYOUTS=1 timeout 600 python tools/gpu_check.py $(python -c "print(*[1<<k for k in range(4,23)])") > gpurun_out/sweep_${TAG}.log 2>&1; tail -3 gpurun_out/sweep_${TAG}.log
timeout 300 python tools/gpu_check.py > gpurun_out/parity_${TAG}.log 2>&1; grep -E "worst|n= " gpurun_out/parity_${TAG}.log | tail -3
echo "--- ncu launch list (${SECONDS}s)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
tail -1 gpurun_out/launches_${TAG}.csv | cut -c1-200
echo "--- ncu full (${SECONDS}s)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 2 -o gpurun_out/prof_c2c1024_${TAG} python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
S=/tmp/ncu_${TAG}; mkdir -p $S
for spec in "c2c 16384" "c2c 4096" "c2c 8192" "r2c 4096" "c2r 4096" "c2c 65536" "c2c 1048576"; do set -- $spec
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"small_kernel|fft_kernel|pipe_kernel" -s 2 -c 1 -o $S/prof_$1_$2 python tools/prof_one.py $1 $2 27 > $S/ncu_$1_$2.log 2>&1
  { python tools/ncu_summary.py $S/prof_$1_$2.ncu-rep; python tools/ncu_hot.py $S/prof_$1_$2.ncu-rep 25; } > gpurun_out/ncu_$1_$2_${TAG}.txt 2>&1
  grep -E "gpu__time_duration|issue_active" gpurun_out/ncu_$1_$2_${TAG}.txt | tr '\n' ' '; echo
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 2 -c 1 -o $S/prof_stft python bench.py --workload stft4096 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > $S/ncu_stft.log 2>&1
{ python tools/ncu_summary.py $S/prof_stft.ncu-rep; python tools/ncu_hot.py $S/prof_stft.ncu-rep 25; } > gpurun_out/ncu_stft4096_${TAG}.txt 2>&1
du -sh gpurun_out; echo "done ${SECONDS}s"
