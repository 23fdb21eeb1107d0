# round 2, visit y: R2C M = 8192 and audio (STFT) plan variants
mkdir -p gpurun_out; TAG=r2y; SECONDS=0
L=$PWD/ckfft_b200/lib
for v in prod r1 r2 r3 r4; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  echo "=== $v"; CKFFT_B200_LIB=$lib timeout 200 python tools/exp_check.py real 16384 2>&1 | grep -E "FAIL|rror"
  CKFFT_B200_LIB=$lib timeout 300 python tools/gpu_check.py 16384 2>&1 | grep -E "r2c" | tee gpurun_out/sweep_${v}_${TAG}.log
done
echo "--- stft (${SECONDS}s)"
for v in prod a1 a2 a3 a4; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  CKFFT_B200_LIB=$lib timeout 200 python bench.py --workload stft4096 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stft4096 $v', d['ms_per_step'], d['roofline']['frac'])" | tee -a gpurun_out/stft_${TAG}.log
done
echo "done ${SECONDS}s"
