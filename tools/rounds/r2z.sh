# round 2, visit z: real-inverse split prefetch (twist straight from the raw half rows), 16-values-per-thread audio plans
mkdir -p gpurun_out; TAG=r2z; SECONDS=0
L=$PWD/ckfft_b200/lib
for v in prod c1 c0; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  echo "=== $v"; CKFFT_B200_LIB=$lib timeout 200 python tools/exp_check.py real 8192 16384 32768 2>&1 | grep -E "FAIL|rror|real"
  CKFFT_B200_LIB=$lib timeout 300 python tools/gpu_check.py 8192 16384 32768 2>&1 | grep -E "c2r" | tee gpurun_out/sweep_${v}_${TAG}.log
done
echo "--- c2r tests with c1 (${SECONDS}s)"
CKFFT_B200_LIB=$L/libckfft_b200_c1.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_layouts_gpu.py -m gpu -x -q -k "real" 2>&1 | tail -4
echo "--- stft (${SECONDS}s)"
for v in prod a5 a6 a7 a8; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  CKFFT_B200_LIB=$lib timeout 200 python bench.py --workload stft4096 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stft4096 $v', d['ms_per_step'], d['roofline']['frac'])" | tee -a gpurun_out/stft_${TAG}.log
  [ $v = prod ] || CKFFT_B200_LIB=$lib timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "power" 2>&1 | tail -1
done
echo "done ${SECONDS}s"
