# round 2, visit t: linear padded addressing (all single-pass kernels) vs the previous build, packed-math / constant-factor /
# register-twiddle A/B libraries, full GPU suite on the new product library
mkdir -p gpurun_out; TAG=r2t; SECONDS=0
L=$PWD/ckfft_b200/lib
echo "--- GPU suite (product library)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_${TAG}.log
echo "--- sweeps (${SECONDS}s)"
for v in base prod pk pkc rtwc twr16; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  [ -f $lib ] || { echo "no $lib"; continue; }
  echo "=== $v"
  CKFFT_B200_LIB=$lib timeout 300 python tools/gpu_check.py 64 128 256 512 1024 2048 4096 8192 16384 32768 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep_${v}_${TAG}.log
done
echo "--- stft (${SECONDS}s)"
for v in base prod pk pkc; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  [ -f $lib ] || continue
  CKFFT_B200_LIB=$lib timeout 200 python bench.py --workload stft4096 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stft4096 $v', d['ms_per_step'], d['roofline']['frac'])" | tee -a gpurun_out/stft_${TAG}.log
done
echo "done ${SECONDS}s"
