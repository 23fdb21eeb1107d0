# round 2, visit w: split-prefetch plans for C2C 2048 / 4096 / 8192 and R2C M = 2048 / 4096, register twiddles for C2R 16384
mkdir -p gpurun_out; TAG=r2w; SECONDS=0
L=$PWD/ckfft_b200/lib
run() { v=$1; kind=$2; shift 2; lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  echo "=== $v"; CKFFT_B200_LIB=$lib timeout 200 python tools/exp_check.py $kind "$@" 2>&1 | grep -E "FAIL|Error|error" ; \
  CKFFT_B200_LIB=$lib timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "$kind" | tee gpurun_out/sweep_${v}_${TAG}.log; }
run prod c2c 2048 4096 8192
run f1 c2c 2048 4096
run f2 c2c 2048 8192
run f3 c2c 4096 8192
echo "--- real (${SECONDS}s)"
run prod r2c 2048 4096 8192 32768
run g1 r2c 4096 8192
run g2 r2c 4096 8192
run g3 r2c 2048 4096
run h0 c2r 4096 32768
run h1 c2r 32768
echo "--- config 3 (${SECONDS}s)"
for v in prod g1 g2 g3 h0; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  for w in r2c4096 c2r4096; do
  CKFFT_B200_LIB=$lib timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $v', d['ms_per_step'], d['roofline']['frac'])" | tee -a gpurun_out/cfg3_${TAG}.log
  done
done
echo "done ${SECONDS}s"
