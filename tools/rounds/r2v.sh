# round 2, visit v: prefetch-distance experiments for C2C 2048 / 4096 (tools/exp_build.sh), config 3 on old and new build
mkdir -p gpurun_out; TAG=r2v; SECONDS=0
L=$PWD/ckfft_b200/lib
for v in base prod e1 e2 e3 e4 e5; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  [ -f $lib ] || { echo "no $lib"; continue; }
  echo "=== $v"
  CKFFT_B200_LIB=$lib timeout 300 python tools/gpu_check.py 2048 4096 2>&1 | grep -E "c2c" | tee gpurun_out/sweep_${v}_${TAG}.log
done
echo "--- config 3 (${SECONDS}s)"
for v in base prod; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  for w in r2c4096 c2r4096; do
  CKFFT_B200_LIB=$lib timeout 200 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $v', d['ms_per_step'], d['roofline']['frac'])" | tee -a gpurun_out/cfg3_${TAG}.log
  done
done
echo "done ${SECONDS}s"
