mkdir -p gpurun_out; TAG=r2ae; SECONDS=0
L=$PWD/ckfft_b200/lib
for v in prod p1; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  echo "=== $v"
  CKFFT_SWEEP_LAYOUTS=1 CKFFT_B200_LIB=$lib timeout 300 python tools/gpu_check.py 2048 4096 8192 16384 2>&1 | grep -E "c2cp" | tee gpurun_out/sweep_${v}_${TAG}.log
done
CKFFT_B200_LIB=$L/libckfft_b200_p1.so timeout 600 python -m pytest tests/test_layouts_gpu.py -m gpu -x -q -k "planar" 2>&1 | tail -3
echo "done ${SECONDS}s"
