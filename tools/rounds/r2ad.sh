# round 2, visit ad: C2R M = 8192 plan variants; GPU suite on the product library (split tile prefetch for L0 >= 512)
mkdir -p gpurun_out; TAG=r2ad; SECONDS=0
L=$PWD/ckfft_b200/lib
for v in prod k1 k2 k3 k4; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  echo "=== $v"; CKFFT_B200_LIB=$lib timeout 200 python tools/exp_check.py real 16384 2>&1 | grep -E "FAIL|rror"
  CKFFT_B200_LIB=$lib timeout 300 python tools/gpu_check.py 16384 2>&1 | grep -E "c2r" | tee gpurun_out/sweep_${v}_${TAG}.log
done
echo "--- suite (${SECONDS}s)"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
echo "done ${SECONDS}s"
