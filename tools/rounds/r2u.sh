# round 2, visit u: why did C2C 2048 / 4096 lose with the leaner loop?  ncu of the old and the new build; GPU suite
mkdir -p gpurun_out; TAG=r2u; SECONDS=0
L=$PWD/ckfft_b200/lib
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_${TAG}.log
echo "--- ncu (${SECONDS}s)"
for v in base prod; do
  lib=$L/libckfft_b200_$v.so; [ $v = prod ] && lib=$L/libckfft_b200.so
  for n in 2048 4096; do
    CKFFT_B200_LIB=$lib timeout 200 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 2 -c 1 -o gpurun_out/prof_c2c_${n}_${v}_${TAG} \
      python tools/prof_one.py c2c $n 27 > gpurun_out/ncu_${n}_${v}_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_${n}_${v}_${TAG}.log
  done
done
echo "done ${SECONDS}s"
