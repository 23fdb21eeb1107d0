# round 2, visit aa: validation of the new single-pass plans: GPU suite, sanitizer memcheck / racecheck, size sweep
mkdir -p gpurun_out; TAG=r2aa; SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
echo "--- sweep (${SECONDS}s)"
timeout 300 python tools/gpu_check.py 256 512 1024 2048 4096 8192 16384 32768 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep_${TAG}.log
echo "--- sanitizer memcheck (${SECONDS}s)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/small_cover.py > gpurun_out/sanitizer_memcheck_${TAG}.log 2>&1; echo "memcheck rc=$? (${SECONDS}s)"; tail -4 gpurun_out/sanitizer_memcheck_${TAG}.log
echo "--- sanitizer racecheck (${SECONDS}s)"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/small_cover.py > gpurun_out/sanitizer_racecheck_${TAG}.log 2>&1; echo "racecheck rc=$? (${SECONDS}s)"; tail -6 gpurun_out/sanitizer_racecheck_${TAG}.log
echo "done ${SECONDS}s"
