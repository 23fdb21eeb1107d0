mkdir -p gpurun_out; TAG=r2n; SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -6 gpurun_out/pytest_gpu_${TAG}.log
echo "--- latency (${SECONDS}s)"
python tools/latency_probe.py 2>&1 | tee gpurun_out/latency_${TAG}.log
CKFFT_B200_SPIN_SYNC=0 python tools/latency_probe.py 1024 4096 2>&1 | tee -a gpurun_out/latency_${TAG}.log
echo "--- small sizes (${SECONDS}s)"
timeout 300 python tools/gpu_check.py 16 32 64 128 256 512 2>&1 | grep -E "c2c|r2c|c2r" | tee gpurun_out/sweep_small_${TAG}.log
echo "--- sanitizer memcheck (${SECONDS}s)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/small_cover.py > gpurun_out/sanitizer_memcheck_${TAG}.log 2>&1; echo "memcheck rc=$? (${SECONDS}s)"; tail -4 gpurun_out/sanitizer_memcheck_${TAG}.log
echo "--- sanitizer racecheck (${SECONDS}s)"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/small_cover.py > gpurun_out/sanitizer_racecheck_${TAG}.log 2>&1; echo "racecheck rc=$? (${SECONDS}s)"; tail -6 gpurun_out/sanitizer_racecheck_${TAG}.log
echo "done ${SECONDS}s"
