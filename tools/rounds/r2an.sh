# round 2, visit an: copy threads per team with the non-temporal staging copies (the default of 8 was measured under memcpy)
mkdir -p gpurun_out; TAG=r2n; SECONDS=0
for t in 4 6 12; do CKFFT_B200_HOST_THREADS=$t PROBE_LOG2_BATCH=17 timeout 15 python tools/pageable_probe.py 1 2>&1 | grep -v "^cpu" | tee -a gpurun_out/pageable_threads_${TAG}.log; done
echo "done ${SECONDS}s"
