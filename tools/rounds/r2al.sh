# round 2, visit al: bench line of the final tree (default command, as the driver runs it)
mkdir -p gpurun_out; TAG=r2l; SECONDS=0
timeout 150 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$? (${SECONDS}s)"
tail -c 400 gpurun_out/bench_${TAG}.json
