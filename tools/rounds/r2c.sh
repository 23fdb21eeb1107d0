mkdir -p gpurun_out; TAG=r2c; SECONDS=0
echo "--- stats (non-intrusive)"; bash tools/pipe_stats.sh 65536 1048576 2>&1 | tee gpurun_out/pipe_stats_${TAG}.log
echo "--- stats lag 60 / 80, nbuf 1 / 2 at 2^16"
export CKFFT_B200_LIB=$PWD/ckfft_b200/lib/libckfft_b200_stats.so
for cfg in "1 60 64" "2 60 64" "2 80 96" "2 120 128" "1 120 128"; do set -- $cfg
  CKFFT_B200_PIPE_NBUF=$1 CKFFT_B200_PIPE_LAG=$2 CKFFT_B200_PIPE_RING_MB=$3 python tools/prof_one.py c2c 65536 2>&1 | grep pipe_stats | tail -1
done
unset CKFFT_B200_LIB
echo "--- grid (${SECONDS}s)"
GRID_LAGS=0,50,60,80,120 timeout 300 python tools/pipe_grid.py 16 2>&1 | tee gpurun_out/pipe_grid16_${TAG}.log
GRID_LAGS=0,100,130,180 timeout 300 python tools/pipe_grid.py 15 2>&1 | tee gpurun_out/pipe_grid15_${TAG}.log
GRID_LAGS=0,8,12,20 timeout 300 python tools/pipe_grid.py 20 2>&1 | tee gpurun_out/pipe_grid20_${TAG}.log
GRID_LAGS=0,20,30,40 timeout 300 python tools/pipe_grid.py 18 2>&1 | tee gpurun_out/pipe_grid18_${TAG}.log
echo "done ${SECONDS}s"
