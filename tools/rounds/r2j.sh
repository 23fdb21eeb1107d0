export CKFFT_B200_LIB=$PWD/ckfft_b200/lib/libckfft_b200_stats.so
for k in c2r r2c c2c; do for n in 131072 1048576; do python tools/prof_one.py $k $n 2>&1 | grep pipe_stats | tail -1; done; done
unset CKFFT_B200_LIB
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 2 -c 1 -o gpurun_out/prof_pipe_c2r_131072_r2j python tools/prof_one.py c2r 131072 > gpurun_out/ncu_c2r_r2j.log 2>&1; tail -1 gpurun_out/ncu_c2r_r2j.log
