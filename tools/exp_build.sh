#!/bin/bash
# Development A/B library: recompile ONE variant of fft_variants.cu with extra flags and link it with the product objects.
#   tools/exp_build.sh <tag> <variant 0..6> <nvcc flags...>     -> ckfft_b200/lib/libckfft_b200_<tag>.so
set -e
tag=$1; var=$2; shift 2
names=(c2c_fwd c2c_inv r2c c2r r2c_audio c2c_fwd_planar c2c_inv_planar)
D=ckfft_b200/build; O=/tmp/exp_build_$tag; mkdir -p $O
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=default -Iinclude -Ickfft_b200/csrc \
     -DCKB_VARIANT=$var "$@" -Xptxas -v -c ckfft_b200/csrc/fft_variants.cu -o $O/${names[$var]}.o > $O/ptxas.log 2>&1
objs=""
for f in $D/*.o; do b=$(basename $f); [ $b = ${names[$var]}.o ] && objs="$objs $O/$b" || objs="$objs $f"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ckfft_b200/lib/libckfft_b200_$tag.so $objs
echo "built ckfft_b200/lib/libckfft_b200_$tag.so"; grep -c "error" $O/ptxas.log || true
