// launch.h -- internal interface between the C ABI layer (api.cu) and the kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include "fft_kernel.cuh"

namespace ckb {

// one variant per translation unit (fft_variants.cu is compiled four times)
cudaError_t launch_c2c_fwd(int M, const KernelParams& p, cudaStream_t s);
cudaError_t launch_c2c_inv(int M, const KernelParams& p, cudaStream_t s);
cudaError_t launch_r2c(int M, const KernelParams& p, cudaStream_t s);
cudaError_t launch_c2r(int M, const KernelParams& p, cudaStream_t s);

// tiny sizes (tiny.cu)
cudaError_t launch_tiny_c2c(int n, bool inverse, const KernelParams& p, cudaStream_t s);
cudaError_t launch_tiny_r2c(int n, const float* in, cf* out, const cf* table, int log2_nt, long long batch,
                            long long in_stride, long long out_stride, cudaStream_t s);
cudaError_t launch_tiny_c2r(int n, const cf* in, float* out, const cf* table, int log2_nt, long long batch,
                            long long in_stride, long long out_stride, cudaStream_t s);

void count_launch();                  // api.cu: global launch counter
int sm_count_of_current_device();     // api.cu: cached multiProcessorCount

struct PlanRow { int M, E, R0, R1, R2, G, MINB, smem_bytes; };
const PlanRow* find_plan(int M);      // launch.cu: nullptr if M is not a single-pass length

}  // namespace ckb
