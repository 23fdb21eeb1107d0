// launch.h -- internal interface between the C ABI layer (api.cu) and the kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include "fft_kernel.cuh"
#include "ckfft/ckfft_b200.h"

struct CUtensorMap_st;                 // <cuda.h>; only pointers to it cross this interface

namespace ckb {

// one variant per translation unit (fft_variants.cu is compiled once per variant)
cudaError_t launch_c2c_fwd(int M, const KernelParams& p, cudaStream_t s);
cudaError_t launch_c2c_inv(int M, const KernelParams& p, cudaStream_t s);
cudaError_t launch_r2c(int M, const KernelParams& p, cudaStream_t s);
cudaError_t launch_c2r(int M, const KernelParams& p, cudaStream_t s);
cudaError_t launch_r2c_audio(int M, const KernelParams& p, cudaStream_t s);   // window + real forward + power spectrum
cudaError_t launch_c2c_fwd_planar(int M, const KernelParams& p, cudaStream_t s);   // split-complex arrays (p.in_im / p.out_im)
cudaError_t launch_c2c_inv_planar(int M, const KernelParams& p, cudaStream_t s);

// tiny sizes (tiny.cu)
cudaError_t launch_tiny_c2c(int n, bool inverse, const KernelParams& p, cudaStream_t s);
cudaError_t launch_tiny_r2c(int n, const float* in, cf* out, const cf* table, int log2_nt, long long batch,
                            long long in_stride, long long out_stride, cudaStream_t s);
cudaError_t launch_tiny_c2r(int n, const cf* in, float* out, const cf* table, int log2_nt, long long batch,
                            long long in_stride, long long out_stride, cudaStream_t s);

// short lengths, one thread per transform over shared-memory tiles (small.cu): complex 8 .. 32, real 16 .. 64 points
bool small_enabled();
cudaError_t launch_small_c2c(int M, bool inverse, const KernelParams& p, cudaStream_t s);
cudaError_t launch_small_r2c(int M, const KernelParams& p, cudaStream_t s);      // strides in 8-byte units
cudaError_t launch_small_c2r(int M, const KernelParams& p, cudaStream_t s);

void count_launch();                  // api.cu: global launch counter
cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t s);   // api.cu: stream-ordered scratch from the library's own pool (free with cudaFreeAsync)
int sm_count_of_current_device();     // api.cu: cached multiProcessorCount

// multi-pass path (four_step.cu)
struct BigTwiddles { const cf* lo; const cf* hi; int h; int log2_tmax; };   // W_Tmax^e = hi[e >> h] * lo[e & (2^h-1)]
void four_step_plan(int log2n, int* npass, int L[3]);
cudaError_t launch_four_step(bool inverse, int log2n, const cf* in, cf* out, cf* scratch, long long batch,
                             const cf* table, int log2_nt, const BigTwiddles& tw, cudaStream_t s);
bool pipe_enabled();                  // two-pass lengths as one L2-resident dataflow kernel (pipe_kernel.cuh)
cudaError_t launch_pipe(bool inverse, int log2n, const cf* in, cf* out, long long batch, const cf* table, int log2_nt,
                        const BigTwiddles& tw, cudaStream_t s);
cudaError_t launch_pipe_r2c(int log2m, const cf* in, cf* out, long long batch, long long out_stride, const cf* table, int log2_nt,
                            const BigTwiddles& tw, cudaStream_t s);
cudaError_t launch_pipe_c2r(int log2m, const cf* in, cf* out, long long batch, long long in_stride, const cf* table, int log2_nt,
                            const BigTwiddles& tw, cudaStream_t s);
cudaError_t launch_real_split(const cf* z, cf* y, int n, long long batch, long long z_stride, long long y_stride,
                              const BigTwiddles& tw, cudaStream_t s);
cudaError_t launch_real_twist(const cf* y, cf* t, int n, long long batch, long long y_stride, long long t_stride,
                              const BigTwiddles& tw, cudaStream_t s);

// distributed six-step helpers (dist_glue.cu)
cudaError_t launch_pack_columns(const cf* in, cf* out, long long rows, int parts, long long w, cudaStream_t s);
cudaError_t launch_unpack_transpose(const cf* in, cf* out, int parts, long long rowsPer, long long w, cudaStream_t s);
cudaError_t launch_twiddle_rows(cf* data, long long rows, long long cols, long long first_row, const BigTwiddles& tw,
                                int log2n, bool inverse, cudaStream_t s);

// fused distributed transform (dist_fused.cu)
#ifndef CKB_MAX_PEERS
#define CKB_MAX_PEERS 8
#endif
struct DistBuffers {
    cf* buf[3][CKB_MAX_PEERS];          // [work, mid, out][rank]
    unsigned* flags[CKB_MAX_PEERS];     // per rank: arrive[CKB_MAX_PEERS] + error word
    cf* in[CKB_MAX_PEERS];              // pull mode: every rank's peer-visible input array
    const ::CUtensorMap_st* pull_maps;   // pull mode: device array of `world` tensor maps over those arrays
    int pull_box_rows;
};
cudaError_t dist_make_pull_maps(const CkFftB200DistLayout& l, int rank, cf* const* in, ::CUtensorMap_st* dmaps, int* box_rows_out);
bool dist_layout(long long n, int world, int prefer, CkFftB200DistLayout* out);
int dist_describe(const CkFftB200DistLayout& l, int rank, CkFftB200DistPass passes[4]);
struct DistMarks {                      // optional per-phase events of one execution (profiling)
    static constexpr int kMax = 12;
    cudaEvent_t ev[kMax];
    const char* name[kMax];
    int count;
};
cudaError_t dist_exec(const CkFftB200DistLayout& l, int rank, const DistBuffers& b, unsigned* epoch, const cf* in_local,
                      bool inverse, const cf* table, int log2_nt, const BigTwiddles& tw, cudaStream_t s, DistMarks* marks);

struct PlanRow { int M, E, R0, R1, R2, G, MINB, smem_bytes; };
const PlanRow* find_plan(int M);      // launch.cu: nullptr if M is not a single-pass length

}  // namespace ckb
