// tiny.cu -- launchers for the thread-per-transform kernels (complex n <= 8, real n <= 16).
#include "launch.h"
#include "tiny_kernel.cuh"
#include "plans.h"

namespace ckb {

static int tiny_grid(long long batch)
{
    long long blocks = (batch + 127) / 128;
    const long long cap = 16LL * sm_count_of_current_device();
    if (blocks > cap) blocks = cap;
    return (int) (blocks < 1 ? 1 : blocks);
}

template <int M>
static cudaError_t tiny_c2c(bool inverse, const KernelParams& p, cudaStream_t s)
{
    if (p.in_im != nullptr) {           // split-complex rows
        if (inverse) tiny_c2c_kernel<M, true, true><<<tiny_grid(p.batch), 128, 0, s>>>(p);
        else         tiny_c2c_kernel<M, false, true><<<tiny_grid(p.batch), 128, 0, s>>>(p);
    } else {
        if (inverse) tiny_c2c_kernel<M, true><<<tiny_grid(p.batch), 128, 0, s>>>(p);
        else         tiny_c2c_kernel<M, false><<<tiny_grid(p.batch), 128, 0, s>>>(p);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_tiny_c2c(int n, bool inverse, const KernelParams& p, cudaStream_t s)
{
    if (p.batch <= 0) return cudaSuccess;
    switch (n) {
        case 1: return tiny_c2c<1>(inverse, p, s);
        case 2: return tiny_c2c<2>(inverse, p, s);
        case 4: return tiny_c2c<4>(inverse, p, s);
        case 8: return tiny_c2c<8>(inverse, p, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_tiny_r2c(int n, const float* in, cf* out, const cf* table, int log2_nt, long long batch,
                            long long in_stride, long long out_stride, cudaStream_t s)
{
    if (batch <= 0) return cudaSuccess;
    const int grid = tiny_grid(batch);
    switch (n) {
#define CASE(N) case N: tiny_r2c_kernel<N><<<grid, 128, 0, s>>>(in, out, table, log2_nt, batch, in_stride, out_stride); break;
        CASE(1) CASE(2) CASE(4) CASE(8) CASE(16)
#undef CASE
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_tiny_c2r(int n, const cf* in, float* out, const cf* table, int log2_nt, long long batch,
                            long long in_stride, long long out_stride, cudaStream_t s)
{
    if (batch <= 0) return cudaSuccess;
    const int grid = tiny_grid(batch);
    switch (n) {
#define CASE(N) case N: tiny_c2r_kernel<N><<<grid, 128, 0, s>>>(in, out, table, log2_nt, batch, in_stride, out_stride); break;
        CASE(1) CASE(2) CASE(4) CASE(8) CASE(16)
#undef CASE
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

static const PlanRow kPlans[] = {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    { M_, E_, R0_, R1_, R2_, G_, MINB_, Cfg<M_, E_, R0_, R1_, R2_, G_, false, MODE_C2C, MINB_, PF_NONE, TWR_ != 0>::SMEM_BYTES },
    CKB_SINGLE_PASS_PLANS(X)
#undef X
};

const PlanRow* find_plan(int M)
{
    for (const PlanRow& r : kPlans)
        if (r.M == M) return &r;
    return nullptr;
}

}  // namespace ckb
