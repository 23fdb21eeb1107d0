// small.cu -- launchers of the thread-per-transform tile kernels (small_kernel.cuh): complex 8 .. 64 points,
// real 16 .. 64 points.
#include "launch.h"
#include "small_kernel.cuh"
#include <stdlib.h>

namespace ckb {

bool small_enabled()
{
    // CKFFT_B200_SMALL=0 sends these lengths back to the cooperative / tiny kernels (development switch, read once)
    static const bool on = [] { const char* e = getenv("CKFFT_B200_SMALL"); return !(e && e[0] == '0'); }();
    return on;
}

template <int M, int MODE, bool INV, bool PLANAR>
static cudaError_t small_launch(const KernelParams& p, cudaStream_t s)
{
    using SC = SmallCfg<M, MODE>;
    auto kern = small_kernel<M, MODE, INV, PLANAR>;
    static int grid_cap[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (grid_cap[dev] == 0) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SC::SMEM_BYTES)) != cudaSuccess) return e;
        int occ = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SC::THREADS, SC::SMEM_BYTES)) != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        grid_cap[dev] = occ * sm_count_of_current_device();
    }
    const long long tiles = (p.batch + SC::ROWS - 1) / SC::ROWS;
    const int grid = (int) (tiles < grid_cap[dev] ? tiles : grid_cap[dev]);
    if (grid <= 0) return cudaSuccess;
    kern<<<grid, SC::THREADS, SC::SMEM_BYTES, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <int MODE, bool INV, bool PLANAR>
static cudaError_t small_dispatch(int M, const KernelParams& p, cudaStream_t s)
{
    switch (M) {
        case 8:  return small_launch<8, MODE, INV, PLANAR>(p, s);
        case 16: return small_launch<16, MODE, INV, PLANAR>(p, s);
        case 32: return small_launch<32, MODE, INV, PLANAR>(p, s);
        case 64:
            if constexpr (MODE == MODE_C2C) return small_launch<64, MODE, INV, PLANAR>(p, s);
            else return cudaErrorInvalidValue;
        default: return cudaErrorInvalidValue;
    }
}

// complex M = 8, 16, 32, 64; p.in_im != nullptr: split-complex rows (strides in floats)
cudaError_t launch_small_c2c(int M, bool inverse, const KernelParams& p, cudaStream_t s)
{
    if (p.in_im != nullptr)
        return inverse ? small_dispatch<MODE_C2C, true, true>(M, p, s) : small_dispatch<MODE_C2C, false, true>(M, p, s);
    return inverse ? small_dispatch<MODE_C2C, true, false>(M, p, s) : small_dispatch<MODE_C2C, false, false>(M, p, s);
}

// real n = 2M = 16, 32, 64; both arrays addressed in 8-byte units (even strides of the real array)
cudaError_t launch_small_r2c(int M, const KernelParams& p, cudaStream_t s) { return small_dispatch<MODE_R2C, false, false>(M, p, s); }
cudaError_t launch_small_c2r(int M, const KernelParams& p, cudaStream_t s) { return small_dispatch<MODE_C2R, true, false>(M, p, s); }

}  // namespace ckb
