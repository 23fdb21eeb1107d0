// fft_variants.cu -- instantiates the single-pass kernel family for ONE (direction, mode) variant.
// Compiled seven times (-DCKB_VARIANT=0..6) so the heavy template instantiations build in parallel:
//   0  complex forward   (CkFftComplexForward,  reference src/ckfft/ckfft.cpp:78-95)
//   1  complex inverse   (CkFftComplexInverse,  :97-114)
//   2  real forward      (CkFftRealForward,     :36-53)   half-length forward FFT + fused split
//   3  real inverse      (CkFftRealInverse,     :55-76)   fused twist + half-length inverse FFT
//   4  audio front end   (no reference counterpart) window * frame -> real forward -> |Y|^2
//   5  complex forward, split-complex ("planar") arrays   (callers that hold re[] / im[] separately)
//   6  complex inverse, split-complex arrays
#include "launch.h"
#include "plans.h"
#include <stdint.h>
#include <stdlib.h>

#ifndef CKB_VARIANT
#error "compile with -DCKB_VARIANT=0..6"
#endif

namespace ckb {

#if CKB_VARIANT == 0
#define CKB_FN launch_c2c_fwd
static constexpr bool kInv = false; static constexpr int kMode = MODE_C2C;
#elif CKB_VARIANT == 1
#define CKB_FN launch_c2c_inv
static constexpr bool kInv = true; static constexpr int kMode = MODE_C2C;
#elif CKB_VARIANT == 2
#define CKB_FN launch_r2c
static constexpr bool kInv = false; static constexpr int kMode = MODE_R2C;
#elif CKB_VARIANT == 3
#define CKB_FN launch_c2r
static constexpr bool kInv = true; static constexpr int kMode = MODE_C2R;
#elif CKB_VARIANT == 4
#define CKB_FN launch_r2c_audio
static constexpr bool kInv = false; static constexpr int kMode = MODE_R2C;
#elif CKB_VARIANT == 5
#define CKB_FN launch_c2c_fwd_planar
static constexpr bool kInv = false; static constexpr int kMode = MODE_C2C;
#else
#define CKB_FN launch_c2c_inv_planar
static constexpr bool kInv = true; static constexpr int kMode = MODE_C2C;
#endif
static constexpr bool kAudio = CKB_VARIANT == 4;
static constexpr bool kPlanar = CKB_VARIANT >= 5;

static int prefetch_mode()
{
    // CKFFT_B200_PREFETCH=0 disables the bulk-prefetch kernels (development switch, read once)
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("CKFFT_B200_PREFETCH");
        mode = (e && e[0] == '0') ? 0 : 1;
    }
    return mode;
}

template <class C>
static cudaError_t launch_cfg(const KernelParams& p, cudaStream_t s)
{
    // persistent grid: as many CTAs as fit on the device at once (a multiple of the SM count),
    // or fewer when the batch is small.  Cached per device.
    static int grid_cap[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (grid_cap[dev] == 0) {
        e = cudaFuncSetAttribute(fft_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_kernel<C>, C::THREADS, C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        grid_cap[dev] = occ * sm_count_of_current_device();
    }
    const long long ctas = (p.batch + C::G - 1) / C::G;
    const int grid = (int) (ctas < grid_cap[dev] ? ctas : grid_cap[dev]);
    if (grid <= 0) return cudaSuccess;
    fft_kernel<C><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t CKB_FN(int M, const KernelParams& p, cudaStream_t s)
{
    // development A/B builds (tools/exp_build.sh): -DCKB_EXP_A_M=2048 -DCKB_EXP_A_R2=2 -DCKB_EXP_A_G=3 -DCKB_EXP_A_MINB=2
    // -DCKB_EXP_A_TWR=1 -DCKB_EXP_A_PF=PF_DOUBLE (and _B_) route one length to an experimental three-stage plan (aligned rows)
#ifndef CKB_EXP_A_E
#define CKB_EXP_A_E 32
#define CKB_EXP_A_R0 32
#define CKB_EXP_A_R1 32
#endif
#ifndef CKB_EXP_B_E
#define CKB_EXP_B_E 32
#define CKB_EXP_B_R0 32
#define CKB_EXP_B_R1 32
#endif
#if defined(CKB_EXP_A_M) && CKB_VARIANT != 3 && CKB_VARIANT < 5
    if (((uintptr_t) p.in & 15) == 0 && (p.in_stride & 1) == 0) {
        if (M == CKB_EXP_A_M)
            return launch_cfg<Cfg<CKB_EXP_A_M, CKB_EXP_A_E, CKB_EXP_A_R0, CKB_EXP_A_R1, CKB_EXP_A_R2, CKB_EXP_A_G, kInv, kMode, CKB_EXP_A_MINB, CKB_EXP_A_PF, CKB_EXP_A_TWR != 0, kAudio>>(p, s);
#ifdef CKB_EXP_B_M
        if (M == CKB_EXP_B_M)
            return launch_cfg<Cfg<CKB_EXP_B_M, CKB_EXP_B_E, CKB_EXP_B_R0, CKB_EXP_B_R1, CKB_EXP_B_R2, CKB_EXP_B_G, kInv, kMode, CKB_EXP_B_MINB, CKB_EXP_B_PF, CKB_EXP_B_TWR != 0, kAudio>>(p, s);
#endif
    }
#endif
#if CKB_VARIANT != 3 && CKB_VARIANT < 5
    // bulk copies need 16-byte aligned rows
    if (prefetch_mode() && ((uintptr_t) p.in & 15) == 0 && (p.in_stride & 1) == 0) {
        switch (M) {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_DOUBLE, TWR_ != 0, kAudio>>(p, s);
            CKB_PREFETCH_PLANS(X)
#undef X
            default: break;
        }
    }
#endif
#if CKB_VARIANT <= 1
    if (prefetch_mode() && ((uintptr_t) p.in & 15) == 0 && (p.in_stride & 1) == 0) {
        switch (M) {           // split prefetch where the plan table has it, else in place
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_SPLIT, TWR_ != 0, kAudio>>(p, s);
            CKB_SPLIT_PREFETCH_PLANS_C2C(X)
#undef X
            default: break;
        }
        switch (M) {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_INPLACE, TWR_ != 0, kAudio>>(p, s);
            CKB_INPLACE_PREFETCH_PLANS(X)
#undef X
            default: break;
        }
    }
#elif CKB_VARIANT == 2
    if (prefetch_mode() && ((uintptr_t) p.in & 15) == 0 && (p.in_stride & 1) == 0) {
        switch (M) {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_SPLIT, TWR_ != 0, kAudio>>(p, s);
            CKB_SPLIT_PREFETCH_PLANS_R2C(X)
#undef X
            default: break;
        }
    }
#elif CKB_VARIANT == 4
    if (prefetch_mode() && ((uintptr_t) p.in & 15) == 0 && (p.in_stride & 1) == 0) {
        switch (M) {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_INPLACE, TWR_ != 0, kAudio>>(p, s);
            CKB_INPLACE_PREFETCH_PLANS_AUDIO(X)
#undef X
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_SPLIT, TWR_ != 0, kAudio>>(p, s);
            CKB_SPLIT_PREFETCH_PLANS_AUDIO(X)
#undef X
            default: break;
        }
    }
#endif
#if CKB_VARIANT >= 5
    // split-complex rows: in-place prefetch when both planes are 16-byte aligned row by row
    if (prefetch_mode() && (((uintptr_t) p.in | (uintptr_t) p.in_im) & 15) == 0 && (p.in_stride & 3) == 0) {
        switch (M) {           // split prefetch (the two planes are the two halves) where the plan table has it, else in place
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_SPLIT, TWR_ != 0, kAudio, true>>(p, s);
            CKB_SPLIT_PREFETCH_PLANS_PLANAR(X)
#undef X
            default: break;
        }
        switch (M) {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_INPLACE, TWR_ != 0, kAudio, true>>(p, s);
            CKB_INPLACE_PREFETCH_PLANS_PLANAR(X)
#undef X
            default: break;
        }
    }
#endif
#if CKB_VARIANT == 3
    KernelParams rest = p;          // what the plain-load kernels below still have to do
    if (prefetch_mode() && p.batch > 0 && ((uintptr_t) p.in & 15) == 0) {      // (first row: nothing to read below it)
        // Rows start on 8-byte boundaries; the kernel copies M+2 values from the 16-byte boundary below each row.
        // If that would run past the end of the LAST row (its pad is 0 and rows are dense), that row goes to the
        // plain kernel instead.
        const uintptr_t last = (uintptr_t) (p.in + (p.batch - 1) * p.in_stride);
        const bool last_overruns = ((last >> 3) & 1) == 0 && p.in_stride <= M + 1;
        KernelParams head = p;
        if (last_overruns) head.batch = p.batch - 1;
        cudaError_t e = cudaErrorInvalidValue;
        switch (M) {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: e = launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_INPLACE, TWR_ != 0, kAudio>>(head, s); break;
            CKB_INPLACE_PREFETCH_PLANS_C2R(X)
#undef X
            default: break;
        }
        if (e == cudaSuccess) {
            if (!last_overruns) return cudaSuccess;
            // the last row through the plain-load kernel of the SAME plan (same twiddle source, so the same bits whichever
            // rows a caller's chunking turns into last rows)
            rest.in = p.in + (p.batch - 1) * p.in_stride;
            rest.out = p.out + (p.batch - 1) * p.out_stride;
            rest.batch = 1;
            switch (M) {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_NONE, TWR_ != 0, kAudio>>(rest, s);
                CKB_INPLACE_PREFETCH_PLANS_C2R(X)
#undef X
                default: return cudaErrorInvalidValue;
            }
        } else if (e != cudaErrorInvalidValue) {
            return e;
        }
    }
    const KernelParams& q = rest;
#else
    const KernelParams& q = p;
#endif
    switch (M) {
#define X(M_, E_, R0_, R1_, R2_, G_, MINB_, TWR_) \
    case M_: return launch_cfg<Cfg<M_, E_, R0_, R1_, R2_, G_, kInv, kMode, MINB_, PF_NONE, TWR_ != 0, kAudio, kPlanar>>(q, s);
        CKB_SINGLE_PASS_PLANS(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ckb
