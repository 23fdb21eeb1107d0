// four_step.cu -- host planner and launchers of the multi-pass path (see four_step.cuh).
#include "four_step.cuh"
#include "launch.h"
#include "plans.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace ckb {

// tile plans: X(L, E, R0, R1, C, MINB)
#define CKB_TILE_PLANS(X) \
    X(128,  16, 16,  8, 16, 4) \
    X(256,  16, 16, 16, 16, 3) \
    X(512,  32, 32, 16, 16, 2) \
    X(1024, 32, 32, 32,  8, 2)

void four_step_plan(int log2n, int* npass, int L[3])
{
    L[0] = L[1] = L[2] = 0;
    if (log2n <= 20) {
        const int a1 = log2n / 2;
        *npass = 2;
        L[0] = 1 << a1;
        L[1] = 1 << (log2n - a1);          // the contiguous last pass takes the longer factor
    } else {
        const int a1 = log2n / 3, a2 = (log2n - a1) / 2;
        *npass = 3;
        L[0] = 1 << a1;
        L[1] = 1 << a2;
        L[2] = 1 << (log2n - a1 - a2);
    }
}

// ---- tensor maps for the TMA-staged column pass ----------------------------------------------------------
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TensorMapEncodeTiledFn tensor_map_encoder()
{
    static TensorMapEncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (TensorMapEncodeTiledFn) f;
    }();
    return fn;
}

// [rows][ncols] array of 8-byte elements, boxes of `box_rows` x `box_cols`
static bool make_tile_map(CUtensorMap* map, const cf* base, long long rows, int ncols, int box_rows, int box_cols)
{
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc || ((uintptr_t) base & 15) || rows <= 0 || rows >= (1LL << 32)) return false;
    const cuuint64_t gdim[2] = { (cuuint64_t) ncols, (cuuint64_t) rows };
    const cuuint64_t gstride[1] = { (cuuint64_t) ncols * 8 };
    const cuuint32_t box[2] = { (cuuint32_t) box_cols, (cuuint32_t) box_rows };
    const cuuint32_t estr[2] = { 1, 1 };
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*) base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int tile_prefetch_mode()
{
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("CKFFT_B200_PREFETCH");
        mode = (e && e[0] == '0') ? 0 : 1;
    }
    return mode;
}

template <class TC>
static cudaError_t launch_tile_cfg(const TileParams& p, const CUtensorMap& tmap, cudaStream_t s)
{
    static int grid_cap[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (grid_cap[dev] == 0) {
        e = cudaFuncSetAttribute(tile_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile_kernel<TC>, TC::THREADS, TC::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        grid_cap[dev] = occ * sm_count_of_current_device();
    }
    if (p.ncols % TC::C != 0) return cudaErrorInvalidValue;
    const long long tiles = p.nproblems * (p.ncols / TC::C);
    const int grid = (int) (tiles < grid_cap[dev] ? tiles : grid_cap[dev]);
    if (grid <= 0) return cudaSuccess;
    tile_kernel<TC><<<grid, TC::THREADS, TC::SMEM_BYTES, s>>>(p, tmap);
    count_launch();
    return cudaGetLastError();
}

template <bool INV, int KIND>
static cudaError_t launch_tile(int L, const TileParams& p, cudaStream_t s)
{
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    // TMA-staged variant: needs 16-byte aligned data and, for the column pass, a tensor map of the whole array
    bool pft = tile_prefetch_mode() && ((uintptr_t) p.in & 15) == 0;
    if (pft && KIND == KIND_COLUMN) {
        switch (L) {
#define X(L_, E_, R0_, R1_, C_, MINB_) \
    case L_: pft = make_tile_map(&tmap, p.in, p.nproblems * L_, p.ncols, TileCfg<L_, E_, R0_, R1_, C_, INV, KIND, MINB_, true>::BOX_ROWS, C_); break;
            CKB_TILE_PLANS(X)
#undef X
            default: pft = false;
        }
    }
    if (pft) {
        switch (L) {
#define X(L_, E_, R0_, R1_, C_, MINB_) \
    case L_: return launch_tile_cfg<TileCfg<L_, E_, R0_, R1_, C_, INV, KIND, MINB_, true>>(p, tmap, s);
            CKB_TILE_PLANS(X)
#undef X
            default: return cudaErrorInvalidValue;
        }
    }
    switch (L) {
#define X(L_, E_, R0_, R1_, C_, MINB_) \
    case L_: return launch_tile_cfg<TileCfg<L_, E_, R0_, R1_, C_, INV, KIND, MINB_, false>>(p, tmap, s);
        CKB_TILE_PLANS(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

static cudaError_t launch_pass(bool inverse, int kind, int L, const TileParams& p, cudaStream_t s)
{
    if (kind == KIND_COLUMN) return inverse ? launch_tile<true, KIND_COLUMN>(L, p, s) : launch_tile<false, KIND_COLUMN>(L, p, s);
    return inverse ? launch_tile<true, KIND_LAST>(L, p, s) : launch_tile<false, KIND_LAST>(L, p, s);
}

// `batch` dense transforms of n = 2^log2n complex points: in -> out, `scratch` holds batch*n complex values.
cudaError_t launch_four_step(bool inverse, int log2n, const cf* in, cf* out, cf* scratch, long long batch,
                             const cf* table, int log2_nt, const BigTwiddles& tw, cudaStream_t s)
{
    int npass, L[3];
    four_step_plan(log2n, &npass, L);
    const long long n = 1LL << log2n;
    TileParams p{};
    p.table = table; p.log2_nt = log2_nt;
    p.tw_lo = tw.lo; p.tw_hi = tw.hi; p.tw_h = tw.h;
    cudaError_t e;
    // Cache policy: the user's input is read once and the final output written once (streaming, evict-first);
    // intermediates are written with the default policy so that the next pass finds them in the 126 MB L2 -- the
    // caller keeps `batch` small enough for that (api.cu, enqueue_large_c2c).
    if (npass == 2) {
        p.in = in; p.out = scratch; p.nproblems = batch; p.ncols = L[1]; p.tw_shift = tw.log2_tmax - log2n; p.P = 1; p.Q = 1;
        p.stream_in = 1; p.stream_out = 0;
        if ((e = launch_pass(inverse, KIND_COLUMN, L[0], p, s)) != cudaSuccess) return e;
        p.in = scratch; p.out = out; p.nproblems = batch; p.ncols = L[0]; p.P = L[0]; p.Q = 1;
        p.stream_in = 0; p.stream_out = 1;
        return launch_pass(inverse, KIND_LAST, L[1], p, s);
    }
    p.in = in; p.out = out; p.nproblems = batch; p.ncols = (int) (n / L[0]); p.tw_shift = tw.log2_tmax - log2n; p.P = 1; p.Q = 1;
    p.stream_in = 1; p.stream_out = 0;
    if ((e = launch_pass(inverse, KIND_COLUMN, L[0], p, s)) != cudaSuccess) return e;
    p.in = out; p.out = scratch; p.nproblems = batch * L[0]; p.ncols = L[2];
    p.tw_shift = tw.log2_tmax - ilog2(L[1] * L[2]);
    p.stream_in = 0; p.stream_out = 0;
    if ((e = launch_pass(inverse, KIND_COLUMN, L[1], p, s)) != cudaSuccess) return e;
    p.in = scratch; p.out = out; p.nproblems = batch; p.ncols = L[0] * L[1]; p.P = L[0]; p.Q = L[1];
    p.stream_in = 0; p.stream_out = 1;
    return launch_pass(inverse, KIND_LAST, L[2], p, s);
}

static int glue_grid(long long items)
{
    long long blocks = (items + 255) / 256;
    const long long cap = 32LL * sm_count_of_current_device();
    if (blocks > cap) blocks = cap;
    return (int) (blocks < 1 ? 1 : blocks);
}

cudaError_t launch_real_split(const cf* z, cf* y, int n, long long batch, long long z_stride, long long y_stride,
                              const BigTwiddles& tw, cudaStream_t s)
{
    RealGlueParams p{ z, y, tw.lo, tw.hi, tw.h, tw.log2_tmax - ilog2(n), batch, n / 2, z_stride, y_stride };
    real_split_kernel<<<glue_grid(batch * (n / 4 + 1)), 256, 0, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_real_twist(const cf* y, cf* t, int n, long long batch, long long y_stride, long long t_stride,
                              const BigTwiddles& tw, cudaStream_t s)
{
    RealGlueParams p{ y, t, tw.lo, tw.hi, tw.h, tw.log2_tmax - ilog2(n), batch, n / 2, y_stride, t_stride };
    real_twist_kernel<<<glue_grid(batch * (n / 4 + 1)), 256, 0, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace ckb
