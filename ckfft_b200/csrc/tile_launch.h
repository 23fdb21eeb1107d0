// tile_launch.h -- host-side launch helpers of the multi-pass tile kernels (four_step.cuh), shared by the local
// four-step path (four_step.cu) and the routed passes of the fused distributed transform (dist_fused.cu).
#pragma once
#include "four_step.cuh"
#include "launch.h"
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

// 3-D view [nz][ny][nx] of 8-byte elements with byte strides {ystride, zstride} (multiples of 16), boxes of 1 x box_rows x box_cols
static bool make_tile_map3(CUtensorMap* map, const void* base, long long nx, long long ny, long long nz, long long ystride_bytes,
                           long long zstride_bytes, int box_rows, int box_cols)
{
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc || ((uintptr_t) base & 15) || (ystride_bytes & 15) || (zstride_bytes & 15) || nx <= 0 || ny <= 0 || nz <= 0) return false;
    const cuuint64_t gdim[3] = { (cuuint64_t) nx, (cuuint64_t) ny, (cuuint64_t) nz };
    const cuuint64_t gstride[2] = { (cuuint64_t) ystride_bytes, (cuuint64_t) zstride_bytes };
    const cuuint32_t box[3] = { (cuuint32_t) box_cols, (cuuint32_t) box_rows, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*) base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
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

template <bool INV, int KIND, bool RT = false>
static cudaError_t launch_tile(int L, const TileParams& p, cudaStream_t s)
{
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    // TMA-staged variant: needs 16-byte aligned data and, for the column pass, a tensor map of the whole array
    bool pft = tile_prefetch_mode() && ((uintptr_t) p.in & 15) == 0;
    if (pft && KIND == KIND_COLUMN) {
        switch (L) {
#define X(L_, E_, R0_, R1_, C_, MINB_) \
    case L_: pft = make_tile_map(&tmap, p.in, p.nproblems * L_, p.ncols, TileCfg<L_, E_, R0_, R1_, C_, INV, KIND, MINB_, true, RT>::BOX_ROWS, C_); break;
            CKB_TILE_PLANS(X)
#undef X
            default: pft = false;
        }
    }
    if (pft) {
        switch (L) {
#define X(L_, E_, R0_, R1_, C_, MINB_) \
    case L_: return launch_tile_cfg<TileCfg<L_, E_, R0_, R1_, C_, INV, KIND, MINB_, true, RT>>(p, tmap, s);
            CKB_TILE_PLANS(X)
#undef X
            default: return cudaErrorInvalidValue;
        }
    }
    switch (L) {
#define X(L_, E_, R0_, R1_, C_, MINB_) \
    case L_: return launch_tile_cfg<TileCfg<L_, E_, R0_, R1_, C_, INV, KIND, MINB_, false, RT>>(p, tmap, s);
        CKB_TILE_PLANS(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

// four_step.cu: the local (non-routed) tile passes, instantiated once
cudaError_t launch_local_pass(bool inverse, int kind, int L, const TileParams& p, cudaStream_t s);

}  // namespace ckb
