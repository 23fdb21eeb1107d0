// dist_glue.cu -- device helpers for the distributed six-step transform of one very large 1-D FFT
// (same index algebra as the vendored ext/fftw-3.3.2/mpi/dft-rank1.c:58-79, which the
// reference ships but never builds).  The exchange itself (all-to-all over NVLink) is done by the caller with
// NCCL; these kernels are the local steps around it:
//   pack    [rows][parts][w]          -> [parts][rows][w]        (slab q = the columns that go to peer q)
//   unpack  [parts][rowsPer][w]       -> [w][parts*rowsPer]      (tiled transpose of what the peers sent)
//   twiddle data[i][k] *= W_n^((firstRow+i)*k)                   (the inter-step twiddle of the six-step)
#include "launch.h"

namespace ckb {

__global__ void __launch_bounds__(256) pack_columns_kernel(const cf* __restrict__ in, cf* __restrict__ out,
                                                           long long rows, int parts, long long w)
{
    const long long total = rows * parts * w;
    for (long long o = blockIdx.x * (long long) blockDim.x + threadIdx.x; o < total; o += (long long) gridDim.x * blockDim.x) {
        const long long c = o % w;
        const long long i = (o / w) % rows;
        const long long q = o / (w * rows);
        out[o] = __ldcs(in + (i * parts + q) * w + c);
    }
}

// in: [parts][rowsPer][w]; out: [w][parts*rowsPer].  32x32 tiles through padded shared memory.
__global__ void __launch_bounds__(256) unpack_transpose_kernel(const cf* __restrict__ in, cf* __restrict__ out,
                                                               int parts, long long rowsPer, long long w)
{
    __shared__ cf tile[32][33];
    const long long tiles_i = (rowsPer + 31) / 32, tiles_c = (w + 31) / 32;
    const long long ntiles = (long long) parts * tiles_i * tiles_c;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const long long R = (long long) parts * rowsPer;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long s = t / (tiles_i * tiles_c);
        const long long i0 = ((t / tiles_c) % tiles_i) * 32, c0 = (t % tiles_c) * 32;
        const cf* src = in + s * rowsPer * w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long i = i0 + ty + 8 * k, c = c0 + tx;
            if (i < rowsPer && c < w) tile[ty + 8 * k][tx] = __ldcs(src + i * w + c);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long c = c0 + ty + 8 * k, i = i0 + tx;
            if (i < rowsPer && c < w) __stcs(out + c * R + s * rowsPer + i, tile[tx][ty + 8 * k]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) twiddle_rows_kernel(cf* __restrict__ data, long long rows, long long cols,
                                                           long long first_row, BigTwiddles tw, int shift, bool inverse)
{
    const long long total = rows * cols;
    for (long long o = blockIdx.x * (long long) blockDim.x + threadIdx.x; o < total; o += (long long) gridDim.x * blockDim.x) {
        const unsigned long long e64 = (unsigned long long) (first_row + o / cols) * (unsigned long long) (o % cols);
        const unsigned e = (unsigned) (e64 << shift);            // (row * col) < n <= 2^30 by contract
        cf w = cmul(__ldg(tw.lo + (e & ((1u << tw.h) - 1u))), __ldg(tw.hi + (e >> tw.h)));
        if (inverse) w.y = -w.y;
        data[o] = cmul(data[o], w);
    }
}

static int glue_blocks(long long items, int per_block)
{
    long long blocks = (items + per_block - 1) / per_block;
    const long long cap = 32LL * sm_count_of_current_device();
    if (blocks > cap) blocks = cap;
    return (int) (blocks < 1 ? 1 : blocks);
}

cudaError_t launch_pack_columns(const cf* in, cf* out, long long rows, int parts, long long w, cudaStream_t s)
{
    pack_columns_kernel<<<glue_blocks(rows * parts * w, 256 * 8), 256, 0, s>>>(in, out, rows, parts, w);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_unpack_transpose(const cf* in, cf* out, int parts, long long rowsPer, long long w, cudaStream_t s)
{
    const long long ntiles = (long long) parts * ((rowsPer + 31) / 32) * ((w + 31) / 32);
    unpack_transpose_kernel<<<glue_blocks(ntiles, 1), 256, 0, s>>>(in, out, parts, rowsPer, w);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_twiddle_rows(cf* data, long long rows, long long cols, long long first_row, const BigTwiddles& tw,
                                int log2n, bool inverse, cudaStream_t s)
{
    twiddle_rows_kernel<<<glue_blocks(rows * cols, 256 * 8), 256, 0, s>>>(data, rows, cols, first_row, tw,
                                                                          tw.log2_tmax - log2n, inverse);
    count_launch();
    return cudaGetLastError();
}

}  // namespace ckb
