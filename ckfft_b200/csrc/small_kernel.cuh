// small_kernel.cuh -- short transforms (complex 8 .. 32 points, real 16 .. 64 points): one thread per transform, the
// CTA's tile of 128 rows staged through shared memory.
//
// The cooperative kernels (fft_kernel.cuh) give a 16- or 32-point transform to 4 threads: every global request then
// covers 32-byte pieces of 8 different rows, the stage twiddles go through a shared-memory table and the exchange
// between the two radix stages costs two more trips -- they reach 0.6-0.8 of the copy peak there (0.25-0.4 for the
// thread-per-transform kernels of tiny_kernel.cuh, whose global accesses are row-strided).  Here
//   * the 128 threads of a CTA copy the tile of 128 consecutive rows global -> shared with fully coalesced 8-byte
//     accesses (256 contiguous bytes per warp request; a dense batch is one contiguous span),
//   * every thread then owns one row: it reads it (row pitch odd in 8-byte units: conflict-free), runs the whole
//     transform as ONE register butterfly network (fft_regs.cuh: every twiddle a compile-time constant, no table, no
//     exchange), applies the real split / twist with compile-time factors W_2M^k = W_64^(k*32/M), and writes the
//     result back over its row,
//   * and the tile leaves shared -> global the way it came.
// 64-point complex rows take TWO threads per row (SmallCfg::TPR = 2): each runs a 32-point network on the even / odd
// samples, the odd half is multiplied by W_64^k (constants) and the final radix-2 step exchanges 16 values per thread
// with one warp shuffle each -- still no table and no shared-memory exchange.
// Same mathematics as the reference's small cases (src/ckfft/fft_default.cpp:22-167 leaves, fft_real_default.cpp:13-114).
#pragma once
#include "fft_kernel.cuh"

namespace ckb {

template <int M, int MODE>
struct SmallCfg {
    static constexpr int THREADS = 128;
    static constexpr int TPR = M > 32 ? M / 32 : 1;                    // threads per row
    static constexpr int ROWS = THREADS / TPR;                         // rows per tile
    static constexpr int NV = M / TPR;                                 // complex values per thread
    static constexpr int IN_E = MODE == MODE_C2R ? M + 1 : M;          // 8-byte elements per input row
    static constexpr int OUT_E = MODE == MODE_R2C ? M + 1 : M;         // 8-byte elements per output row
    // row pitch in the tile (8-byte words).  One thread per row: odd, so that 16 rows spread over the 16 bank pairs.
    // Two threads per row (adjacent lanes read adjacent words of the same row): 2 mod 16, so that 8 rows x 2 words do.
    static constexpr int PITCH = TPR == 1 ? ((M + 1) | 1) : M + 2;
    static constexpr int SMEM_BYTES = ROWS * PITCH * 8;
    // 32 points + split / twist want ~170 registers: two CTAs per SM without spills (.87 / .85 of the copy peak for real
    // n = 64) beat three with 32-40 bytes of spills (.79 / .77) and four with ~200 (slower still)
    static constexpr int MINB = (M == 32 && MODE != MODE_C2C) ? 2 : 4;
    static_assert(M >= 4 && M <= 64 && (M & (M - 1)) == 0, "one or two register networks per transform");
    static_assert(TPR == 1 || (TPR == 2 && MODE == MODE_C2C && PITCH % 16 == 2), "two threads per row: complex transforms");
    static_assert((ROWS * IN_E) % THREADS == 0 || TPR == 1, "whole copy rounds");
};

// tile <-> global, 8-byte elements, consecutive threads on consecutive elements of the [rows][E] tile
template <int E, int PITCH, int ROWS, bool LOAD>
__device__ __forceinline__ void small_copy(cf* tile, const cf* gin, cf* gout, long long stride, int rows, int tid)
{
    constexpr int THREADS = 128;
    constexpr int PER = (ROWS * E + THREADS - 1) / THREADS;       // elements per thread (full tile)
    if (rows == ROWS) {
        if constexpr (LOAD) {
            cf tmp[PER];                      // all loads in flight before the first shared-memory store
            static_for<0, PER>([&](auto i_) {
                const int e = decltype(i_)::value * THREADS + tid;
                if ((decltype(i_)::value + 1) * THREADS <= ROWS * E || e < ROWS * E)
                    tmp[decltype(i_)::value] = __ldcs(gin + (long long) (e / E) * stride + e % E);
            });
            static_for<0, PER>([&](auto i_) {
                const int e = decltype(i_)::value * THREADS + tid;
                if ((decltype(i_)::value + 1) * THREADS <= ROWS * E || e < ROWS * E)
                    tile[(e / E) * PITCH + e % E] = tmp[decltype(i_)::value];
            });
        } else {
            static_for<0, PER>([&](auto i_) {
                const int e = decltype(i_)::value * THREADS + tid;
                if ((decltype(i_)::value + 1) * THREADS <= ROWS * E || e < ROWS * E)
                    __stcs(gout + (long long) (e / E) * stride + e % E, tile[(e / E) * PITCH + e % E]);
            });
        }
    } else {
        for (int e = tid; e < rows * E; e += THREADS) {
            if constexpr (LOAD) tile[(e / E) * PITCH + e % E] = __ldcs(gin + (long long) (e / E) * stride + e % E);
            else                __stcs(gout + (long long) (e / E) * stride + e % E, tile[(e / E) * PITCH + e % E]);
        }
    }
}

// split-complex rows: the planes of real and imaginary parts are separate float arrays (strides in floats)
template <int M, int PITCH, int ROWS, bool LOAD>
__device__ __forceinline__ void small_copy_planar(cf* tile, const float* gre, const float* gim, float* ore, float* oim,
                                                  long long stride, int rows, int tid)
{
    constexpr int THREADS = 128;
    constexpr int PER = ROWS * M / THREADS;
    static_assert(ROWS * M % THREADS == 0, "whole copy rounds");
    if (rows == ROWS) {
        // 4-byte global accesses (128 contiguous bytes per plane and warp request), 8-byte shared-memory accesses
        if constexpr (LOAD) {
            cf tmp[PER];
            static_for<0, PER>([&](auto i_) {
                const int e = decltype(i_)::value * THREADS + tid;
                const long long g = (long long) (e / M) * stride + e % M;
                tmp[decltype(i_)::value] = make_float2(__ldcs(gre + g), __ldcs(gim + g));
            });
            static_for<0, PER>([&](auto i_) {
                const int e = decltype(i_)::value * THREADS + tid;
                tile[(e / M) * PITCH + e % M] = tmp[decltype(i_)::value];
            });
        } else {
            static_for<0, PER>([&](auto i_) {
                const int e = decltype(i_)::value * THREADS + tid;
                const long long g = (long long) (e / M) * stride + e % M;
                const cf x = tile[(e / M) * PITCH + e % M];
                __stcs(ore + g, x.x);
                __stcs(oim + g, x.y);
            });
        }
    } else {
        for (int e = tid; e < rows * M; e += THREADS) {
            const int r = e / M, c = e % M;
            const long long g = (long long) r * stride + c;
            if constexpr (LOAD) {
                tile[r * PITCH + c] = make_float2(__ldcs(gre + g), __ldcs(gim + g));
            } else {
                const cf x = tile[r * PITCH + c];
                __stcs(ore + g, x.x);
                __stcs(oim + g, x.y);
            }
        }
    }
}

template <int M, int MODE, bool INV, bool PLANAR>
__global__ void __launch_bounds__(SmallCfg<M, MODE>::THREADS, SmallCfg<M, MODE>::MINB) small_kernel(const KernelParams p)
{
    using SC = SmallCfg<M, MODE>;
    constexpr int PITCH = SC::PITCH;
    static_assert(!PLANAR || MODE == MODE_C2C, "planar rows: complex transforms");
    extern __shared__ __align__(16) unsigned char small_smem[];
    cf* tile = reinterpret_cast<cf*>(small_smem);
    const int tid = threadIdx.x;

    for (long long base = (long long) blockIdx.x * SC::ROWS; base < p.batch; base += (long long) gridDim.x * SC::ROWS) {
        const long long left = p.batch - base;
        const int rows = left < SC::ROWS ? (int) left : SC::ROWS;

        if constexpr (PLANAR)
            small_copy_planar<M, PITCH, SC::ROWS, true>(tile, reinterpret_cast<const float*>(p.in) + base * p.in_stride, p.in_im + base * p.in_stride,
                                              nullptr, nullptr, p.in_stride, rows, tid);
        else
            small_copy<SC::IN_E, PITCH, SC::ROWS, true>(tile, p.in + base * p.in_stride, nullptr, p.in_stride, rows, tid);
        __syncthreads();

        if constexpr (SC::TPR == 2) {
            // every thread computes (rows past the end of the batch work on whatever the buffer holds and are not stored):
            // the shuffles below then always find their partner
            const int t = tid & 1;
            cf* row = tile + (tid >> 1) * PITCH;
            cf v[32];
            static_for<0, 32>([&](auto i_) { constexpr int i = decltype(i_)::value; v[bitrev<32>(i)] = row[2 * i + t]; });
            fft_regs<32, 0, INV>(v);                                  // A_t[k] = sum_i x[2i + t] W_32^(ik)
            if (t) {                                                  // odd half: times W_64^k (conjugate for the inverse)
                static_for<1, 32>([&](auto k_) {
                    constexpr int k = decltype(k_)::value;
                    v[k] = cmul_w64<INV ? 64 - k : k>(v[k]);
                });
            }
            // X[k] = A_0[k] + B[k], X[k + 32] = A_0[k] - B[k]: thread t takes the k of its own parity, so each thread
            // hands over the 16 values of the other parity and both write runs of adjacent words
            static_for<0, 16>([&](auto i_) {
                constexpr int i = decltype(i_)::value;
                const cf send = t ? v[2 * i] : v[2 * i + 1];
                cf recv;
                recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                const cf a = t ? recv : v[2 * i];
                const cf b = t ? v[2 * i + 1] : recv;
                row[2 * i + t] = make_float2(a.x + b.x, a.y + b.y);
                row[2 * i + t + 32] = make_float2(a.x - b.x, a.y - b.y);
            });
        } else if (tid < rows) {
            cf* row = tile + tid * PITCH;
            cf v[M];
            if constexpr (MODE == MODE_C2R) {
                // twist (fft_real_default.cpp:65-111): T[k] = (Y[k] + conj Y[M-k]) + i conj(W_2M^k) (Y[k] - conj Y[M-k]),
                // pairwise: with c = i conj(w) * dif,  T[k] = sum + c,  T[M-k] = conj(sum - c)
                static_for<0, M / 2>([&](auto k_) {
                    constexpr int k = decltype(k_)::value;
                    constexpr float wx = cos64(k * (32 / M)), wy = -sin64(k * (32 / M));      // forward W_2M^k
                    const cf y0 = row[k], y1 = row[M - k];
                    const cf sum = make_float2(y0.x + y1.x, y0.y - y1.y);
                    const cf dif = make_float2(y0.x - y1.x, y0.y + y1.y);
                    const cf c = cmul(make_float2(wy, wx), dif);
                    v[bitrev<M>(k)] = make_float2(sum.x + c.x, sum.y + c.y);
                    if constexpr (k != 0) v[bitrev<M>(M - k)] = make_float2(sum.x - c.x, -(sum.y - c.y));
                });
                const cf ym = row[M / 2];
                v[bitrev<M>(M / 2)] = make_float2(2.0f * ym.x, -2.0f * ym.y);
            } else {
                static_for<0, M>([&](auto t_) { constexpr int t = decltype(t_)::value; v[bitrev<M>(t)] = row[t]; });
            }
            fft_regs<M, 0, INV>(v);
            if constexpr (MODE == MODE_R2C) {
                // split (fft_real_default.cpp:23-62): Y[k] = (Z[k] + conj Z[M-k]) - i W_2M^k (Z[k] - conj Z[M-k]); Z[M] == Z[0]
                static_for<0, M / 2>([&](auto k_) {
                    constexpr int k = decltype(k_)::value;
                    constexpr float wx = cos64(k * (32 / M)), wy = -sin64(k * (32 / M));
                    const cf z0 = v[k], z1 = v[(M - k) & (M - 1)];
                    const cf sum = make_float2(z0.x + z1.x, z0.y - z1.y);
                    const cf dif = make_float2(z0.x - z1.x, z0.y + z1.y);
                    const cf c = cmul(make_float2(-wy, wx), dif);
                    row[k] = make_float2(sum.x - c.x, sum.y - c.y);
                    row[M - k] = make_float2(sum.x + c.x, -(sum.y + c.y));
                });
                row[M / 2] = make_float2(2.0f * v[M / 2].x, -2.0f * v[M / 2].y);
            } else {
                static_for<0, M>([&](auto u_) { constexpr int u = decltype(u_)::value; row[u] = v[u]; });
            }
        }
        __syncthreads();

        if constexpr (PLANAR)
            small_copy_planar<M, PITCH, SC::ROWS, false>(tile, nullptr, nullptr, reinterpret_cast<float*>(p.out) + base * p.out_stride,
                                               p.out_im + base * p.out_stride, p.out_stride, rows, tid);
        else
            small_copy<SC::OUT_E, PITCH, SC::ROWS, false>(tile, nullptr, p.out + base * p.out_stride, p.out_stride, rows, tid);
        __syncthreads();                      // the next tile overwrites the buffer
    }
}

}  // namespace ckb
