// four_step.cuh -- multi-pass ("four-step") transforms for lengths above the single-pass limit.
//
// A transform of N = L1*L2 (two passes, N <= 2^20) or N = L1*L2*L3 (three passes, N <= 2^30) points is
// computed as passes of L-point FFTs (L = 128 .. 1024) over TILES of C adjacent columns:
//
//   column pass (KIND_COLUMN): problem of TN = L*ncols points viewed as [L][ncols]; the CTA stages the
//       [L][C] tile through shared memory with row-chunk accesses (C*8 = 128 contiguous bytes per row,
//       lanes run along the columns), transforms every column with the register/Stockham stages of
//       fft_kernel.cuh, multiplies output (k, c) by the inter-pass twiddle W_TN^(c*k) -- fused into the
//       store -- and writes the tile back in the same geometry;
//   last pass (KIND_LAST): every column is a contiguous run of L points (lanes run along the transform),
//       the result is scattered in natural order: output k of column c goes to c + ncols*k, again through
//       a shared-memory tile so that each store is a 128-byte row chunk.
//
// Each pass is one HBM read + one HBM write of the whole array; intermediate results live in a
// stream-ordered scratch buffer.  The reference has no equivalent (it recurses on one core,
// src/ckfft/fft_default.cpp:168-266); the decomposition is the classic four-step / six-step algorithm
// (cf. the vendored ext/fftw-3.3.2/mpi/dft-rank1.c:21-79 for the same index algebra).
//
// Inter-pass twiddles are W_Tmax^e with e up to 2^30: a full table is out of the question, so the context
// carries a two-level table built in double precision, W^e = hi[e >> h] * lo[e & (2^h - 1)].
#pragma once
#include <cuda.h>          // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "fft_kernel.cuh"

namespace ckb {

enum TileKind { KIND_COLUMN = 0, KIND_LAST = 1 };
#ifndef CKB_MAX_PEERS
#define CKB_MAX_PEERS 8     /* GPUs of one NVSwitch domain a routed pass can store to */
#endif

struct TileParams {
    const cf* in;
    cf* out;
    const cf* table;        // W_Nt^k for the stage LUT
    int log2_nt;
    const cf* tw_lo;        // two-level inter-pass twiddles, forward sign
    const cf* tw_hi;
    int tw_h;               // log2(entries of tw_lo)
    int tw_shift;           // log2(Tmax) - log2(TN): scales an exponent of W_TN to one of W_Tmax
    long long nproblems;    // independent problems of TN = L * ncols points, back to back
    int ncols;              // columns per problem
    int P, Q;               // KIND_LAST: input column c starts at ((c % P) * Q + c / P) * L
    int stream_in;          // 1: the input is read once (ld.global.cs); 0: it was just written by the previous pass and
    int stream_out;         //    should be found in L2 (default policy).  Same for the output.
    int tw_col_shift;       // column pass: the inter-pass twiddle is W_TN^(((tw_col_base + c) >> tw_col_shift) * k)
    int tw_col_base;        //    (a column index that carries batch digits below the twiddle digits / a rank offset above them)
    // ---- routed kernels (TileCfg::RT, the passes of the fused distributed transform, dist_fused.cu) ----
    // Output bin k of column c of problem `prob` is row kk = prob * k_prob_mul + k * k_mul of a row-distributed
    // [rows][out_row_stride] array: it goes to rank kk >> rank_shift, which may be a peer GPU (NVLink store), at
    //     peer[kk >> rank_shift] + (kk & (2^rank_shift - 1)) * out_row_stride + out_col_base + c ,
    // and the column pass multiplies it by W_TN^(((tw_col_base + c) >> tw_col_shift) * kk).
    // Last pass: column c of problem `prob` is the contiguous run starting at in + c*in_col_stride + prob*in_prob_stride.
    cf* peer[CKB_MAX_PEERS];
    int k_prob_mul, k_mul;
    int rank_shift;
    long long out_row_stride;
    long long out_col_base;
    long long in_col_stride, in_prob_stride;
    // pull (first pass of the fused distributed transform): the [L][ncols] problem is spread over the ranks' input
    // arrays, rank s holding rows [s*pull_rows, (s+1)*pull_rows) as a 2-D tensor described by pull_maps[s] (device
    // memory); tile column c0 is that tensor's column (c0 >> pull_w_shift) * pull_n2 + pull_col0 + (c0 & (w - 1)).
    const CUtensorMap* pull_maps;
    int pull_rows, pull_box_rows, pull_w_shift;
    long long pull_n2, pull_col0;
};

__device__ __forceinline__ cf ld_sel(const cf* p, int stream) { return stream ? __ldcs(p) : __ldg(p); }
__device__ __forceinline__ void st_sel(cf* p, cf v, int stream) { if (stream) __stcs(p, v); else *p = v; }

// PFT_ = true: the tile is staged by the TMA unit and prefetched.  Column pass: the [L][C] tile (rows ncols apart
// in global memory) is fetched with cp.async.bulk.tensor.2d boxes described by a CUtensorMap; last pass: each of
// the C contiguous columns with a 1-D cp.async.bulk.  The copy of the NEXT tile is issued as soon as the second
// radix stage has gathered its inputs, i.e. it overlaps that stage's arithmetic, the inter-pass twiddles and the
// stores; it lands in the same shared-memory buffer the exchange uses (no extra shared memory).
template <int L_, int E_, int R0_, int R1_, int C_, bool INV_, int KIND_, int MINB_, bool PFT_ = false, bool RT_ = false>
struct TileCfg {
    static constexpr bool PFT = PFT_;
    static constexpr bool RT = RT_;     // routed: generalised output addressing (stores may target peer GPUs)
    static constexpr int L = L_, E = E_, R0 = R0_, R1 = R1_, C = C_, KIND = KIND_, MINB = MINB_;
    static constexpr bool INV = INV_;
    static constexpr int T = L / E;
    static constexpr int THREADS = C * T;
    static constexpr int LOGPAD = ilog2(R0);
    static_assert(R1 % 4 == 0, "the inter-pass twiddle factorisation works on groups of four bins");
    static constexpr int XRAW = L + (L >> LOGPAD) + 1;
    // Column pitch of the exchange buffer.  With the along-the-columns thread map a half-warp (one 128-byte wavefront of
    // 8-byte accesses) holds 16 adjacent columns of one row (C >= 16), or 8 columns of two adjacent rows (C = 8), and a row
    // step moves the address by 1 (mod 16, in 8-byte words): an odd pitch spreads 16 columns over the 16 bank pairs; for
    // C = 8 the pitch must be 2 (mod 4) so that the two rows interleave (an odd pitch there is a 2-way conflict on almost
    // every access: ncu counted 35 % of the shared-memory wavefronts of the 1024 x 1024 kernel as conflicts).
    static constexpr int XBUF = C_ >= 16 ? (XRAW | 1) : XRAW + ((6 - XRAW % 4) % 4);
    static constexpr int LUT1 = (R1 - 1) * R0;
    static constexpr int SMEM_BYTES = 8 * (LUT1 + C * XBUF) + (PFT_ ? 16 : 0);
    static constexpr int BOX_ROWS = L < 256 ? L : 256;         // TMA boxes are at most 256 elements per dimension
    static_assert((LUT1 * 8) % 128 == 0, "the tile buffer must start on a 128-byte boundary for the tensor copies");
    static_assert(R0 * R1 == L && T <= 32 && THREADS <= 1024, "tile plan");
    static_assert(C_ >= 16 ? XBUF % 2 == 1 : (C_ == 8 && XBUF % 4 == 2), "bank-conflict-free column pitch");
    static_assert((L * C) % THREADS == 0 && (L * C) / THREADS == E, "one tile = E elements per thread");
};

__device__ __forceinline__ cf big_twiddle(const TileParams& p, unsigned c, unsigned k, bool inverse)
{
    const unsigned e = (c * k) << p.tw_shift;              // c*k < TN <= 2^30
    const cf lo = __ldg(p.tw_lo + (e & ((1u << p.tw_h) - 1u)));
    const cf hi = __ldg(p.tw_hi + (e >> p.tw_h));
    cf w = cmul(lo, hi);
    if (inverse) w.y = -w.y;
    return w;
}

// Thread maps.  "along the transform": group = tid / T, j = tid % T  (a group's lanes are adjacent: contiguous columns
// are read with unit stride).  "along the columns": group = tid % C, j = tid / C  (adjacent lanes hold adjacent
// columns: a warp request covers whole 128-byte row chunks of the [L][ncols] array).  The shared-memory exchange
// between the two radix stages is where a kernel may switch from one map to the other, so every pass moves its
// data through shared memory exactly once.
// one 2-D box of a tiled tensor map -> shared memory; completion is signalled on `bar`
__device__ __forceinline__ void tensor_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

template <class TC>
__global__ void __launch_bounds__(TC::THREADS, TC::MINB) tile_kernel(const TileParams p, const __grid_constant__ CUtensorMap tmap_in)
{
    constexpr int L = TC::L, E = TC::E, T = TC::T, C = TC::C, R0 = TC::R0, R1 = TC::R1;
    constexpr int LOGPAD = TC::LOGPAD, XBUF = TC::XBUF, THREADS = TC::THREADS;
    constexpr bool INV = TC::INV;

    extern __shared__ __align__(128) unsigned char tile_smem[];
    cf* lut1 = reinterpret_cast<cf*>(tile_smem);
    cf* xall = lut1 + TC::LUT1;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(xall + C * XBUF);
    const int tid = threadIdx.x;
    // stage 0 map: along the columns for the column pass (strided rows), along the transform for the last pass
    const int g0 = TC::KIND == KIND_COLUMN ? tid % C : tid / T;
    const int j0 = TC::KIND == KIND_COLUMN ? tid / C : tid % T;
    // stage 1 map: always along the columns (the results leave as 128-byte row chunks)
    const int g1 = tid % C;
    const int j1 = tid / C;

    {
        const int sh1 = p.log2_nt - ilog2(L);
        for (int i = tid; i < TC::LUT1; i += THREADS)
            lut1[i] = table_w(p.table, ((i / R0 + 1) * (i % R0)) << sh1, INV);
    }
    unsigned long long l2pol = 0;
    if constexpr (TC::PFT) {
        if (tid == 0) mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        l2pol = l2_evict_first_policy();
    }
    __syncthreads();

    const int blocks_per_problem = p.ncols / C;
    const long long ntiles = p.nproblems * blocks_per_problem;
    const long long tn = (long long) L * p.ncols;

    // start of (contiguous) input column c of a last-pass problem
    auto last_src = [&](long long prob, int c) -> const cf* {
        if constexpr (TC::RT) return p.in + (long long) c * p.in_col_stride + prob * p.in_prob_stride;
        else return p.in + prob * tn + (long long) ((c % p.P) * p.Q + c / p.P) * L;
    };
    // where output bin k of column c of problem `prob` goes
    auto out_ptr = [&](long long prob, int k, int c) -> cf* {
        if constexpr (TC::RT) {
            const long long kk = prob * p.k_prob_mul + (long long) k * p.k_mul;
            return p.peer[kk >> p.rank_shift] + (kk & ((1LL << p.rank_shift) - 1)) * p.out_row_stride + p.out_col_base + c;
        } else {
            return p.out + prob * tn + c + (long long) k * p.ncols;
        }
    };

    // thread 0 asks the TMA unit for a whole tile: dense [L][C] (column pass) or [C][L] (last pass) in `xall`
    auto issue_tile = [&](long long t) {
        if (tid != 0 || t >= ntiles) return;
        const long long prob = t / blocks_per_problem;
        const int c0 = (int) (t % blocks_per_problem) * C;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(mbar, L * C * 8);
        if constexpr (TC::KIND == KIND_COLUMN) {
            if (TC::RT && p.pull_maps) {                       // rows come from the peers' input arrays, over NVLink
                const int x0 = (int) ((c0 >> p.pull_w_shift) * p.pull_n2 + p.pull_col0) + (c0 & ((1 << p.pull_w_shift) - 1));
                for (int r0 = 0; r0 < L; r0 += p.pull_box_rows)
                    tensor_load_2d(xall + r0 * C, p.pull_maps + r0 / p.pull_rows, x0, r0 % p.pull_rows, mbar);
            } else {
#pragma unroll
                for (int r0 = 0; r0 < L; r0 += TC::BOX_ROWS)
                    tensor_load_2d(xall + r0 * C, &tmap_in, c0, (int) (prob * L + r0), mbar);
            }
        } else {
#pragma unroll
            for (int g = 0; g < C; ++g) bulk_load(xall + g * L, last_src(prob, c0 + g), L * 8, mbar, l2pol);
        }
    };
    unsigned phase = 0;
    if constexpr (TC::PFT) issue_tile(blockIdx.x);

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long prob = tile / blocks_per_problem;
        const int c0 = (int) (tile % blocks_per_problem) * C;
        const cf* __restrict__ pin = p.in + prob * tn;
        cf v[E];

        if constexpr (TC::PFT) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
            constexpr int B0 = E / R0, STR0 = L / R0;
            static_for<0, B0>([&](auto q_) {
                constexpr int q = decltype(q_)::value;
                const int jq = j0 + q * T;
                static_for<0, R0>([&](auto t_) {
                    constexpr int t = decltype(t_)::value;
                    const int row = jq + t * STR0;
                    v[q * R0 + bitrev<R0>(t)] = TC::KIND == KIND_COLUMN ? xall[row * C + g0] : xall[g0 * L + row];
                });
            });
            __syncthreads();        // the staged tile is consumed: the buffer now serves the exchange
        } else if constexpr (TC::KIND == KIND_COLUMN) {
            gather_rows<L, T, E, R0>(v, pin + c0 + g0, p.ncols, j0, p.stream_in);
        } else {
            const cf* src = last_src(prob, c0 + g0);
            if (p.stream_in) stage_gather<L, T, E, R0, LOGPAD, SRC_GLOBAL>(v, src, nullptr, j0, true);
            else             stage_gather<L, T, E, R0, LOGPAD, SRC_GLOBAL_KEEP>(v, src, nullptr, j0, true);
        }
        stage_math<T, E, R0, 1, INV, TW_NONE>(v, nullptr, p.table, 0, j0);
        stage_scatter<L, T, E, R0, 1, LOGPAD, DST_XCHG>(v, nullptr, xall + g0 * XBUF, j0, true);
        __syncthreads();
        stage_gather<L, T, E, R1, LOGPAD, SRC_XBUF>(v, nullptr, xall + g1 * XBUF, j1, true);
        if constexpr (TC::PFT) {
            __syncthreads();        // exchange consumed: the buffer is free for the next tile
            issue_tile(tile + gridDim.x);
        }
        // Results: this thread will hold bins k = jq + u * (L / R1) of column c0 + g1, in natural u order.
        constexpr int B1 = E / R1, STR1 = L / R1;
        const int ocol = c0 + g1;
        // column pass: the two inter-pass twiddle look-ups of every butterfly (see below) are fetched BEFORE the butterfly
        // arithmetic, so that the table loads' latency hides behind it
        cf tw_s[B1], tw_b[B1];
        if constexpr (TC::KIND == KIND_COLUMN) {
            const unsigned cc = (unsigned) (p.tw_col_base + ocol) >> p.tw_col_shift;
            const unsigned kmul = TC::RT ? (unsigned) p.k_mul : 1u;
            const unsigned koff = TC::RT ? (unsigned) (prob * p.k_prob_mul) : 0u;
            static_for<0, B1>([&](auto q_) {
                constexpr int q = decltype(q_)::value;
                tw_s[q] = big_twiddle(p, cc, kmul * (unsigned) STR1, INV);
                tw_b[q] = big_twiddle(p, cc, koff + kmul * (unsigned) (j1 + q * T), INV);
            });
        }
        stage_math<T, E, R1, R0, INV, TW_LUT>(v, lut1, p.table, 0, j1);

        static_for<0, B1>([&](auto q_) {
            constexpr int q = decltype(q_)::value;
            const int jq = j1 + q * T;
            if constexpr (TC::KIND == KIND_COLUMN) {
                // inter-pass twiddle W_TN^(cc * k), k = jq + u * STR1: geometric in u; with u = 4a + b it factors as
                // A_a * B_b,  B_b = s^b,  A_a = W^(cc * jq) * (s^4)^a,  s = W^(cc * STR1): two table look-ups and six
                // multiplications per butterfly (the same arithmetic as the dataflow kernel, pipe_kernel.cuh, so that the two
                // paths stay bit-identical).  Routed passes: k -> kk = prob*k_prob_mul + k*k_mul (linear, so the same
                // factorisation holds with s = W^(cc * k_mul*STR1)).
                cf bb[3], aav[R1 / 4];
                {
                    const cf s1 = tw_s[q];
                    bb[0] = s1; bb[1] = cmul(s1, s1); bb[2] = cmul(bb[1], s1);
                    const cf s4 = cmul(bb[1], bb[1]);
                    aav[0] = tw_b[q];
                    static_for<1, R1 / 4>([&](auto a_) { constexpr int a = decltype(a_)::value; aav[a] = cmul(aav[a - 1], s4); });
                }
                static_for<0, R1 / 4>([&](auto a_) {
                    constexpr int a = decltype(a_)::value;
                    const cf aa = aav[a];
                    static_for<0, 4>([&](auto b_) {
                        constexpr int b = decltype(b_)::value;
                        constexpr int u = 4 * a + b;
                        cf val = cmul(v[q * R1 + u], aa);
                        if constexpr (b > 0) val = cmul(val, bb[b - 1]);
                        st_sel(out_ptr(prob, jq + u * STR1, ocol), val, p.stream_out);
                    });
                });
            } else {
                static_for<0, R1>([&](auto u_) {
                    constexpr int u = decltype(u_)::value;
                    st_sel(out_ptr(prob, jq + u * STR1, ocol), v[q * R1 + u], p.stream_out);
                });
            }
        });
        if constexpr (!TC::PFT) __syncthreads();     // the next tile's scatter must not overtake this tile's gather
    }
}

// ---- real <-> half-length complex glue for n above the fused single-pass limit -------------------------
// (the same split / twist as fft_real_default.cpp:13-114, as separate element-wise passes)
struct RealGlueParams {
    const cf* in;
    cf* out;
    const cf* tw_lo;
    const cf* tw_hi;
    int tw_h;
    int tw_shift;          // log2(Tmax) - log2(n)
    long long batch;
    int M;                 // n / 2
    long long in_stride, out_stride;   // complex elements per frame
};

__device__ __forceinline__ cf glue_twiddle(const RealGlueParams& p, unsigned k)
{
    const unsigned e = k << p.tw_shift;
    return cmul(__ldg(p.tw_lo + (e & ((1u << p.tw_h) - 1u))), __ldg(p.tw_hi + (e >> p.tw_h)));
}

// Z[0..M) -> Y[0..M]   (forward split)
static __global__ void __launch_bounds__(256) real_split_kernel(const RealGlueParams p)
{
    const long long per = p.M / 2 + 1;
    const long long total = p.batch * per;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const long long b = i / per;
        const int k = (int) (i % per);
        const cf* z = p.in + b * p.in_stride;
        cf* y = p.out + b * p.out_stride;
        if (k == p.M / 2) {
            const cf m = z[k];
            y[k] = make_float2(2.0f * m.x, -2.0f * m.y);
            continue;
        }
        const cf z0 = z[k], z1 = z[(p.M - k) & (p.M - 1)];
        const cf w = glue_twiddle(p, (unsigned) k);
        const cf sum = make_float2(z0.x + z1.x, z0.y - z1.y);
        const cf dif = make_float2(z0.x - z1.x, z0.y + z1.y);
        const cf c = cmul(make_float2(-w.y, w.x), dif);
        y[k] = make_float2(sum.x - c.x, sum.y - c.y);
        y[p.M - k] = make_float2(sum.x + c.x, -(sum.y + c.y));
    }
}

// Y[0..M] -> T[0..M)   (inverse twist)
static __global__ void __launch_bounds__(256) real_twist_kernel(const RealGlueParams p)
{
    const long long per = p.M / 2 + 1;
    const long long total = p.batch * per;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const long long b = i / per;
        const int k = (int) (i % per);
        const cf* y = p.in + b * p.in_stride;
        cf* t = p.out + b * p.out_stride;
        if (k == p.M / 2) {
            const cf m = y[k];
            t[k] = make_float2(2.0f * m.x, -2.0f * m.y);
            continue;
        }
        const cf y0 = y[k], y1 = y[p.M - k];
        const cf w = glue_twiddle(p, (unsigned) k);
        const cf sum = make_float2(y0.x + y1.x, y0.y - y1.y);
        const cf dif = make_float2(y0.x - y1.x, y0.y + y1.y);
        const cf c = cmul(make_float2(w.y, w.x), dif);
        t[k] = make_float2(sum.x + c.x, sum.y + c.y);
        if (k != 0) t[p.M - k] = make_float2(sum.x - c.x, -(sum.y - c.y));
    }
}

}  // namespace ckb
