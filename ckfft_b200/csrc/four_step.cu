// four_step.cu -- host planner and launchers of the multi-pass path (see four_step.cuh).
#include "tile_launch.h"
#include "pipe_kernel.cuh"
#include "plans.h"
#include <stdio.h>

namespace ckb {

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

cudaError_t launch_local_pass(bool inverse, int kind, int L, const TileParams& p, cudaStream_t s)
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
        if ((e = launch_local_pass(inverse, KIND_COLUMN, L[0], p, s)) != cudaSuccess) return e;
        p.in = scratch; p.out = out; p.nproblems = batch; p.ncols = L[0]; p.P = L[0]; p.Q = 1;
        p.stream_in = 0; p.stream_out = 1;
        return launch_local_pass(inverse, KIND_LAST, L[1], p, s);
    }
    p.in = in; p.out = out; p.nproblems = batch; p.ncols = (int) (n / L[0]); p.tw_shift = tw.log2_tmax - log2n; p.P = 1; p.Q = 1;
    p.stream_in = 1; p.stream_out = 0;
    if ((e = launch_local_pass(inverse, KIND_COLUMN, L[0], p, s)) != cudaSuccess) return e;
    p.in = out; p.out = scratch; p.nproblems = batch * L[0]; p.ncols = L[2];
    p.tw_shift = tw.log2_tmax - ilog2(L[1] * L[2]);
    p.stream_in = 0; p.stream_out = 0;
    if ((e = launch_local_pass(inverse, KIND_COLUMN, L[1], p, s)) != cudaSuccess) return e;
    p.in = scratch; p.out = out; p.nproblems = batch; p.ncols = L[0] * L[1]; p.P = L[0]; p.Q = L[1];
    p.stream_in = 0; p.stream_out = 1;
    return launch_local_pass(inverse, KIND_LAST, L[2], p, s);
}

// ---- two-pass lengths as one L2-resident dataflow kernel (pipe_kernel.cuh) --------------------------------
// pipe plans: X(L0, L1, MINB) with the tile plans A = column pass over L0, B = last pass over L1 (256 threads each)
template <int L, bool INV, int KIND> struct PipeTile;
template <bool INV, int KIND> struct PipeTile<128, INV, KIND>  { using type = TileCfg<128, 16, 16, 8, 32, INV, KIND, 3, true>; };
template <bool INV, int KIND> struct PipeTile<256, INV, KIND>  { using type = TileCfg<256, 16, 16, 16, 16, INV, KIND, 3, true>; };
template <bool INV, int KIND> struct PipeTile<512, INV, KIND>  { using type = TileCfg<512, 32, 32, 16, 16, INV, KIND, 2, true>; };
template <bool INV, int KIND> struct PipeTile<1024, INV, KIND> { using type = TileCfg<1024, 32, 32, 32, 8, INV, KIND, 2, true>; };

#define CKB_PIPE_PLANS(X) \
    X(128, 256, 3) \
    X(256, 256, 3) \
    X(256, 512, 2) \
    X(512, 512, 2) \
    X(512, 1024, 2) \
    X(1024, 1024, 2)

static long long env_ll(const char* name, long long dflt)
{
    const char* e = getenv(name);
    return e && *e ? atoll(e) : dflt;
}

bool pipe_enabled()
{
    return env_ll("CKFFT_B200_PIPE", 1) != 0 && tensor_map_encoder() != nullptr;   // read per call: tests flip it
}

template <int L0, int L1, int MINB, bool INV, int NBUF, int MODE = PIPE_C2C, bool SPLIT = false>
static cudaError_t launch_pipe_cfg(const cf* in, cf* out, long long batch, const cf* table, int log2_nt, const BigTwiddles& tw,
                                   cudaStream_t s, long long out_stride = 0, long long in_stride = 0)
{
    using A = typename PipeTile<L0, INV, KIND_COLUMN>::type;
    using B = typename PipeTile<L1, INV, KIND_LAST>::type;
    using PC = PipeCfg<A, B, MINB, NBUF, MODE, SPLIT>;
    constexpr bool REAL = MODE == PIPE_R2C, TWIST = MODE == PIPE_C2R;
    auto kern = pipe_kernel<PC, A, B>;
    constexpr int CTA = PC::THREADS + 64;                    // consumers + loader warp + signaller warp
    static int grid_cap[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (grid_cap[dev] == 0) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PC::SMEM_BYTES)) != cudaSuccess) return e;
        int occ = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA, PC::SMEM_BYTES)) != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        grid_cap[dev] = occ * sm_count_of_current_device();
    }
    constexpr long long N = (long long) L0 * L1;
    constexpr int S = PC::T1 + PC::T2;
    const long long items = batch * S;
    if (items >= (1LL << 32) - (1 << 20) || batch * L0 >= (1LL << 32)) return cudaErrorInvalidValue;
    const int grid = (int) (items < grid_cap[dev] ? items : grid_cap[dev]);

    // Pipeline depth.  Every CTA holds up to 2 + NBUF tickets (computing, requested, drawn), so W = (2 + NBUF) * grid
    // tickets are in flight; a pass-2 item finds its dependencies met without waiting if it trails its pass-1 items by
    // at least W tickets (lag problems of S tickets), and a pass-1 item finds its ring slot drained if the ring is
    // another W tickets longer.  The ring is capped so that it stays resident in L2.
    // Measured (tools/pipe_grid.py): with ping-pong buffers the loader runs a whole item ahead, so its dependencies must be
    // met a whole item earlier: 1.4 x the window is the optimum at 2^15 / 2^16 (0.539 -> 0.554, 0.523 -> 0.550 of the copy
    // peak); with one buffer the bare window is (deeper lags only push the ring out of L2).
    const long long window = ((long long) (2 + NBUF) * grid + S - 1) / S;
    long long lag = env_ll("CKFFT_B200_PIPE_LAG", NBUF == 2 ? window * 7 / 5 : window + 1);
    long long slots = env_ll("CKFFT_B200_PIPE_RING", 2 * lag);
    const long long cap = (env_ll("CKFFT_B200_PIPE_RING_MB", 64) << 20) / (N * 8);
    if (slots > cap) slots = cap;
    if (slots < 2) slots = 2;
    if (lag > slots - 1) lag = slots - 1;
    if (lag > batch) lag = batch;
    if (slots > batch) slots = batch;                        // never more slots than problems (then no slot is reused)

    const size_t ring_bytes = (size_t) slots * N * sizeof(cf);
    const size_t ctr_bytes = ((size_t) (2 * batch + 4) * sizeof(unsigned) + 127) & ~size_t(127);
    unsigned char* ws = nullptr;
    if ((e = scratch_alloc((void**) &ws, ring_bytes + ctr_bytes, s)) != cudaSuccess) return e;
    unsigned* ctr = (unsigned*) (ws + ring_bytes);
    if ((e = cudaMemsetAsync(ctr, 0, ctr_bytes, s)) != cudaSuccess) { cudaFreeAsync(ws, s); return e; }

    CUtensorMap tmap, tmap2;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap2, 0, sizeof(tmap2));
    int xshift[2] = { 0, 0 };
    if (TWIST) {
        // Half spectra: rows of in_stride (= M + 1 when dense) values, so the frames are not all 16-byte aligned.  Frames of
        // parity q form a 3-D tensor [ceil((batch - q) / 2)][L0][L1 + shift] with frame stride 2 * in_stride elements (a multiple
        // of 16 bytes whatever in_stride is) whose base is the 16-byte boundary at or below frame q; a frame that starts 8
        // bytes above its boundary has its columns shifted by one.
        const long long istr = in_stride ? in_stride : N + 1;
        for (int q = 0; q < 2; ++q) {
            const uintptr_t first = (uintptr_t) (in + q * istr);
            xshift[q] = (int) ((first >> 3) & 1);
            const long long frames = batch > q ? (batch - q + 1) / 2 : 1;
            if (!make_tile_map3(q ? &tmap2 : &tmap, (const void*) (first & ~uintptr_t(15)), L1 + xshift[q], L0, frames, (long long) L1 * 8,
                                2 * istr * 8, A::BOX_ROWS, PC::TWIST_W)) {
                cudaFreeAsync(ws, s);
                return cudaErrorNotSupported;              // the caller falls back to the separate twist pass
            }
        }
    } else if (!make_tile_map(&tmap, in, batch * L0, L1, PC::BOXR, A::C)) { cudaFreeAsync(ws, s); return cudaErrorInvalidValue; }

    PipeParams p{};
    p.in = in; p.out = out; p.ring = (cf*) ws;
    p.table = table; p.log2_nt = log2_nt;
    p.tw_lo = tw.lo; p.tw_hi = tw.hi; p.tw_h = tw.h; p.tw_shift = tw.log2_tmax - ilog2(L0) - ilog2(L1);
    p.batch = batch; p.ring_slots = (int) slots; p.lag = (int) lag;
    p.ticket = ctr; p.done1 = ctr + 4; p.done2 = ctr + 4 + batch;
    p.out_stride = out_stride ? out_stride : N + 1;
    p.in_stride = in_stride ? in_stride : N + 1;
    // Measured (B200, tools/gpu_check.py): one bulk copy per pass-2 tile instead of C: + 0.5 - 1 point everywhere; discarding the
    // dead ring lines: 2^18 0.550 -> 0.560, 2^20 0.469 -> 0.492 (less write-back traffic at the power cap), but - 0.5 point
    // at 2^15 / 2^16, where the ring is small enough to be overwritten in L2 before it is ever evicted.
    p.flags = (int) env_ll("CKFFT_B200_PIPE_FLAGS", N >= (1 << 17) ? 3 : 1);
    p.twist_xshift[0] = xshift[0]; p.twist_xshift[1] = xshift[1];
    p.tw_shift_real = tw.log2_tmax - ilog2(L0) - ilog2(L1) - 1;
    if ((REAL || TWIST) && p.tw_shift_real < 0) { cudaFreeAsync(ws, s); return cudaErrorInvalidValue; }
#if CKB_PIPE_STATS
    unsigned long long* dstats = nullptr;
    cudaMalloc((void**) &dstats, 16 * sizeof(unsigned long long));
    cudaMemsetAsync(dstats, 0, 16 * sizeof(unsigned long long), s);
    p.stats = dstats;
#endif
    kern<<<grid, CTA, PC::SMEM_BYTES, s>>>(p, tmap, tmap2);
    count_launch();
    e = cudaGetLastError();
#if CKB_PIPE_STATS
    {
        unsigned long long h[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, dstats, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(dstats);
        auto avg = [&](int sum, int cnt) { return h[cnt] ? (double) h[sum] / (double) h[cnt] : 0.0; };
        fprintf(stderr, "pipe_stats L0=%d L1=%d mode=%d nbuf=%d grid=%d lag=%lld slots=%lld batch=%lld | per item (cycles): consumer wait %.0f; "
                        "pass1: tile latency %.0f (%llu waits of %llu), compute %.0f; pass2: tile latency %.0f (%llu waits of %llu), compute %.0f | "
                        "loader per item: ticket %.0f deps %.0f free-wait %.0f sig-wait %.0f\n",
                L0, L1, MODE, NBUF, grid, lag, slots, batch, avg(0, 1), avg(2, 3), h[3], h[7], avg(6, 7), avg(4, 5), h[5], h[9], avg(8, 9),
                avg(10, 14), avg(11, 14), avg(12, 14), avg(13, 14));
    }
#endif
    cudaError_t e2 = cudaFreeAsync(ws, s);
    return e != cudaSuccess ? e : e2;
}

// `batch` dense transforms of 2^log2n points (2^15 .. 2^20), 16-byte aligned input
cudaError_t launch_pipe(bool inverse, int log2n, const cf* in, cf* out, long long batch, const cf* table, int log2_nt,
                        const BigTwiddles& tw, cudaStream_t s)
{
    int npass, L[3];
    four_step_plan(log2n, &npass, L);
    if (npass != 2 || ((uintptr_t) in & 15)) return cudaErrorNotSupported;
    const bool two = env_ll("CKFFT_B200_PIPE_NBUF", 2) == 2 && L[1] <= 256;     // ping-pong tile buffers where three CTAs still fit (CKFFT_B200_PIPE_NBUF=1: A/B)
    // one buffer + half-tile staging area for the 512 / 1024-point tile plans (CKFFT_B200_PIPE_SPLIT=0: A/B).  Measured, one
    // buffer -> split: 2^18 .572 -> .591, 2^19 .551 -> .569, 2^20 .515 -> .534; 2^17 (256 x 512) .518 -> .503 and keeps one buffer
    const bool split = env_ll("CKFFT_B200_PIPE_SPLIT", 1) != 0 && L[0] >= 512;
#define X(L0_, L1_, MINB_) \
    if (L[0] == L0_ && L[1] == L1_) { \
        if (two && L1_ <= 256) \
            return inverse ? launch_pipe_cfg<L0_, L1_, MINB_, true, (L1_ <= 256 ? 2 : 1)>(in, out, batch, table, log2_nt, tw, s) \
                           : launch_pipe_cfg<L0_, L1_, MINB_, false, (L1_ <= 256 ? 2 : 1)>(in, out, batch, table, log2_nt, tw, s); \
        if (split && L0_ >= 512) \
            return inverse ? launch_pipe_cfg<L0_, L1_, MINB_, true, 1, PIPE_C2C, (L0_ >= 512)>(in, out, batch, table, log2_nt, tw, s) \
                           : launch_pipe_cfg<L0_, L1_, MINB_, false, 1, PIPE_C2C, (L0_ >= 512)>(in, out, batch, table, log2_nt, tw, s); \
        return inverse ? launch_pipe_cfg<L0_, L1_, MINB_, true, 1>(in, out, batch, table, log2_nt, tw, s) \
                       : launch_pipe_cfg<L0_, L1_, MINB_, false, 1>(in, out, batch, table, log2_nt, tw, s); \
    }
    CKB_PIPE_PLANS(X)
#undef X
    return cudaErrorNotSupported;
}

// `batch` real forward transforms of n = 2 * 2^log2m points: in = the real input viewed as 2^log2m complex values per
// frame (dense), out = half spectra of out_stride complex each.  The split is fused into pass 2 of the dataflow kernel.
cudaError_t launch_pipe_r2c(int log2m, const cf* in, cf* out, long long batch, long long out_stride, const cf* table, int log2_nt,
                            const BigTwiddles& tw, cudaStream_t s)
{
    int npass, L[3];
    four_step_plan(log2m, &npass, L);
    if (npass != 2 || ((uintptr_t) in & 15)) return cudaErrorNotSupported;
    const bool two = env_ll("CKFFT_B200_PIPE_NBUF", 2) == 2 && L[1] <= 256;
    const bool split = env_ll("CKFFT_B200_PIPE_SPLIT", 1) != 0 && L[0] >= 512;      // real forward: 2^19 .431 -> .438, 2^20 .362 -> .385
#define X(L0_, L1_, MINB_) \
    if (L[0] == L0_ && L[1] == L1_) { \
        if (two && L1_ <= 256) return launch_pipe_cfg<L0_, L1_, MINB_, false, (L1_ <= 256 ? 2 : 1), PIPE_R2C>(in, out, batch, table, log2_nt, tw, s, out_stride); \
        if (split && L0_ >= 512) return launch_pipe_cfg<L0_, L1_, MINB_, false, 1, PIPE_R2C, (L0_ >= 512)>(in, out, batch, table, log2_nt, tw, s, out_stride); \
        return launch_pipe_cfg<L0_, L1_, MINB_, false, 1, PIPE_R2C>(in, out, batch, table, log2_nt, tw, s, out_stride); \
    }
    CKB_PIPE_PLANS(X)
#undef X
    return cudaErrorNotSupported;
}

// `batch` real inverse transforms of n = 2 * 2^log2m points: in = half spectra (rows of in_stride complex values, 8-byte
// aligned), out = the real output viewed as 2^log2m complex values per frame (dense).  The twist is fused into pass 1.
cudaError_t launch_pipe_c2r(int log2m, const cf* in, cf* out, long long batch, long long in_stride, const cf* table, int log2_nt,
                            const BigTwiddles& tw, cudaStream_t s)
{
    int npass, L[3];
    four_step_plan(log2m, &npass, L);
    if (npass != 2) return cudaErrorNotSupported;
    const bool two = false;      // one (wider) tile buffer: two would not leave room for three CTAs per SM
#define X(L0_, L1_, MINB_) \
    if (L[0] == L0_ && L[1] == L1_) { \
        if (two && L1_ <= 256) return launch_pipe_cfg<L0_, L1_, MINB_, true, (L1_ <= 256 ? 2 : 1), PIPE_C2R>(in, out, batch, table, log2_nt, tw, s, 0, in_stride); \
        return launch_pipe_cfg<L0_, L1_, MINB_, true, 1, PIPE_C2R>(in, out, batch, table, log2_nt, tw, s, 0, in_stride); \
    }
    CKB_PIPE_PLANS(X)
#undef X
    return cudaErrorNotSupported;
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
