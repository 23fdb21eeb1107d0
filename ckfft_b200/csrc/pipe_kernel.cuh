// pipe_kernel.cuh -- two-pass ("four-step") transforms of N = L0*L1 points (2^15 .. 2^20) as ONE persistent
// dataflow kernel whose intermediate array never leaves the 126 MB L2.
//
// The two tile passes of four_step.cuh (column pass over [L0][L1] with the fused inter-pass twiddle, then the
// contiguous last pass) each move the whole array through HBM: 2 reads + 2 writes per point, i.e. at most half the
// copy roofline.  Here both passes are work items of one kernel:
//   * items are handed out by a global ticket counter in a fixed order: the pass-1 tiles of problem i, then the
//     pass-2 tiles of problem i - LAG.  A CTA only ever waits for items with smaller tickets, which are held by
//     running CTAs, so the scheme cannot deadlock whatever subset of the grid is resident;
//   * pass 1 writes its (twiddled) tile into a RING of a few problem slots instead of a full-size scratch array,
//     pass 2 reads the slot back LAG problems later: the ring (<= 48 MiB) stays in L2, it is overwritten in place
//     before its dirty lines are ever evicted, so HBM sees one read of the input and one write of the output;
//   * dependencies are per-problem counters in global memory: done1[p] (pass-1 tiles stored; released with
//     __threadfence + atomicAdd, acquired by the thread that issues the pass-2 bulk copies, followed by a
//     generic->async proxy fence) and done2[p] (pass-2 tiles whose copies have landed in shared memory, which frees
//     the ring slot for problem p + RING);
//   * warp specialisation: 8 consumer warps do the arithmetic; a *loader* thread draws the ticket of the NEXT item,
//     polls its dependencies and requests its tile from the TMA unit as soon as the consumers have drained the
//     buffer (mbarrier `free`); a *signaller* thread raises the completion counter of a finished item once the
//     consumers have issued their stores (mbarrier ring `stored`): no global round trip sits on the consumers' path,
//     and they synchronise among themselves only twice per item (named barrier).
//   * the service threads never invalidate the SM's L1: dependency polls are `ld.relaxed.gpu` (the data they guard is
//     read by the async proxy, i.e. from L2, after a proxy fence) and the completion counter is raised with
//     `red.release.gpu` instead of __threadfence + atomicAdd.  ptxas puts a CCTL.IVALL behind every gpu-scope ACQUIRE
//     (ld.acquire, fence.sc / acq_rel): round 1 did one per poll and per signalled item, i.e. the L1 of every SM was
//     flushed about once a microsecond and the consumers' inter-pass twiddle loads went to L2 every time.
//   * every wait inside the CTA is a blocking mbarrier try_wait (the hardware suspends the thread), every global spin
//     has a watchdog (globaltimer) that traps instead of hanging the GPU.
// The arithmetic of a tile is exactly that of tile_kernel (same stage functions, same twiddle factorisation), so
// results are bit-identical to the two-kernel path.
#pragma once
#include "four_step.cuh"

namespace ckb {

struct PipeParams {
    const cf* in;
    cf* out;
    cf* ring;               // ring_slots problem slots of N complex values
    const cf* table;
    int log2_nt;
    const cf* tw_lo;
    const cf* tw_hi;
    int tw_h;
    int tw_shift;
    long long batch;
    int ring_slots;         // problem slots in the ring (> lag)
    int lag;                // pass-2 items of problem i - lag follow the pass-1 items of problem i (0 <= lag <= batch)
    unsigned* ticket;
    unsigned* done1;        // [batch]
    unsigned* done2;        // [batch]
    // real-forward kernels (PipeCfg::REAL): `in` is the real input viewed as M complex, `out` the half spectrum
    long long out_stride;   // complex elements per output row (M + 1 by default)
    int tw_shift_real;      // scales an exponent of W_(2M) to one of W_Tmax
    // real-inverse kernels (PipeCfg::TWIST): `in` is the half spectrum, rows of in_stride complex (8-byte aligned
    // only), `out` the real output viewed as M complex per frame
    long long in_stride;
    int twist_xshift[2];    // frames of parity q are described by tensor map q, whose columns are shifted by this (0 / 1)
    // development build only (-DCKB_PIPE_STATS=1, tools/pipe_stats.sh): cycle counters, see the bottom of launch_pipe_cfg
    unsigned long long* stats;
    // CKFFT_B200_PIPE_FLAGS (defaults chosen per length by measurement, four_step.cu): 1 = fetch a complex pass-2 tile (C
    // adjacent ring rows = one contiguous block) with ONE bulk copy instead of C; 2 = discard the ring lines from L2 once the
    // tile has landed (discard.global.L2: no write-back of dead data)
    int flags;
};

#ifndef CKB_PIPE_STATS
#define CKB_PIPE_STATS 0
#endif
#if CKB_PIPE_STATS
#define CKB_STAT_ADD(i, v) (stat_acc[i] += (unsigned long long) (v))
#define CKB_STAT_CLOCK() clock64()
#else
#define CKB_STAT_ADD(i, v) ((void) 0)
#define CKB_STAT_CLOCK() 0LL
#endif

// gpu-scope poll that does not touch the L1 (no CCTL.IVALL behind it, unlike ld.acquire.gpu): the value comes from L2
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// counter += 1 with release semantics at gpu scope: everything this thread has observed (the consumers' stores, through
// the `stored` mbarrier) is visible to whoever reads the new value.  A release needs no L1 invalidation.
__device__ __forceinline__ void red_release_gpu_inc(unsigned* p)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// blocking wait with a suspend-time hint: the hardware parks the thread until the phase completes (or ~the hint
// elapses), so a waiting service thread costs no issue slots.  Bounded: a protocol bug traps instead of hanging.
constexpr unsigned long long kPipeWatchdogNs = 10000000000ULL;      // a wait that does not resolve in 10 s is a bug

__device__ __forceinline__ void mbar_wait_parked(unsigned long long* bar, unsigned parity)
{
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
        if (ok) return;
        if ((spins & 63u) == 63u) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kPipeWatchdogNs) __trap();
        }
    }
}


__device__ __forceinline__ unsigned long long l2_evict_normal_policy()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ void tensor_load_2d_hint(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar,
                                                    unsigned long long policy)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ void tensor_load_3d_hint(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar,
                                                    unsigned long long policy)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ cf pipe_twiddle(const PipeParams& p, unsigned c, unsigned k, bool inverse)
{
    const unsigned e = (c * k) << p.tw_shift;              // c*k < N <= 2^20
    cf w = cmul(__ldg(p.tw_lo + (e & ((1u << p.tw_h) - 1u))), __ldg(p.tw_hi + (e >> p.tw_h)));
    if (inverse) w.y = -w.y;
    return w;
}

// A: column-pass plan (L0, C0 columns of the [L0][L1] problem per tile), B: last-pass plan (L1, C1 contiguous columns)
enum PipeMode { PIPE_C2C = 0, PIPE_R2C = 1, PIPE_C2R = 2 };

// SPLIT_ (one tile buffer, NBUF = 1): the first half of the NEXT tile -- rows [0, L0/2) of a pass-1 tile, columns [0, C/2) of a
// pass-2 tile -- is fetched into a half-tile staging area S of its own as soon as the consumers have read S, i.e. right
// after their first barrier; the second half lands in the exchange buffer once they have drained it, as before.  The
// 512- and 1024-point tile plans have no room for a second whole buffer (ping-pong), and with one buffer the tile copy was
// only in flight during the last radix stage and the stores: 2 000 of 7 800 cycles per item spent waiting for it.
template <class A, class B, int MINB_, int NBUF_ = 1, int MODE_ = PIPE_C2C, bool SPLIT_ = false>
struct PipeCfg {
    static constexpr bool SPLIT = SPLIT_;
    static_assert(!SPLIT_ || (NBUF_ == 1 && MODE_ != PIPE_C2R), "split tile prefetch: one buffer; the real inverse keeps its paired tiles");
    // REAL: real forward transform of 2M points = this M-point complex transform + the split
    //   Y[k] = (Z[k] + conj Z[M-k]) - i W_2M^k (Z[k] - conj Z[M-k])        (src/ckfft/fft_real_default.cpp:13-63)
    // fused into pass 2: a tile holds C/2 columns c and their mirrors L0 - c (columns 0 and L0/2 mirror themselves),
    // so Z[k] and Z[M-k] meet in the tile and the half spectrum is the only thing written to HBM.
    static constexpr bool REAL = MODE_ == PIPE_R2C;
    // TWIST: real inverse transform of 2M points = the twist
    //   T[k] = (Y[k] + conj Y[M-k]) + i conj(W_2M^k) (Y[k] - conj Y[M-k])     (src/ckfft/fft_real_default.cpp:65-111)
    // followed by this M-point inverse complex transform, fused into pass 1.  Y[k] pairs with Y[M-k]: element (row r,
    // column c) of the [L0][L1] view with (row L0-1-r, column L1-c).  A pass-1 tile therefore holds C/2 columns
    // [ca, ca + C/2) and their C/2 mirror columns (slot g pairs with slot C-1-g), so that both members of every pair
    // meet in the tile and every element is fetched from HBM exactly once.  The rows of the half spectrum hold M+1 values,
    // i.e. every other frame starts 8 bytes off a 16-byte boundary; the tile is still fetched by TMA: frames of each
    // parity get their own 3-D tensor map [frames/2][L0][L1 (+1)] whose base is the 16-byte boundary at or below the first
    // such frame.  A TMA box must START on a 16-byte boundary as well (measured: an odd element coordinate raises "illegal
    // instruction", tools/probes/tma3d_probe.cu) and the mirror columns of an aligned box begin 8 bytes off one, so every box
    // is C/2 + 2 columns wide and the wanted columns sit at offset 0 or 1 inside it.  Columns 0
    // and L1/2 pair with themselves: column 0 one row down (its r = 0 partner is Y[M], read with a plain load), column
    // L1/2 takes the slot of the (non-existent) column L1 in the first tile and is read with plain loads.
    // (First version: the consumers read Y[k] and Y[M-k] with plain 8-byte loads -- 19 000 cycles per pass-1 item instead
    // of 7 000, no faster than the separate twist pass.)
    static constexpr bool TWIST = MODE_ == PIPE_C2R;
    static_assert(MODE_ != PIPE_C2R || A::INV, "the real inverse runs the inverse complex transform");
    static_assert(MODE_ != PIPE_R2C || !A::INV, "the real forward runs the forward complex transform");
    static constexpr int NBUF = NBUF_;            // tile buffers (each is staging area, then exchange buffer, of one item)
    static_assert(NBUF_ == 1 || NBUF_ == 2, "one buffer, or two in ping-pong");
    static_assert(A::THREADS == B::THREADS, "both passes run in the same CTA");
    static_assert(A::INV == B::INV, "one direction");
    static constexpr int THREADS = A::THREADS;
    static constexpr int MINB = MINB_;
    static constexpr int L0 = A::L, L1 = B::L;
    static constexpr int T1 = L1 / A::C;          // pass-1 tiles per problem
    static constexpr int T2 = L0 / B::C;          // pass-2 tiles per problem
    static constexpr int XA = A::C * A::XBUF, XB = B::C * B::XBUF;
    // real inverse: the staged pass-1 tile is two halves of [L0][C/2 + 2] values (TMA boxes must start on 16-byte boundaries;
    // the mirror columns of an aligned box start 8 bytes off one, so every box is fetched two columns wider)
    static constexpr int TWIST_W = A::C / 2 + 2;
    static constexpr int XT = MODE_ == PIPE_C2R ? 2 * A::L * TWIST_W : 0;
    static constexpr int XALL = (((XA > XB ? XA : XB) > XT ? (XA > XB ? XA : XB) : XT) + 15) & ~15;   // whole 128-byte lines
    static constexpr bool SHARE_LUT = SPLIT_ && A::L == B::L && A::R0 == B::R0 && A::R1 == B::R1;     // one copy of the stage LUT when both passes have the same plan
    static constexpr int LUTA = A::LUT1, LUTB = SHARE_LUT ? 0 : B::LUT1;
    static constexpr int TILE_A = A::L * A::C, TILE_B = B::L * B::C;       // dense (staged) tiles
    static constexpr int SHALF = SPLIT_ ? (TILE_A > TILE_B ? TILE_A : TILE_B) / 2 : 0;
    static constexpr int BOXR = SPLIT_ && A::BOX_ROWS > A::L / 2 ? A::L / 2 : A::BOX_ROWS;     // rows per TMA box of a pass-1 tile
    static constexpr int SMEM_BYTES = 8 * (LUTA + LUTB + NBUF * XALL + SHALF) + 384;
    static_assert((XALL * 8) % 128 == 0 && (SHALF * 8) % 128 == 0, "tile buffer alignment");
    static_assert(((LUTA + LUTB) * 8) % 128 == 0, "tile buffer alignment");
    static_assert(!SPLIT_ || (A::R0 % 2 == 0 && B::C % 2 == 0 && (A::E / A::R0) * A::T == A::L / A::R0), "split tile prefetch: halves by butterfly input / by column");
};

template <class PC, class A, class B>
__global__ void __launch_bounds__(PC::THREADS + 64, PC::MINB) pipe_kernel(const PipeParams p, const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_in2)
{
    constexpr int THREADS = PC::THREADS;          // consumer threads; two more warps serve them (loader, signaller)
    constexpr bool INV = A::INV;
    constexpr int L0 = PC::L0, L1 = PC::L1;
    constexpr int T1 = PC::T1, T2 = PC::T2, S = T1 + T2;
    constexpr long long N = (long long) L0 * L1;

    extern __shared__ __align__(128) unsigned char pipe_smem[];
    cf* lutA = reinterpret_cast<cf*>(pipe_smem);
    cf* lutB = PC::SHARE_LUT ? lutA : lutA + PC::LUTA;
    cf* xall = lutA + PC::LUTA + PC::LUTB;
    constexpr int NBUF = PC::NBUF;
    cf* sbuf = xall + NBUF * PC::XALL;            // SPLIT: staging area of the first half tile
    (void) sbuf;
    // control block (128 bytes behind the tile buffers).  SPLIT: index 0 = the exchange buffer (second half tile), 1 = S (first half)
    unsigned long long* bar_full = reinterpret_cast<unsigned long long*>(xall + NBUF * PC::XALL + PC::SHALF);   // [2] tile landed (TMA complete_tx)
    unsigned long long* bar_free = bar_full + 2;       // [2] every consumer has drained the buffer (stage-1 gather done)
    unsigned long long* bar_stored = bar_full + 4;     // [4] ring: every consumer has issued its global stores of item k (k & 3)
    unsigned long long* bar_sig = bar_full + 8;        // [4] ring: the signaller has handled item k (k & 3)
    // [8] item descriptors, item k at k & 7: {pass (0 = sentinel), problem, first column, ring slot}.  The loader decodes the
    // ticket once; the 256 consumers and the signaller read one 16-byte word instead of each redoing the divisions.
    uint4* item_desc = reinterpret_cast<uint4*>(bar_full + 12);
    long long* req_clock = reinterpret_cast<long long*>(bar_full + 28);                     // [8] (statistics builds): when item k was requested
    (void) req_clock;
    const int tid = threadIdx.x;
#if CKB_PIPE_STATS
    unsigned long long stat_acc[16] = {0};        // per thread (only consumer 0 and the loader use them), written out once at exit
    auto stat_flush = [&] { for (int i = 0; i < 16; ++i) if (stat_acc[i]) atomicAdd(p.stats + i, stat_acc[i]); };
#endif

    {
        const int shA = p.log2_nt - ilog2(L0), shB = p.log2_nt - ilog2(L1);
        for (int i = tid; i < PC::LUTA; i += THREADS + 64) lutA[i] = table_w(p.table, ((i / A::R0 + 1) * (i % A::R0)) << shA, INV);
        for (int i = tid; i < PC::LUTB; i += THREADS + 64) lutB[i] = table_w(p.table, ((i / B::R0 + 1) * (i % B::R0)) << shB, INV);
    }
    if (tid == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_full + 1, 1);
        mbar_init(bar_free, THREADS);
        mbar_init(bar_free + 1, THREADS);
        for (int i = 0; i < 4; ++i) { mbar_init(bar_stored + i, THREADS); mbar_init(bar_sig + i, 1); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const long long lag = p.lag;
    const unsigned long long head = (unsigned long long) lag * T1;                       // pass-1 only
    const unsigned long long steady = (unsigned long long) (p.batch - lag) * S;          // alternating
    const unsigned long long total = (unsigned long long) p.batch * S;

    // ticket -> (pass, problem, first column)
    auto decode = [&](unsigned long long t, int& pass, long long& prob, int& c0) {
        if (t < head) { pass = 1; prob = (long long) (t / T1); c0 = (int) (t % T1) * A::C; return; }
        t -= head;
        if (t < steady) {
            const long long step = (long long) (t / S);
            const int r = (int) (t % S);
            if (r < T1) { pass = 1; prob = lag + step; c0 = r * A::C; }
            else        { pass = 2; prob = step; c0 = (r - T1) * B::C; }
            return;
        }
        t -= steady;
        pass = 2; prob = (p.batch - lag) + (long long) (t / T2); c0 = (int) (t % T2) * B::C;
    };

    // pass-2 column held by slot g of the tile that starts at c0.  Complex kernels: c0 + g.  Real kernels: slots
    // [0, C/2) hold columns c0/2 + g, slots [C/2, C) their mirrors L0 - (c0/2 + g); the first tile pairs the two
    // self-mirrored columns, 0 and L0/2.
    auto pass2_column = [&](int c0, int g) -> int {
        if constexpr (!PC::REAL) return c0 + g;
        constexpr int H = B::C / 2;
        const int c = c0 / 2 + (g & (H - 1));
        if (g < H) return c;
        return c == 0 ? L0 / 2 : L0 - c;
    };

    if (tid >= THREADS) {
        // ---------------- two service threads (lane 0 of two warps): loader and signaller ----------------
        // No global round trip (ticket atomic, dependency poll, the release of a completion counter) sits on the
        // consumers' critical path, and the two chains do not serialise each other: the loader draws the next ticket
        // and polls its dependencies while the consumers work, then requests the tile the moment a buffer is free;
        // the signaller raises done1 once the consumers have issued an item's stores.  Only the loader ever waits for
        // other CTAs, and never for anything this CTA still has to signal (that is the signaller's job), so the
        // ticket-order argument for deadlock freedom is unchanged.
        // Item k of this CTA uses tile buffer k % NBUF, `stored` / `sig` barrier k & 3 and ticket slot k & 7.  The loader
        // requests item j only after the signaller has handled item j - 4 (bar_sig), which keeps every barrier of the two
        // rings within one phase of its waiter and every ticket slot alive until the signaller has read it.
        if (tid == THREADS + 32) {
            for (unsigned k = 0;; ++k) {                                   // signaller
                mbar_wait_parked(bar_stored + (k & 3u), (k >> 2) & 1u);
                const uint4 d = item_desc[k & 7u];
                if (d.x == 0u) return;                                     // sentinel (the consumers arrive for it, too)
                if (d.x == 1u) red_release_gpu_inc(p.done1 + d.y);         // the consumers' stores are visible before the count
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_sig + (k & 3u))) : "memory");
            }
        }
        if (tid != THREADS) return;
        const unsigned long long pol_stream = l2_evict_first_policy();
        const unsigned long long pol_keep = l2_evict_normal_policy();
        auto spin_until = [&](const unsigned* ctr, unsigned want) {
            if (ld_relaxed_gpu(ctr) >= want) return;
            const unsigned long long t0 = global_timer_ns();
            while (ld_relaxed_gpu(ctr) < want) {
                __nanosleep(200);
                if (global_timer_ns() - t0 > kPipeWatchdogNs) __trap();
            }
        };
        auto wait_deps = [&](unsigned long long t) {
            if (t >= total) return;
            int pass; long long prob; int c0;
            decode(t, pass, prob, c0);
            if (pass == 1) {
                if (prob >= p.ring_slots) spin_until(p.done2 + (prob - p.ring_slots), (unsigned) T2);
            } else {
                spin_until(p.done1 + prob, (unsigned) T1);
            }
        };
        // part 2: the whole tile into buffer k % NBUF.  SPLIT: part 0 = item descriptor + first half tile into S, part 1 = second
        // half into the exchange buffer.
        auto request = [&](unsigned long long t, unsigned k, int part) {        // item number k of this CTA carries ticket t
            const unsigned b = PC::SPLIT ? (part == 0 ? 1u : 0u) : k % NBUF;
            cf* stage = PC::SPLIT && part == 0 ? sbuf : xall + (PC::SPLIT ? 0u : b) * PC::XALL;
            unsigned long long* full = bar_full + b;
            if (t >= total) {                                          // sentinel: complete the phase without a copy
                if (part != 1) item_desc[k & 7u] = make_uint4(0u, 0u, 0u, 0u);
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full)) : "memory");
                return;
            }
            int pass; long long prob; int c0;
            decode(t, pass, prob, c0);
            if (part != 1) item_desc[k & 7u] = make_uint4((unsigned) pass, (unsigned) prob, (unsigned) c0, (unsigned) (prob % p.ring_slots));
            if constexpr (PC::SPLIT) {
                if (pass == 1) {
                    // rows [0, L0/2) or [L0/2, L0) of the [L0][C] tile
                    constexpr int HR = L0 / 2;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(full, HR * A::C * 8);
#pragma unroll
                    for (int r0 = 0; r0 < HR; r0 += PC::BOXR)
                        tensor_load_2d_hint(stage + r0 * A::C, &tmap_in, c0, (int) (prob * L0 + part * HR + r0), full, pol_stream);
                } else {
                    // columns [0, C/2) or [C/2, C) of the tile, L1 contiguous values each
                    constexpr int HC = B::C / 2;
                    asm volatile("fence.proxy.async;" ::: "memory");       // other CTAs' generic stores -> our async-proxy reads
                    mbar_expect_tx(full, L1 * HC * 8);
                    const cf* slot = p.ring + (prob % p.ring_slots) * N;
                    if (!PC::REAL && (p.flags & 1)) {
                        bulk_load(stage, slot + (long long) (c0 + part * HC) * L1, L1 * HC * 8, full, pol_keep);
                    } else {
#pragma unroll
                        for (int g = 0; g < HC; ++g) bulk_load(stage + g * L1, slot + (long long) pass2_column(c0, part * HC + g) * L1, L1 * 8, full, pol_keep);
                    }
                }
                return;
            }
            if (pass == 1) {
                if constexpr (PC::TWIST) {
                    // C/2 columns from ca and their C/2 mirror columns, as two [L0][C/2] halves of the buffer
                    // Both boxes start at an even element coordinate.  Aligned frame (xs = 0): own columns from x = ca (offset 0
                    // in the box), mirror columns L1-ca-H+1 .. from x = L1-ca-H (offset 1).  Frame 8 bytes above its boundary
                    // (xs = 1, x = column + 1): own from x = ca (offset 1), mirror from x = L1-ca-H+2 (offset 0).
                    constexpr int H = A::C / 2, W = PC::TWIST_W;
                    const int q = (int) (prob & 1);
                    const CUtensorMap* map = q ? &tmap_in2 : &tmap_in;
                    const int xs = p.twist_xshift[q], ca = c0 / 2, z = (int) (prob >> 1);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(full, 2 * L0 * W * 8);
#pragma unroll
                    for (int r0 = 0; r0 < L0; r0 += A::BOX_ROWS) {
                        tensor_load_3d_hint(stage + r0 * W, map, ca, r0, z, full, pol_stream);
                        tensor_load_3d_hint(stage + (L0 + r0) * W, map, L1 - ca - H + 2 * xs, r0, z, full, pol_stream);
                    }
                } else {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(full, L0 * A::C * 8);
#pragma unroll
                    for (int r0 = 0; r0 < L0; r0 += A::BOX_ROWS)
                        tensor_load_2d_hint(stage + r0 * A::C, &tmap_in, c0, (int) (prob * L0 + r0), full, pol_stream);
                }
            } else {
                asm volatile("fence.proxy.async;" ::: "memory");       // other CTAs' generic stores -> our async-proxy reads
                mbar_expect_tx(full, L1 * B::C * 8);
                const cf* slot = p.ring + (prob % p.ring_slots) * N;
                if (!PC::REAL && (p.flags & 1)) {
                    bulk_load(stage, slot + (long long) c0 * L1, L1 * B::C * 8, full, pol_keep);      // C adjacent rows of the slot are contiguous
                } else {
#pragma unroll
                    for (int g = 0; g < B::C; ++g) bulk_load(stage + g * L1, slot + (long long) pass2_column(c0, g) * L1, L1 * 8, full, pol_keep);
                }
            }
        };
        for (unsigned j = 0;; ++j) {
            const long long s0 = CKB_STAT_CLOCK();
            const unsigned long long t = atomicAdd(p.ticket, 1u);
            const long long s1 = CKB_STAT_CLOCK();
            wait_deps(t);
            const long long s2 = CKB_STAT_CLOCK();
            if constexpr (PC::SPLIT) {
                if (j >= 1u) mbar_wait_parked(bar_free + 1, (j - 1u) & 1u);                           // item j - 1 has read S
            } else {
                if (j >= (unsigned) NBUF) mbar_wait_parked(bar_free + j % NBUF, (j / NBUF - 1) & 1u);     // item j - NBUF has drained the buffer
            }
            const long long s3 = CKB_STAT_CLOCK();
            if (j >= 4u) mbar_wait_parked(bar_sig + (j & 3u), ((j >> 2) - 1) & 1u);                  // item j - 4 has been signalled
            const long long s4 = CKB_STAT_CLOCK();
#if CKB_PIPE_STATS
            req_clock[j & 7u] = s4;
            CKB_STAT_ADD(10, s1 - s0); CKB_STAT_ADD(11, s2 - s1); CKB_STAT_ADD(12, s3 - s2); CKB_STAT_ADD(13, s4 - s3); CKB_STAT_ADD(14, 1);
#endif
            if constexpr (PC::SPLIT) {
                request(t, j, 0);
                if (j >= 1u) mbar_wait_parked(bar_free, (j - 1u) & 1u);                               // item j - 1 has drained the exchange buffer
                request(t, j, 1);
            } else {
                request(t, j, 2);
            }
            if (t >= total) {
#if CKB_PIPE_STATS
                stat_flush();
#endif
                return;
            }
        }
    }

    // ---------------- consumers: 256 threads, one named barrier between the radix stages ----------------
    auto consumer_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); };
    for (unsigned k = 0;; ++k) {
        const unsigned b = k % NBUF;
        cf* xb = xall + b * PC::XALL;                           // this item's buffer: staged tile first, exchange buffer after
        unsigned long long* bfree = bar_free + b;
        const long long c0clk = CKB_STAT_CLOCK();
        if constexpr (PC::SPLIT) mbar_wait_parked(bar_full + 1, k & 1u);       // first half tile (S)
        mbar_wait_parked(bar_full + b, (k / NBUF) & 1u);
        const long long c1clk = CKB_STAT_CLOCK();
        const uint4 desc = item_desc[k & 7u];
        if (desc.x == 0u) {                                     // sentinel: tell the signaller, then leave
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_stored + (k & 3u))) : "memory");
#if CKB_PIPE_STATS
            if (tid == 0) stat_flush();
#endif
            break;
        }
        const int pass = (int) desc.x, c0 = (int) desc.z;
        const long long prob = (long long) desc.y;
#if CKB_PIPE_STATS
        if (tid == 0) {
            // [0] cycles the consumers waited for tiles, [1] items; per pass p: [2p] request -> landed (only when the consumer
            // was already waiting, i.e. the true arrival time), [2p+1] count.  Loader: [10] ticket atomic, [11] dependency
            // polls, [12] wait for a free buffer, [13] wait for the signaller, [14] items.
            CKB_STAT_ADD(0, c1clk - c0clk); CKB_STAT_ADD(1, 1);
            if (c1clk - c0clk > 300) { CKB_STAT_ADD(2 * pass, c1clk - req_clock[k & 7u]); CKB_STAT_ADD(2 * pass + 1, 1); }
        }
#endif
        if (pass == 1) {
            constexpr int L = A::L, E = A::E, T = A::T, C = A::C, R0 = A::R0, R1 = A::R1, LOGPAD = A::LOGPAD, XBUF = A::XBUF;
            const int g = tid % C, j = tid / C;               // along the columns, both stages
            cf v[E];
            constexpr int B0 = E / R0, STR0 = L / R0;
            int jj = j;                                       // butterfly set this thread works on in stage 0
            int ocolumn = c0 + g;                             // column of the [L0][L1] problem held by slot g
            if constexpr (PC::TWIST) {
                // element (row r, column c) of the [L0][L1] view is T[i], i = r * L1 + c.  Pairwise definition of the twist pass
                // (four_step.cuh real_twist_kernel; same operations, the factor W_2M^k one rounding apart -- see below): with
                // k = min(i, M - i), y0 = Y[k], y1 = Y[M-k], c = f_k * dif,
                //   T[k] = sum + c,   T[M-k] = conj(sum - c),   T[M/2] = 2 conj(Y[M/2]).
                // Slots [0, H) hold columns ca + g, slots H + m the mirror columns L1 - ca - (H-1) + m; slot g pairs with slot
                // C-1-g, row r with row L0-1-r.  The threads of the mirror slots work on the mirrored butterfly set (T-1-j holds
                // exactly the rows L0-1-r of j's rows), so the two halves of a warp request always read rows of opposite parity:
                // conflict-free 64-byte rows.
                constexpr int M = (int) N, H = C / 2, W = PC::TWIST_W;
                const cf* __restrict__ yrow = p.in + prob * p.in_stride;
                const int ca = c0 / 2, half = g / H, pos = g % H;
                const int xs = p.twist_xshift[prob & 1];             // where the wanted columns start inside the two boxes
                const int o_own = half == 0 ? xs : 1 - xs, o_mir = half == 0 ? 1 - xs : xs;
                const bool self_mid = c0 == 0 && g == C - 1;       // column L1/2 (pairs with itself): straight from global memory
                const bool self_zero = c0 == 0 && g == 0;          // column 0 pairs with itself, one row down; its row 0 with Y[M]
                ocolumn = self_mid ? L1 / 2 : (half == 0 ? ca + pos : L1 - ca - (H - 1) + pos);
                jj = half == 0 ? j : T - 1 - j;
                const cf* own_base = xb + half * (L * W) + pos + o_own;                   // [row][W]
                const cf* mir_base = xb + (1 - half) * (L * W) + (H - 1 - pos) + o_mir;
                static_assert(32 % R0 == 0, "the row step of a stage-0 butterfly is a multiple of 1/64 turn of W_2M");
                static_for<0, B0>([&](auto q_) {
                    constexpr int q = decltype(q_)::value;
                    // Twist factors.  The thread's R0 elements of this butterfly are L1 * STR0 = M / R0 apart, so
                    //   W_2M^(i0 + t * M/R0) = W_2M^i0 * W_(2 R0)^t :
                    // ONE two-level table look-up per butterfly (i0 = the t = 0 element) times compile-time constants instead
                    // of a look-up per element (that was 10 % + 13 % of the kernel's stall samples, all waiting on table loads).
                    // Elements of the upper half use k = M - i:  W_2M^(M-i) = -conj(W_2M^i), exact.
                    const int i0 = (jj + q * T) * L1 + ocolumn;
                    const unsigned e0 = (unsigned) i0 << p.tw_shift_real;
                    const cf wbase = cmul(__ldg(p.tw_lo + (e0 & ((1u << p.tw_h) - 1u))), __ldg(p.tw_hi + (e0 >> p.tw_h)));
                    static_for<0, R0>([&](auto t_) {
                        constexpr int t = decltype(t_)::value;
                        const int r = jj + q * T + t * STR0;
                        cf own, mir;
                        if (self_mid) {
                            own = __ldg(yrow + r * L1 + L1 / 2);
                            mir = __ldg(yrow + (L - 1 - r) * L1 + L1 / 2);
                        } else if (self_zero) {
                            own = own_base[r * W];
                            mir = r == 0 ? __ldg(yrow + M) : own_base[(L - r) * W];
                        } else {
                            own = own_base[r * W];
                            mir = mir_base[(L - 1 - r) * W];
                        }
                        const int i = r * L1 + ocolumn;
                        const bool upper = 2 * i > M;
                        const cf y0 = upper ? mir : own, y1 = upper ? own : mir;
                        const cf wi = cmul_w64<t * (32 / R0)>(wbase);                       // W_2M^i
                        const cf w = upper ? make_float2(-wi.x, wi.y) : wi;                  // W_2M^k, k = min(i, M - i)
                        const cf sum = make_float2(y0.x + y1.x, y0.y - y1.y);
                        const cf dif = make_float2(y0.x - y1.x, y0.y + y1.y);
                        const cf cc = cmul(make_float2(w.y, w.x), dif);
                        cf res = upper ? make_float2(sum.x - cc.x, -(sum.y - cc.y)) : make_float2(sum.x + cc.x, sum.y + cc.y);
                        if (2 * i == M) res = make_float2(2.0f * own.x, -2.0f * own.y);
                        v[q * R0 + bitrev<R0>(t)] = res;
                    });
                });
                consumer_sync();                              // every pair has been read: the buffer now serves the exchange
            } else {
                static_for<0, B0>([&](auto q_) {
                    constexpr int q = decltype(q_)::value;
                    static_for<0, R0>([&](auto t_) {
                        constexpr int t = decltype(t_)::value;
                        if constexpr (PC::SPLIT)      // rows below L/2 are in S, the others in the exchange buffer (t * STR0 = t * L / R0)
                            v[q * R0 + bitrev<R0>(t)] = t < R0 / 2 ? sbuf[(j + q * T + t * STR0) * C + g] : xb[(j + q * T + (t - R0 / 2) * STR0) * C + g];
                        else
                        v[q * R0 + bitrev<R0>(t)] = xb[(j + q * T + t * STR0) * C + g];
                    });
                });
                consumer_sync();                              // the staged tile is consumed: the buffer now serves the exchange
                if constexpr (PC::SPLIT) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_free + 1)) : "memory");
            }
            stage_math<T, E, R0, 1, INV, TW_NONE>(v, nullptr, p.table, 0, jj);
            stage_scatter<L, T, E, R0, 1, LOGPAD, DST_XCHG>(v, nullptr, xb + g * XBUF, jj, true);
            consumer_sync();
            stage_gather<L, T, E, R1, LOGPAD, SRC_XBUF>(v, nullptr, xb + g * XBUF, j, true);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bfree)) : "memory");
            constexpr int B1 = E / R1, STR1 = L / R1;
            const unsigned cc = (unsigned) ocolumn;
            cf tw_s[B1], tw_b[B1];                            // the two inter-pass twiddle look-ups of each butterfly (see below)
            static_for<0, B1>([&](auto q_) {
                constexpr int q = decltype(q_)::value;
                tw_s[q] = pipe_twiddle(p, cc, (unsigned) STR1, INV);
                tw_b[q] = pipe_twiddle(p, cc, (unsigned) (j + q * T), INV);
            });
            stage_math<T, E, R1, R0, INV, TW_LUT>(v, lutA, p.table, 0, j);
            cf* ocol = p.ring + (long long) desc.w * N + ocolumn;
            static_for<0, B1>([&](auto q_) {
                constexpr int q = decltype(q_)::value;
                const int jq = j + q * T;
                // inter-pass twiddle W_N^(cc * k), k = jq + u * STR1: geometric in u; with u = 4a + b it factors as A_a * B_b,
                // B_b = s^b, A_a = W^(cc*jq) * (s^4)^a, s = W^(cc*STR1): TWO table look-ups per butterfly (tw_b[q], fetched before
                // the butterfly arithmetic so that their latency hides behind it) and six multiplications.  (Round 1 looked up
                // all R1/4 + 3 factors: 14 dependent L1 loads per thread at the end of the item; 2^16: 0.556 -> 0.618.)
                cf bb[3], aa[R1 / 4];
                bb[0] = tw_s[q]; bb[1] = cmul(tw_s[q], tw_s[q]); bb[2] = cmul(bb[1], tw_s[q]);
                const cf s4 = cmul(bb[1], bb[1]);
                aa[0] = tw_b[q];
                static_for<1, R1 / 4>([&](auto a_) { constexpr int a = decltype(a_)::value; aa[a] = cmul(aa[a - 1], s4); });
                static_for<0, R1 / 4>([&](auto a_) {
                    constexpr int a = decltype(a_)::value;
                    static_for<0, 4>([&](auto b_) {
                        constexpr int b = decltype(b_)::value;
                        constexpr int u = 4 * a + b;
                        cf val = cmul(v[q * R1 + u], aa[a]);
                        if constexpr (b > 0) val = cmul(val, bb[b - 1]);
                        ocol[(long long) (jq + u * STR1) * L1] = val;          // stays in L2 (default policy)
                    });
                });
            });
        } else {
            constexpr int L = B::L, E = B::E, T = B::T, C = B::C, R0 = B::R0, R1 = B::R1, LOGPAD = B::LOGPAD, XBUF = B::XBUF;
            const int g0 = tid / T, j0 = tid % T;             // stage 0 along the transform (contiguous columns)
            const int g1 = tid % C, j1 = tid / C;             // stage 1 along the columns (row-chunk stores)
            if (tid == 0) atomicAdd(p.done2 + prob, 1u);      // the copies have landed: the ring slot is no longer needed
            if (!PC::REAL && (p.flags & 2)) {
                // the tile's ring lines are dead now: drop them from L2 instead of letting them be written back to HBM
                const char* dead = reinterpret_cast<const char*>(p.ring + (long long) desc.w * N + (long long) c0 * L1);
                for (int ofs = tid * 128; ofs < L1 * C * 8; ofs += THREADS * 128)
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(dead + ofs) : "memory");
            }
            cf v[E];
            constexpr int B0 = E / R0, STR0 = L / R0;
            const cf* srow = xb + g0 * L;                      // this thread's staged column
            if constexpr (PC::SPLIT) srow = g0 < C / 2 ? sbuf + g0 * L : xb + (g0 - C / 2) * L;
            static_for<0, B0>([&](auto q_) {
                constexpr int q = decltype(q_)::value;
                static_for<0, R0>([&](auto t_) {
                    constexpr int t = decltype(t_)::value;
                    v[q * R0 + bitrev<R0>(t)] = srow[j0 + q * T + t * STR0];
                });
            });
            consumer_sync();
            if constexpr (PC::SPLIT) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_free + 1)) : "memory");
            stage_math<T, E, R0, 1, INV, TW_NONE>(v, nullptr, p.table, 0, j0);
            stage_scatter<L, T, E, R0, 1, LOGPAD, DST_XCHG>(v, nullptr, xb + g0 * XBUF, j0, true);
            consumer_sync();
            stage_gather<L, T, E, R1, LOGPAD, SRC_XBUF>(v, nullptr, xb + g1 * XBUF, j1, true);
            constexpr int B1 = E / R1, STR1 = L / R1;
            if constexpr (!PC::REAL) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bfree)) : "memory");
                stage_math<T, E, R1, R0, INV, TW_LUT>(v, lutB, p.table, 0, j1);
                cf* ocol = p.out + prob * N + c0 + g1;
                static_for<0, B1>([&](auto q_) {
                    constexpr int q = decltype(q_)::value;
                    const int jq = j1 + q * T;
                    static_for<0, R1>([&](auto u_) {
                        constexpr int u = decltype(u_)::value;
                        __stcs(ocol + (long long) (jq + u * STR1) * L0, v[q * R1 + u]);
                    });
                });
            } else {
                // real forward.  Bin k = col + L0*row pairs with M - k = (L0 - col) + L0*(L - 1 - row): partner slot,
                // mirrored row.  A thread's rows u < R1/2 are bins below M/2, the others above: the upper half of Z goes
                // through the tile buffer ([slot][row], odd pitch), every thread then evaluates its lower-half bins
                // together with their mirrors -- each pair once, with the arithmetic of the separate split pass
                // (four_step.cuh real_split_kernel) -- and stores Y[k] and Y[M-k].
                consumer_sync();                               // every thread has gathered: the buffer can be rewritten
                stage_math<T, E, R1, R0, INV, TW_LUT>(v, lutB, p.table, 0, j1);
                constexpr int ZP = L + 1;
                constexpr int H = C / 2;
                static_assert(R1 % 2 == 0, "the last radix splits into a lower and an upper half");
                static_for<0, B1>([&](auto q_) {
                    constexpr int q = decltype(q_)::value;
                    static_for<R1 / 2, R1>([&](auto u_) {
                        constexpr int u = decltype(u_)::value;
                        xb[g1 * ZP + j1 + q * T + u * STR1] = v[q * R1 + u];
                    });
                });
                consumer_sync();
                const int col = pass2_column(c0, g1);
                const bool self0 = col == 0, selfh = col == L0 / 2;
                const int gm = (self0 || selfh) ? g1 : (g1 ^ H);              // partner slot
                constexpr long long M = N;
                cf* yrow = p.out + prob * p.out_stride;
                static_assert(32 % R1 == 0, "the row step of a last-stage butterfly is a multiple of 1/64 turn of W_2M");
                static_for<0, B1>([&](auto q_) {
                    constexpr int q = decltype(q_)::value;
                    // split factors W_2M^k, k = col + L0 * (j1 + q*T + u*STR1): rows u are L0 * STR1 = M / R1 apart, so one table
                    // look-up per butterfly (u = 0) times the constants W_(2 R1)^u serves all of them (as in the inverse's twist)
                    const unsigned k0 = (unsigned) col + (unsigned) L0 * (unsigned) (j1 + q * T);
                    const unsigned e0 = k0 << p.tw_shift_real;
                    const cf wbase = cmul(__ldg(p.tw_lo + (e0 & ((1u << p.tw_h) - 1u))), __ldg(p.tw_hi + (e0 >> p.tw_h)));
                    static_for<0, R1 / 2>([&](auto u_) {
                        constexpr int u = decltype(u_)::value;
                        const int kr = j1 + q * T + u * STR1;                  // row < L/2: bin k = col + L0 * kr < M/2
                        const int km = self0 ? (L - kr) : (L - 1 - kr);
                        const cf z0 = v[q * R1 + u];
                        const cf z1 = (self0 && kr == 0) ? z0 : xb[gm * ZP + km];
                        const unsigned k = (unsigned) col + (unsigned) L0 * (unsigned) kr;
                        const cf w = cmul_w64<u * (32 / R1)>(wbase);
                        const cf sum = make_float2(z0.x + z1.x, z0.y - z1.y);
                        const cf dif = make_float2(z0.x - z1.x, z0.y + z1.y);
                        const cf cc = cmul(make_float2(-w.y, w.x), dif);
                        __stcs(yrow + k, make_float2(sum.x - cc.x, sum.y - cc.y));
                        __stcs(yrow + (M - k), make_float2(sum.x + cc.x, -(sum.y + cc.y)));
                    });
                });
                if (self0 && j1 == 0) {                                        // bin M/2 (column 0, row L/2) mirrors itself
                    const cf m = v[R1 / 2];
                    __stcs(yrow + M / 2, make_float2(2.0f * m.x, -2.0f * m.y));
                }
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bfree)) : "memory");
            }
        }
#if CKB_PIPE_STATS
        if (tid == 0) { CKB_STAT_ADD(4 + 2 * pass, CKB_STAT_CLOCK() - c1clk); CKB_STAT_ADD(5 + 2 * pass, 1); }     // [6],[7] pass 1; [8],[9] pass 2: tile landed -> stores issued
#endif
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_stored + (k & 3u))) : "memory");
    }
}

}  // namespace ckb
