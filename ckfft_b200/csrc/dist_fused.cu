// dist_fused.cu -- fused distributed transform of one very large 1-D FFT over the GPUs of one NVSwitch domain
// (BASELINE config 5, SURVEY.md 8e: "local FFT kernels write directly into per-peer slabs").
//
// The six-step algorithm (same index algebra as the vendored ext/fftw-3.3.2/mpi/dft-rank1.c:58-79, which the
// reference ships but never builds) needs three global transposes.  Here none of them is a separate pass and no
// collective library is on the data path: every array is kept in a layout whose CONTIGUOUS axis is the one that
// stays on the rank, so the FFT passes are strided-column tile passes (four_step.cuh) and the two inner transposes
// reduce to "row kk of the result lives on rank kk >> s": the routed tile kernels (TileCfg::RT) store their
// 128-byte row chunks straight into the destination GPU's memory over NVLink while the next tile is being fetched
// by TMA.  Only the first exchange is a copy kernel of its own (it has no arithmetic to hide behind); it, too,
// pushes straight into the peers' arrays.
//
//   n = n1*n2, x[n1][n2]; rank r owns rows n1 in block r (h = n1/P rows) = its natural-order slice.
//   exchange:  work_q[n1][n2 - q*w]           <- x_r[n1 - r*h][n2],  n2 in column block q (w = n2/P columns)
//   pass A:    (n1 = a*lb + b)  la-point FFTs over a,  * W_n1^(b*ka)            work -> out (as scratch), local
//   pass B:    lb-point FFTs over b, k1 = ka + la*kb,  * W_n^(n2*k1)            -> mid_{rank(k1)}[k1 mod h][n2]
//   pass C:    (n2 = c*ld + d)  lc-point FFTs over c,  * W_n2^(d*kc)            mid -> work, local
//   pass D:    ld-point FFTs over d, k2 = kc + lc*kd                            -> out_{rank(k2)}[k2 mod w][k1]
//   flag barrier (release/acquire words in peer memory) after the exchange, after pass B and after pass D.
#include "tile_launch.h"
#include "ckfft/ckfft_b200.h"

namespace ckb {

static int host_ilog2(long long x) { int l = 0; while ((1LL << l) < x) ++l; return l; }

bool dist_layout(long long n, int world, int prefer, CkFftB200DistLayout* out)
{
    if (!out || n <= 0 || (n & (n - 1)) || world < 1 || world > CKB_MAX_PEERS || (world & (world - 1))) return false;
    const int lg = host_ilog2(n);
    if (lg < 14 || lg > 30) return false;
    CkFftB200DistLayout l{};
    l.log2n = lg;
    l.world = world;
    if (lg <= 20) {
        l.log2n1 = lg / 2;
        l.log2n2 = lg - l.log2n1;
        l.la = l.lc = 1;
        l.lb = 1 << l.log2n1;
        l.ld = 1 << l.log2n2;
        l.passes = 2;
    } else if ((lg <= 27 || prefer == 3) && !(prefer == 4 && lg >= 28)) {
        int lg1 = lg - 14 < 9 ? lg - 14 : 9;
        if (lg1 < lg - 20) lg1 = lg - 20;
        const int lg2 = lg - lg1;
        l.log2n1 = lg1;
        l.log2n2 = lg2;
        l.la = 1;
        l.lb = 1 << lg1;
        l.ld = 1 << (lg2 / 2);            // the pass that stores to the peers takes the shorter factor (16-column tiles)
        l.lc = 1 << (lg2 - lg2 / 2);
        l.passes = 3;
    } else {
        const int lg1 = lg / 2, lg2 = lg - lg1;
        l.log2n1 = lg1;
        l.log2n2 = lg2;
        l.lb = 1 << (lg1 / 2);
        l.la = 1 << (lg1 - lg1 / 2);
        l.ld = 1 << (lg2 / 2);
        l.lc = 1 << (lg2 - lg2 / 2);
        l.passes = 4;
    }
    // every rank needs whole 16-column tiles of its row block and of its column block
    if (((1LL << l.log2n1) / world) < 16 || ((1LL << l.log2n2) / world) < 16) return false;
    *out = l;
    return true;
}

int dist_describe(const CkFftB200DistLayout& l, int rank, CkFftB200DistPass passes[4])
{
    const long long n1 = 1LL << l.log2n1, n2 = 1LL << l.log2n2;
    const long long h = n1 / l.world, w = n2 / l.world;
    int np = 0;
    auto set_pull = [&](CkFftB200DistPass& f) {            // the first pass reads the ranks' input arrays directly
        if (!l.pull) return;
        f.pull = 1;
        f.pullRows = f.L / l.world;
        f.pullRowLen = (long long) (f.ncols / w) * n2;
        f.pullN2 = n2; f.pullCol0 = rank * w; f.pullW = (int) w;
        f.src = 3;
    };
    if (l.la > 1) {
        CkFftB200DistPass a{};
        a.kind = KIND_COLUMN; a.routed = 0; a.L = l.la; a.nproblems = 1; a.ncols = (int) (l.lb * w);
        a.twLog2 = l.log2n1; a.twColBase = 0; a.twColShift = host_ilog2(w);
        a.src = 0; a.dst = 2;
        set_pull(a);
        passes[np++] = a;
    }
    {
        CkFftB200DistPass b{};
        b.kind = KIND_COLUMN; b.routed = 1; b.L = l.lb; b.nproblems = l.la; b.ncols = (int) w;
        b.twLog2 = l.log2n; b.twColBase = (int) (rank * w); b.twColShift = 0;
        b.kProbMul = l.la > 1 ? 1 : 0; b.kMul = l.la; b.rankShift = host_ilog2(h);
        b.outRowStride = n2; b.outColBase = rank * w;
        b.src = l.la > 1 ? 2 : 0; b.dst = 1;
        if (l.la == 1) set_pull(b);
        passes[np++] = b;
    }
    if (l.lc > 1) {
        CkFftB200DistPass c{};
        c.kind = KIND_COLUMN; c.routed = 0; c.L = l.lc; c.nproblems = h; c.ncols = l.ld;
        c.twLog2 = l.log2n2; c.twColBase = 0; c.twColShift = 0;
        c.src = 1; c.dst = 0;
        passes[np++] = c;
    }
    {
        CkFftB200DistPass d{};
        d.kind = KIND_LAST; d.routed = 1; d.L = l.ld; d.nproblems = l.lc; d.ncols = (int) h;
        d.twLog2 = -1;
        d.kProbMul = l.lc > 1 ? 1 : 0; d.kMul = l.lc; d.rankShift = host_ilog2(w);
        d.outRowStride = n1; d.outColBase = rank * h;
        d.inColStride = n2; d.inProbStride = l.ld;
        d.src = l.lc > 1 ? 0 : 1; d.dst = 2;
        passes[np++] = d;
    }
    return np;
}

// ---- the first exchange: rank s pushes the column blocks of its rows into the peers' work arrays ----------
struct PeerPtrs { cf* p[CKB_MAX_PEERS]; };

// x: [h][n2] complex (this rank's rows);  work_q: [n1][w];  row (rank*h + i), columns [q*w, (q+1)*w) -> work_q
// 16-byte units; consecutive threads walk a row, so every peer receives runs of w*8 >= 128 contiguous bytes.
__global__ void __launch_bounds__(256) exchange_push_kernel(const float4* __restrict__ x, PeerPtrs work, long long h, long long n2_2,
                                                            long long w_2, int w2_shift, long long row0)
{
    const long long total = h * n2_2;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const long long row = i / n2_2, col = i - row * n2_2;
        const int q = (int) (col >> w2_shift);
        const long long j = col & (w_2 - 1);
        float4* dst = reinterpret_cast<float4*>(work.p[q]) + (row0 + row) * w_2 + j;
        *dst = __ldcs(x + i);
    }
}

// ---- flag barrier over peer memory -------------------------------------------------------------------------
// Every rank owns a block of words: arrive[q] (written by rank q) and an error word.  Barrier number `epoch`
// (monotonic, the same on every rank): store `epoch` into my slot on every rank with release semantics at system
// scope (the kernel runs after this rank's stores of the preceding pass have completed), then wait until every
// slot of my own block has reached it.  A rank that waits longer than ~20 s gives up and raises the error word
// instead of hanging the GPU.
struct FlagPtrs { unsigned* f[CKB_MAX_PEERS]; };
constexpr int kErrSlot = CKB_MAX_PEERS;

// `poison` (last barrier of an execution only): if this or an earlier barrier of the plan has timed out, the rank's
// result is overwritten with NaNs so that a caller who never asks CkFftB200DistPlanStatus cannot mistake the garbage
// for a spectrum.
__global__ void __launch_bounds__(256) dist_barrier_kernel(FlagPtrs fl, int rank, int world, unsigned epoch, unsigned long long timeout_ns,
                                                           float* poison, long long poison_floats)
{
    const int q = threadIdx.x;
    if (q < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(fl.f[q] + rank), "r"(epoch) : "memory");
        const unsigned* mine = fl.f[rank] + q;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int) (v - epoch) >= 0) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                fl.f[rank][kErrSlot] = 1u;
                break;
            }
            __nanosleep(200);
        }
        __threadfence_system();
    }
    if (poison == nullptr) return;
    __syncthreads();
    if (*reinterpret_cast<volatile unsigned*>(fl.f[rank] + kErrSlot) == 0u) return;
    const float nan = __int_as_float(0x7fc00000);
    for (long long i = threadIdx.x; i < poison_floats; i += blockDim.x) poison[i] = nan;
}

// Barrier timeout: 20 s by default (first-launch module loads, a debugger or an oversubscribed host can delay a peer
// by seconds); CKFFT_B200_DIST_TIMEOUT_MS overrides it.
static unsigned long long barrier_timeout_ns()
{
    static const unsigned long long ns = [] {
        const char* e = getenv("CKFFT_B200_DIST_TIMEOUT_MS");
        const long long ms = e && *e ? atoll(e) : 0;
        return (unsigned long long) (ms > 0 ? ms : 20000) * 1000000ULL;
    }();
    return ns;
}

static cudaError_t launch_barrier(const DistBuffers& b, int rank, int world, unsigned epoch, cudaStream_t s,
                                  cf* poison = nullptr, long long poison_elems = 0)
{
    FlagPtrs fl{};
    for (int q = 0; q < world; ++q) fl.f[q] = b.flags[q];
    dist_barrier_kernel<<<1, poison ? 256 : 32, 0, s>>>(fl, rank, world, epoch, barrier_timeout_ns(), reinterpret_cast<float*>(poison),
                                                        2 * poison_elems);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t launch_routed_pass(bool inverse, int kind, int L, const TileParams& p, cudaStream_t s)
{
    if (kind == KIND_COLUMN)
        return inverse ? launch_tile<true, KIND_COLUMN, true>(L, p, s) : launch_tile<false, KIND_COLUMN, true>(L, p, s);
    return inverse ? launch_tile<true, KIND_LAST, true>(L, p, s) : launch_tile<false, KIND_LAST, true>(L, p, s);
}

// C columns / TMA box rows of the tile plan for length L (tile_launch.h)
static bool tile_plan_of(int L, int* C, int* box_rows)
{
#define X(L_, E_, R0_, R1_, C_, MINB_) \
    if (L == L_) { *C = C_; *box_rows = TileCfg<L_, E_, R0_, R1_, C_, false, KIND_COLUMN, MINB_, true, true>::BOX_ROWS; return true; }
    CKB_TILE_PLANS(X)
#undef X
    return false;
}

// Pull mode: one tensor map per rank over its input array as the first pass sees it, [L/P][rowLen], boxes of
// [min(L/P, box rows)][C]; uploaded once to `dmaps` (device memory, world entries).
cudaError_t dist_make_pull_maps(const CkFftB200DistLayout& l, int rank, cf* const* in, CUtensorMap* dmaps, int* box_rows_out)
{
    CkFftB200DistPass passes[4];
    dist_describe(l, rank, passes);
    const CkFftB200DistPass& f = passes[0];
    if (!f.pull) return cudaErrorInvalidValue;
    int C = 0, box = 0;
    if (!tile_plan_of(f.L, &C, &box)) return cudaErrorInvalidValue;
    if (box > f.pullRows) box = f.pullRows;
    CUtensorMap host[CKB_MAX_PEERS];
    memset(host, 0, sizeof(host));
    for (int q = 0; q < l.world; ++q)
        if (f.pullRowLen >= (1LL << 31) || !make_tile_map(&host[q], in[q], f.pullRows, (int) f.pullRowLen, box, C)) return cudaErrorInvalidValue;
    *box_rows_out = box;
    return cudaMemcpy(dmaps, host, sizeof(CUtensorMap) * l.world, cudaMemcpyHostToDevice);
}

cudaError_t dist_exec(const CkFftB200DistLayout& l, int rank, const DistBuffers& b, unsigned* epoch, const cf* in_local,
                      bool inverse, const cf* table, int log2_nt, const BigTwiddles& tw, cudaStream_t s, DistMarks* marks)
{
    const long long n1 = 1LL << l.log2n1, n2 = 1LL << l.log2n2;
    const long long h = n1 / l.world, w = n2 / l.world;
    cudaError_t e;
    if (marks) marks->count = 0;
    auto mark = [&](const char* name) {
        if (!marks || marks->count >= DistMarks::kMax) return;
        cudaEventRecord(marks->ev[marks->count], s);
        marks->name[marks->count++] = name;
    };
    mark("start");
    if (l.pull) {
        // no exchange kernel: the input only has to sit in the peer-visible array before the barrier
        if (in_local != b.in[rank]) {
            if ((e = cudaMemcpyAsync(b.in[rank], in_local, (size_t) (h * n2) * sizeof(cf), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
            mark("copy-in");
        }
    } else {
        PeerPtrs work{};
        for (int q = 0; q < l.world; ++q) work.p[q] = b.buf[0][q];
        const long long units = h * n2 / 2;
        long long blocks = (units + 256 * 8 - 1) / (256 * 8);
        const long long cap = 16LL * sm_count_of_current_device();
        if (blocks > cap) blocks = cap;
        exchange_push_kernel<<<(int) blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(in_local), work, h, n2 / 2, w / 2,
                                                          host_ilog2(w / 2), (long long) rank * h);
        count_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        mark("exchange");
    }
    if ((e = launch_barrier(b, rank, l.world, ++*epoch, s)) != cudaSuccess) return e;
    mark("barrier");

    CkFftB200DistPass passes[4];
    const int np = dist_describe(l, rank, passes);
    for (int i = 0; i < np; ++i) {
        const CkFftB200DistPass& d = passes[i];
        TileParams p{};
        p.in = d.src == 3 ? b.in[rank] : b.buf[d.src][rank];
        p.out = b.buf[d.dst][rank];
        p.table = table; p.log2_nt = log2_nt;
        p.tw_lo = tw.lo; p.tw_hi = tw.hi; p.tw_h = tw.h;
        p.tw_shift = d.twLog2 >= 0 ? tw.log2_tmax - d.twLog2 : 0;
        p.nproblems = d.nproblems; p.ncols = d.ncols; p.P = 1; p.Q = 1;
        p.stream_in = 1; p.stream_out = d.routed ? 1 : 0;
        p.tw_col_shift = d.twColShift; p.tw_col_base = d.twColBase;
        const char* name = d.routed ? (d.kind == KIND_COLUMN ? "passB(push)" : "passD(push)") : (d.dst == 2 ? "passA" : "passC");
        if (d.pull) {
            p.pull_maps = b.pull_maps; p.pull_rows = d.pullRows; p.pull_box_rows = b.pull_box_rows;
            p.pull_w_shift = host_ilog2(d.pullW); p.pull_n2 = d.pullN2; p.pull_col0 = d.pullCol0;
            name = d.routed ? "passB(pull+push)" : "passA(pull)";
        }
        if (d.routed || d.pull) {
            if (d.routed) {
                for (int q = 0; q < l.world; ++q) p.peer[q] = b.buf[d.dst][q];
                p.k_prob_mul = d.kProbMul; p.k_mul = d.kMul; p.rank_shift = d.rankShift;
                p.out_row_stride = d.outRowStride; p.out_col_base = d.outColBase;
                p.in_col_stride = d.inColStride; p.in_prob_stride = d.inProbStride;
            } else {
                // a local pass run by the routed kernel (it is the one that can pull): every row stays here
                p.peer[0] = p.out; p.k_prob_mul = 0; p.k_mul = 1; p.rank_shift = 30;
                p.out_row_stride = d.ncols; p.out_col_base = 0;
            }
            e = launch_routed_pass(inverse, d.kind, d.L, p, s);
            mark(name);
            if (e == cudaSuccess && d.routed) {
                const bool last = i == np - 1;          // the barrier that completes the execution also guards the result
                e = launch_barrier(b, rank, l.world, ++*epoch, s, last ? b.buf[2][rank] : nullptr, last ? h * n2 : 0);
                mark("barrier");
            }
        } else {
            e = launch_local_pass(inverse, d.kind, d.L, p, s);
            mark(name);
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace ckb
