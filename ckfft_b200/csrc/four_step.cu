// four_step.cu -- host planner and launchers of the multi-pass path (see four_step.cuh).
#include "tile_launch.h"
#include "plans.h"

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
