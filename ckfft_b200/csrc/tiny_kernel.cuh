// tiny_kernel.cuh -- one thread per transform for the sizes below the cooperative kernels:
// complex n = 1, 2, 4, 8 and real n = 1 .. 16.  The closed forms for n = 1, 2 (complex) and
// n = 1, 2, 4 (real) are the reference's own (src/ckfft/fft.cpp:20-31, src/ckfft/fft_real.cpp:20-48,
// 69-93); the rest is a register DFT plus the split / twist of fft_real_default.cpp:13-114.
#pragma once
#include "fft_kernel.cuh"

namespace ckb {

template <int M, bool INV, bool PLANAR = false>
__global__ void __launch_bounds__(128) tiny_c2c_kernel(const KernelParams p)
{
    for (long long item = blockIdx.x * (long long) blockDim.x + threadIdx.x; item < p.batch;
         item += (long long) gridDim.x * blockDim.x) {
        cf v[M];
        if constexpr (PLANAR) {
            // split-complex rows (strides in floats)
            const float* re = reinterpret_cast<const float*>(p.in) + item * p.in_stride;
            const float* im = p.in_im + item * p.in_stride;
            static_for<0, M>([&](auto t_) { constexpr int t = decltype(t_)::value; v[bitrev<M>(t)] = make_float2(re[t], im[t]); });
            fft_regs<M, 0, INV>(v);
            float* ore = reinterpret_cast<float*>(p.out) + item * p.out_stride;
            float* oim = p.out_im + item * p.out_stride;
            static_for<0, M>([&](auto u_) { constexpr int u = decltype(u_)::value; ore[u] = v[u].x; oim[u] = v[u].y; });
        } else {
            const cf* src = p.in + item * p.in_stride;
            cf* dst = p.out + item * p.out_stride;
            static_for<0, M>([&](auto t_) { constexpr int t = decltype(t_)::value; v[bitrev<M>(t)] = src[t]; });
            fft_regs<M, 0, INV>(v);
            static_for<0, M>([&](auto u_) { constexpr int u = decltype(u_)::value; dst[u] = v[u]; });
        }
    }
}

// real forward, n = N floats -> N/2+1 complex.  Strides in floats (input) / complex (output).
template <int N>
__global__ void __launch_bounds__(128) tiny_r2c_kernel(const float* in, cf* out,       // may alias (in-place calls)
                                                       const cf* __restrict__ table, int log2_nt, long long batch,
                                                       long long in_stride, long long out_stride)
{
    for (long long item = blockIdx.x * (long long) blockDim.x + threadIdx.x; item < batch;
         item += (long long) gridDim.x * blockDim.x) {
        const float* x = in + item * in_stride;
        cf* y = out + item * out_stride;
        if constexpr (N == 1) {
            y[0] = make_float2(x[0] * 2.0f, 0.0f);
        } else if constexpr (N == 2) {
            const float x0 = x[0], x1 = x[1];
            y[0] = make_float2((x0 + x1) * 2.0f, 0.0f);
            y[1] = make_float2((x0 - x1) * 2.0f, 0.0f);
        } else if constexpr (N == 4) {
            const float s02 = (x[0] + x[2]) * 2.0f, d02 = (x[0] - x[2]) * 2.0f;
            const float s13 = (x[1] + x[3]) * 2.0f, d13 = (x[1] - x[3]) * 2.0f;
            y[0] = make_float2(s02 + s13, 0.0f);
            y[1] = make_float2(d02, -d13);
            y[2] = make_float2(s02 - s13, 0.0f);
        } else {
            constexpr int M = N / 2;
            const int sh = log2_nt - ilog2(N);
            cf v[M];
            static_for<0, M>([&](auto t_) {
                constexpr int t = decltype(t_)::value;
                v[bitrev<M>(t)] = make_float2(x[2 * t], x[2 * t + 1]);
            });
            fft_regs<M, 0, false>(v);
            static_for<0, M / 2>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                const cf z0 = v[k], z1 = v[(M - k) & (M - 1)];
                const cf w = __ldg(table + (k << sh));
                const cf sum = make_float2(z0.x + z1.x, z0.y - z1.y);
                const cf dif = make_float2(z0.x - z1.x, z0.y + z1.y);
                const cf c = cmul(make_float2(-w.y, w.x), dif);
                y[k] = make_float2(sum.x - c.x, sum.y - c.y);
                y[M - k] = make_float2(sum.x + c.x, -(sum.y + c.y));
            });
            y[M / 2] = make_float2(2.0f * v[M / 2].x, -2.0f * v[M / 2].y);
        }
    }
}

// real inverse, N/2+1 complex -> N floats.  Strides in complex (input) / floats (output).
template <int N>
__global__ void __launch_bounds__(128) tiny_c2r_kernel(const cf* in, float* out,       // may alias (in-place calls)
                                                       const cf* __restrict__ table, int log2_nt, long long batch,
                                                       long long in_stride, long long out_stride)
{
    for (long long item = blockIdx.x * (long long) blockDim.x + threadIdx.x; item < batch;
         item += (long long) gridDim.x * blockDim.x) {
        const cf* y = in + item * in_stride;
        float* x = out + item * out_stride;
        if constexpr (N == 1) {
            x[0] = y[0].x;
        } else if constexpr (N == 2) {
            const float y0 = y[0].x, y1 = y[1].x;
            x[0] = y0 + y1;
            x[1] = y0 - y1;
        } else if constexpr (N == 4) {
            const float s02 = y[0].x + y[2].x, s13 = 2.0f * y[1].x;
            const float d02 = y[0].x - y[2].x, d13 = 2.0f * y[1].y;
            x[0] = s02 + s13;
            x[1] = d02 - d13;
            x[2] = s02 - s13;
            x[3] = d02 + d13;
        } else {
            constexpr int M = N / 2;
            const int sh = log2_nt - ilog2(N);
            cf v[M];
            static_for<0, M / 2>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                const cf y0 = y[k], y1 = y[M - k];
                const cf w = __ldg(table + (k << sh));
                const cf sum = make_float2(y0.x + y1.x, y0.y - y1.y);
                const cf dif = make_float2(y0.x - y1.x, y0.y + y1.y);
                const cf c = cmul(make_float2(w.y, w.x), dif);
                v[bitrev<M>(k)] = make_float2(sum.x + c.x, sum.y + c.y);
                if constexpr (k != 0) v[bitrev<M>(M - k)] = make_float2(sum.x - c.x, -(sum.y - c.y));
            });
            v[bitrev<M>(M / 2)] = make_float2(2.0f * y[M / 2].x, -2.0f * y[M / 2].y);
            fft_regs<M, 0, true>(v);
            static_for<0, M>([&](auto u_) {
                constexpr int u = decltype(u_)::value;
                x[2 * u] = v[u].x;
                x[2 * u + 1] = v[u].y;
            });
        }
    }
}

}  // namespace ckb
