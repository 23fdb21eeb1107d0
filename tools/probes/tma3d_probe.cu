// Probe: does a 3-D tiled tensor map accept (a) a z-stride that is not a multiple of the y-stride, (b) dim0 * elemsize > y-stride,
// (c) a box that starts at an odd element (8 bytes off a 16-byte boundary)?  nvcc -arch=sm_100a tma3d_probe.cu -o tma3d_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned long long* out, int n)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + 65536);
    unsigned sb = (unsigned) __cvta_generic_to_shared(bar), sd = (unsigned) __cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(n * 8));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(sd), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(z), "r"(sb) : "memory");
        unsigned ok = 0;
        for (int i = 0; i < 100000 && !ok; ++i)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sb) : "memory");
        for (int i = 0; i < n; ++i) out[i] = ok ? reinterpret_cast<unsigned long long*>(smem)[i] : 0xdeadULL;
    }
}

int main()
{
    const int L0 = 16, L1 = 32, frames = 6;
    const long long istr = (long long) L0 * L1 + 1;            // odd frame stride, like half spectra
    unsigned long long* d;
    cudaMalloc(&d, (frames * istr + 64) * 8);
    unsigned long long* h = new unsigned long long[frames * istr + 64];
    for (long long i = 0; i < frames * istr + 64; ++i) h[i] = (unsigned long long) i;
    cudaMemcpy(d, h, (frames * istr + 64) * 8, cudaMemcpyHostToDevice);
    unsigned long long* out;
    cudaMallocManaged(&out, 4096 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66000);
    struct Case { const char* name; long long base_elem; long long nx; long long zstride_elems; int x, y, z; } cases[] = {
        { "nested strides, aligned", 0, L1, (long long) L0 * L1, 8, 2, 1 },
        { "z-stride 2*(M+1) (not a multiple of the y-stride)", 0, L1, 2 * istr, 8, 2, 1 },
        { "same, odd x", 0, L1, 2 * istr, 9, 2, 1 },
        { "dim0 = L1+1 > y-stride/8, odd x", 0, L1 + 1, 2 * istr, 9, 2, 1 },
        { "base = frame 1 - 1 element (aligned), x = col + 1", istr - 1, L1 + 1, 2 * istr, 8 + 1, 2, 1 },
    };
    for (auto& c : cases) {
        CUtensorMap map;
        memset(&map, 0, sizeof(map));
        cuuint64_t gdim[3] = { (cuuint64_t) c.nx, (cuuint64_t) L0, 2 };
        cuuint64_t gstr[2] = { (cuuint64_t) L1 * 8, (cuuint64_t) c.zstride_elems * 8 };
        cuuint32_t box[3] = { 8, 4, 1 }, es[3] = { 1, 1, 1 };
        CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, d + c.base_elem, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-60s encode FAILED (%d)\n", c.name, (int) r); continue; }
        probe<<<1, 32, 66000>>>(map, c.x, c.y, c.z, out, 32);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-60s kernel FAILED: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        const long long want0 = c.base_elem + c.z * c.zstride_elems + (long long) c.y * L1 + c.x;
        bool ok = true;
        for (int r2 = 0; r2 < 4; ++r2) for (int cc = 0; cc < 8; ++cc) ok = ok && out[r2 * 8 + cc] == (unsigned long long) (want0 + r2 * L1 + cc);
        printf("%-60s %s (first = %llu, want %lld)\n", c.name, ok ? "ok" : "WRONG DATA", out[0], want0);
    }
    return 0;
}
