// tma_probe.cu — how fast does one SM's TMA unit fill a 16 KB operand tile (128 rows x 128 B, SWIZZLE_128B) from an L2-resident
// tensor, per tensor-map flavour?  (a) 2-D tiled box {32, 128} of a row-major matrix; (b) 4-D tiled box {32 c, W, 128/W, 1} of an
// NHWC image (the first implicit-GEMM producer); (c) im2col-mode map (channelsPerPixel 32, pixelsPerColumn 128).
// Also checks the im2col load against a host reference (values + zero padding).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tma_probe scripts/tma_probe.cu -lcuda && scripts/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// mode 0: 2-D {k0, m0}; mode 1: 4-D tiled {c0, x, y, n}; mode 2: im2col {c0, w, h, n} + offsets {kx, ky}
template <int MODE>
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, int iters, int span, long long* cycles, float* dump) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar[8];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int s = i & 7;
            if (i >= 8) mbar_wait(smem_u32(&bar[s]), ((i >> 3) - 1) & 1);
            const uint32_t dst = base + s * 16384, b = smem_u32(&bar[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(16384) : "memory");
            const int j = (i * 7 + blockIdx.x * 13) % span;
            if (MODE == 0) {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"((i % 9) * 32), "r"(j * 128) : "memory");
            } else if (MODE == 1) {
                const int tap = i % 9;
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(0), "r"(tap % 3 - 1), "r"(tap / 3 - 1 + 8 * (j & 1)), "r"(j >> 1) : "memory");
            } else {
                const int tap = i % 9;
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(0), "r"(-1), "r"(-1 + 8 * (j & 1)), "r"(j >> 1),
                               "h"((uint16_t)(tap % 3)), "h"((uint16_t)(tap / 3)) : "memory");
            }
        }
        for (int i = iters - 8 < 0 ? 0 : iters - 8; i < iters; ++i) mbar_wait(smem_u32(&bar[i & 7]), (i >> 3) & 1);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (dump != nullptr && blockIdx.x == 0) {   // last tile written to stage (iters-1)&7, raw (swizzled) bytes
        const float* src = reinterpret_cast<const float*>(raw + (base - smem_u32(raw)) + ((iters - 1) & 7) * 16384);
        for (int i = threadIdx.x; i < 4096; i += blockDim.x) dump[i] = src[i];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F>
static F entry(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess) { printf("no %s\n", name); exit(1); }
    return reinterpret_cast<F>(p);
}

int main() {
    const int Nimg = 1024, H = 16, W = 16, C = 64;          // conv2 of config 3
    const size_t elems = (size_t)Nimg * H * W * C;
    std::vector<float> h(elems);
    for (size_t i = 0; i < elems; ++i) h[i] = (float)(i % 65521) * 0.001f + 1.0f;
    float* d; CK(cudaMalloc(&d, elems * 4)); CK(cudaMemcpy(d, h.data(), elems * 4, cudaMemcpyHostToDevice));
    long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
    float* dump; CK(cudaMalloc(&dump, 16384));
    auto tiled = entry<EncodeTiledFn>("cuTensorMapEncodeTiled");
    auto im2col = entry<EncodeIm2colFn>("cuTensorMapEncodeIm2col");
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMap m2, m4, mi;
    {   // the same buffer as a [rows = Nimg*H*W/9.., K = 9*32] matrix: row pitch 9*32*4 B
        cuuint64_t dims[2] = {288, elems / 288}; cuuint64_t str[1] = {288 * 4}; cuuint32_t box[2] = {32, 128};
        CUresult r = tiled(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("tiled 2d failed %d\n", r); return 1; }
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
        cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        cuuint32_t box[4] = {32, 16, 8, 1};
        CUresult r = tiled(&m4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("tiled 4d failed %d\n", r); return 1; }
        int lo[2] = {-1, -1}, hi[2] = {-1, -1};   // 3x3, pad 1: lower = -pad, upper = pad - (k - 1)
        r = im2col(&mi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, lo, hi, 32, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("im2col encode failed %d\n", r); return 1; }
    }
    const int smem = 8 * 16384 + 1024, iters = 2000;
    CK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int grid : {1, 148}) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<grid, 128, smem>>>(m2, iters, 2048, cyc, nullptr);
                if (mode == 1) probe<1><<<grid, 128, smem>>>(m4, iters, 2048, cyc, nullptr);
                if (mode == 2) probe<2><<<grid, 128, smem>>>(mi, iters, 2048, cyc, nullptr);
                CK(cudaDeviceSynchronize());
            }
            long long hc[148]; CK(cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
            printf("grid %3d mode %d (%s): %.1f cycles per 16 KB tile  (%.1f B/clk/SM)\n", grid, mode,
                   mode == 0 ? "2-D tiled {32,128}" : mode == 1 ? "4-D tiled {32,16,8,1}" : "im2col 32 ch x 128 px", (double)mx / iters,
                   16384.0 * iters / mx);
        }
    }
    // correctness of the im2col load: one tile, tap (kx, ky) = ((iters-1)%9 %3, /3) with iters = 1 -> tap 0 = (-1,-1) shift; base pixel (w,h,n) = (-1,-1,0)
    probe<2><<<1, 128, smem>>>(mi, 1, 1, cyc, dump);
    CK(cudaDeviceSynchronize());
    std::vector<float> t(4096);
    CK(cudaMemcpy(t.data(), dump, 16384, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int p = 0; p < 128; ++p) {          // output pixel p = (oy, ox) of image 0; tap 0 reads input (oy - 1, ox - 1)
        const int oy = p / W, ox = p % W, iy = oy - 1, ix = ox - 1;
        for (int c = 0; c < 32; ++c) {
            const float want = (iy < 0 || ix < 0) ? 0.0f : h[((size_t)iy * W + ix) * C + c];
            const int chunk = (c / 4) ^ (p & 7);                      // SWIZZLE_128B
            const float got = t[p * 32 + chunk * 4 + (c & 3)];
            if (got != want && bad++ < 5) printf("mismatch p=%d c=%d got %f want %f\n", p, c, got, want);
        }
    }
    printf("im2col tile check: %s (%d mismatches)\n", bad ? "FAILED" : "ok", bad);
    return 0;
}
