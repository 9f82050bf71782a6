// ffma2_probe.cu — FP32 FMA throughput of one B200: scalar FFMA vs packed FFMA2 (fma.rn.f32x2 / __ffma2_rn, sm_100+), 16 independent
// accumulator chains per thread, 8 warps x 4 CTAs per SM.  Prints TFLOP/s for both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ffma2_probe scripts/ffma2_probe.cu && scripts/ffma2_probe
#include <cuda_runtime.h>
#include <stdio.h>

template <int PACKED>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a0, float b0) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 a = make_float2(a0, a0 * 0.999f), b = make_float2(b0, b0 * 1.001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (PACKED) acc[i] = __ffma2_rn(acc[i], a, b);
            else { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, grid = 148 * 8;
    for (int rep = 0; rep < 2; ++rep)
        for (int packed = 0; packed < 2; ++packed) {
            cudaEventRecord(e0);
            if (packed) probe<1><<<grid, 256>>>(out, iters, 0.999f, 1e-3f); else probe<0><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double flop = 2.0 * 32 * (double)iters * grid * 256;
            printf("%s: %.3f ms, %.1f TFLOP/s (%s)\n", packed ? "FFMA2" : "FFMA ", ms, flop / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
