// common.cuh — shared device helpers for libb200nn (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/b200nn.h"

namespace b200nn {

void set_error(const char* fmt, ...);
void count_launch();  // bumps the process-wide kernel-launch counter (b200nn_launch_count)

#define B200NN_REQUIRE(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            ::b200nn::set_error(__VA_ARGS__);     \
            return B200NN_EINVAL;                 \
        }                                         \
    } while (0)

#define B200NN_CUDA(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::b200nn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return B200NN_ECUDA;                                                               \
        }                                                                                      \
    } while (0)

#define B200NN_LAUNCH_CHECK(name)                                                               \
    do {                                                                                       \
        ::b200nn::count_launch();                                                              \
        cudaError_t _e = cudaPeekAtLastError();                                                \
        if (_e != cudaSuccess) {                                                               \
            cudaGetLastError();                                                                \
            ::b200nn::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
            return B200NN_ECUDA;                                                               \
        }                                                                                      \
    } while (0)

constexpr int kNumSMs = 148;  // B200

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- lazy clip chain (see b200nn.h) -------------------------------------------------------------
__device__ __forceinline__ float chain_scale(const double* __restrict__ ss, uint64_t chain) {
    if (ss == nullptr) return 1.0f;
    if (chain >> 63) {   // range form: slots [lo, lo + cnt) in order (deferred clips: every slot since the loss is still pending)
        const int lo = static_cast<int>(chain & 0xffff), cnt = static_cast<int>((chain >> 16) & 0xffff);
        double acc = 1.0;
        for (int i = 0; i < cnt; ++i) {
            const double n2 = acc * acc * ss[lo + i];
            if (n2 > 1.0) acc *= rsqrt(n2);
        }
        return static_cast<float>(acc);
    }
    const int n = static_cast<int>(chain & 0xff);
    if (n == 0) return 1.0f;
    double acc = 1.0;
    for (int i = 0; i < n; ++i) {
        const int slot = static_cast<int>((chain >> (8 + 12 * i)) & 0xfff);
        const double n2 = acc * acc * ss[slot];
        if (n2 > 1.0) acc *= rsqrt(n2);
    }
    return static_cast<float>(acc);
}

// ---- activations ---------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(int act, float x, float alpha) {
    switch (act) {
        case B200NN_ACT_RELU: return fmaxf(x, 0.0f);
        case B200NN_ACT_SIGMOID: {
            const float c = fminf(fmaxf(x, -50.0f), 50.0f);
            return 1.0f / (1.0f + expf(-c));
        }
        case B200NN_ACT_TANH: return tanhf(x);
        case B200NN_ACT_LEAKYRELU: return fmaxf(x, x * alpha);
        case B200NN_ACT_ELU: return x > 0.0f ? x : expm1f(x);
        case B200NN_ACT_SELU: return 1.0507009873554805f * (x > 0.0f ? x : alpha * expm1f(x));
        case B200NN_ACT_GELU: {
            const float z = 0.7978845608028654f * (x + 0.044715f * x * x * x);
            return 0.5f * x * (1.0f + tanhf(z));
        }
        default: return x;
    }
}
// derivative evaluated from the activation OUTPUT y (valid for all kinds above; leaky needs alpha > 0)
__device__ __forceinline__ float act_grad_from_y(int act, float y, float alpha) {
    switch (act) {
        case B200NN_ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
        case B200NN_ACT_SIGMOID: return y * (1.0f - y);
        case B200NN_ACT_TANH: return 1.0f - y * y;
        case B200NN_ACT_LEAKYRELU: return y > 0.0f ? 1.0f : alpha;
        case B200NN_ACT_ELU: return y > 0.0f ? 1.0f : y + 1.0f;                                   // exp(x) = y + 1 for x <= 0
        case B200NN_ACT_SELU: return y > 0.0f ? 1.0507009873554805f : y + 1.0507009873554805f * alpha;   // s*a*exp(x) = y + s*a
        case B200NN_ACT_GELU: {   // the argument is the INPUT x here (activations.py:157-161)
            const float x = y;
            const float z = 0.7978845608028654f * (x + 0.044715f * x * x * x), t = tanhf(z);
            const float g = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
            return 0.5f * (1.0f + t + x * (1.0f - t * t) * g);
        }
        default: return 1.0f;
    }
}

// ---- BatchNormalization helpers (layers.py:1230-1256) ------------------------------------------------
__device__ __forceinline__ float clip10(float v) { return fminf(fmaxf(v, -10.0f), 10.0f); }   // layers.py:1234

// mean / batch_var / invstd of one position from the rank-summed moments (layers.py:1237-1238, :1252)
__device__ __forceinline__ void bn_finalize(double S1, double S2, double n_total, float eps, float& mean, float& batch_var,
                                            float& invstd) {
    const double m = S1 / n_total;
    double var = S2 / n_total - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    batch_var = (float)var + eps;
    invstd = rsqrtf(batch_var + eps);
}

// ---- reductions ----------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of per-thread fp32 partials, promoted to fp64 across warps; result valid in thread 0.
// `scratch` must hold >= 32 doubles of shared memory. Contains __syncthreads(): call from all threads.
__device__ __forceinline__ double block_sum_to_double(float v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
    double d = warp_sum(static_cast<double>(v));
    __syncthreads();
    if (lane == 0) scratch[wid] = d;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// one atomicAdd(double) per CTA into a sum-of-squares slot
__device__ __forceinline__ void block_accumulate(float partial, double* __restrict__ out, double* scratch) {
    const double r = block_sum_to_double(partial, scratch);
    if (threadIdx.x == 0 && out != nullptr) atomicAdd(out, r);
}

}  // namespace b200nn
