// conv_smallk3.cuh — small-K convolution (K = kh*kw*C <= 32: the first layer of a CNN, layers.py:442-530 with
// preprocessing.py:34-126) as register-tiled fp32 shared-memory GEMMs, for large batches (config 3's conv 1 at batch 1024:
// 1 M output pixels x 64 filters x 27 taps; 268 MB of output / error against 12.6 MB of input).
//
// Both kernels are persistent and walk tiles of 128 consecutive output pixels. The im2col patch tile lives in shared memory
// K-MAJOR, Xs[tap (ky,kx,c)][pixel]: the gather writes consecutive pixels (conflict-free) and the GEMM reads four pixels of one
// tap as one 128-bit broadcast. The next tile's taps are fetched into registers BEFORE the current tile's FMAs and stored after
// them (one __syncthreads per tile).
//   forward : thread = 4 pixels x 8 (F = 64) or 4 (F = 32) filters; 3 LDS.128 per 32 FFMA; every store instruction writes
//             full 128-byte lines. The first-generation kernel (thread per pixel, 64 accumulators, 16-byte stores 256 B apart)
//             ran at 1.1 TB/s; this one is FFMA-issue bound.
//   wgrad   : dW[(tap | ones), f] = sum_pix X[tap, pix] * E[pix, f]; thread = (filter quad, tap group of <= 9 rows, pixel-quad
//             subset), the error tile E[128][F] arrives by 16-byte cp.async (double-buffered, zero-filled tails), 11 LDS.128 per
//             112 FFMA; the pixel subsets are folded through shared memory at the end; partial[cta][K + 1][F] -> splitk_reduce.
//             The first-generation kernel issued 2 LDS per FFMA (585 us on config 3).
#pragma once
#include "common.cuh"

namespace b200nn {
namespace smallk3 {

constexpr int TP = 128;        // pixels per tile
constexpr int THREADS = 256;

__device__ __forceinline__ void cp16_zfill(void* smem_dst, const void* gmem_src, bool ok) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(ok ? 16 : 0));
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

struct Taps {
    int off[33], dy[33], dx[33];
};

// tap table in the (ky,kx,c) order; rows >= K are never fetched
__device__ __forceinline__ void fill_taps(Taps& t, const b200nn_conv_geom& g, int K) {
    if (threadIdx.x < K) {
        const int k = threadIdx.x, c = k % g.C, tt = k / g.C;
        t.dy[k] = tt / g.kw; t.dx[k] = tt % g.kw;
        t.off[k] = (t.dy[k] * g.W + t.dx[k]) * g.C + c;
    }
}

// the calling thread's share of a tile's patch matrix: pixel (tile * TP + tid % TP), rows (tid / TP) + 2 j.
// Row K is the "ones" row when ONES (bias gradient), rows beyond are zero.
template <int ROWS, bool ONES>
__device__ __forceinline__ void gather(float (&pre)[(ROWS + 1) / 2], const b200nn_conv_geom& g, const Taps& t, const float* __restrict__ x,
                                       int K, int tile, int n_pix) {
    const int pp = threadIdx.x & (TP - 1), kk = threadIdx.x >> 7;
    const int pix = tile * TP + pp;
    const bool pok = pix < n_pix;
    const int ox = pix % g.OW; const int q = pix / g.OW; const int oy = q % g.OH; const int n = q / g.OH;
    const int iy0 = oy * g.sh - g.ph, ix0 = ox * g.sw - g.pw;
    const float* xb = x + ((long long)n * g.H + iy0) * g.W * g.C + (long long)ix0 * g.C;
#pragma unroll
    for (int j = 0; j < (ROWS + 1) / 2; ++j) {
        const int k = kk + 2 * j;
        float v = 0.0f;
        if (pok && k < K) {
            const int iy = iy0 + t.dy[k], ix = ix0 + t.dx[k];
            if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = __ldg(xb + t.off[k]);
        } else if (ONES && pok && k == K) {
            v = 1.0f;
        }
        pre[j] = v;
    }
}
template <int ROWS>
__device__ __forceinline__ void scatter(const float (&pre)[(ROWS + 1) / 2], float* __restrict__ Xs) {
    const int pp = threadIdx.x & (TP - 1), kk = threadIdx.x >> 7;
#pragma unroll
    for (int j = 0; j < (ROWS + 1) / 2; ++j) {
        const int k = kk + 2 * j;
        if (k < ROWS) Xs[k * TP + pp] = pre[j];
    }
}

template <int KT, int NF4>   // KT: taps held (9, 27 or 32 = any K <= 32); NF4 = F / 32
__global__ void __launch_bounds__(THREADS, 2) conv2d_smallk_fwd3_kernel(b200nn_conv_geom g, const float* __restrict__ x,
                                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                                        float* __restrict__ y, int act, float alpha, int n_pix) {
    constexpr int F = NF4 * 32;
    __shared__ __align__(16) float Ws[KT * F];        // [tap (ky,kx,c)][F]
    __shared__ __align__(16) float Xs[2][KT * TP];    // [tap][pixel]
    __shared__ Taps taps;
    const int K = g.kh * g.kw * g.C;
    for (int i = threadIdx.x; i < KT * F; i += THREADS) {
        const int f = i % F, k = i / F;
        float v = 0.0f;
        if (k < K) {
            const int c = k % g.C, tt = k / g.C, kx = tt % g.kw, ky = tt / g.kw;
            v = w[(((long long)c * g.kh + ky) * g.kw + kx) * F + f];   // reference K order (c,ky,kx), Q1
        }
        Ws[i] = v;
    }
    fill_taps(taps, g, K);
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    float4 bv[NF4];
#pragma unroll
    for (int j = 0; j < NF4; ++j)
        bv[j] = bias != nullptr ? *reinterpret_cast<const float4*>(bias + j * 32 + tx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int n_tiles = (n_pix + TP - 1) / TP;
    __syncthreads();
    float pre[(KT + 1) / 2];
    int tile = blockIdx.x;
    if (tile >= n_tiles) return;
    gather<KT, false>(pre, g, taps, x, K, tile, n_pix);
    scatter<KT>(pre, Xs[0]);
    __syncthreads();
    int buf = 0;
    for (; tile < n_tiles; tile += gridDim.x) {
        const int next = tile + gridDim.x;
        if (next < n_tiles) gather<KT, false>(pre, g, taps, x, K, next, n_pix);
        float4 acc[4][NF4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NF4; ++j) acc[i][j] = bv[j];
        const float* xs = Xs[buf] + ty * 4;
        const float* wsp = Ws + tx * 4;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + k * TP);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int j = 0; j < NF4; ++j) {
                const float4 wv = *reinterpret_cast<const float4*>(wsp + k * F + j * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][j].x = fmaf(xa[i], wv.x, acc[i][j].x); acc[i][j].y = fmaf(xa[i], wv.y, acc[i][j].y);
                    acc[i][j].z = fmaf(xa[i], wv.z, acc[i][j].z); acc[i][j].w = fmaf(xa[i], wv.w, acc[i][j].w);
                }
            }
        }
        const long long p0 = (long long)tile * TP + ty * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (p0 + i < n_pix) {
#pragma unroll
                for (int j = 0; j < NF4; ++j) {
                    float4 v = acc[i][j];
                    v.x = act_apply(act, v.x, alpha); v.y = act_apply(act, v.y, alpha);
                    v.z = act_apply(act, v.z, alpha); v.w = act_apply(act, v.w, alpha);
                    *reinterpret_cast<float4*>(y + (p0 + i) * F + j * 32 + tx * 4) = v;
                }
            }
        }
        if (next < n_tiles) scatter<KT>(pre, Xs[buf ^ 1]);
        __syncthreads();
        buf ^= 1;
    }
}

template <int KT>
struct WgShape {
    static constexpr int ROWS = KT + 1;                         // taps + the ones row
    static constexpr int KG = KT <= 9 ? 2 : 4;                  // tap groups
    static constexpr int KR = (ROWS + KG - 1) / KG;             // rows per group (5, 7, 9)
    static constexpr int XROWS = KG * KR;
};

template <int KT, int NF4>
__global__ void __launch_bounds__(THREADS, 2) conv2d_smallk_wgrad3_kernel(b200nn_conv_geom g, const float* __restrict__ x,
                                                                          const float* __restrict__ err, float* __restrict__ partial,
                                                                          int n_pix) {
    using S = WgShape<KT>;
    constexpr int F = NF4 * 32, F4 = F / 4, KG = S::KG, KR = S::KR, XROWS = S::XROWS;
    constexpr int TPS = F4 * KG, PS = THREADS / TPS;            // threads per pixel subset, pixel subsets
    constexpr int QUADS = TP / 4;
    extern __shared__ __align__(16) float smem3[];              // Es[2][TP * F] | Xs[2][XROWS * TP]
    __shared__ Taps taps;
    float* Es = smem3;
    float* Xs = smem3 + 2 * TP * F;
    const int K = g.kh * g.kw * g.C;
    fill_taps(taps, g, K);
    const int fq = threadIdx.x % F4, kg = (threadIdx.x / F4) % KG, ps = threadIdx.x / TPS;
    const int n_tiles = (n_pix + TP - 1) / TP;
    float4 acc[KR];
#pragma unroll
    for (int r = 0; r < KR; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    auto stage_err = [&](int tile, float* dst) {
        const long long base = (long long)tile * TP * F;
        const long long lim = (long long)n_pix * F;
#pragma unroll
        for (int i = threadIdx.x; i < TP * F4; i += THREADS) {
            const long long e = base + (long long)i * 4;
            const bool ok = e < lim;
            cp16_zfill(dst + i * 4, ok ? err + e : err, ok);
        }
    };
    float pre[(XROWS + 1) / 2];
    int tile = blockIdx.x;
    if (tile < n_tiles) {
        stage_err(tile, Es);
        gather<XROWS, true>(pre, g, taps, x, K, tile, n_pix);
        scatter<XROWS>(pre, Xs);
        cp_wait_all();
    }
    __syncthreads();
    int buf = 0;
    for (; tile < n_tiles; tile += gridDim.x) {
        const int next = tile + gridDim.x;
        if (next < n_tiles) {
            stage_err(next, Es + (buf ^ 1) * TP * F);
            gather<XROWS, true>(pre, g, taps, x, K, next, n_pix);
        }
        const float* es = Es + buf * TP * F + fq * 4;
        const float* xs = Xs + buf * XROWS * TP + kg * KR * TP;
#pragma unroll 2
        for (int q = ps; q < QUADS; q += PS) {
            float4 e[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = *reinterpret_cast<const float4*>(es + (q * 4 + i) * F);
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + r * TP + q * 4);
                const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[r].x = fmaf(xa[i], e[i].x, acc[r].x); acc[r].y = fmaf(xa[i], e[i].y, acc[r].y);
                    acc[r].z = fmaf(xa[i], e[i].z, acc[r].z); acc[r].w = fmaf(xa[i], e[i].w, acc[r].w);
                }
            }
        }
        if (next < n_tiles) {
            scatter<XROWS>(pre, Xs + (buf ^ 1) * XROWS * TP);
            cp_wait_all();
        }
        __syncthreads();
        buf ^= 1;
    }
    // fold the pixel subsets: red[ps][row][F] over the (now idle) error stages
    float* red = Es;
    static_assert(PS * XROWS * F <= 2 * TP * F + 2 * XROWS * TP, "reduction scratch exceeds the tile stages");
#pragma unroll
    for (int r = 0; r < KR; ++r)
        *reinterpret_cast<float4*>(red + ((ps * XROWS) + kg * KR + r) * F + fq * 4) = acc[r];
    __syncthreads();
    float* dst = partial + (long long)blockIdx.x * (K + 1) * F;
    for (int i = threadIdx.x; i < (K + 1) * F; i += THREADS) {
        float t = 0.0f;
#pragma unroll
        for (int s = 0; s < PS; ++s) t += red[s * XROWS * F + i];
        dst[i] = t;
    }
}

template <int KT>
constexpr size_t wgrad_smem(int F) { return (size_t)(2 * TP * F + 2 * WgShape<KT>::XROWS * TP) * sizeof(float); }

static inline bool ok(const b200nn_conv_geom* g) {
    const int K = g->kh * g->kw * g->C;
    return K <= 32 && (g->F == 32 || g->F == 64) && (long long)g->N * g->OH * g->OW >= 4096;
}
static inline int blocks(const b200nn_conv_geom* g) {
    const long long n_tiles = ((long long)g->N * g->OH * g->OW + TP - 1) / TP;
    const long long b = n_tiles < kNumSMs * 2 ? n_tiles : kNumSMs * 2;
    return (int)(b < 1 ? 1 : b);
}

template <int KT, int NF4>
static int launch_fwd_t(const b200nn_conv_geom* g, const float* x, const float* w, const float* b, float* y, int act, float alpha, int n_pix,
                        cudaStream_t st) {
    conv2d_smallk_fwd3_kernel<KT, NF4><<<blocks(g), THREADS, 0, st>>>(*g, x, w, b, y, act, alpha, n_pix);
    return 0;
}
static inline void launch_fwd(const b200nn_conv_geom* g, const float* x, const float* w, const float* b, float* y, int act, float alpha,
                              int n_pix, cudaStream_t st) {
    const int K = g->kh * g->kw * g->C;
    if (g->F == 64) {
        if (K == 9) launch_fwd_t<9, 2>(g, x, w, b, y, act, alpha, n_pix, st);
        else if (K == 27) launch_fwd_t<27, 2>(g, x, w, b, y, act, alpha, n_pix, st);
        else launch_fwd_t<32, 2>(g, x, w, b, y, act, alpha, n_pix, st);
    } else {
        if (K == 9) launch_fwd_t<9, 1>(g, x, w, b, y, act, alpha, n_pix, st);
        else if (K == 27) launch_fwd_t<27, 1>(g, x, w, b, y, act, alpha, n_pix, st);
        else launch_fwd_t<32, 1>(g, x, w, b, y, act, alpha, n_pix, st);
    }
}

template <int KT, int NF4>
static int launch_wgrad_t(const b200nn_conv_geom* g, const float* x, const float* err, float* partial, int n_pix, cudaStream_t st) {
    const size_t smem = wgrad_smem<KT>(NF4 * 32);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv2d_smallk_wgrad3_kernel<KT, NF4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev & 63] = true;
    }
    conv2d_smallk_wgrad3_kernel<KT, NF4><<<blocks(g), THREADS, smem, st>>>(*g, x, err, partial, n_pix);
    return 0;
}
static inline int launch_wgrad(const b200nn_conv_geom* g, const float* x, const float* err, float* partial, int n_pix, cudaStream_t st) {
    const int K = g->kh * g->kw * g->C;
    if (g->F == 64) {
        if (K == 9) return launch_wgrad_t<9, 2>(g, x, err, partial, n_pix, st);
        if (K == 27) return launch_wgrad_t<27, 2>(g, x, err, partial, n_pix, st);
        return launch_wgrad_t<32, 2>(g, x, err, partial, n_pix, st);
    }
    if (K == 9) return launch_wgrad_t<9, 1>(g, x, err, partial, n_pix, st);
    if (K == 27) return launch_wgrad_t<27, 1>(g, x, err, partial, n_pix, st);
    return launch_wgrad_t<32, 1>(g, x, err, partial, n_pix, st);
}

}  // namespace smallk3
}  // namespace b200nn
