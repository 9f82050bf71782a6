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
                for (int i = 0; i < 4; ++i) {      // packed FFMA2: two filters per instruction, the roundings of scalar FFMAs
                    const float2 x2 = make_float2(xa[i], xa[i]);
                    const float2 lo = __ffma2_rn(x2, make_float2(wv.x, wv.y), make_float2(acc[i][j].x, acc[i][j].y));
                    const float2 hi = __ffma2_rn(x2, make_float2(wv.z, wv.w), make_float2(acc[i][j].z, acc[i][j].w));
                    acc[i][j] = make_float4(lo.x, lo.y, hi.x, hi.y);
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

// ---- tensor-core forms (math modes tf32 / tf32x3): the same tiles on mma.sync.m16n8k8 TF32 ------------------------------------------
// K is padded to 32 taps (zero rows). The FFMA kernels above spend 864 FFMA + ~670 other instructions per warp and tile; here a warp
// owns 16 pixels x all F filters of a tile: 4 k-steps x F/8 column blocks of m16n8k8, A fragments straight from the K-major patch tile
// (row pitch TP + 8: conflict-free), B fragments from the weights kept in shared memory as {hi, lo} TF32 pairs (pitch F + 8).
// X3: x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi); lo.hi + hi.lo + hi.hi into one fp32 accumulator (chains of 4 k-steps: the
// truncating accumulate of the tensor core stays below 1e-6) — the 22-bit products of the 3xTF32 scheme, so the result meets the FFMA
// path's 1e-5. Single-pass tf32 issues the hi.hi product only.
constexpr int XP = TP + 8;     // patch-tile row pitch (floats)

__device__ __forceinline__ uint32_t rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int ROWS>
__device__ __forceinline__ void scatter_p(const float (&pre)[(ROWS + 1) / 2], float* __restrict__ Xs) {
    const int pp = threadIdx.x & (TP - 1), kk = threadIdx.x >> 7;
#pragma unroll
    for (int j = 0; j < (ROWS + 1) / 2; ++j) {
        const int k = kk + 2 * j;
        if (k < ROWS) Xs[k * XP + pp] = pre[j];
    }
}

template <int NF4, bool X3>
__global__ void __launch_bounds__(THREADS, 2) conv2d_smallk_fwd3_mma_kernel(b200nn_conv_geom g, const float* __restrict__ x,
                                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                                            float* __restrict__ y, int act, float alpha, int n_pix) {
    constexpr int F = NF4 * 32, WP = F + 8, NB = F / 8;
    extern __shared__ __align__(16) float smem3m[];                  // Ws[32][WP] as {hi, lo} pairs | Xs[2][32][XP]
    __shared__ Taps taps;
    uint2* Ws = reinterpret_cast<uint2*>(smem3m);
    float* Xs = smem3m + 2 * 32 * WP;
    const int K = g.kh * g.kw * g.C;
    for (int i = threadIdx.x; i < 32 * F; i += THREADS) {
        const int f = i % F, k = i / F;
        float v = 0.0f;
        if (k < K) {
            const int c = k % g.C, tt = k / g.C, kx = tt % g.kw, ky = tt / g.kw;
            v = w[(((long long)c * g.kh + ky) * g.kw + kx) * F + f];   // reference K order (c,ky,kx), Q1
        }
        const uint32_t hi = rna_tf32(v);
        Ws[k * WP + f] = make_uint2(hi, rna_tf32(v - __uint_as_float(hi)));
    }
    fill_taps(taps, g, K);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gq = lane >> 2, t = lane & 3;
    float2 bv[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
        bv[nb] = bias != nullptr ? *reinterpret_cast<const float2*>(bias + nb * 8 + 2 * t) : make_float2(0.f, 0.f);
    const int n_tiles = (n_pix + TP - 1) / TP;
    __syncthreads();
    float pre[16];
    int tile = blockIdx.x;
    if (tile >= n_tiles) return;
    gather<32, false>(pre, g, taps, x, K, tile, n_pix);
    scatter_p<32>(pre, Xs);
    __syncthreads();
    int buf = 0;
    for (; tile < n_tiles; tile += gridDim.x) {
        const int next = tile + gridDim.x;
        if (next < n_tiles) gather<32, false>(pre, g, taps, x, K, next, n_pix);
        float acc[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) { acc[nb][0] = bv[nb].x; acc[nb][1] = bv[nb].y; acc[nb][2] = bv[nb].x; acc[nb][3] = bv[nb].y; }
        const float* xs = Xs + buf * 32 * XP + warp * 16 + gq;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float af[4] = {xs[(ks * 8 + t) * XP], xs[(ks * 8 + t) * XP + 8], xs[(ks * 8 + t + 4) * XP], xs[(ks * 8 + t + 4) * XP + 8]};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ah[i] = rna_tf32(af[i]);
                al[i] = X3 ? rna_tf32(af[i] - __uint_as_float(ah[i])) : 0u;
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const uint2 b0 = Ws[(ks * 8 + t) * WP + nb * 8 + gq], b1 = Ws[(ks * 8 + t + 4) * WP + nb * 8 + gq];
                if (X3) {
                    mma_tf32_16x8x8(acc[nb], al, b0.x, b1.x);
                    mma_tf32_16x8x8(acc[nb], ah, b0.y, b1.y);
                }
                mma_tf32_16x8x8(acc[nb], ah, b0.x, b1.x);
            }
        }
        // C fragment: rows gq / gq + 8 of the warp's 16 pixels, columns nb * 8 + 2 t, + 1: four lanes cover one 32-byte sector
        const long long p0 = (long long)tile * TP + warp * 16 + gq;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long p = p0 + 8 * h;
            if (p < n_pix) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
                    *reinterpret_cast<float2*>(y + p * F + nb * 8 + 2 * t) =
                        make_float2(act_apply(act, acc[nb][2 * h], alpha), act_apply(act, acc[nb][2 * h + 1], alpha));
            }
        }
        if (next < n_tiles) scatter_p<32>(pre, Xs + (buf ^ 1) * 32 * XP);
        __syncthreads();
        buf ^= 1;
    }
}
template <int NF4>
constexpr size_t fwd_mma_smem() { return (size_t)(2 * 32 * (NF4 * 32 + 8) + 2 * 32 * XP) * sizeof(float); }

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
                for (int i = 0; i < 4; ++i) {      // packed FFMA2 (see the forward)
                    const float2 x2 = make_float2(xa[i], xa[i]);
                    const float2 lo = __ffma2_rn(x2, make_float2(e[i].x, e[i].y), make_float2(acc[r].x, acc[r].y));
                    const float2 hi = __ffma2_rn(x2, make_float2(e[i].z, e[i].w), make_float2(acc[r].z, acc[r].w));
                    acc[r] = make_float4(lo.x, lo.y, hi.x, hi.y);
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

// weight gradient on mma.sync (tf32 / tf32x3): dW[(tap | ones), f] = sum_pix X[tap, pix] * E[pix, f] as M = 32 rows (K taps, the ones row
// K for the bias gradient, zero rows) x N = F x K = pixels. A warp owns 16 pixels of every tile (2 k-steps) and keeps the whole 32 x F
// accumulator (64 registers at F = 64) across the CTA's tiles; patch tile pitch TP + 4 and error tile pitch F + 8 make both fragment
// reads conflict-free; X3 splits both fragments in registers. The warps' accumulators are folded through shared memory at the end into
// partial[cta][K + 1][F] (the layout of the FFMA kernel: same splitk_reduce). Needs K <= 31 (the ones row lives in the 32-row block).
constexpr int XW = TP + 4;
template <int NF4, bool X3>
__global__ void __launch_bounds__(THREADS, 2) conv2d_smallk_wgrad3_mma_kernel(b200nn_conv_geom g, const float* __restrict__ x,
                                                                              const float* __restrict__ err, float* __restrict__ partial,
                                                                              int n_pix) {
    constexpr int F = NF4 * 32, F4 = F / 4, EP = F + 8, NB = F / 8;
    extern __shared__ __align__(16) float smem3w[];              // Es[2][TP][EP] | Xs[2][32][XW]
    __shared__ Taps taps;
    float* Es = smem3w;
    float* Xs = smem3w + 2 * TP * EP;
    const int K = g.kh * g.kw * g.C;
    fill_taps(taps, g, K);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gq = lane >> 2, t = lane & 3;
    const int n_tiles = (n_pix + TP - 1) / TP;
    float acc[2][NB][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mb][nb][i] = 0.0f;
    __syncthreads();
    auto stage_err = [&](int tile, float* dst) {
        const long long base = (long long)tile * TP * F;
        const long long lim = (long long)n_pix * F;
#pragma unroll
        for (int i = threadIdx.x; i < TP * F4; i += THREADS) {
            const long long e = base + (long long)i * 4;
            const bool ok = e < lim;
            cp16_zfill(dst + (i / F4) * EP + (i % F4) * 4, ok ? err + e : err, ok);
        }
    };
    auto scatter_w = [&](const float (&pre)[16], float* dst) {
        const int pp = threadIdx.x & (TP - 1), kk = threadIdx.x >> 7;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[(kk + 2 * j) * XW + pp] = pre[j];
    };
    float pre[16];
    int tile = blockIdx.x;
    if (tile < n_tiles) {
        stage_err(tile, Es);
        gather<32, true>(pre, g, taps, x, K, tile, n_pix);
        scatter_w(pre, Xs);
        cp_wait_all();
    }
    __syncthreads();
    int buf = 0;
    for (; tile < n_tiles; tile += gridDim.x) {
        const int next = tile + gridDim.x;
        if (next < n_tiles) {
            stage_err(next, Es + (buf ^ 1) * TP * EP);
            gather<32, true>(pre, g, taps, x, K, next, n_pix);
        }
        const float* es = Es + buf * TP * EP + (warp * 16 + t) * EP + gq;
        const float* xs = Xs + buf * 32 * XW + gq * XW + warp * 16 + t;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const float* xa = xs + mb * 16 * XW + ks * 8;
                const float af[4] = {xa[0], xa[8 * XW], xa[4], xa[8 * XW + 4]};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ah[mb][i] = rna_tf32(af[i]);
                    al[mb][i] = X3 ? rna_tf32(af[i] - __uint_as_float(ah[mb][i])) : 0u;
                }
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const float bf0 = es[(ks * 8) * EP + nb * 8], bf1 = es[(ks * 8 + 4) * EP + nb * 8];
                const uint32_t bh0 = rna_tf32(bf0), bh1 = rna_tf32(bf1);
                if (X3) {
                    const uint32_t bl0 = rna_tf32(bf0 - __uint_as_float(bh0)), bl1 = rna_tf32(bf1 - __uint_as_float(bh1));
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) {
                        mma_tf32_16x8x8(acc[mb][nb], al[mb], bh0, bh1);
                        mma_tf32_16x8x8(acc[mb][nb], ah[mb], bl0, bl1);
                    }
                }
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) mma_tf32_16x8x8(acc[mb][nb], ah[mb], bh0, bh1);
            }
        }
        if (next < n_tiles) {
            scatter_w(pre, Xs + (buf ^ 1) * 32 * XW);
            cp_wait_all();
        }
        __syncthreads();
        buf ^= 1;
    }
    // fold the warps: red[warp][32][F] over the (now idle) error stages
    float* red = Es;
    static_assert(8 * 32 * F <= 2 * TP * EP, "reduction scratch exceeds the error stages");
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            float* r0 = red + (warp * 32 + mb * 16 + gq) * F + nb * 8 + 2 * t;
            *reinterpret_cast<float2*>(r0) = make_float2(acc[mb][nb][0], acc[mb][nb][1]);
            *reinterpret_cast<float2*>(r0 + 8 * F) = make_float2(acc[mb][nb][2], acc[mb][nb][3]);
        }
    __syncthreads();
    float* dst = partial + (long long)blockIdx.x * (K + 1) * F;
    for (int i = threadIdx.x; i < (K + 1) * F; i += THREADS) {
        float v = 0.0f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) v += red[w8 * 32 * F + i];
        dst[i] = v;
    }
}
template <int NF4>
constexpr size_t wgrad_mma_smem() { return (size_t)(2 * TP * (NF4 * 32 + 8) + 2 * 32 * XW) * sizeof(float); }

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
template <int NF4, bool X3>
static int launch_fwd_mma_t(const b200nn_conv_geom* g, const float* x, const float* w, const float* b, float* y, int act, float alpha, int n_pix,
                            cudaStream_t st) {
    constexpr size_t smem = fwd_mma_smem<NF4>();
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        if (cudaFuncSetAttribute(conv2d_smallk_fwd3_mma_kernel<NF4, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
        configured[dev & 63] = true;
    }
    conv2d_smallk_fwd3_mma_kernel<NF4, X3><<<blocks(g), THREADS, smem, st>>>(*g, x, w, b, y, act, alpha, n_pix);
    return 0;
}
// Which math modes take the mma.sync kernels: B200NN_SMALLK_MMA = 0 none, 1 (default) single-pass tf32 only, 2 tf32 and tf32x3.
// Measured on config 3's first convolution (batch 1024, forward + weight gradient together): tf32 -60 us against the FFMA kernels,
// 3xTF32 +11 us — the legacy mma.sync TF32 path of sm_100 is slow enough that three MMAs per k-step cost what 27 FFMAs per output do,
// so tf32x3 keeps the exact-fp32 FFMA kernels by default.
static inline bool mma_enabled(int tc) {
    const char* e = getenv("B200NN_SMALLK_MMA");
    const int mode = e != nullptr ? atoi(e) : 1;
    return tc != 0 && tc <= mode;
}
// tc: 0 = FFMA (math mode fp32), 1 = single-pass TF32, 2 = 3xTF32
static inline void launch_fwd(const b200nn_conv_geom* g, const float* x, const float* w, const float* b, float* y, int act, float alpha,
                              int n_pix, cudaStream_t st, int tc = 0) {
    const int K = g->kh * g->kw * g->C;
    if (mma_enabled(tc) && (reinterpret_cast<uintptr_t>(b) & 7) == 0) {
        int rc;
        if (g->F == 64) rc = tc == 2 ? launch_fwd_mma_t<2, true>(g, x, w, b, y, act, alpha, n_pix, st) : launch_fwd_mma_t<2, false>(g, x, w, b, y, act, alpha, n_pix, st);
        else rc = tc == 2 ? launch_fwd_mma_t<1, true>(g, x, w, b, y, act, alpha, n_pix, st) : launch_fwd_mma_t<1, false>(g, x, w, b, y, act, alpha, n_pix, st);
        if (rc == 0) return;
    }
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
template <int NF4, bool X3>
static int launch_wgrad_mma_t(const b200nn_conv_geom* g, const float* x, const float* err, float* partial, int n_pix, cudaStream_t st) {
    constexpr size_t smem = wgrad_mma_smem<NF4>();
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv2d_smallk_wgrad3_mma_kernel<NF4, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev & 63] = true;
    }
    conv2d_smallk_wgrad3_mma_kernel<NF4, X3><<<blocks(g), THREADS, smem, st>>>(*g, x, err, partial, n_pix);
    return 0;
}
// tc: 0 = FFMA (math mode fp32), 1 = single-pass TF32, 2 = 3xTF32
static inline int launch_wgrad(const b200nn_conv_geom* g, const float* x, const float* err, float* partial, int n_pix, cudaStream_t st,
                               int tc = 0) {
    const int K = g->kh * g->kw * g->C;
    if (K <= 31 && mma_enabled(tc)) {
        if (g->F == 64) return tc == 2 ? launch_wgrad_mma_t<2, true>(g, x, err, partial, n_pix, st) : launch_wgrad_mma_t<2, false>(g, x, err, partial, n_pix, st);
        return tc == 2 ? launch_wgrad_mma_t<1, true>(g, x, err, partial, n_pix, st) : launch_wgrad_mma_t<1, false>(g, x, err, partial, n_pix, st);
    }
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
