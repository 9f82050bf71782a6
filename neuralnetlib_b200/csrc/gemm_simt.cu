// gemm_simt.cu — fp32 FFMA contraction engine (the rel-1e-5 parity path) for Dense, Conv2D and
// Conv2DTranspose. One register-tiled kernel C[M,N] = sum_k A(m,k) * B(k,n), parameterised by operand
// *functors* that gather straight from the NHWC / row-major tensors (implicit GEMM: no im2col buffer is
// ever materialised, replacing preprocessing.py:34-126), a split-K variant for skinny shapes
// (deterministic: partials to a workspace, reduced in fixed order), and a shared epilogue that fuses
// bias + activation (forward), the lazy clip multiplier, the preceding activation's backward and the
// fp64 sum-of-squares slots (backward).
//
// Tile: 64x64x16, 256 threads, 4x4 micro-tile, register prefetch of the next k-slab.
#include "common.cuh"
#include "epilogue.cuh"
#include "gemm_tc_api.cuh"
#include "conv_smallk3.cuh"

namespace b200nn {

constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

// ------------------------------------------------------------------------------------------------
// operand functors. A(m,k): row(m), col(k), fetch(row, col). B(k,n): row(k), col(n), fetch(row, col).
// Invalid indices (out of range or -1) decode to descriptors that fetch 0.
// ------------------------------------------------------------------------------------------------
struct RowMajorA {  // A(m,k) = p[m*ld + k]
    static constexpr bool kContigK = true;
    const float* p; long long ld; int M, K;
    struct Row { const float* base; };
    struct Col { int k; };
    __device__ Row row(int m) const { return Row{(m >= 0 && m < M) ? p + (long long)m * ld : nullptr}; }
    __device__ Col col(int k) const { return Col{(k >= 0 && k < K) ? k : -1}; }
    __device__ float fetch(const Row& r, const Col& c) const { return (r.base != nullptr && c.k >= 0) ? r.base[c.k] : 0.0f; }
};

struct ColMajorOnesA {  // A(m,k) = p[k*ld + m] for m < Mreal; A(Mreal,k) = 1 when `ones` (db row)
    static constexpr bool kContigK = false;
    const float* p; long long ld; int Mreal, K; int ones;
    struct Row { int m; };  // -1 invalid, Mreal = ones row
    struct Col { const float* base; };
    __device__ Row row(int m) const { return Row{(m >= 0 && m < Mreal + ones) ? m : -1}; }
    __device__ Col col(int k) const { return Col{(k >= 0 && k < K) ? p + (long long)k * ld : nullptr}; }
    __device__ float fetch(const Row& r, const Col& c) const {
        if (r.m < 0 || c.base == nullptr) return 0.0f;
        return r.m == Mreal ? 1.0f : c.base[r.m];
    }
};

struct RowMajorB {  // B(k,n) = p[k*ld + n]
    static constexpr bool kContigN = true;
    const float* p; long long ld; int K, N;
    struct Row { const float* base; };
    struct Col { int n; };
    __device__ Row row(int k) const { return Row{(k >= 0 && k < K) ? p + (long long)k * ld : nullptr}; }
    __device__ Col col(int n) const { return Col{(n >= 0 && n < N) ? n : -1}; }
    __device__ float fetch(const Row& r, const Col& c) const { return (r.base != nullptr && c.n >= 0) ? r.base[c.n] : 0.0f; }
};

struct ColMajorB {  // B(k,n) = p[n*ld + k]   (W^T)
    static constexpr bool kContigN = false;
    const float* p; long long ld; int K, N;
    struct Row { int k; };
    struct Col { const float* base; };
    __device__ Row row(int k) const { return Row{(k >= 0 && k < K) ? k : -1}; }
    __device__ Col col(int n) const { return Col{(n >= 0 && n < N) ? p + (long long)n * ld : nullptr}; }
    __device__ float fetch(const Row& r, const Col& c) const { return (r.k >= 0 && c.base != nullptr) ? c.base[r.k] : 0.0f; }
};

// im2col window gather: pixel = (n, oy, ox) over an OHxOW grid, tap = (ky, kx, c) with c fastest.
// value = src[n, oy*sh + ky - ph, ox*sw + kx - pw, c] (0 outside). `skip_truncated`: a window that does not
// fit entirely inside the source (no padding) yields 0 for every tap (Conv2DTranspose.backward_pass,
// layers.py:2651).
struct PatchGather {
    const float* src; int SH, SW, SC;       // source spatial dims and channels
    int OH, OW;                             // pixel grid
    int kh, kw, sh, sw, ph, pw;
    int n_pix, n_tap;                       // N*OH*OW, kh*kw*SC
    int skip_truncated;
    struct Pix { const float* base; int iy0, ix0; };   // base = image n; nullptr invalid
    struct Tap { int ky, kx, c; };                      // ky = -1 invalid
    __device__ Pix pix(int m) const {
        if (m < 0 || m >= n_pix) return Pix{nullptr, 0, 0};
        const int ox = m % OW; const int t = m / OW; const int oy = t % OH; const int n = t / OH;
        const int iy0 = oy * sh - ph, ix0 = ox * sw - pw;
        if (skip_truncated && (iy0 + kh > SH || ix0 + kw > SW)) return Pix{nullptr, 0, 0};
        return Pix{src + (long long)n * SH * SW * SC, iy0, ix0};
    }
    __device__ Tap tap(int k) const {
        if (k < 0 || k >= n_tap) return Tap{-1, 0, 0};
        const int c = k % SC; const int t = k / SC;
        return Tap{t / kw, t % kw, c};
    }
    __device__ float get(const Pix& p, const Tap& t) const {
        if (p.base == nullptr || t.ky < 0) return 0.0f;
        const int iy = p.iy0 + t.ky, ix = p.ix0 + t.kx;
        if (iy < 0 || iy >= SH || ix < 0 || ix >= SW) return 0.0f;
        return p.base[((long long)iy * SW + ix) * SC + t.c];
    }
};

struct PatchA {  // A(m = pixel, k = tap)
    static constexpr bool kContigK = true;
    PatchGather g;
    using Row = PatchGather::Pix; using Col = PatchGather::Tap;
    __device__ Row row(int m) const { return g.pix(m); }
    __device__ Col col(int k) const { return g.tap(k); }
    __device__ float fetch(const Row& r, const Col& c) const { return g.get(r, c); }
};

struct PatchTOnesA {  // A(m = tap (+ ones row), k = pixel)  — weight gradients
    static constexpr bool kContigK = false;
    PatchGather g; int ones;
    struct Row { PatchGather::Tap t; int one; };
    using Col = PatchGather::Pix;
    __device__ Row row(int m) const {
        if (ones && m == g.n_tap) return Row{PatchGather::Tap{-1, 0, 0}, 1};
        return Row{g.tap(m), 0};
    }
    __device__ Col col(int k) const {
        if (k < 0 || k >= g.n_pix) return Col{nullptr, -1, 0};
        Col c = g.pix(k);
        if (c.base == nullptr) { c.iy0 = -2; }  // in-range pixel whose window is skipped: contributes 0 (and 0 to ones)
        return c;
    }
    __device__ float fetch(const Row& r, const Col& c) const {
        if (r.one) return (c.base != nullptr || c.iy0 == -2) ? 1.0f : 0.0f;
        return g.get(c, r.t);
    }
};

// transposed-window gather (col2im as a gather / Conv2DTranspose forward): pixel = (n, Y, X) over a DHxDW
// grid, tap = (ky, kx, ch): value = src[n, (Y + py - ky)/sh, (X + px - kx)/sw, ch] when divisible and in range.
struct StridedGatherA {
    static constexpr bool kContigK = true;
    const float* src; int SH, SW, SC;       // source (coarse) dims and channels
    int DH, DW;                             // destination (fine) grid
    int kh, kw, sh, sw, py, px;
    int n_pix, n_tap;
    struct Row { const float* base; int y, x; };
    struct Col { int ky, kx, c; };
    __device__ Row row(int m) const {
        if (m < 0 || m >= n_pix) return Row{nullptr, 0, 0};
        const int X = m % DW; const int t = m / DW; const int Y = t % DH; const int n = t / DH;
        return Row{src + (long long)n * SH * SW * SC, Y + py, X + px};
    }
    __device__ Col col(int k) const {
        if (k < 0 || k >= n_tap) return Col{-1, 0, 0};
        const int c = k % SC; const int t = k / SC;
        return Col{t / kw, t % kw, c};
    }
    __device__ float fetch(const Row& r, const Col& c) const {
        if (r.base == nullptr || c.ky < 0) return 0.0f;
        const int ny = r.y - c.ky, nx = r.x - c.kx;
        if (ny < 0 || nx < 0) return 0.0f;
        const int iy = ny / sh, ix = nx / sw;
        if (iy * sh != ny || ix * sw != nx || iy >= SH || ix >= SW) return 0.0f;
        return r.base[((long long)iy * SW + ix) * SC + c.c];
    }
};

// weights of Conv2D seen from the forward: B(k = (ky,kx,c), f) = w[((c*kh + ky)*kw + kx)*F + f]   (Q1)
struct ConvWFwdB {
    static constexpr bool kContigN = true;
    const float* w; int kh, kw, C, F;
    struct Row { const float* base; };
    struct Col { int n; };
    __device__ Row row(int k) const {
        if (k < 0 || k >= kh * kw * C) return Row{nullptr};
        const int c = k % C; const int t = k / C; const int kx = t % kw, ky = t / kw;
        return Row{w + (((long long)c * kh + ky) * kw + kx) * F};
    }
    __device__ Col col(int n) const { return Col{(n >= 0 && n < F) ? n : -1}; }
    __device__ float fetch(const Row& r, const Col& c) const { return (r.base != nullptr && c.n >= 0) ? r.base[c.n] : 0.0f; }
};

// weights of Conv2D seen from dgrad: B(k = (ky,kx,f), c) = w[((c*kh + ky)*kw + kx)*F + f]
struct ConvWDgradB {
    static constexpr bool kContigN = false;
    const float* w; int kh, kw, C, F;
    struct Row { int off; };   // (ky*kw + kx)*F + f ; -1 invalid
    struct Col { const float* base; };
    __device__ Row row(int k) const { return Row{(k >= 0 && k < kh * kw * F) ? k : -1}; }
    __device__ Col col(int n) const { return Col{(n >= 0 && n < C) ? w + (long long)n * kh * kw * F : nullptr}; }
    __device__ float fetch(const Row& r, const Col& c) const { return (r.off >= 0 && c.base != nullptr) ? c.base[r.off] : 0.0f; }
};

// weights of Conv2DTranspose seen from the forward: B(k = (ky,kx,c), f) = w[((ky*kw + kx)*F + f)*C + c]
struct ConvTWFwdB {
    static constexpr bool kContigN = false;
    const float* w; int kh, kw, C, F;
    struct Row { int tap, c; };
    struct Col { int f; };
    __device__ Row row(int k) const {
        if (k < 0 || k >= kh * kw * C) return Row{-1, 0};
        return Row{k / C, k % C};
    }
    __device__ Col col(int n) const { return Col{(n >= 0 && n < F) ? n : -1}; }
    __device__ float fetch(const Row& r, const Col& c) const {
        return (r.tap >= 0 && c.f >= 0) ? w[((long long)r.tap * F + c.f) * C + r.c] : 0.0f;
    }
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <class AF, class BF>
__global__ void __launch_bounds__(256) gemm_simt_kernel(AF af, BF bf, Epilogue epi, int M, int N, int K, int k_chunk,
                                                        float* __restrict__ ws) {
    __shared__ __align__(16) float As[BK][BM + PADM];
    __shared__ __align__(16) float Bs[BK][BN + PADM];
    __shared__ double scratch[32];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int tx = tid & 15, ty = tid >> 4;

    // ---- loader thread mapping (hoist the decode of the index that stays fixed across k-slabs)
    typename AF::Row a_rows[4];
    typename BF::Col b_cols[4];
    if constexpr (AF::kContigK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a_rows[i] = af.row(m0 + (tid >> 4) + i * 16);
    } else {
        a_rows[0] = af.row(m0 + (tid & 63));
    }
    if constexpr (BF::kContigN) {
        b_cols[0] = bf.col(n0 + (tid & 63));
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) b_cols[i] = bf.col(n0 + (tid >> 4) + i * 16);
    }

    float ra[4], rb[4];
    auto load_tile = [&](int k0) {
        if constexpr (AF::kContigK) {
            const int k = k0 + (tid & 15);
            const typename AF::Col c = af.col(k < k_end ? k : -1);
#pragma unroll
            for (int i = 0; i < 4; ++i) ra[i] = af.fetch(a_rows[i], c);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + (tid >> 6) + i * 4;
                ra[i] = af.fetch(a_rows[0], af.col(k < k_end ? k : -1));
            }
        }
        if constexpr (BF::kContigN) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + (tid >> 6) + i * 4;
                rb[i] = bf.fetch(bf.row(k < k_end ? k : -1), b_cols[0]);
            }
        } else {
            const int k = k0 + (tid & 15);
            const typename BF::Row r = bf.row(k < k_end ? k : -1);
#pragma unroll
            for (int i = 0; i < 4; ++i) rb[i] = bf.fetch(r, b_cols[i]);
        }
    };
    auto store_tile = [&]() {
        if constexpr (AF::kContigK) {
#pragma unroll
            for (int i = 0; i < 4; ++i) As[tid & 15][(tid >> 4) + i * 16] = ra[i];
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) As[(tid >> 6) + i * 4][tid & 63] = ra[i];
        }
        if constexpr (BF::kContigN) {
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[(tid >> 6) + i * 4][tid & 63] = rb[i];
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[tid & 15][(tid >> 4) + i * 16] = rb[i];
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    if (k_begin < k_end) {
        load_tile(k_begin);
        store_tile();
        __syncthreads();
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
            const bool has_next = k0 + BK < k_end;
            if (has_next) load_tile(k0 + BK);
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
            if (has_next) {
                store_tile();
                __syncthreads();
            }
        }
    }

    if (ws != nullptr) {  // split-K: raw partials, reduced by splitk_reduce_kernel
        float* dst = ws + (long long)blockIdx.z * M * N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n < N) dst[(long long)m * N + n] = acc[i][j];
            }
        }
        return;
    }
    EpiState st;
    epi_begin(epi, st);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) epi_store(epi, st, m, n, acc[i][j]);
        }
    }
    epi_finish(epi, st, scratch);
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                                            Epilogue epi) {
    __shared__ double scratch[32];
    EpiState st;
    epi_begin(epi, st);
    const long long total = (long long)M * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float v = 0.0f;
        for (int s = 0; s < splits; ++s) v += ws[(long long)s * total + i];
        epi_store(epi, st, (int)(i / N), (int)(i % N), v);
    }
    epi_finish(epi, st, scratch);
}

// The same for plain row-major outputs with N % 4 == 0 (activations and input errors of the tensor-core GEMMs: up to
// hundreds of MB of partials): float4 traffic, one index division per four outputs, activation kinds at compile time
// (AM / MM: 0 none, 1 ReLU, 2 run time). The scalar kernel above spends ~47 instructions per element on its generic epilogue
// and ran at 0.9 TB/s.
template <int AM, int MM>
__global__ void __launch_bounds__(256) splitk_reduce_vec_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                                                Epilogue epi) {
    __shared__ double scratch[32];
    EpiState st;
    epi_begin(epi, st);
    const int act = AM == 0 ? B200NN_ACT_NONE : (AM == 1 ? B200NN_ACT_RELU : epi.act);
    const int mact = MM == 1 ? B200NN_ACT_RELU : epi.mask_act;
    const long long total4 = (long long)M * N / 4;
    const int N4 = N / 4;
    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(ws);
    float a0 = 0.0f, a1 = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = w4[i];
        int s = 1;
        for (; s + 1 < splits; s += 2) {
            const float4 p = w4[(long long)s * total4 + i], q = w4[(long long)(s + 1) * total4 + i];
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        }
        if (s < splits) {
            const float4 p = w4[(long long)s * total4 + i];
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        const long long m = i / N4;
        const int n = (int)(i - m * N4) * 4;
        const float4 b = epi.bias != nullptr ? *reinterpret_cast<const float4*>(epi.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        float o[4] = {fmaf(v.x, st.scale, b.x), fmaf(v.y, st.scale, b.y), fmaf(v.z, st.scale, b.z), fmaf(v.w, st.scale, b.w)};
        float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MM != 0) mk = *reinterpret_cast<const float4*>(epi.mask_y + m * epi.ldo + n);
        const float mkv[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x = act_apply(act, o[j], epi.alpha);
            a0 = fmaf(x, x, a0);
            if (MM != 0) {
                x *= act_grad_from_y(mact, mkv[j], epi.mask_alpha);
                a1 = fmaf(x, x, a1);
            }
            o[j] = x;
        }
        *reinterpret_cast<float4*>(epi.out + m * epi.ldo + n) = make_float4(o[0], o[1], o[2], o[3]);
    }
    st.a0 = a0; st.a1 = a1;
    epi_finish(epi, st, scratch);
}


// Same reduction for FEW outputs and MANY splits (weight gradients of small-K convolutions: a few hundred
// outputs, hundreds of per-CTA partials): 32 outputs per CTA, the 8 warps stride over the splits with coalesced
// 128-byte reads, shared-memory combine, warp 0 applies the epilogue.
__global__ void __launch_bounds__(256) splitk_reduce_wide_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                                                 Epilogue epi) {
    __shared__ float sm[8][32];
    __shared__ double scratch[32];
    EpiState st;
    epi_begin(epi, st);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long total = (long long)M * N;
    const long long i = (long long)blockIdx.x * 32 + lane;
    float v = 0.0f;
    if (i < total) {
#pragma unroll 4
        for (int s = w; s < splits; s += 8) v += ws[(long long)s * total + i];
    }
    sm[w][lane] = v;
    __syncthreads();
    if (w == 0 && i < total) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][lane];
        epi_store(epi, st, (int)(i / N), (int)(i % N), t);
    }
    epi_finish(epi, st, scratch);
}

int launch_splitk_reduce(const float* ws, int splits, int M, int N, const Epilogue& epi, cudaStream_t st) {
    const long long total = (long long)M * N;
    if (total <= 32768 && splits >= 16) {
        splitk_reduce_wide_kernel<<<(int)((total + 31) / 32), 256, 0, st>>>(ws, splits, M, N, epi);
        B200NN_LAUNCH_CHECK("splitk_reduce_wide");
        return B200NN_OK;
    }
    const uintptr_t al = reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(epi.out) |
                         reinterpret_cast<uintptr_t>(epi.bias) | reinterpret_cast<uintptr_t>(epi.mask_y);
    if (N % 4 == 0 && epi.ldo % 4 == 0 && (al & 15) == 0 && epi.row_mode == 0 && epi.n_rows >= M && total >= (1 << 16)) {
        long long blocks = (total / 4 + 255) / 256;
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        const int am = epi.act == B200NN_ACT_NONE ? 0 : (epi.act == B200NN_ACT_RELU ? 1 : 2);
        const int mm = epi.mask_y == nullptr ? 0 : (epi.mask_act == B200NN_ACT_RELU ? 1 : 2);
#define RV(A, B) splitk_reduce_vec_kernel<A, B><<<(int)blocks, 256, 0, st>>>(ws, splits, M, N, epi); break;
        switch (am * 3 + mm) {
            case 0: RV(0, 0) case 1: RV(0, 1) case 2: RV(0, 2) case 3: RV(1, 0) case 4: RV(1, 1) case 5: RV(1, 2)
            case 6: RV(2, 0) case 7: RV(2, 1) default: RV(2, 2)
        }
#undef RV
        B200NN_LAUNCH_CHECK("splitk_reduce_vec");
        return B200NN_OK;
    }
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    splitk_reduce_kernel<<<(int)blocks, 256, 0, st>>>(ws, splits, M, N, epi);
    B200NN_LAUNCH_CHECK("splitk_reduce");
    return B200NN_OK;
}

// column sums, 128-bit form (F % 4 == 0, 16-byte aligned rows): thread = a quad of columns, eight row loads in flight (the scalar form
// below keeps four 4-byte loads in flight: 67 us for 134 MB at config 4)
// maxbits (optional): the CTA's max |x| over finite elements is folded into *maxbits with atomicMax (bit pattern; zeroed by the caller) --
// the range scale of the fp16 operand planes (f16_split.cu) as a by-product of the pass that reads the error matrix anyway
__device__ __forceinline__ float absmax4(float m, const float4& v) {
    return fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
}
__device__ __forceinline__ void publish_absmax(float m, unsigned* maxbits) {
    __shared__ unsigned sm_max[8];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm_max[threadIdx.x >> 5] = __float_as_uint(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned r = 0;
        for (int i = 0; i < 8; ++i) r = max(r, sm_max[i]);        // non-negative floats order like their bit patterns
        if (r != 0 && r < 0x7f800000u) atomicMax(maxbits, r);      // NaN / Inf never become the scale (as f16::maxabs_kernel)
    }
}
__global__ void __launch_bounds__(256) colsum_partial4_kernel(const float* __restrict__ x, long long R, int F, float* __restrict__ partial,
                                                              unsigned* maxbits = nullptr) {
    const int F4 = F >> 2;
    float mx = 0.0f;
    const long long rows_per_block = (R + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int c = threadIdx.x; c < F4; c += blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long r = r0; r < r1; r += 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = r + j < r1 ? x4[(r + j) * F4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; mx = absmax4(mx, v[j]); }
        }
        reinterpret_cast<float4*>(partial + (long long)blockIdx.x * F)[c] = s;
    }
    if (maxbits != nullptr) publish_absmax(mx, maxbits);
}

// column sums, 128-bit form for NARROW matrices (F / 4 divides 256: conv bias gradients with 32..512 filters): thread = (column quad,
// row lane), 256 / (F / 4) row lanes per CTA, eight row loads in flight per thread, row lanes folded through shared memory
__global__ void __launch_bounds__(256) colsum_partial4r_kernel(const float* __restrict__ x, long long R, int F, float* __restrict__ partial,
                                                               unsigned* maxbits = nullptr) {
    __shared__ float4 red[256];
    float mx = 0.0f;
    const int F4 = F >> 2, RL = 256 / F4;
    const int c = threadIdx.x % F4, ty = threadIdx.x / F4;
    const long long rows_per_block = (R + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r = r0 + ty; r < r1; r += 8LL * RL) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = r + (long long)j * RL < r1 ? x4[(r + (long long)j * RL) * F4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; mx = absmax4(mx, v[j]); }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    if (ty == 0) {
        for (int l = 1; l < RL; ++l) { const float4 t = red[l * F4 + c]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        reinterpret_cast<float4*>(partial + (long long)blockIdx.x * F)[c] = s;
    }
    if (maxbits != nullptr) publish_absmax(mx, maxbits);
}
static bool colsum4r_ok(const void* x, const void* ws, int F) {
    const int F4 = F >> 2;
    return (F & 3) == 0 && F4 >= 1 && F4 <= 256 && 256 % F4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0;
}

// column sums of a [R, F] matrix with the lazy clip multiplier: out[f] = scale * sum_r x[r, f].
// partial: the 256 threads of a CTA form (256 / FT) row lanes x FT columns (FT = min(F rounded up to 32, 256)); each thread
// keeps four independent accumulators over its rows, the row lanes are folded through shared memory.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, long long R, int F,
                                                             float* __restrict__ partial) {
    __shared__ float red[256];
    const int FT = F >= 256 ? 256 : ((F + 31) & ~31);
    const int RL = 256 / FT;
    const int tx = threadIdx.x % FT, ty = threadIdx.x / FT;
    const long long rows_per_block = (R + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(R, r0 + rows_per_block);
    for (int f0 = 0; f0 < F; f0 += FT) {
        const int f = f0 + tx;
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        if (f < F && ty < RL) {
            long long r = r0 + ty;
            for (; r + 3LL * RL < r1; r += 4LL * RL) {
                s0 += x[r * F + f];
                s1 += x[(r + RL) * F + f];
                s2 += x[(r + 2LL * RL) * F + f];
                s3 += x[(r + 3LL * RL) * F + f];
            }
            for (; r < r1; r += RL) s0 += x[r * F + f];
        }
        red[threadIdx.x] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (ty == 0 && f < F) {
            float s = red[tx];
            for (int l = 1; l < RL; ++l) s += red[l * FT + tx];
            partial[(long long)blockIdx.x * F + f] = s;
        }
        __syncthreads();
    }
}
// final: 32 columns x 8 lanes per CTA over the per-block partials
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int nblocks, int F, const double* ss,
                                                           uint64_t chain, float* __restrict__ out) {
    __shared__ float red[8][33];
    const float scale = chain_scale(ss, chain);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + tx;
    float s = 0.0f;
    if (f < F)
        for (int b = ty; b < nblocks; b += 8) s += partial[(long long)b * F + f];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && f < F) {
        for (int l = 1; l < 8; ++l) s += red[l][tx];
        out[f] = s * scale;
    }
}


// ------------------------------------------------------------------------------------------------
// small-K convolution fast paths (K = C*kh*kw <= 32: first layers on 1- or 3-channel images).
// With so few taps the contraction is HBM-bound, so instead of the 64x64 GEMM tile
// (which wastes >80 % of its FMAs on padding) these kernels stream the big tensor exactly once with
// 128-bit coalesced accesses and keep the tiny operand on chip.
// ------------------------------------------------------------------------------------------------
constexpr int SMALLK_MAX = 32;

// y[pix, :] = act(b + sum_tap x[pix, tap] * w[tap, :]).  One thread per output pixel computes ALL F filters: the
// K input taps are loaded once per pixel (not once per filter group), the weights come from shared memory as
// warp-broadcast 128-bit reads, and the F outputs leave as F/4 float4 stores. ~2 instructions per FMA instead of ~8.
template <int FV>   // F / 4
__global__ void __launch_bounds__(128) conv2d_smallk_fwd_kernel(b200nn_conv_geom g, const float* __restrict__ x,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ y,
                                                                int act, float alpha, int n_pix) {
    constexpr int F = FV * 4;
    __shared__ __align__(16) float ws[SMALLK_MAX * F];  // [tap (ky,kx,c)][F]
    __shared__ __align__(16) float bs[F];
    const int K = g.kh * g.kw * g.C;
    for (int i = threadIdx.x; i < K * F; i += blockDim.x) {
        const int f = i % F, k = i / F;
        const int c = k % g.C, t = k / g.C, kx = t % g.kw, ky = t / g.kw;
        ws[i] = w[(((long long)c * g.kh + ky) * g.kw + kx) * F + f];   // reference K order (c,ky,kx), Q1
    }
    for (int i = threadIdx.x; i < F; i += blockDim.x) bs[i] = bias != nullptr ? bias[i] : 0.0f;
    __syncthreads();
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n_pix; pix += gridDim.x * blockDim.x) {
        const int ox = pix % g.OW; const int t = pix / g.OW; const int oy = t % g.OH; const int n = t / g.OH;
        const float* xi = x + (long long)n * g.H * g.W * g.C;
        float4 acc[FV];
#pragma unroll
        for (int j = 0; j < FV; ++j) acc[j] = *reinterpret_cast<const float4*>(&bs[j * 4]);
        int k = 0;
        for (int ky = 0; ky < g.kh; ++ky) {
            const int iy = oy * g.sh + ky - g.ph;
            for (int kx = 0; kx < g.kw; ++kx) {
                const int ix = ox * g.sw + kx - g.pw;
                const bool ok = iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
                const float* xp = xi + ((long long)iy * g.W + ix) * g.C;
                for (int c = 0; c < g.C; ++c, ++k) {
                    const float xv = ok ? xp[c] : 0.0f;
                    const float4* wk = reinterpret_cast<const float4*>(&ws[k * F]);
#pragma unroll
                    for (int j = 0; j < FV; ++j) {
                        const float4 wv = wk[j];
                        acc[j].x = fmaf(xv, wv.x, acc[j].x); acc[j].y = fmaf(xv, wv.y, acc[j].y);
                        acc[j].z = fmaf(xv, wv.z, acc[j].z); acc[j].w = fmaf(xv, wv.w, acc[j].w);
                    }
                }
            }
        }
        float4* yo = reinterpret_cast<float4*>(y + (long long)pix * F);
#pragma unroll
        for (int j = 0; j < FV; ++j) {
            float4 v = acc[j];
            v.x = act_apply(act, v.x, alpha); v.y = act_apply(act, v.y, alpha);
            v.z = act_apply(act, v.z, alpha); v.w = act_apply(act, v.w, alpha);
            yo[j] = v;
        }
    }
}

// Weight gradient of a small-K convolution as a shared-memory mini-GEMM:  dW[tap, f] = sum_pix X[tap, pix] * R[pix, f]
// (+ a "ones" tap that yields db). A CTA walks a contiguous range of output pixels in tiles: the error tile
// R[TP][F] is staged with coalesced float4 loads (read from HBM exactly once), the im2col patch tile X[K+1][TP] is
// gathered once per pixel, and thread (tap group, f) accumulates its few (tap, f) outputs from shared memory
// (1 conflict-free + 1 broadcast LDS per FMA). Per-CTA partials go to `partial[block][K+1][F]`.
constexpr int WG_TILE_FLOATS = 8192;   // error tile: TP * F floats = 32 KB
template <int NACC>
__global__ void __launch_bounds__(256) conv2d_smallk_wgrad_kernel(b200nn_conv_geom g, const float* __restrict__ x,
                                                                  const float* __restrict__ err,
                                                                  float* __restrict__ partial, int n_pix, int TP) {
    extern __shared__ __align__(16) float smem[];   // Rs[TP][F] | Xs[K+1][TP]
    __shared__ int tap_off[SMALLK_MAX], tap_dy[SMALLK_MAX], tap_dx[SMALLK_MAX];
    const int K = g.kh * g.kw * g.C, F = g.F;
    float* Rs = smem;
    float* Xs = smem + TP * F;
    if (threadIdx.x < K) {
        const int k = threadIdx.x;
        const int c = k % g.C, tt = k / g.C;
        tap_dy[k] = tt / g.kw; tap_dx[k] = tt % g.kw;
        tap_off[k] = (tap_dy[k] * g.W + tap_dx[k]) * g.C + c;
    }
    const int TG = blockDim.x / F;            // tap groups
    const int f = threadIdx.x % F, tg = threadIdx.x / F;
    float acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0f;
    int chunk = (n_pix + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + TP - 1) / TP * TP;
    const int p_begin = blockIdx.x * chunk, p_end = min(n_pix, p_begin + chunk);
    for (int p0 = p_begin; p0 < p_end; p0 += TP) {
        __syncthreads();   // previous tile fully consumed (also orders the tap table on the first pass)
        const int np = min(TP, p_end - p0);
        // error tile, float4
        const float4* src = reinterpret_cast<const float4*>(err + (long long)p0 * F);
        const int n4 = np * F / 4, t4 = TP * F / 4;
        for (int i = threadIdx.x; i < t4; i += blockDim.x)
            reinterpret_cast<float4*>(Rs)[i] = i < n4 ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        // patch tile: TP divides the block size, so a thread owns ONE pixel of the tile (decoded once) and a
        // strided subset of its taps
        {
            const int pp = threadIdx.x % TP, kstep = blockDim.x / TP;
            const int pix = p0 + pp;
            const bool pok = pp < np;
            const int ox = pix % g.OW, t = pix / g.OW, oy = t % g.OH, n = t / g.OH;
            const int iy0 = oy * g.sh - g.ph, ix0 = ox * g.sw - g.pw;
            const float* xb = x + ((long long)n * g.H + iy0) * g.W * g.C + (long long)ix0 * g.C;
            for (int k = threadIdx.x / TP; k <= K; k += kstep) {
                float v = 0.0f;
                if (pok) {
                    if (k == K) {
                        v = 1.0f;
                    } else {
                        const int iy = iy0 + tap_dy[k], ix = ix0 + tap_dx[k];
                        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = xb[tap_off[k]];
                    }
                }
                Xs[k * TP + pp] = v;
            }
        }
        __syncthreads();
        for (int pp = 0; pp < TP; pp += 4) {
            const float r0 = Rs[(pp + 0) * F + f], r1 = Rs[(pp + 1) * F + f], r2 = Rs[(pp + 2) * F + f], r3 = Rs[(pp + 3) * F + f];
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
                const int t = tg + a * TG;
                if (t <= K) {
                    const float4 xv = *reinterpret_cast<const float4*>(&Xs[t * TP + pp]);   // warp-broadcast
                    acc[a] = fmaf(xv.x, r0, acc[a]); acc[a] = fmaf(xv.y, r1, acc[a]);
                    acc[a] = fmaf(xv.z, r2, acc[a]); acc[a] = fmaf(xv.w, r3, acc[a]);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        const int t = tg + a * TG;
        if (t <= K) partial[((long long)blockIdx.x * (K + 1) + t) * F + f] = acc[a];
    }
}


static bool smallk_fwd_ok(const b200nn_conv_geom* g) {
    const int K = g->kh * g->kw * g->C;
    return K <= SMALLK_MAX && (g->F == 32 || g->F == 64 || g->F == 16);
}
static bool smallk_wgrad_ok(const b200nn_conv_geom* g) {
    const int K = g->kh * g->kw * g->C;
    return K <= SMALLK_MAX && g->F >= 4 && g->F <= 256 && 256 % g->F == 0;
}
static int smallk_wgrad_tile(const b200nn_conv_geom* g) {   // pixels per tile: divides 256, multiple of 4
    const int tp = WG_TILE_FLOATS / g->F;
    return tp > 256 ? 256 : tp;
}
static int smallk_wgrad_blocks(const b200nn_conv_geom* g) {
    const long long n_pix = (long long)g->N * g->OH * g->OW;
    const int TP = smallk_wgrad_tile(g);
    long long blocks = (n_pix + TP - 1) / TP;       // at least one tile per CTA
    if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;
    return (int)(blocks < 1 ? 1 : blocks);
}


// ------------------------------------------------------------------------------------------------
// Small-N Dense (classifier heads: Dense(10), Dense(1)).  With N <= 32 a 64x64 / 128xN MMA tile is almost all
// padding and the layer is pure latency, so these kernels map threads to the long dimensions instead.
// ------------------------------------------------------------------------------------------------
constexpr int SMALLN_MAX = 32;
constexpr int HEAD_MAXK = 256;   // the fused head can also emit per-CTA weight-gradient partials for inputs this wide

// y[m, :] = act(b + x[m, :] . w)   — one warp per row, lanes stride over k, N accumulators per lane, shuffle-reduce
__global__ void __launch_bounds__(256) dense_smalln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ b, float* __restrict__ y, int M,
                                                               int N, int K, int act, float alpha) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
        float acc[SMALLN_MAX];
#pragma unroll
        for (int n = 0; n < SMALLN_MAX; ++n) acc[n] = 0.0f;
        const float* xr = x + (long long)m * K;
        for (int k = lane; k < K; k += 32) {
            const float xv = xr[k];
            const float* wr = w + (long long)k * N;
#pragma unroll
            for (int n = 0; n < SMALLN_MAX; ++n)
                if (n < N) acc[n] = fmaf(xv, wr[n], acc[n]);
        }
        float mine = 0.0f;
#pragma unroll
        for (int n = 0; n < SMALLN_MAX; ++n) {
            if (n < N) {
                const float t = warp_sum(acc[n]);
                if (lane == n) mine = t;
            }
        }
        if (lane < N) y[(long long)m * N + lane] = act_apply(act, mine + (b != nullptr ? b[lane] : 0.0f), alpha);
    }
}

// dx[m, k] = scale * sum_n e[m, n] * w[k, n]  (+ fused activation backward and the two sums of squares)
__global__ void __launch_bounds__(256) dense_smalln_dx_kernel(const float* __restrict__ e, const float* __restrict__ w,
                                                              Epilogue epi, int M, int N, int K) {
    __shared__ double scratch[32];
    EpiState st;
    epi_begin(epi, st);
    const long long total = (long long)M * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / K), k = (int)(i % K);
        const float* er = e + (long long)m * N;
        const float* wr = w + (long long)k * N;
        float v = 0.0f;
        for (int n = 0; n < N; ++n) v = fmaf(er[n], wr[n], v);
        const float mk = epi.mask_y != nullptr ? epi.mask_y[i] : 0.0f;
        epi_store_premask(epi, st, m, k, v, mk);
    }
    epi_finish(epi, st, scratch);
}

// partial[chunk][k (+ ones row K)][n] = sum over the chunk's rows of x[m, k] * e[m, n]. The error rows are staged zero-padded to NMAX
// columns (128-bit broadcast reads, no per-class bound check: the branchy form was instruction-bound, 192 us at 8192 x 4096 x 10).
template <int NMAX>
__global__ void __launch_bounds__(256) dense_smalln_dw_kernel(const float* __restrict__ x, const float* __restrict__ e,
                                                              float* __restrict__ partial, int M, int N, int K,
                                                              int rows_per_chunk) {
    __shared__ __align__(16) float es[64 * NMAX];   // the chunk's error rows, staged 64 rows at a time
    const int k = blockIdx.x * blockDim.x + threadIdx.x;    // 0..K (K = ones row -> db)
    const int m_begin = blockIdx.y * rows_per_chunk, m_end = min(M, m_begin + rows_per_chunk);
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.0f;
    for (int mb = m_begin; mb < m_end; mb += 64) {
        const int nr = min(64, m_end - mb);
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * NMAX; i += blockDim.x) {
            const int r = i / NMAX, n = i - r * NMAX;
            es[i] = (r < nr && n < N) ? e[(long long)(mb + r) * N + n] : 0.0f;     // rows past the chunk and columns past N: zeros
        }
        __syncthreads();
        if (k <= K) {
            for (int r0 = 0; r0 < nr; r0 += 16) {
                float xv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)   // 16 independent loads in flight
                    xv[j] = (r0 + j < nr) ? (k < K ? x[(long long)(mb + r0 + j) * K + k] : 1.0f) : 0.0f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4* er = reinterpret_cast<const float4*>(&es[min(r0 + j, 63) * NMAX]);
#pragma unroll
                    for (int q = 0; q < NMAX / 4; ++q) {
                        const float4 ev = er[q];
                        acc[4 * q] = fmaf(xv[j], ev.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv[j], ev.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(xv[j], ev.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv[j], ev.w, acc[4 * q + 3]);
                    }
                }
            }
        }
    }
    if (k <= K) {
        float* dst = partial + ((long long)blockIdx.y * (K + 1) + k) * N;
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < N) dst[n] = acc[n];
    }
}

// Classifier head in ONE kernel: logits = x . w + b (N <= 32), Softmax over axis 1 (activations.py:69-71), CCE value
// (losses.py:106-109), err = p - y (models.py:200-205), argmax accuracy (metrics.py:84-89) and sum(err^2) for the first clip.
// One warp per row: lanes stride over k, then lane n owns class n.
__global__ void __launch_bounds__(256) dense_softmax_cce_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ b, const float* __restrict__ y,
                                                                float* __restrict__ p, float* __restrict__ err,
                                                                double* loss_acc, double* ss_out, int M, int N, int K,
                                                                float* __restrict__ dwp) {
    __shared__ double scratch[32];
    __shared__ float xs[8][HEAD_MAXK];
    __shared__ float es[8][SMALLN_MAX];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    float lsum = 0.0f, esq = 0.0f, correct = 0.0f;
    if (dwp != nullptr) {                                                    // rows past M contribute nothing to the partial
        es[threadIdx.x >> 5][lane] = 0.0f;
        for (int k = lane; k < K; k += 32) xs[threadIdx.x >> 5][k] = 0.0f;
    }
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
        float acc[SMALLN_MAX];
#pragma unroll
        for (int n = 0; n < SMALLN_MAX; ++n) acc[n] = 0.0f;
        const float* xr = x + (long long)m * K;
        for (int k = lane; k < K; k += 32) {
            const float xv = xr[k];
            const float* wr = w + (long long)k * N;
#pragma unroll
            for (int n = 0; n < SMALLN_MAX; ++n)
                if (n < N) acc[n] = fmaf(xv, wr[n], acc[n]);
        }
        float logit = 0.0f;
#pragma unroll
        for (int n = 0; n < SMALLN_MAX; ++n) {
            if (n < N) {
                const float t = warp_sum(acc[n]);
                if (lane == n) logit = t;
            }
        }
        const bool on = lane < N;
        logit = on ? logit + (b != nullptr ? b[lane] : 0.0f) : -INFINITY;
        const float mx = warp_max(logit);
        const float ex = on ? expf(logit - mx) : 0.0f;
        const float pv = ex / warp_sum(ex);
        const float yv = on ? y[(long long)m * N + lane] : 0.0f;
        if (on) {
            p[(long long)m * N + lane] = pv;
            const float pc = fminf(fmaxf(pv, 1e-15f), 1.0f);
            lsum -= yv * logf(pc);
            const float e = pv - yv;
            err[(long long)m * N + lane] = e;
            esq = fmaf(e, e, esq);
            if (dwp != nullptr) es[threadIdx.x >> 5][lane] = e;
        }
        if (dwp != nullptr)
            for (int k = lane; k < K; k += 32) xs[threadIdx.x >> 5][k] = xr[k];
        float bp = on ? pv : -INFINITY, by = on ? yv : -INFINITY;
        int ip = on ? lane : 0x7fffffff, iy = ip;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {                                   // first-index tie break (np.argmax)
            const float obp = __shfl_xor_sync(0xffffffffu, bp, o);
            const int oip = __shfl_xor_sync(0xffffffffu, ip, o);
            if (obp > bp || (obp == bp && oip < ip)) { bp = obp; ip = oip; }
            const float oby = __shfl_xor_sync(0xffffffffu, by, o);
            const int oiy = __shfl_xor_sync(0xffffffffu, iy, o);
            if (oby > by || (oby == by && oiy < iy)) { by = oby; iy = oiy; }
        }
        if (lane == 0 && ip == iy) correct += 1.0f;
    }
    if (dwp != nullptr) {
        // this CTA's 8 rows of the weight gradient of the head (x^T . err, + the "ones" row for db), fixed order: the backward
        // kernel of the layer only has to add the per-CTA partials (layers.py:196-197)
        __syncthreads();
        float* dst = dwp + (long long)blockIdx.x * (K + 1) * N;
        for (int o = threadIdx.x; o < (K + 1) * N; o += blockDim.x) {
            const int k = o / N, n = o % N;
            float t = 0.0f;
#pragma unroll
            for (int r = 0; r < 8; ++r) t = fmaf(k < K ? xs[r][k] : 1.0f, es[r][n], t);
            dst[o] = t;
        }
    }
    const double l = block_sum_to_double(lsum, scratch);
    if (threadIdx.x == 0 && loss_acc != nullptr) atomicAdd(&loss_acc[0], l);
    const double c = block_sum_to_double(correct, scratch);
    if (threadIdx.x == 0 && loss_acc != nullptr) atomicAdd(&loss_acc[1], c);
    if (ss_out != nullptr) block_accumulate(esq, ss_out, scratch);
}


// Large-batch form of the classifier head (config 4: 8192 rows x 4096 features x 10 classes). The one-warp-per-row kernel above
// re-reads the whole weight matrix from L2 for every row (8192 x 164 KB = 1.3 GB, 207 us); here a CTA owns 8 warps x R rows,
// stages the weights one 40 KB chunk of features at a time in shared memory (padded to an odd pitch: conflict-free for lanes striding
// over k) and every warp applies a chunk to its R rows: per 32 features R coalesced x loads + N LDS for R * N FMAs.
template <int R, int NMAX>
__global__ void __launch_bounds__(256) dense_softmax_cce_rows_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ b, const float* __restrict__ y,
                                                                     float* __restrict__ p, float* __restrict__ err,
                                                                     double* loss_acc, double* ss_out, int M, int N, int K) {
    __shared__ double scratch[32];
    constexpr int PITCH = NMAX | 1;
    constexpr int HR_KC = ((10240 / PITCH) / 32) * 32;       // 40 KB of staged weights per chunk
    constexpr int NL = NMAX <= 16 ? 16 : 32;                 // staging lanes per feature row
    __shared__ float ws[HR_KC * PITCH];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int m0 = (blockIdx.x * 8 + wid) * R;
    float acc[R][NMAX];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int n = 0; n < NMAX; ++n) acc[r][n] = 0.0f;
    const float* xr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) xr[r] = x + (long long)(m0 + r < M ? m0 + r : M - 1) * K;   // clamped: rows >= M are never stored
    for (int k0 = 0; k0 < K; k0 += HR_KC) {
        const int kc = min(HR_KC, K - k0);
        __syncthreads();
        {   // columns >= N are staged as zeros, so the FMA loop below carries no bound checks (no branches between its loads).
            // Eight loads in flight per thread: a one-load-per-iteration loop made this staging (48 dependent L2 round trips per
            // chunk) the whole kernel: 137-190 us whatever the batch (ncu: long_scoreboard 18.7 per issue).
            const int n = threadIdx.x % NL;
            constexpr int KSTEP = 256 / NL;
            for (int kb = threadIdx.x / NL; kb < kc; kb += KSTEP * 8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = kb + j * KSTEP;
                    v[j] = (k < kc && n < N) ? __ldg(w + (long long)(k0 + k) * N + n) : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = kb + j * KSTEP;
                    if (k < kc && n < NMAX) ws[k * PITCH + n] = v[j];
                }
            }
        }
        __syncthreads();
        // unrolled: 4 x R (8 x R for one or two rows per warp: small batches have one CTA per SM, and four loads per warp left the
        // kernel waiting on HBM round trips) independent 128-byte row loads in flight per warp
        constexpr int UNR = R <= 2 ? 8 : 4;
#pragma unroll UNR
        for (int kk = lane; kk < kc; kk += 32) {
            float xv[R];
#pragma unroll
            for (int r = 0; r < R; ++r) xv[r] = __ldg(xr[r] + k0 + kk);
#pragma unroll
            for (int n = 0; n < NMAX; ++n) {
                const float wv = ws[kk * PITCH + n];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r][n] = fmaf(xv[r], wv, acc[r][n]);
            }
        }
    }
    float lsum = 0.0f, esq = 0.0f, correct = 0.0f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int m = m0 + r;
        if (m >= M) break;                                                   // warp-uniform
        float logit = 0.0f;
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {
            if (n < N) {
                const float t = warp_sum(acc[r][n]);
                if (lane == n) logit = t;
            }
        }
        const bool on = lane < N;
        logit = on ? logit + (b != nullptr ? b[lane] : 0.0f) : -INFINITY;
        const float mx = warp_max(logit);
        const float ex = on ? expf(logit - mx) : 0.0f;
        const float pv = ex / warp_sum(ex);
        const float yv = on ? y[(long long)m * N + lane] : 0.0f;
        if (on) {
            p[(long long)m * N + lane] = pv;
            const float pc = fminf(fmaxf(pv, 1e-15f), 1.0f);
            lsum -= yv * logf(pc);
            const float e = pv - yv;
            err[(long long)m * N + lane] = e;
            esq = fmaf(e, e, esq);
        }
        float bp = on ? pv : -INFINITY, by = on ? yv : -INFINITY;
        int ip = on ? lane : 0x7fffffff, iy = ip;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {                                   // first-index tie break (np.argmax)
            const float obp = __shfl_xor_sync(0xffffffffu, bp, o);
            const int oip = __shfl_xor_sync(0xffffffffu, ip, o);
            if (obp > bp || (obp == bp && oip < ip)) { bp = obp; ip = oip; }
            const float oby = __shfl_xor_sync(0xffffffffu, by, o);
            const int oiy = __shfl_xor_sync(0xffffffffu, iy, o);
            if (oby > by || (oby == by && oiy < iy)) { by = oby; iy = oiy; }
        }
        if (lane == 0 && ip == iy) correct += 1.0f;
    }
    const double l = block_sum_to_double(lsum, scratch);
    if (threadIdx.x == 0 && loss_acc != nullptr) atomicAdd(&loss_acc[0], l);
    const double c = block_sum_to_double(correct, scratch);
    if (threadIdx.x == 0 && loss_acc != nullptr) atomicAdd(&loss_acc[1], c);
    if (ss_out != nullptr) block_accumulate(esq, ss_out, scratch);
}

// Large-batch dx of a small Dense: thread = one input feature k with w[k, :] in registers, CTA = 256 features x DXR rows whose
// error rows sit in shared memory (broadcast reads); the mask load and the dx store are coalesced 1 KB runs per row.
constexpr int DXR = 128;   // rows per CTA: the prologue (error rows, w[k, :] into registers) is amortised over 128 KB of traffic
template <int NMAX>
__global__ void __launch_bounds__(256) dense_smalln_dx_rows_kernel(const float* __restrict__ e, const float* __restrict__ w,
                                                                   Epilogue epi, int M, int N, int K) {
    __shared__ double scratch[32];
    __shared__ __align__(16) float es[DXR][NMAX];
    EpiState st;
    epi_begin(epi, st);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int m0 = blockIdx.y * DXR, nr = min(DXR, M - m0);
    for (int i = threadIdx.x; i < DXR * NMAX; i += blockDim.x) {
        const int r = i / NMAX, n = i - r * NMAX;
        es[r][n] = (r < nr && n < N) ? e[(long long)(m0 + r) * N + n] : 0.0f;     // zero-padded: 128-bit broadcast reads below
    }
    float wr[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) wr[n] = (k < K && n < N) ? w[(long long)k * N + n] : 0.0f;
    __syncthreads();
    if (k < K) {
        constexpr int RB = 8;                        // mask loads in flight per thread
        for (int r0 = 0; r0 < nr; r0 += RB) {
            float mk[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j)
                mk[j] = (epi.mask_y != nullptr && r0 + j < nr) ? epi.mask_y[(long long)(m0 + r0 + j) * K + k] : 0.0f;
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                if (r0 + j < nr) {
                    const float4* er = reinterpret_cast<const float4*>(es[r0 + j]);
                    float v = 0.0f;
#pragma unroll
                    for (int q = 0; q < NMAX / 4; ++q) {
                        const float4 ev = er[q];
                        v = fmaf(ev.x, wr[4 * q], v); v = fmaf(ev.y, wr[4 * q + 1], v);
                        v = fmaf(ev.z, wr[4 * q + 2], v); v = fmaf(ev.w, wr[4 * q + 3], v);
                    }
                    epi_store_premask(epi, st, m0 + r0 + j, k, v, mk[j]);
                }
            }
        }
    }
    epi_finish(epi, st, scratch);
}

// The same with a QUAD of features per thread (K % 4 == 0, aligned rows; N <= 16): 128-bit mask loads and dx stores, i.e. four times
// the bytes in flight per thread (the scalar form above keeps 24 KB per SM in flight and runs at 2 TB/s on config 4's 268 MB).
template <int NMAX>
__global__ void __launch_bounds__(256) dense_smalln_dx_rows4_kernel(const float* __restrict__ e, const float* __restrict__ w,
                                                                    Epilogue epi, int M, int N, int K, int rpc) {
    __shared__ double scratch[32];
    __shared__ __align__(16) float es[DXR][NMAX];
    EpiState st;
    epi_begin(epi, st);
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int m0 = blockIdx.y * rpc, nr = min(rpc, M - m0);      // rpc <= DXR rows per CTA (fewer at small batches: see the launch)
    for (int i = threadIdx.x; i < rpc * NMAX; i += blockDim.x) {
        const int r = i / NMAX, n = i - r * NMAX;
        es[r][n] = (r < nr && n < N) ? e[(long long)(m0 + r) * N + n] : 0.0f;
    }
    float wr[4][NMAX];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < NMAX; ++n) wr[j][n] = (k + j < K && n < N) ? w[(long long)(k + j) * N + n] : 0.0f;
    __syncthreads();
    float a0 = 0.0f, a1 = 0.0f;
    if (k < K) {
        constexpr int RB = 4;
        const bool has_mask = epi.mask_y != nullptr;
        for (int r0 = 0; r0 < nr; r0 += RB) {
            float4 mk[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j)
                mk[j] = (has_mask && r0 + j < nr) ? *reinterpret_cast<const float4*>(epi.mask_y + (long long)(m0 + r0 + j) * K + k)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                if (r0 + j < nr) {
                    const float4* er = reinterpret_cast<const float4*>(es[r0 + j]);
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int q = 0; q < NMAX / 4; ++q) {
                        const float4 ev = er[q];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            v[c] = fmaf(ev.x, wr[c][4 * q], v[c]); v[c] = fmaf(ev.y, wr[c][4 * q + 1], v[c]);
                            v[c] = fmaf(ev.z, wr[c][4 * q + 2], v[c]); v[c] = fmaf(ev.w, wr[c][4 * q + 3], v[c]);
                        }
                    }
                    const float m[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {        // epi_store_premask without bias / forward activation (a data gradient has neither)
                        v[c] *= st.scale;
                        a0 = fmaf(v[c], v[c], a0);
                        if (has_mask) {
                            v[c] *= act_grad_from_y(epi.mask_act, m[c], epi.mask_alpha);
                            a1 = fmaf(v[c], v[c], a1);
                        }
                    }
                    *reinterpret_cast<float4*>(epi.out + (long long)(m0 + r0 + j) * epi.ldo + k) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }
    st.a0 += a0; st.a1 += a1;
    epi_finish(epi, st, scratch);
}

// Backward of a small Dense (N <= 32, M <= 512) in ONE kernel: CTAs [0, dx_blocks) compute dx (+ fused activation backward
// and the two clip slots), the remaining CTAs compute dw / db directly (thread = one (k, n) output, the whole batch summed
// in fixed order: no partials, no reduce pass).
__global__ void __launch_bounds__(256) dense_smalln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ e, Epilogue epi, float* __restrict__ dw,
                                                               float* __restrict__ db, int M, int N, int K, int dx_blocks,
                                                               const float* __restrict__ dwp, int n_partial) {
    __shared__ double scratch[32];
    if ((int)blockIdx.x >= dx_blocks) {
        const float scale = chain_scale(epi.ss, epi.chain);
        const int o = ((int)blockIdx.x - dx_blocks) * blockDim.x + threadIdx.x;   // output index over (K + 1) x N
        if (o >= (K + 1) * N) return;
        const int k = o / N, n = o % N;
        if (dwp != nullptr) {                            // the fused head already formed per-CTA partials: add them in order
            float t = 0.0f;
#pragma unroll 8
            for (int c = 0; c < n_partial; ++c) t += dwp[(long long)c * (K + 1) * N + o];
            t *= scale;
            if (k < K) { if (dw != nullptr) dw[o] = t; }
            else if (db != nullptr) db[n] = t;
            return;
        }
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        int m = 0;
        if (k < K) {
            for (; m + 4 <= M; m += 4) {                                    // 8 independent loads in flight
                a0 = fmaf(x[(long long)m * K + k], e[(long long)m * N + n], a0);
                a1 = fmaf(x[(long long)(m + 1) * K + k], e[(long long)(m + 1) * N + n], a1);
                a2 = fmaf(x[(long long)(m + 2) * K + k], e[(long long)(m + 2) * N + n], a2);
                a3 = fmaf(x[(long long)(m + 3) * K + k], e[(long long)(m + 3) * N + n], a3);
            }
            for (; m < M; ++m) a0 = fmaf(x[(long long)m * K + k], e[(long long)m * N + n], a0);
            if (dw != nullptr) dw[(long long)k * N + n] = scale * ((a0 + a1) + (a2 + a3));   // layers.py:196
        } else {
            for (; m + 4 <= M; m += 4) {
                a0 += e[(long long)m * N + n]; a1 += e[(long long)(m + 1) * N + n];
                a2 += e[(long long)(m + 2) * N + n]; a3 += e[(long long)(m + 3) * N + n];
            }
            for (; m < M; ++m) a0 += e[(long long)m * N + n];
            if (db != nullptr) db[n] = scale * ((a0 + a1) + (a2 + a3));                       // layers.py:197
        }
        return;
    }
    EpiState st;
    epi_begin(epi, st);
    const long long total = (long long)M * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)dx_blocks * blockDim.x) {
        const int m = (int)(i / K), k = (int)(i % K);
        const float* er = e + (long long)m * N;
        const float* wr = w + (long long)k * N;
        float v = 0.0f;
        for (int n = 0; n < N; ++n) v = fmaf(er[n], wr[n], v);
        const float mk = epi.mask_y != nullptr ? epi.mask_y[i] : 0.0f;
        epi_store_premask(epi, st, m, k, v, mk);
    }
    epi_finish(epi, st, scratch);
}

static int smalln_chunks(int M) {
    int c = (M + 7) / 8;                // >= 8 rows per chunk
    return c > 64 ? 64 : (c < 1 ? 1 : c);
}
static size_t smalln_ws_bytes(int M, int N, int K) { return (size_t)smalln_chunks(M) * (K + 1) * N * sizeof(float); }

}  // namespace b200nn
#include "dense_skinny.cuh"   // needs Epilogue / launch_splitk_reduce; lives in b200nn::skinny
namespace b200nn {

// ------------------------------------------------------------------------------------------------
// host-side launch helpers
// ------------------------------------------------------------------------------------------------
static int choose_splits(long long M, long long N, long long K) {
    const long long tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    if (tiles >= 2LL * kNumSMs || K < 512) return 1;   // a short reduction is cheaper than the extra reduce pass
    long long want = (2LL * kNumSMs + tiles - 1) / tiles;
    long long maxs = K / 128;
    if (maxs < 1) maxs = 1;
    long long s = want < maxs ? want : maxs;
    if (s > 1024) s = 1024;
    return (int)(s < 1 ? 1 : s);
}

static size_t ws_bytes_for(long long M, long long N, long long K) {
    const int s = choose_splits(M, N, K);
    return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

template <class AF, class BF>
static int launch_gemm(const char* name, const AF& af, const BF& bf, const Epilogue& epi, int M, int N, int K,
                       float* ws, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return B200NN_OK;
    const int splits = choose_splits(M, N, K);
    const dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
    if (splits == 1) {
        gemm_simt_kernel<AF, BF><<<grid, 256, 0, stream>>>(af, bf, epi, M, N, K, K > 0 ? K : 1, nullptr);
        B200NN_LAUNCH_CHECK(name);
        return B200NN_OK;
    }
    B200NN_REQUIRE(ws != nullptr, "%s: split-K needs a workspace (M=%d N=%d K=%d)", name, M, N, K);
    int k_chunk = (K + splits - 1) / splits;
    k_chunk = (k_chunk + BK - 1) / BK * BK;
    gemm_simt_kernel<AF, BF><<<grid, 256, 0, stream>>>(af, bf, epi, M, N, K, k_chunk, ws);
    B200NN_LAUNCH_CHECK(name);
    return launch_splitk_reduce(ws, splits, M, N, epi, stream);
}

static int check_act(const char* who, int act, float alpha) {
    B200NN_REQUIRE(act >= 0 && act <= B200NN_ACT_MAX_FUSED, "%s: unknown activation %d", who, act);
    B200NN_REQUIRE(act != B200NN_ACT_LEAKYRELU || alpha > 0.0f, "%s: LeakyReLU needs alpha > 0", who);
    return B200NN_OK;
}

// the tensor-core contraction: the persistent CTA-pair kernel (gemm_tc2.cuh) unless B200NN_TC2=0 selects the first-generation one
// premax_a / premax_b (optional): device words that already hold the bit pattern of max |A| / max |B| (see publish_absmax)
static int tc_gemm_any(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi,
                       int M, int N, int K, float* ws, cudaStream_t st, bool x3, const unsigned* premax_a = nullptr,
                       const unsigned* premax_b = nullptr) {
    // large error-compensated contractions: fp16 hi / lo planes at twice the TF32 rate (f16_split.cu)
    if (x3 && ws != nullptr && tcapi::f16_enabled() && tcapi::f16_worth(M, N, K))
        return tcapi::gemm2_f16x3(A, lda, a_mn, B, ldb, b_mn, epi, M, N, K, ws, st, premax_a, premax_b);
    if (tcapi::tc2_enabled())
        return x3 ? tcapi::gemm2_x3(A, lda, a_mn, B, ldb, b_mn, epi, M, N, K, ws, st)
                  : tcapi::gemm2_plain(A, lda, a_mn, B, ldb, b_mn, epi, M, N, K, ws, st);
    return tcapi::gemm1(A, lda, a_mn, B, ldb, b_mn, epi, M, N, K, ws, st, x3);
}
// which: 1 forward (a = x, b = w), 2 data gradient (a = err, b = w), 3 weight gradient (a = x, b = err)
static int tc_conv_any(int which, const b200nn_conv_geom* g, const float* a, const float* b, const Epilogue& epi, float* ws,
                       cudaStream_t st, bool x3, const unsigned* premax_a = nullptr) {
    // large error-compensated forward / data-gradient convolutions: fp16 hi / lo planes at twice the TF32 rate (f16_split.cu)
    if (x3 && ws != nullptr && tcapi::tc2_enabled() && tcapi::conv_f16_worth(which, g))
        return tcapi::conv2_f16x3(which, g, a, b, epi, ws, st, premax_a);
    if (tcapi::tc2_enabled()) return x3 ? tcapi::conv2_x3(which, g, a, b, epi, ws, st) : tcapi::conv2_plain(which, g, a, b, epi, ws, st);
    return tcapi::conv1(which, g, a, b, epi, ws, st, x3);
}

static void dx_epilogue(Epilogue& e, const double* ss, uint64_t chain, const float* mask_y, int prev_act,
                        float prev_alpha, double* ss0, double* ss1) {
    e.ss = ss; e.chain = chain; e.ss0 = ss0;
    if (prev_act != B200NN_ACT_NONE) {
        e.mask_y = mask_y; e.mask_act = prev_act; e.mask_alpha = prev_alpha; e.ss1 = ss1;
    }
}

}  // namespace b200nn

using namespace b200nn;

extern "C" {

// ---------------------------------------------------------------------------------------------- Dense
size_t b200nn_dense_ws_bytes_math(int math, int M, int N, int K) {
    size_t r = b200nn_dense_ws_bytes(M, N, K);
    if (math == B200NN_MATH_TF32X3) {             // chain-limited split-K partials of the compensated tensor-core path
        auto mx = [&](size_t v) { if (v > r) r = v; };
        mx(tcapi::ws1(M, N, K, true));
        mx(tcapi::ws1(M, K, N, true));
        mx(tcapi::ws1(K, N, M, true));
        mx(tcapi::ws2(M, N, K, true));
        mx(tcapi::ws2(M, K, N, true));
        mx(tcapi::ws2(K, N, M, true));
        auto f16 = [&](long long m, long long n, long long k) {       // partials + operand planes of the fp16-plane path
            if (tcapi::f16_enabled() && tcapi::f16_worth(m, n, k))
                mx((tcapi::ws2(m, n, (k + 1) / 2, true) + 255) / 256 * 256 + tcapi::f16_extra_bytes(m, n, k));
        };
        f16(M, N, K); f16(M, K, N); f16(K, N, M);
        r = (r + 255) / 256 * 256 + 256;      // the last 256 bytes: max word of the error matrix (publish_absmax -> operand planes)
    }
    return r;
}
// where b200nn_dense_bwd keeps max |err| between its column-sum pass and the fp16-plane contractions: the tail of the workspace the
// sizing call above promises (every other user of the workspace stays below it)
static unsigned* dense_max_word(int math, int M, int N, int K, float* ws) {
    if (math != B200NN_MATH_TF32X3 || ws == nullptr) return nullptr;
    return reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + b200nn_dense_ws_bytes_math(math, M, N, K) - 256);
}

size_t b200nn_dense_ws_bytes(int M, int N, int K) {
    size_t r = 0;
    auto mx = [&](size_t v) { if (v > r) r = v; };
    mx(ws_bytes_for(M, N, K));                    // fwd  [M,N] over K
    mx(ws_bytes_for(M, K, N));                    // dx   [M,K] over N
    mx(ws_bytes_for((long long)K + 1, N, M));     // dw+db [K+1,N] over M
    mx(tcapi::ws1(M, N, K));                 // the same three on the tcgen05 path
    mx(tcapi::ws1(M, K, N));
    mx(tcapi::ws1(K, N, M));
    mx(tcapi::ws2(M, N, K, false));
    mx(tcapi::ws2(M, K, N, false));
    mx(tcapi::ws2(K, N, M, false));
    mx((size_t)kNumSMs * 2 * N * sizeof(float));  // column-sum partials for db on the tcgen05 path
    if (N <= SMALLN_MAX) mx(smalln_ws_bytes(M, N, K));
    if (M <= skinny::MAXM && N <= skinny::MAXN) mx(skinny::fwd_ws_bytes(M, N, K));
    return r;
}

static bool dense_tc_ok(const float* x, const float* w, const float* other, int M, int N, int K) {
    // TMA needs 16-byte aligned bases and row pitches; tiny problems stay on the FFMA kernel
    // below ~8 M multiply-adds the persistent kernel's fixed cost (cluster launch, TMEM allocation, tensor-map fetch) exceeds the
    // FFMA kernel's whole run time (config 1's 32 x 784 x 128 layer: 3.2 M)
    return tcapi::tma_ok(x, K) && tcapi::tma_ok(w, N) && tcapi::tma_ok(other, N) && (long long)M * N * K >= (1LL << 23);
}

int b200nn_dense_fwd(int math, const float* x, const float* w, const float* b, float* y, int M, int N, int K,
                     int act, float alpha, float* ws, void* stream) {
    B200NN_REQUIRE(math >= B200NN_MATH_FP32 && math <= B200NN_MATH_TF32X3, "dense_fwd: unknown math mode %d", math);
    B200NN_REQUIRE(M > 0 && N > 0 && K > 0, "dense_fwd: bad shape M=%d N=%d K=%d", M, N, K);
    if (int rc = check_act("dense_fwd", act, alpha)) return rc;
    if (N <= SMALLN_MAX) {
        int blocks = (M + 7) / 8;
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        dense_smalln_fwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, w, b, y, M, N, K, act, alpha);
        B200NN_LAUNCH_CHECK("dense_smalln_fwd");
        return B200NN_OK;
    }
    Epilogue e = plain_epilogue(y, N, M);
    e.bias = b; e.act = act; e.alpha = alpha;
    // K >> M, N (config 2's Dense(64) on 6272 features): the slab kernels beat a tensor-core tile that is mostly padding, and they
    // are exact fp32 — so they serve every math mode except single-pass tf32 (where the caller asked for the tensor cores)
    if (math != B200NN_MATH_TF32 && ws != nullptr && skinny::applicable(x, w, y, M, N, K))
        return skinny::launch_fwd(x, w, e, M, N, K, ws, as_stream(stream));
    if (math != B200NN_MATH_FP32 && dense_tc_ok(x, w, y, M, N, K))
        return tc_gemm_any(x, K, false, w, N, true, e, M, N, K, ws, as_stream(stream), math == B200NN_MATH_TF32X3);
    RowMajorA af{x, K, M, K};
    RowMajorB bf{w, N, K, N};
    return launch_gemm("dense_fwd", af, bf, e, M, N, K, ws, as_stream(stream));
}

int b200nn_dense_softmax_cce(const float* x, const float* w, const float* b, const float* y_true, float* p, float* err,
                             double* loss_acc, double* ss_out, int M, int N, int K, float* dw_partial, void* stream) {
    B200NN_REQUIRE(M > 0 && K > 0 && N > 0 && N <= SMALLN_MAX, "dense_softmax_cce: needs N <= %d (M=%d N=%d K=%d)", SMALLN_MAX, M, N, K);
    B200NN_REQUIRE(x != nullptr && w != nullptr && y_true != nullptr && p != nullptr && err != nullptr, "dense_softmax_cce: null buffer");
    if (dw_partial == nullptr && M >= 1024 && K >= 256) {   // large batch: weights staged per CTA, several rows per warp
        cudaStream_t st = as_stream(stream);
        // rows per warp: as many as keep >= one CTA (8 warps) per SM
        const int R = M >= 8 * 8 * kNumSMs ? 8 : M >= 8 * 4 * kNumSMs ? 4 : M >= 8 * 2 * kNumSMs ? 2 : 1;
#define B200NN_HEAD_ROWS(RR, NN) dense_softmax_cce_rows_kernel<RR, NN><<<(M + 8 * RR - 1) / (8 * RR), 256, 0, st>>>(x, w, b, y_true, p, err, loss_acc, ss_out, M, N, K)
        if (N <= 12) { if (R == 8) B200NN_HEAD_ROWS(8, 12); else if (R == 4) B200NN_HEAD_ROWS(4, 12); else if (R == 2) B200NN_HEAD_ROWS(2, 12); else B200NN_HEAD_ROWS(1, 12); }
        else if (N <= 16) { if (R >= 4) B200NN_HEAD_ROWS(4, 16); else if (R == 2) B200NN_HEAD_ROWS(2, 16); else B200NN_HEAD_ROWS(1, 16); }
        else { if (R >= 2) B200NN_HEAD_ROWS(2, 32); else B200NN_HEAD_ROWS(1, 32); }
#undef B200NN_HEAD_ROWS
        B200NN_LAUNCH_CHECK("dense_softmax_cce_rows");
        return B200NN_OK;
    }
    int blocks = (M + 7) / 8;
    B200NN_REQUIRE(dw_partial == nullptr || (K <= HEAD_MAXK && blocks <= kNumSMs * 8), "dense_softmax_cce: partials need K <= %d", HEAD_MAXK);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    dense_softmax_cce_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, w, b, y_true, p, err, loss_acc, ss_out, M, N, K, dw_partial);
    B200NN_LAUNCH_CHECK("dense_softmax_cce");
    return B200NN_OK;
}

int b200nn_dense_head_partials(int M, int K) { return (K <= HEAD_MAXK && M <= 512) ? (M + 7) / 8 : 0; }

int b200nn_dense_bwd_head(const float* x, const float* w, const float* err, const double* ss, uint64_t chain, float* dx, float* dw,
                          float* db, int M, int N, int K, int prev_act, float prev_alpha, double* ss_out0, double* ss_out1,
                          const float* dw_partial, void* stream) {
    B200NN_REQUIRE(M > 0 && K > 0 && N > 0 && N <= SMALLN_MAX && M <= 512 && K <= HEAD_MAXK && dw_partial != nullptr && dw != nullptr,
                   "dense_bwd_head: unsupported shape M=%d N=%d K=%d", M, N, K);
    if (int rc = check_act("dense_bwd_head", prev_act, prev_alpha)) return rc;
    long long dx_blocks = dx != nullptr ? ((long long)M * K + 255) / 256 : 0;
    if (dx_blocks > kNumSMs * 8) dx_blocks = kNumSMs * 8;
    const int dw_blocks = ((K + 1) * N + 255) / 256;
    Epilogue e = plain_epilogue(dx, K, M);
    e.ss = ss; e.chain = chain;
    if (dx != nullptr) dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
    dense_smalln_bwd_kernel<<<(int)dx_blocks + dw_blocks, 256, 0, as_stream(stream)>>>(x, w, err, e, dw, db, M, N, K, (int)dx_blocks,
                                                                                     dw_partial, (M + 7) / 8);
    B200NN_LAUNCH_CHECK("dense_bwd_head");
    return B200NN_OK;
}

int b200nn_dense_bwd(int math, const float* x, const float* w, const float* err, const double* ss, uint64_t chain,
                     float* dx, float* dw, float* db, int M, int N, int K, int prev_act, float prev_alpha,
                     double* ss_out0, double* ss_out1, float* ws, void* stream) {
    B200NN_REQUIRE(math >= B200NN_MATH_FP32 && math <= B200NN_MATH_TF32X3, "dense_bwd: unknown math mode %d", math);
    B200NN_REQUIRE(M > 0 && N > 0 && K > 0, "dense_bwd: bad shape M=%d N=%d K=%d", M, N, K);
    if (int rc = check_act("dense_bwd", prev_act, prev_alpha)) return rc;
    cudaStream_t st = as_stream(stream);
    if (N <= SMALLN_MAX && M <= 512) {
        long long dx_blocks = dx != nullptr ? ((long long)M * K + 255) / 256 : 0;
        if (dx_blocks > kNumSMs * 8) dx_blocks = kNumSMs * 8;
        const int dw_blocks = dw != nullptr ? ((K + 1) * N + 255) / 256 : 0;
        Epilogue e = plain_epilogue(dx, K, M);
        e.ss = ss; e.chain = chain;
        if (dx != nullptr) dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
        if (dx_blocks + dw_blocks > 0) {
            dense_smalln_bwd_kernel<<<(int)dx_blocks + dw_blocks, 256, 0, st>>>(x, w, err, e, dw, db, M, N, K, (int)dx_blocks, nullptr, 0);
            B200NN_LAUNCH_CHECK("dense_smalln_bwd");
        }
        return B200NN_OK;
    }
    if (N <= SMALLN_MAX && ws != nullptr) {
        if (dw != nullptr) {
            const int chunks = smalln_chunks(M);
            const int rpc = (M + chunks - 1) / chunks;
            const dim3 grid((K + 1 + 255) / 256, chunks);
            if (N <= 12) dense_smalln_dw_kernel<12><<<grid, 256, 0, st>>>(x, err, ws, M, N, K, rpc);
            else if (N <= 16) dense_smalln_dw_kernel<16><<<grid, 256, 0, st>>>(x, err, ws, M, N, K, rpc);
            else dense_smalln_dw_kernel<32><<<grid, 256, 0, st>>>(x, err, ws, M, N, K, rpc);
            B200NN_LAUNCH_CHECK("dense_smalln_dw");
            Epilogue e = plain_epilogue(dw, N, K);
            e.ss = ss; e.chain = chain; e.db = db;
            if (int rc = launch_splitk_reduce(ws, chunks, K + 1, N, e, st)) return rc;
        }
        if (dx != nullptr) {
            Epilogue e = plain_epilogue(dx, K, M);
            dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
            if (M >= 256 && K >= 256) {
                const dim3 grid((K + 255) / 256, (M + DXR - 1) / DXR);
                const bool quad = (K & 3) == 0 && N <= 16 && e.bias == nullptr && e.act == B200NN_ACT_NONE && (e.ldo & 3) == 0 &&
                                  ((reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(e.mask_y)) & 15) == 0;
                if (quad) {
                    // 128 rows per CTA amortise the prologue at config 4's batch; a small batch (config 3: 1024 rows -> 32 CTAs) gets
                    // fewer rows per CTA so that two CTAs per SM are in flight
                    const int gx = (K / 4 + 255) / 256;
                    int rpc = DXR;
                    if ((long long)gx * ((M + DXR - 1) / DXR) < 2 * kNumSMs) {
                        const int want = (2 * kNumSMs + gx - 1) / gx;                 // row blocks wanted
                        rpc = ((M + want - 1) / want + 15) / 16 * 16;
                        rpc = rpc < 16 ? 16 : (rpc > DXR ? DXR : rpc);
                    }
                    const dim3 grid4(gx, (M + rpc - 1) / rpc);
                    if (N <= 12) dense_smalln_dx_rows4_kernel<12><<<grid4, 256, 0, st>>>(err, w, e, M, N, K, rpc);
                    else dense_smalln_dx_rows4_kernel<16><<<grid4, 256, 0, st>>>(err, w, e, M, N, K, rpc);
                } else if (N <= 12) dense_smalln_dx_rows_kernel<12><<<grid, 256, 0, st>>>(err, w, e, M, N, K);
                else if (N <= 16) dense_smalln_dx_rows_kernel<16><<<grid, 256, 0, st>>>(err, w, e, M, N, K);
                else dense_smalln_dx_rows_kernel<32><<<grid, 256, 0, st>>>(err, w, e, M, N, K);
                B200NN_LAUNCH_CHECK("dense_smalln_dx_rows");
            } else {
                long long blocks = ((long long)M * K + 255) / 256;
                if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
                dense_smalln_dx_kernel<<<(int)blocks, 256, 0, st>>>(err, w, e, M, N, K);
                B200NN_LAUNCH_CHECK("dense_smalln_dx");
            }
        }
        return B200NN_OK;
    }
    if (math != B200NN_MATH_TF32 && skinny::applicable(x, w, err, M, N, K) &&
        ((reinterpret_cast<uintptr_t>(dw) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0)
        return skinny::launch_bwd(x, w, err, ss, chain, dx, dw, db, M, N, K, prev_act, prev_alpha, ss_out0, ss_out1, st);
    const bool use_tc = math != B200NN_MATH_FP32 && dense_tc_ok(x, w, err, M, N, K) && ws != nullptr;
    const bool x3 = math == B200NN_MATH_TF32X3;
    // max |err| for the fp16 operand planes of the weight / data gradient, taken by the column-sum pass that reads err anyway
    unsigned* errmax = nullptr;
    if (dw != nullptr && use_tc) {
        if (db != nullptr) {   // layers.py:197
            int nb = (M + 7) / 8;   // >= 8 rows per CTA; threads run over the columns
            if (nb > kNumSMs * 2) nb = kNumSMs * 2;
            const bool quad = (N & 3) == 0 && ((reinterpret_cast<uintptr_t>(err) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0;
            if (x3 && quad && tcapi::f16_enabled() && (tcapi::f16_worth(K, N, M) || (dx != nullptr && tcapi::f16_worth(M, K, N)))) {
                errmax = dense_max_word(math, M, N, K, ws);
                B200NN_CUDA(cudaMemsetAsync(errmax, 0, sizeof(unsigned), st));
            }
            if (colsum4r_ok(err, ws, N))
                colsum_partial4r_kernel<<<nb, 256, 0, st>>>(err, M, N, ws, errmax);
            else if (quad)
                colsum_partial4_kernel<<<nb, 256, 0, st>>>(err, M, N, ws, errmax);
            else
                colsum_partial_kernel<<<nb, 256, 0, st>>>(err, M, N, ws);
            B200NN_LAUNCH_CHECK("dense_db_partial");
            colsum_final_kernel<<<(N + 31) / 32, 256, 0, st>>>(ws, nb, N, ss, chain, db);
            B200NN_LAUNCH_CHECK("dense_db_final");
        }
        Epilogue e = plain_epilogue(dw, N, K);   // dw[K,N] = x^T . err : both operands MN-major, no transposed copies
        e.ss = ss; e.chain = chain;
        if (int rc = tc_gemm_any(x, K, true, err, N, true, e, K, N, M, ws, st, x3, nullptr, errmax)) return rc;
    } else if (dw != nullptr) {  // dw[K,N] (+ db row) = x^T[K(+1), M] . err[M,N]
        ColMajorOnesA af{x, K, K, M, db != nullptr ? 1 : 0};
        RowMajorB bf{err, N, M, N};
        Epilogue e = plain_epilogue(dw, N, K);
        e.ss = ss; e.chain = chain; e.db = db;
        if (int rc = launch_gemm("dense_dw", af, bf, e, K + (db != nullptr ? 1 : 0), N, M, ws, st)) return rc;
    }
    if (dx != nullptr) {  // dx[M,K] = err[M,N] . w^T
        Epilogue e = plain_epilogue(dx, K, M);
        dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
        if (use_tc) {
            if (int rc = tc_gemm_any(err, N, false, w, N, false, e, M, K, N, ws, st, x3, errmax, nullptr)) return rc;
        } else {
            RowMajorA af{err, N, M, N};
            ColMajorB bf{w, N, N, K};
            if (int rc = launch_gemm("dense_dx", af, bf, e, M, K, N, ws, st)) return rc;
        }
    }
    return B200NN_OK;
}

// ---------------------------------------------------------------------------------------------- Conv2D

// ------------------------------------------------------------------------------------------------
// Convolutions the TMA implicit-GEMM path does not cover (strides, 'valid', any image size, any filter count incl. F = 1) in the
// tensor-core math modes: im2col into a column buffer, then the Dense kernels — col[(n,oy,ox), (ky,kx,c)] . Wp[(ky,kx,c), f] —
// i.e. tcgen05 (TF32 / 3xTF32), the skinny kernels, or the small-N kernels, whichever the shape selects (replaces
// preprocessing.py:34-126 for those geometries). The column order (ky,kx,c) keeps both the gather and the column stores full
// 16-byte accesses; the weights are permuted once per call from the reference's (c,ky,kx) order (Q1) and the weight gradient back.
__global__ void __launch_bounds__(256) conv_im2col_kernel(b200nn_conv_geom g, const float* __restrict__ x, float* __restrict__ col,
                                                          long long total4) {
    const int C4 = g.C >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4; long long t = i / C4;
        const int kx = (int)(t % g.kw); t /= g.kw;
        const int ky = (int)(t % g.kh); t /= g.kh;
        const int ox = (int)(t % g.OW); t /= g.OW;
        const int oy = (int)(t % g.OH); const int n = (int)(t / g.OH);
        const int iy = oy * g.sh - g.ph + ky, ix = ox * g.sw - g.pw + kx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = *reinterpret_cast<const float4*>(x + (((long long)n * g.H + iy) * g.W + ix) * g.C + c);
        *reinterpret_cast<float4*>(col + i * 4) = v;
    }
}

// to_ref == 0: wp[(ky,kx,c), f] = w_flat[(c,ky,kx), f];  to_ref == 1: the inverse (weight gradient back to the reference order)
__global__ void __launch_bounds__(256) conv_weight_permute_kernel(const float* __restrict__ src, float* __restrict__ dst, int kh, int kw,
                                                                  int C, int F, int to_ref) {
    const int total = kh * kw * C * F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int f = i % F; int t = i / F;
        const int c = t % C; t /= C;
        const int kx = t % kw, ky = t / kw;                       // i indexes the (ky,kx,c) order
        const int r = ((c * kh + ky) * kw + kx) * F + f;          // the same element in the reference order
        if (to_ref) dst[r] = src[i]; else dst[i] = src[r];
    }
}

// dx[n,iy,ix,c] = sum over the taps that reach (iy,ix) of dcol[(n,oy,ox), (ky,kx,c)]  (col2im, preprocessing.py:82-126), then the
// preceding activation's backward and the two sum-of-squares slots
__global__ void __launch_bounds__(256) conv_col2im_kernel(b200nn_conv_geom g, const float* __restrict__ dcol, const float* __restrict__ mask_y,
                                                          int mask_act, float mask_alpha, float* __restrict__ dx, double* ss0, double* ss1,
                                                          long long total4) {
    __shared__ double scratch[32];
    const int C4 = g.C >> 2, K = g.kh * g.kw * g.C;
    float a0 = 0.0f, a1 = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4; long long t = i / C4;
        const int ix = (int)(t % g.W); t /= g.W;
        const int iy = (int)(t % g.H); const int n = (int)(t / g.H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ky = 0; ky < g.kh; ++ky) {
            const int ty = iy + g.ph - ky;
            if (ty < 0 || ty % g.sh != 0) continue;
            const int oy = ty / g.sh;
            if (oy >= g.OH) continue;
            for (int kx = 0; kx < g.kw; ++kx) {
                const int tx = ix + g.pw - kx;
                if (tx < 0 || tx % g.sw != 0) continue;
                const int ox = tx / g.sw;
                if (ox >= g.OW) continue;
                const float4 v = *reinterpret_cast<const float4*>(dcol + (((long long)n * g.OH + oy) * g.OW + ox) * K +
                                                                  (long long)(ky * g.kw + kx) * g.C + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        a0 += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
        if (mask_y != nullptr) {
            const float4 m = *reinterpret_cast<const float4*>(mask_y + i * 4);
            acc.x *= act_grad_from_y(mask_act, m.x, mask_alpha); acc.y *= act_grad_from_y(mask_act, m.y, mask_alpha);
            acc.z *= act_grad_from_y(mask_act, m.z, mask_alpha); acc.w *= act_grad_from_y(mask_act, m.w, mask_alpha);
            a1 += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
        }
        *reinterpret_cast<float4*>(dx + i * 4) = acc;
    }
    if (ss0 != nullptr) block_accumulate(a0, ss0, scratch);
    if (mask_y != nullptr && ss1 != nullptr) block_accumulate(a1, ss1, scratch);
}

constexpr long long CONV_COL_MAX_BYTES = 1536LL << 20;   // per column buffer
static bool conv_col_ok(const b200nn_conv_geom* g) {
    const long long M = (long long)g->N * g->OH * g->OW, K = (long long)g->kh * g->kw * g->C;
    return g->C % 4 == 0 && K >= 32 && M * K * 4 <= CONV_COL_MAX_BYTES && M * K * g->F >= (1LL << 20) && M < (1LL << 31) / 2;
}
struct ConvColWs {
    float *col, *dcol, *wp, *dwp, *rest;
};
static size_t align64(size_t n) { return (n + 63) / 64 * 64; }
static ConvColWs conv_col_ws(const b200nn_conv_geom* g, float* ws) {
    const size_t MK = align64((size_t)g->N * g->OH * g->OW * g->kh * g->kw * g->C), KF = align64((size_t)g->kh * g->kw * g->C * g->F);
    ConvColWs r;
    r.col = ws; r.dcol = ws + MK; r.wp = r.dcol + MK; r.dwp = r.wp + KF; r.rest = r.dwp + KF;
    return r;
}


}  // extern "C" (templates below)

// ------------------------------------------------------------------------------------------------
// Convolutions with a handful of filters (F <= 4: the generator's output layer, 28x28x32 -> 1) or a handful of input channels on
// the data-gradient side (C <= 4: the discriminator's first layer). As GEMMs they have N <= 4, i.e. a tile that is almost all
// padding; they are bandwidth problems (read the wide tensor once) and get direct kernels. Weights in the reference's effective
// order W[c,ky,kx,f] = flat[((c*kh+ky)*kw+kx)*F + f] (Q1), staged in shared memory.
constexpr int FEWF_MAX = 4;
constexpr int FEW_SMEM_FLOATS = 8192;

// y[n,oy,ox,f] = act(b[f] + sum_{ky,kx,c} x[n,iy,ix,c] * W[c,ky,kx,f]); thread = output pixel, float4 over c
__global__ void __launch_bounds__(256) conv_fewf_fwd_kernel(b200nn_conv_geom g, const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, int act, float alpha,
                                                            long long n_pix) {
    __shared__ float ws[FEW_SMEM_FLOATS];            // [(ky,kx)][c][f]
    const int K = g.kh * g.kw * g.C, F = g.F;
    for (int i = threadIdx.x; i < K * F; i += blockDim.x) {
        const int f = i % F; int t = i / F;
        const int c = t % g.C, tap = t / g.C;
        ws[i] = w[((long long)(c * g.kh + tap / g.kw) * g.kw + tap % g.kw) * F + f];
    }
    __syncthreads();
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(p % g.OW); long long t = p / g.OW;
        const int oy = (int)(t % g.OH); const int n = (int)(t / g.OH);
        float acc[FEWF_MAX];
#pragma unroll
        for (int f = 0; f < FEWF_MAX; ++f) acc[f] = (bias != nullptr && f < F) ? bias[f] : 0.0f;
        for (int ky = 0; ky < g.kh; ++ky) {
            const int iy = oy * g.sh - g.ph + ky;
            if (iy < 0 || iy >= g.H) continue;
            for (int kx = 0; kx < g.kw; ++kx) {
                const int ix = ox * g.sw - g.pw + kx;
                if (ix < 0 || ix >= g.W) continue;
                const float4* xp = reinterpret_cast<const float4*>(x + (((long long)n * g.H + iy) * g.W + ix) * g.C);
                const float* wt = ws + (ky * g.kw + kx) * g.C * F;
                for (int c4 = 0; c4 < g.C / 4; ++c4) {
                    const float4 v = xp[c4];
#pragma unroll
                    for (int f = 0; f < FEWF_MAX; ++f)
                        if (f < F) {
                            acc[f] = fmaf(v.x, wt[(c4 * 4 + 0) * F + f], acc[f]); acc[f] = fmaf(v.y, wt[(c4 * 4 + 1) * F + f], acc[f]);
                            acc[f] = fmaf(v.z, wt[(c4 * 4 + 2) * F + f], acc[f]); acc[f] = fmaf(v.w, wt[(c4 * 4 + 3) * F + f], acc[f]);
                        }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < FEWF_MAX; ++f)
            if (f < F) y[p * F + f] = act_apply(act, acc[f], alpha);
    }
}

// The same with lane <-> CHANNEL QUAD (C / 4 a power of two <= 32, 3 x 3 taps): the C / 4 lanes of a pixel read its 4 C bytes as one
// contiguous run, so a warp load covers whole lines (the thread-per-pixel form above reads 16 bytes of 32 different lines per
// instruction: 132 us per launch for the generator's 28x28x32 -> 1 output layer at batch 512); the lane's 9 x 4 x F weights sit in
// registers, all nine tap loads of a pixel are in flight together, the lanes of a pixel are folded with log2(C / 4) shuffles.
template <int FM>
__global__ void __launch_bounds__(256) conv_fewf_fwd2_kernel(b200nn_conv_geom g, const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y, int act, float alpha,
                                                             long long n_pix) {
    const int C4 = g.C >> 2, F = g.F;
    const int cq = threadIdx.x % C4;                       // channel quad of this lane
    const int ppw = 32 / C4;                               // pixels per warp
    float wr[9][4][FM];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < FM; ++f)
                wr[tap][j][f] = f < F ? w[((long long)((cq * 4 + j) * g.kh + tap / 3) * g.kw + tap % 3) * F + f] : 0.0f;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int sub = (threadIdx.x & 31) / C4;               // which of the warp's pixels
    for (long long p0 = warp * ppw; p0 < n_pix; p0 += n_warps * ppw) {
        const long long p = p0 + sub;
        const bool pok = p < n_pix;
        const int ox = (int)(p % g.OW); long long t = p / g.OW;
        const int oy = (int)(t % g.OH); const int n = (int)(t / g.OH);
        float4 v[9];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int iy = oy * g.sh - g.ph + tap / 3, ix = ox * g.sw - g.pw + tap % 3;
            v[tap] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pok && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
                v[tap] = *reinterpret_cast<const float4*>(x + (((long long)n * g.H + iy) * g.W + ix) * g.C + cq * 4);
        }
        float acc[FM];
#pragma unroll
        for (int f = 0; f < FM; ++f) acc[f] = 0.0f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int f = 0; f < FM; ++f) {
                acc[f] = fmaf(v[tap].x, wr[tap][0][f], acc[f]); acc[f] = fmaf(v[tap].y, wr[tap][1][f], acc[f]);
                acc[f] = fmaf(v[tap].z, wr[tap][2][f], acc[f]); acc[f] = fmaf(v[tap].w, wr[tap][3][f], acc[f]);
            }
#pragma unroll
        for (int f = 0; f < FM; ++f)
            for (int o = 1; o < C4; o <<= 1) acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], o);
        if (pok && cq == 0) {
#pragma unroll
            for (int f = 0; f < FM; ++f)
                if (f < F) y[p * F + f] = act_apply(act, acc[f] + (bias != nullptr ? bias[f] : 0.0f), alpha);
        }
    }
}
// ---- band kernels: 3 x 3, stride 1, 'same' (pad 1), F <= 2, C / 4 a power of two <= 32 (the generator's output layer) -------------
// A CTA stages a band of FB_RB output rows of one image plus its one-row / one-column halo in shared memory, so the nine taps of a
// pixel are shared-memory reads: HBM / L2 see x 1.5 times instead of 9 (the lane <-> channel-quad form above is L2-bound at 9x).
constexpr int FB_RB = 4;

template <int FM>
__global__ void __launch_bounds__(256) conv_fewf_band_fwd_kernel(b200nn_conv_geom g, const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y, int act, float alpha) {
    extern __shared__ __align__(16) float fb_tile[];       // [(FB_RB + 2)][(W + 2)][C]
    const int C = g.C, C4 = C >> 2, F = g.F, W = g.W, H = g.H, WP = W + 2;
    const int bands = (H + FB_RB - 1) / FB_RB;
    const int n = blockIdx.x / bands, y0 = (blockIdx.x % bands) * FB_RB;
    const int cq = threadIdx.x % C4, ppw = 32 / C4, sub = (threadIdx.x & 31) / C4, warp = threadIdx.x >> 5;
    float wr[9][4][FM];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < FM; ++f)
                wr[tap][j][f] = f < F ? w[((long long)((cq * 4 + j) * 3 + tap / 3) * 3 + tap % 3) * F + f] : 0.0f;
    const int row4 = WP * C4;                               // float4 per tile row
    for (int i = threadIdx.x; i < (FB_RB + 2) * row4; i += blockDim.x) {
        const int r = i / row4, rem = i - r * row4, px = rem / C4, c4 = rem - px * C4;
        const int iy = y0 - 1 + r, ix = px - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = *reinterpret_cast<const float4*>(x + (((long long)n * H + iy) * W + ix) * C + c4 * 4);
        reinterpret_cast<float4*>(fb_tile)[i] = v;
    }
    __syncthreads();
    const int rows = min(FB_RB, H - y0), npix = rows * W;
    for (int p0 = warp * ppw; p0 < npix; p0 += 8 * ppw) {
        const int p = p0 + sub;
        const bool pok = p < npix;
        const int oy = p / W, ox = p - oy * W;
        float acc[FM];
#pragma unroll
        for (int f = 0; f < FM; ++f) acc[f] = 0.0f;
        if (pok) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float4 v = reinterpret_cast<const float4*>(fb_tile)[((oy + tap / 3) * WP + ox + tap % 3) * C4 + cq];
#pragma unroll
                for (int f = 0; f < FM; ++f) {
                    acc[f] = fmaf(v.x, wr[tap][0][f], acc[f]); acc[f] = fmaf(v.y, wr[tap][1][f], acc[f]);
                    acc[f] = fmaf(v.z, wr[tap][2][f], acc[f]); acc[f] = fmaf(v.w, wr[tap][3][f], acc[f]);
                }
            }
        }
#pragma unroll
        for (int f = 0; f < FM; ++f)
            for (int o = 1; o < C4; o <<= 1) acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], o);
        if (pok && cq == 0) {
            const long long po = ((long long)n * H + y0 + oy) * W + ox;
#pragma unroll
            for (int f = 0; f < FM; ++f)
                if (f < F) y[po * F + f] = act_apply(act, acc[f] + (bias != nullptr ? bias[f] : 0.0f), alpha);
        }
    }
}

// weight + bias gradient: persistent CTAs walk (image, band) items; lane (pixel slot, channel quad) accumulates its 9 x 4 x F outputs
// in registers over all its pixels; one partial per CTA in the reference's row order (c,ky,kx) (+ the bias row) -> splitk_reduce.
template <int FM>
__global__ void __launch_bounds__(256) conv_fewf_band_wgrad_kernel(b200nn_conv_geom g, const float* __restrict__ x, const float* __restrict__ err,
                                                                   float* __restrict__ partial) {
    extern __shared__ __align__(16) float fb_tile[];       // x tile [(FB_RB + 2)][(W + 2)][C] | e tile [FB_RB][W][FM] | red[8][9 C FM]
    const int C = g.C, C4 = C >> 2, F = g.F, W = g.W, H = g.H, WP = W + 2;
    const int bands = (H + FB_RB - 1) / FB_RB, items = g.N * bands;
    const int cq = threadIdx.x % C4, ppw = 32 / C4, sub = (threadIdx.x & 31) / C4, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* et = fb_tile + (FB_RB + 2) * WP * C;
    float* red = et + FB_RB * W * FM;
    float acc[9][4][FM], bsum[FM];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < FM; ++f) acc[tap][j][f] = 0.0f;
#pragma unroll
    for (int f = 0; f < FM; ++f) bsum[f] = 0.0f;
    const int row4 = WP * C4;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / bands, y0 = (item % bands) * FB_RB;
        const int rows = min(FB_RB, H - y0), npix = rows * W;
        __syncthreads();
        for (int i = threadIdx.x; i < (FB_RB + 2) * row4; i += blockDim.x) {
            const int r = i / row4, rem = i - r * row4, px = rem / C4, c4 = rem - px * C4;
            const int iy = y0 - 1 + r, ix = px - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = *reinterpret_cast<const float4*>(x + (((long long)n * H + iy) * W + ix) * C + c4 * 4);
            reinterpret_cast<float4*>(fb_tile)[i] = v;
        }
        for (int i = threadIdx.x; i < npix * FM; i += blockDim.x) {
            const int p = i / FM, f = i - p * FM;
            et[i] = f < F ? err[(((long long)n * H + y0) * W + p) * F + f] : 0.0f;
        }
        __syncthreads();
        for (int p0 = warp * ppw; p0 < npix; p0 += 8 * ppw) {
            const int p = p0 + sub;
            if (p < npix) {
                const int oy = p / W, ox = p - oy * W;
                float e[FM];
#pragma unroll
                for (int f = 0; f < FM; ++f) e[f] = et[p * FM + f];
                if (cq == 0) {
#pragma unroll
                    for (int f = 0; f < FM; ++f) bsum[f] += e[f];
                }
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const float4 v = reinterpret_cast<const float4*>(fb_tile)[((oy + tap / 3) * WP + ox + tap % 3) * C4 + cq];
#pragma unroll
                    for (int f = 0; f < FM; ++f) {
                        acc[tap][0][f] = fmaf(v.x, e[f], acc[tap][0][f]); acc[tap][1][f] = fmaf(v.y, e[f], acc[tap][1][f]);
                        acc[tap][2][f] = fmaf(v.z, e[f], acc[tap][2][f]); acc[tap][3][f] = fmaf(v.w, e[f], acc[tap][3][f]);
                    }
                }
            }
        }
    }
    // fold the pixel slots of a warp (lanes C4 apart), then the warps through shared memory
    __syncthreads();
    const int K = 9 * C;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < FM; ++f) {
                float v = acc[tap][j][f];
                for (int o = C4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane < C4) red[(warp * K + ((cq * 4 + j) * 9 + tap)) * FM + f] = v;      // row (c, ky, kx): the reference order
            }
    __shared__ float bred[8][FM];
#pragma unroll
    for (int f = 0; f < FM; ++f) {
        float v = bsum[f];
        for (int o = 1; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) bred[warp][f] = v;
    }
    __syncthreads();
    float* pb = partial + (long long)blockIdx.x * (K + 1) * F;
    for (int i = threadIdx.x; i < K * FM; i += blockDim.x) {
        const int row = i / FM, f = i - row * FM;
        float t = 0.0f;
        for (int wv = 0; wv < 8; ++wv) t += red[(wv * K + row) * FM + f];
        if (f < F) pb[(long long)row * F + f] = t;
    }
    if (threadIdx.x < FM && (int)threadIdx.x < F) {
        float t = 0.0f;
        for (int wv = 0; wv < 8; ++wv) t += bred[wv][threadIdx.x];
        pb[(long long)K * F + threadIdx.x] = t;
    }
}

// data gradient: thread = (input pixel, channel quad), the nine error taps are scalar loads that hit L1, no stride arithmetic
template <int FM>
__global__ void __launch_bounds__(256) conv_fewf_band_dx_kernel(b200nn_conv_geom g, const float* __restrict__ err, const float* __restrict__ w,
                                                                Epilogue epi, long long total4) {
    __shared__ double scratch[32];
    const int C4 = g.C >> 2, F = g.F, W = g.W, H = g.H;
    const int cq = threadIdx.x % C4;
    float wr[9][4][FM];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < FM; ++f)
                wr[tap][j][f] = f < F ? w[((long long)((cq * 4 + j) * 3 + tap / 3) * 3 + tap % 3) * F + f] : 0.0f;
    EpiState st;
    epi_begin(epi, st);
    // blockDim.x is a multiple of C4, so a thread keeps its channel quad over the grid-stride loop
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        long long t = i / C4;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H); const long long n = t / H;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int oy = iy + 1 - tap / 3, ox = ix + 1 - tap % 3;
            if (oy >= 0 && oy < H && ox >= 0 && ox < W) {
                const float* ep = err + ((n * H + oy) * W + ox) * F;
#pragma unroll
                for (int f = 0; f < FM; ++f) {
                    if (f < F) {
                        const float e = __ldg(ep + f);
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[j] = fmaf(e, wr[tap][j][f], acc[j]);
                    }
                }
            }
        }
        float v[4] = {acc[0] * st.scale, acc[1] * st.scale, acc[2] * st.scale, acc[3] * st.scale};
        float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
        if (epi.mask_y != nullptr) mk = *reinterpret_cast<const float4*>(epi.mask_y + i * 4);
        const float m[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st.a0 = fmaf(v[j], v[j], st.a0);
            if (epi.mask_y != nullptr) {
                v[j] *= act_grad_from_y(epi.mask_act, m[j], epi.mask_alpha);
                st.a1 = fmaf(v[j], v[j], st.a1);
            }
        }
        *reinterpret_cast<float4*>(epi.out + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    epi_finish(epi, st, scratch);
}

static bool conv_fewf_band_ok(const b200nn_conv_geom* g) {
    const int C4 = g->C >> 2;
    return g->kh == 3 && g->kw == 3 && g->sh == 1 && g->sw == 1 && g->ph == 1 && g->pw == 1 && g->OH == g->H && g->OW == g->W &&
           g->C % 4 == 0 && C4 >= 1 && C4 <= 32 && (C4 & (C4 - 1)) == 0 && g->F <= 2 &&
           (size_t)((FB_RB + 2) * (g->W + 2) * g->C + FB_RB * g->W * 2 + 8 * 9 * g->C * 2) * sizeof(float) <= 46 * 1024;
}
static size_t conv_fewf_band_smem(const b200nn_conv_geom* g, bool wgrad) {
    size_t fl = (size_t)(FB_RB + 2) * (g->W + 2) * g->C;
    if (wgrad) fl += (size_t)FB_RB * g->W * 2 + (size_t)8 * 9 * g->C * 2;
    return fl * sizeof(float);
}
static int conv_fewf_band_groups() { return kNumSMs * 2; }

static bool conv_fewf_fwd2_ok(const b200nn_conv_geom* g) {
    const int C4 = g->C >> 2;
    return g->kh == 3 && g->kw == 3 && g->C % 4 == 0 && C4 >= 1 && C4 <= 32 && (C4 & (C4 - 1)) == 0 && g->F <= 2;
}

// dx[n,iy,ix,c] = scale * sum_{ky,kx,f} err[n,oy,ox,f] * W[c,ky,kx,f] (+ activation backward, sum-of-squares slots).
// thread = (input pixel, channel quad); works for few filters (scalar err loads) and any C % 4 == 0
__global__ void __launch_bounds__(256) conv_fewf_dx_kernel(b200nn_conv_geom g, const float* __restrict__ err, const float* __restrict__ w,
                                                           Epilogue epi, long long total4) {
    __shared__ float ws[FEW_SMEM_FLOATS];            // [(ky,kx)][f][c]
    __shared__ double scratch[32];
    const int K = g.kh * g.kw * g.C, F = g.F, C4 = g.C >> 2;
    for (int i = threadIdx.x; i < K * F; i += blockDim.x) {
        const int c = i % g.C; int t = i / g.C;
        const int f = t % F, tap = t / F;
        ws[i] = w[((long long)(c * g.kh + tap / g.kw) * g.kw + tap % g.kw) * F + f];
    }
    EpiState st;
    epi_begin(epi, st);
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4; long long t = i / C4;
        const int ix = (int)(t % g.W); t /= g.W;
        const int iy = (int)(t % g.H); const int n = (int)(t / g.H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ky = 0; ky < g.kh; ++ky) {
            const int ty = iy + g.ph - ky;
            if (ty < 0 || ty % g.sh != 0) continue;
            const int oy = ty / g.sh;
            if (oy >= g.OH) continue;
            for (int kx = 0; kx < g.kw; ++kx) {
                const int tx = ix + g.pw - kx;
                if (tx < 0 || tx % g.sw != 0) continue;
                const int ox = tx / g.sw;
                if (ox >= g.OW) continue;
                const float* ep = err + (((long long)n * g.OH + oy) * g.OW + ox) * F;
                const float* wt = ws + (ky * g.kw + kx) * F * g.C + c;
                for (int f = 0; f < F; ++f) {
                    const float e = ep[f];
                    const float4 wv = *reinterpret_cast<const float4*>(wt + f * g.C);
                    acc.x = fmaf(e, wv.x, acc.x); acc.y = fmaf(e, wv.y, acc.y); acc.z = fmaf(e, wv.z, acc.z); acc.w = fmaf(e, wv.w, acc.w);
                }
            }
        }
        float v[4] = {acc.x * st.scale, acc.y * st.scale, acc.z * st.scale, acc.w * st.scale};
        float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
        if (epi.mask_y != nullptr) mk = *reinterpret_cast<const float4*>(epi.mask_y + i * 4);
        const float m[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st.a0 = fmaf(v[j], v[j], st.a0);
            if (epi.mask_y != nullptr) {
                v[j] *= act_grad_from_y(epi.mask_act, m[j], epi.mask_alpha);
                st.a1 = fmaf(v[j], v[j], st.a1);
            }
        }
        *reinterpret_cast<float4*>(epi.out + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    epi_finish(epi, st, scratch);
}

// dx for a few INPUT channels (C <= 4, F % 4 == 0): thread = (input pixel, c), float4 over the filters of the error
__global__ void __launch_bounds__(256) conv_fewc_dx_kernel(b200nn_conv_geom g, const float* __restrict__ err, const float* __restrict__ w,
                                                           Epilogue epi, long long total) {
    __shared__ float ws[FEW_SMEM_FLOATS];            // [c][(ky,kx)][f] == the flat weight buffer
    __shared__ double scratch[32];
    const int K = g.kh * g.kw * g.C, F = g.F;
    for (int i = threadIdx.x; i < K * F; i += blockDim.x) ws[i] = w[i];
    EpiState st;
    epi_begin(epi, st);
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % g.C); long long t = i / g.C;
        const int ix = (int)(t % g.W); t /= g.W;
        const int iy = (int)(t % g.H); const int n = (int)(t / g.H);
        float acc = 0.0f;
        for (int ky = 0; ky < g.kh; ++ky) {
            const int ty = iy + g.ph - ky;
            if (ty < 0 || ty % g.sh != 0) continue;
            const int oy = ty / g.sh;
            if (oy >= g.OH) continue;
            for (int kx = 0; kx < g.kw; ++kx) {
                const int tx = ix + g.pw - kx;
                if (tx < 0 || tx % g.sw != 0) continue;
                const int ox = tx / g.sw;
                if (ox >= g.OW) continue;
                const float4* ep = reinterpret_cast<const float4*>(err + (((long long)n * g.OH + oy) * g.OW + ox) * F);
                const float4* wt = reinterpret_cast<const float4*>(ws + ((c * g.kh + ky) * g.kw + kx) * F);
                for (int f4 = 0; f4 < F / 4; ++f4) {
                    const float4 e = ep[f4], wv = wt[f4];
                    acc = fmaf(e.x, wv.x, acc); acc = fmaf(e.y, wv.y, acc); acc = fmaf(e.z, wv.z, acc); acc = fmaf(e.w, wv.w, acc);
                }
            }
        }
        const float mk = epi.mask_y != nullptr ? epi.mask_y[i] : 0.0f;
        float v = acc * st.scale;
        st.a0 = fmaf(v, v, st.a0);
        if (epi.mask_y != nullptr) {
            v *= act_grad_from_y(epi.mask_act, mk, epi.mask_alpha);
            st.a1 = fmaf(v, v, st.a1);
        }
        epi.out[i] = v;
    }
    epi_finish(epi, st, scratch);
}

// Weight / bias gradient for a few filters: dW[c,ky,kx,f] = sum_pix x[n,iy,ix,c] * err[n,oy,ox,f]. A warp owns one channel quad and
// a strided set of INPUT pixels (lane = pixel); each lane accumulates its pixel's contribution to the (tap, 4 channels, f) outputs of
// the quad in registers, the warp folds them with shuffles and writes one partial row block: partial[group][(c,ky,kx) | bias][f].
template <int TAPS, int FM>
__global__ void __launch_bounds__(256) conv_fewf_wgrad_kernel(b200nn_conv_geom g, const float* __restrict__ x, const float* __restrict__ err,
                                                              float* __restrict__ partial, long long n_in_pix, long long n_out_pix) {
    const int lane = threadIdx.x & 31;
    const int C4 = g.C >> 2, F = g.F, K = g.kh * g.kw * g.C;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, total_warps = (gridDim.x * blockDim.x) >> 5;
    const int cq = wg % C4, group = wg / C4, n_groups = total_warps / C4;
    if (group >= n_groups) return;
    float acc[TAPS][4][FM];
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < FM; ++f) acc[t][j][f] = 0.0f;
    for (long long p = (long long)group * 32 + lane; p < n_in_pix; p += (long long)n_groups * 32) {
        const int ix = (int)(p % g.W); long long t = p / g.W;
        const int iy = (int)(t % g.H); const int n = (int)(t / g.H);
        const float4 xv = *reinterpret_cast<const float4*>(x + p * g.C + cq * 4);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int tap = 0; tap < TAPS; ++tap) {
            const int ky = tap / g.kw, kx = tap % g.kw;
            const int ty = iy + g.ph - ky, tx = ix + g.pw - kx;
            if (ty < 0 || tx < 0 || ty % g.sh != 0 || tx % g.sw != 0) continue;
            const int oy = ty / g.sh, ox = tx / g.sw;
            if (oy >= g.OH || ox >= g.OW) continue;
            const float* ep = err + (((long long)n * g.OH + oy) * g.OW + ox) * F;
#pragma unroll
            for (int f = 0; f < FM; ++f) {
                if (f < F) {
                    const float e = ep[f];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[tap][j][f] = fmaf(xs[j], e, acc[tap][j][f]);
                }
            }
        }
    }
    float* pb = partial + (long long)group * (K + 1) * F;
#pragma unroll
    for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < FM; ++f) {
                const float r = warp_sum(acc[tap][j][f]);
                if (lane == 0 && f < F) {
                    const int c = cq * 4 + j;
                    pb[((long long)(c * g.kh + tap / g.kw) * g.kw + tap % g.kw) * F + f] = r;   // reference row order (c,ky,kx)
                }
            }
    if (cq == 0) {   // bias row: sum of the error over this group's share of the OUTPUT pixels
        float bs[FM];
#pragma unroll
        for (int f = 0; f < FM; ++f) bs[f] = 0.0f;
        for (long long p = (long long)group * 32 + lane; p < n_out_pix; p += (long long)n_groups * 32)
#pragma unroll
            for (int f = 0; f < FM; ++f)
                if (f < F) bs[f] += err[p * F + f];
#pragma unroll
        for (int f = 0; f < FM; ++f) {
            const float r = warp_sum(bs[f]);
            if (lane == 0 && f < F) pb[(long long)K * F + f] = r;
        }
    }
}

static bool conv_fewf_ok(const b200nn_conv_geom* g) {
    return g->F <= FEWF_MAX && g->C % 4 == 0 && g->kh * g->kw * g->C * g->F <= FEW_SMEM_FLOATS && g->kh * g->kw == 9;
}
static bool conv_fewc_dx_ok(const b200nn_conv_geom* g) {
    return g->C <= 4 && g->F % 4 == 0 && g->kh * g->kw * g->C * g->F <= FEW_SMEM_FLOATS;
}
static int conv_fewf_groups(const b200nn_conv_geom* g) {
    const int C4 = g->C / 4;
    int warps = kNumSMs * 2 * 8;
    int groups = warps / C4;
    return groups < 1 ? 1 : groups;
}

extern "C" {

static int check_conv_geom(const char* who, const b200nn_conv_geom* g) {
    B200NN_REQUIRE(g != nullptr, "%s: null geometry", who);
    B200NN_REQUIRE(g->N > 0 && g->H > 0 && g->W > 0 && g->C > 0 && g->F > 0 && g->kh > 0 && g->kw > 0 && g->sh > 0 &&
                       g->sw > 0 && g->ph >= 0 && g->pw >= 0,
                   "%s: bad geometry", who);
    B200NN_REQUIRE(g->OH == (g->H + 2 * g->ph - g->kh) / g->sh + 1 && g->OW == (g->W + 2 * g->pw - g->kw) / g->sw + 1 &&
                       g->OH > 0 && g->OW > 0,
                   "%s: output size %dx%d inconsistent with input %dx%d k%dx%d s%dx%d p%dx%d", who, g->OH, g->OW, g->H,
                   g->W, g->kh, g->kw, g->sh, g->sw, g->ph, g->pw);
    B200NN_REQUIRE((long long)g->N * g->OH * g->OW < (1LL << 31) && (long long)g->N * g->H * g->W < (1LL << 31),
                   "%s: pixel count overflows int32", who);
    return B200NN_OK;
}

size_t b200nn_conv2d_ws_bytes(const b200nn_conv_geom* g);
size_t b200nn_conv2d_ws_bytes_math(int math, const b200nn_conv_geom* g) {
    size_t r = b200nn_conv2d_ws_bytes(g);
    if (math != B200NN_MATH_FP32 && !tcapi::conv_ok(g, nullptr, nullptr, nullptr) && conv_col_ok(g)) {
        const long long M = (long long)g->N * g->OH * g->OW, K = (long long)g->kh * g->kw * g->C;
        const size_t cw = (size_t)(conv_col_ws(g, nullptr).rest - (float*)nullptr) * sizeof(float) +
                          b200nn_dense_ws_bytes_math(math, (int)M, g->F, (int)K) + 256;
        r = r > cw ? r : cw;
    }
    if (math != B200NN_MATH_FP32 && tcapi::conv_ok(g, nullptr, nullptr, nullptr)) {
        size_t t = tcapi::conv_ws1(g, math == B200NN_MATH_TF32X3);
        const size_t t2 = tcapi::conv_ws2(g, math == B200NN_MATH_TF32X3);
        t = t > t2 ? t : t2;
        if (math == B200NN_MATH_TF32X3 && tcapi::tc2_enabled()) {      // partials + operand planes of the fp16-plane path
            const long long P = (long long)g->N * g->H * g->W, taps = (long long)g->kh * g->kw;
            for (int which = 1; which <= 2; ++which) {
                if (!tcapi::conv_f16_worth(which, g)) continue;
                const long long Ck = which == 1 ? g->C : g->F, Nn = which == 1 ? g->F : g->C;
                const size_t f = (tcapi::ws2(P, Nn, taps * Ck / 2, true) + 255) / 256 * 256 + tcapi::conv_f16_extra_bytes(g);
                t = t > f ? t : f;
            }
        }
        const size_t d = (size_t)kNumSMs * 2 * g->F * sizeof(float);   // column-sum partials for db
        r = r > t ? r : t;
        r = r > d ? r : d;
        if (math == B200NN_MATH_TF32X3) r = (r + 255) / 256 * 256 + 256;   // the last 256 bytes: max word of the error tensor (conv2d_bwd)
    }
    return r;
}

size_t b200nn_conv2d_ws_bytes(const b200nn_conv_geom* g) {
    const long long Mo = (long long)g->N * g->OH * g->OW, Mi = (long long)g->N * g->H * g->W;
    const long long Kf = (long long)g->kh * g->kw * g->C, Kd = (long long)g->kh * g->kw * g->F;
    size_t a = ws_bytes_for(Mo, g->F, Kf);
    size_t b = ws_bytes_for(Kf + 1, g->F, Mo);
    size_t c = ws_bytes_for(Mi, g->C, Kd);
    size_t r = a > b ? a : b;
    r = r > c ? r : c;
    if (smallk_wgrad_ok(g)) {
        const size_t d = (size_t)smallk_wgrad_blocks(g) * (Kf + 1) * g->F * sizeof(float);
        r = r > d ? r : d;
    }
    if (conv_fewf_ok(g)) {
        size_t d = (size_t)conv_fewf_groups(g) * (Kf + 1) * g->F * sizeof(float);
        r = r > d ? r : d;
        d = (size_t)conv_fewf_band_groups() * (Kf + 1) * g->F * sizeof(float);
        r = r > d ? r : d;
    }
    if (smallk3::ok(g)) {
        const size_t d = (size_t)smallk3::blocks(g) * (Kf + 1) * g->F * sizeof(float);
        r = r > d ? r : d;
    }
    return r;
}

static PatchGather conv_patches(const b200nn_conv_geom* g, const float* x) {
    PatchGather p;
    p.src = x; p.SH = g->H; p.SW = g->W; p.SC = g->C; p.OH = g->OH; p.OW = g->OW;
    p.kh = g->kh; p.kw = g->kw; p.sh = g->sh; p.sw = g->sw; p.ph = g->ph; p.pw = g->pw;
    p.n_pix = g->N * g->OH * g->OW; p.n_tap = g->kh * g->kw * g->C; p.skip_truncated = 0;
    return p;
}

int b200nn_conv2d_fwd(int math, const b200nn_conv_geom* g, const float* x, const float* w, const float* b, float* y,
                      int act, float alpha, float* ws, void* stream) {
    B200NN_REQUIRE(math >= B200NN_MATH_FP32 && math <= B200NN_MATH_TF32X3, "conv2d_fwd: unknown math mode %d", math);
    if (int rc = check_conv_geom("conv2d_fwd", g)) return rc;
    if (int rc = check_act("conv2d_fwd", act, alpha)) return rc;
    const int M = g->N * g->OH * g->OW, K = g->kh * g->kw * g->C;
    if (math != B200NN_MATH_FP32 && tcapi::conv_ok(g, x, w, y)) {   // implicit GEMM on tcgen05, operands fetched by TMA
        Epilogue e = plain_epilogue(y, g->F, M);
        e.bias = b; e.act = act; e.alpha = alpha;
        return tc_conv_any(1, g, x, w, e, ws, as_stream(stream), math == B200NN_MATH_TF32X3);
    }
    if (conv_fewf_ok(g) && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {       // a handful of filters: direct kernel, every math mode
        long long blocks = ((long long)M + 255) / 256;
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        if (conv_fewf_band_ok(g)) {
            const int bands = (g->H + FB_RB - 1) / FB_RB;
            const size_t smem = conv_fewf_band_smem(g, false);
            if (g->F == 1) conv_fewf_band_fwd_kernel<1><<<g->N * bands, 256, smem, as_stream(stream)>>>(*g, x, w, b, y, act, alpha);
            else conv_fewf_band_fwd_kernel<2><<<g->N * bands, 256, smem, as_stream(stream)>>>(*g, x, w, b, y, act, alpha);
            B200NN_LAUNCH_CHECK("conv_fewf_band_fwd");
            return B200NN_OK;
        }
        if (conv_fewf_fwd2_ok(g) && M >= 4096) {
            const long long warps = ((long long)M + (32 / (g->C >> 2)) - 1) / (32 / (g->C >> 2));
            long long b2 = (warps + 7) / 8;
            if (b2 > kNumSMs * 8) b2 = kNumSMs * 8;
            if (g->F == 1) conv_fewf_fwd2_kernel<1><<<(int)b2, 256, 0, as_stream(stream)>>>(*g, x, w, b, y, act, alpha, M);
            else conv_fewf_fwd2_kernel<2><<<(int)b2, 256, 0, as_stream(stream)>>>(*g, x, w, b, y, act, alpha, M);
            B200NN_LAUNCH_CHECK("conv_fewf_fwd2");
            return B200NN_OK;
        }
        conv_fewf_fwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(*g, x, w, b, y, act, alpha, M);
        B200NN_LAUNCH_CHECK("conv_fewf_fwd");
        return B200NN_OK;
    }
    if (math != B200NN_MATH_FP32 && ws != nullptr && conv_col_ok(g) &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0) {
        // im2col + the Dense kernels (tcgen05 / skinny / small-N, by shape)
        cudaStream_t st = as_stream(stream);
        const ConvColWs cw = conv_col_ws(g, ws);
        const long long total4 = (long long)M * K / 4;
        long long blocks = (total4 + 255) / 256;
        if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
        conv_im2col_kernel<<<(int)blocks, 256, 0, st>>>(*g, x, cw.col, total4);
        B200NN_LAUNCH_CHECK("conv_im2col");
        conv_weight_permute_kernel<<<(K * g->F + 255) / 256, 256, 0, st>>>(w, cw.wp, g->kh, g->kw, g->C, g->F, 0);
        B200NN_LAUNCH_CHECK("conv_weight_permute");
        return b200nn_dense_fwd(math, cw.col, cw.wp, b, y, M, g->F, K, act, alpha, cw.rest, stream);
    }
    if (smallk3::ok(g) && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
        smallk3::launch_fwd(g, x, w, b, y, act, alpha, M, as_stream(stream), math == B200NN_MATH_TF32X3 ? 2 : (math == B200NN_MATH_TF32 ? 1 : 0));
        B200NN_LAUNCH_CHECK("conv2d_smallk_fwd3");
        return B200NN_OK;
    }
    if (smallk_fwd_ok(g) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
        long long blocks = ((long long)M + 127) / 128;
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        cudaStream_t st = as_stream(stream);
        if (g->F == 32) conv2d_smallk_fwd_kernel<8><<<(int)blocks, 128, 0, st>>>(*g, x, w, b, y, act, alpha, M);
        else if (g->F == 64) conv2d_smallk_fwd_kernel<16><<<(int)blocks, 128, 0, st>>>(*g, x, w, b, y, act, alpha, M);
        else conv2d_smallk_fwd_kernel<4><<<(int)blocks, 128, 0, st>>>(*g, x, w, b, y, act, alpha, M);
        B200NN_LAUNCH_CHECK("conv2d_smallk_fwd");
        return B200NN_OK;
    }
    PatchA af{conv_patches(g, x)};
    ConvWFwdB bf{w, g->kh, g->kw, g->C, g->F};
    Epilogue e = plain_epilogue(y, g->F, M);
    e.bias = b; e.act = act; e.alpha = alpha;
    return launch_gemm("conv2d_fwd", af, bf, e, M, g->F, K, ws, as_stream(stream));
}

int b200nn_conv2d_bwd(int math, const b200nn_conv_geom* g, const float* x, const float* w, const float* err,
                      const double* ss, uint64_t chain, float* dx, float* dw, float* db, int prev_act,
                      float prev_alpha, double* ss_out0, double* ss_out1, float* ws, void* stream) {
    B200NN_REQUIRE(math >= B200NN_MATH_FP32 && math <= B200NN_MATH_TF32X3, "conv2d_bwd: unknown math mode %d", math);
    if (int rc = check_conv_geom("conv2d_bwd", g)) return rc;
    if (int rc = check_act("conv2d_bwd", prev_act, prev_alpha)) return rc;
    cudaStream_t st = as_stream(stream);
    const int Mo = g->N * g->OH * g->OW, Mi = g->N * g->H * g->W;
    const int Kf = g->kh * g->kw * g->C, Kd = g->kh * g->kw * g->F;
    if (math != B200NN_MATH_FP32 && ws != nullptr && tcapi::conv_ok(g, x, w, err) &&
        ((reinterpret_cast<uintptr_t>(dw) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
        const bool x3 = math == B200NN_MATH_TF32X3;
        unsigned* errmax = nullptr;     // max |err| for the fp16 image planes of the data gradient, from the column-sum pass
        if (dw != nullptr) {
            if (db != nullptr) {   // layers.py:513
                int nb = (Mo + 7) / 8;
                if (nb > kNumSMs * 2) nb = kNumSMs * 2;
                if (x3 && dx != nullptr && colsum4r_ok(err, ws, g->F) && tcapi::tc2_enabled() && tcapi::conv_f16_worth(2, g)) {
                    errmax = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + b200nn_conv2d_ws_bytes_math(math, g) - 256);
                    B200NN_CUDA(cudaMemsetAsync(errmax, 0, sizeof(unsigned), st));
                }
                if (colsum4r_ok(err, ws, g->F)) colsum_partial4r_kernel<<<nb, 256, 0, st>>>(err, Mo, g->F, ws, errmax);
                else colsum_partial_kernel<<<nb, 256, 0, st>>>(err, Mo, g->F, ws);
                B200NN_LAUNCH_CHECK("conv2d_db_partial");
                colsum_final_kernel<<<(g->F + 31) / 32, 256, 0, st>>>(ws, nb, g->F, ss, chain, db);
                B200NN_LAUNCH_CHECK("conv2d_db_final");
            }
            Epilogue e = plain_epilogue(dw, g->F, Kf);
            e.ss = ss; e.chain = chain;
            e.row_mode = 1; e.kh = g->kh; e.kw = g->kw; e.C = g->C;
            if (int rc = tc_conv_any(3, g, x, err, e, ws, st, x3)) return rc;
        }
        if (dx != nullptr) {
            Epilogue e = plain_epilogue(dx, g->C, Mi);
            dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
            if (int rc = tc_conv_any(2, g, err, w, e, ws, st, x3, errmax)) return rc;
        }
        return B200NN_OK;
    }
    if (conv_fewf_ok(g) && ws != nullptr && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
        // a handful of filters: direct bandwidth kernels in every math mode
        const bool band = conv_fewf_band_ok(g);
        if (dw != nullptr && band) {
            const int groups = conv_fewf_band_groups();
            Epilogue e = plain_epilogue(dw, g->F, Kf);
            e.ss = ss; e.chain = chain; e.db = db;
            const size_t smem = conv_fewf_band_smem(g, true);
            if (g->F == 1) conv_fewf_band_wgrad_kernel<1><<<groups, 256, smem, st>>>(*g, x, err, ws);
            else conv_fewf_band_wgrad_kernel<2><<<groups, 256, smem, st>>>(*g, x, err, ws);
            B200NN_LAUNCH_CHECK("conv_fewf_band_wgrad");
            if (int rc = launch_splitk_reduce(ws, groups, Kf + 1, g->F, e, st)) return rc;
        } else if (dw != nullptr) {
            const int groups = conv_fewf_groups(g);
            const int warps = groups * (g->C / 4);
            Epilogue e = plain_epilogue(dw, g->F, Kf);
            e.ss = ss; e.chain = chain; e.db = db;
            if (g->F == 1) conv_fewf_wgrad_kernel<9, 1><<<(warps + 7) / 8, 256, 0, st>>>(*g, x, err, ws, Mi, Mo);
            else if (g->F == 2) conv_fewf_wgrad_kernel<9, 2><<<(warps + 7) / 8, 256, 0, st>>>(*g, x, err, ws, Mi, Mo);
            else conv_fewf_wgrad_kernel<9, 4><<<(warps + 7) / 8, 256, 0, st>>>(*g, x, err, ws, Mi, Mo);
            B200NN_LAUNCH_CHECK("conv_fewf_wgrad");
            if (int rc = launch_splitk_reduce(ws, groups, Kf + 1, g->F, e, st)) return rc;
        }
        if (dx != nullptr) {
            Epilogue e = plain_epilogue(dx, g->C, Mi);
            dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
            const long long t4 = (long long)Mi * g->C / 4;
            long long b2 = (t4 + 255) / 256;
            if (b2 > kNumSMs * 16) b2 = kNumSMs * 16;
            if (band && g->F == 1) conv_fewf_band_dx_kernel<1><<<(int)b2, 256, 0, st>>>(*g, err, w, e, t4);
            else if (band) conv_fewf_band_dx_kernel<2><<<(int)b2, 256, 0, st>>>(*g, err, w, e, t4);
            else conv_fewf_dx_kernel<<<(int)b2, 256, 0, st>>>(*g, err, w, e, t4);
            B200NN_LAUNCH_CHECK("conv_fewf_dx");
        }
        return B200NN_OK;
    }
    if (math != B200NN_MATH_FP32 && ws != nullptr && conv_col_ok(g) &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(err) | reinterpret_cast<uintptr_t>(ws) |
          reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0) {
        // im2col + the Dense backward on the column matrix; the data gradient comes back through col2im (+ activation backward)
        const ConvColWs cw = conv_col_ws(g, ws);
        const long long total4 = (long long)Mo * Kf / 4;
        long long blocks = (total4 + 255) / 256;
        if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
        if (dw != nullptr) {
            conv_im2col_kernel<<<(int)blocks, 256, 0, st>>>(*g, x, cw.col, total4);
            B200NN_LAUNCH_CHECK("conv_im2col");
        }
        conv_weight_permute_kernel<<<(Kf * g->F + 255) / 256, 256, 0, st>>>(w, cw.wp, g->kh, g->kw, g->C, g->F, 0);
        B200NN_LAUNCH_CHECK("conv_weight_permute");
        if (int rc = b200nn_dense_bwd(math, cw.col, cw.wp, err, ss, chain, dx != nullptr ? cw.dcol : nullptr, dw != nullptr ? cw.dwp : nullptr, db,
                                      Mo, g->F, Kf, B200NN_ACT_NONE, 0.0f, nullptr, nullptr, cw.rest, stream)) return rc;
        if (dw != nullptr) {
            conv_weight_permute_kernel<<<(Kf * g->F + 255) / 256, 256, 0, st>>>(cw.dwp, dw, g->kh, g->kw, g->C, g->F, 1);
            B200NN_LAUNCH_CHECK("conv_weight_permute_back");
        }
        if (dx != nullptr) {
            const long long t4 = (long long)Mi * g->C / 4;
            long long b2 = (t4 + 255) / 256;
            if (b2 > kNumSMs * 16) b2 = kNumSMs * 16;
            conv_col2im_kernel<<<(int)b2, 256, 0, st>>>(*g, cw.dcol, prev_act != B200NN_ACT_NONE ? x : nullptr, prev_act, prev_alpha, dx,
                                                       ss_out0, ss_out1, t4);
            B200NN_LAUNCH_CHECK("conv_col2im");
        }
        return B200NN_OK;
    }
    if (dw != nullptr && smallk3::ok(g) && ws != nullptr && (reinterpret_cast<uintptr_t>(err) & 15) == 0) {
        const int nb = smallk3::blocks(g);
        Epilogue e = plain_epilogue(dw, g->F, Kf);
        e.ss = ss; e.chain = chain; e.db = db;
        e.row_mode = 1; e.kh = g->kh; e.kw = g->kw; e.C = g->C;
        B200NN_REQUIRE(smallk3::launch_wgrad(g, x, err, ws, Mo, st, math == B200NN_MATH_TF32X3 ? 2 : (math == B200NN_MATH_TF32 ? 1 : 0)) == 0,
                       "conv2d_smallk_wgrad3: cannot configure shared memory");
        B200NN_LAUNCH_CHECK("conv2d_smallk_wgrad3");
        if (int rc = launch_splitk_reduce(ws, nb, Kf + 1, g->F, e, st)) return rc;
    } else if (dw != nullptr && smallk_wgrad_ok(g) && ws != nullptr && (reinterpret_cast<uintptr_t>(err) & 15) == 0) {
        const int nb = smallk_wgrad_blocks(g), TP = smallk_wgrad_tile(g);
        Epilogue e = plain_epilogue(dw, g->F, Kf);
        e.ss = ss; e.chain = chain; e.db = db;
        e.row_mode = 1; e.kh = g->kh; e.kw = g->kw; e.C = g->C;
        const int TG = 256 / g->F;
        const int nacc = (Kf + 1 + TG - 1) / TG;
        const size_t smem = ((size_t)TP * g->F + (size_t)(Kf + 1) * TP) * sizeof(float);
        static thread_local bool configured = false;
        if (!configured) {
            B200NN_CUDA(cudaFuncSetAttribute(conv2d_smallk_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            B200NN_CUDA(cudaFuncSetAttribute(conv2d_smallk_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            B200NN_CUDA(cudaFuncSetAttribute(conv2d_smallk_wgrad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            B200NN_CUDA(cudaFuncSetAttribute(conv2d_smallk_wgrad_kernel<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            configured = true;
        }
        if (nacc <= 2) conv2d_smallk_wgrad_kernel<2><<<nb, 256, smem, st>>>(*g, x, err, ws, Mo, TP);
        else if (nacc <= 4) conv2d_smallk_wgrad_kernel<4><<<nb, 256, smem, st>>>(*g, x, err, ws, Mo, TP);
        else if (nacc <= 8) conv2d_smallk_wgrad_kernel<8><<<nb, 256, smem, st>>>(*g, x, err, ws, Mo, TP);
        else conv2d_smallk_wgrad_kernel<33><<<nb, 256, smem, st>>>(*g, x, err, ws, Mo, TP);
        B200NN_LAUNCH_CHECK("conv2d_smallk_wgrad");
        if (int rc = launch_splitk_reduce(ws, nb, Kf + 1, g->F, e, st)) return rc;
    } else if (dw != nullptr) {  // rows = taps (ky,kx,c) + ones row; reduce over output pixels
        PatchTOnesA af{conv_patches(g, x), db != nullptr ? 1 : 0};
        RowMajorB bf{err, g->F, Mo, g->F};
        Epilogue e = plain_epilogue(dw, g->F, Kf);
        e.ss = ss; e.chain = chain; e.db = db;
        e.row_mode = 1; e.kh = g->kh; e.kw = g->kw; e.C = g->C;
        if (int rc = launch_gemm("conv2d_dw", af, bf, e, Kf + (db != nullptr ? 1 : 0), g->F, Mo, ws, st)) return rc;
    }
    if (dx != nullptr && conv_fewc_dx_ok(g) && (reinterpret_cast<uintptr_t>(err) & 15) == 0) {   // a handful of input channels
        Epilogue e = plain_epilogue(dx, g->C, Mi);
        dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
        const long long total = (long long)Mi * g->C;
        long long b2 = (total + 255) / 256;
        if (b2 > kNumSMs * 16) b2 = kNumSMs * 16;
        conv_fewc_dx_kernel<<<(int)b2, 256, 0, st>>>(*g, err, w, e, total);
        B200NN_LAUNCH_CHECK("conv_fewc_dx");
    } else if (dx != nullptr) {  // gather form of col2im(err . W^T), preprocessing.py:82-126
        StridedGatherA af;
        af.src = err; af.SH = g->OH; af.SW = g->OW; af.SC = g->F; af.DH = g->H; af.DW = g->W;
        af.kh = g->kh; af.kw = g->kw; af.sh = g->sh; af.sw = g->sw; af.py = g->ph; af.px = g->pw;
        af.n_pix = Mi; af.n_tap = Kd;
        ConvWDgradB bf{w, g->kh, g->kw, g->C, g->F};
        Epilogue e = plain_epilogue(dx, g->C, Mi);
        dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
        if (int rc = launch_gemm("conv2d_dx", af, bf, e, Mi, g->C, Kd, ws, st)) return rc;
    }
    return B200NN_OK;
}

// ---------------------------------------------------------------------------------------------- Conv2DTranspose

// ------------------------------------------------------------------------------------------------
// Conv2DTranspose on the tensor cores (layers.py:2588-2659). The scatter form of the reference IS a plain contraction per INPUT
// pixel: Ycol[(n,h,w), (i,j,f)] = sum_c x[n,h,w,c] * W[i,j,f,c] — A = x [pixels, C] and B = the weight buffer [(i,j,f), C], both
// K-major as they lie in memory — followed by the overlap-add (each output pixel gathers its <= ceil(k/s)^2 contributions, bias
// and activation fused). The backward is two more plain contractions on Ecol[(n,h,w), (i,j,f)] = the stride-s window of the
// cropped error (no pad offset; a window truncated at the border is SKIPPED entirely, layers.py:2651, Q12), pending clips folded
// in: dX = Ecol . W[(i,j,f), c] and dW = Ecol^T . x. The column buffers (k*k/s*s times the output) live in the workspace.
__global__ void __launch_bounds__(256) convt_overlap_add_kernel(b200nn_convt_geom g, const float* __restrict__ ycol,
                                                                const float* __restrict__ bias, float* __restrict__ y, int act,
                                                                float alpha, long long total4) {
    const int F4 = g.F >> 2, NC = g.kh * g.kw * g.F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F4) * 4; long long t = i / F4;
        const int ox = (int)(t % g.OW); t /= g.OW;
        const int oy = (int)(t % g.OH); const int n = (int)(t / g.OH);
        float4 acc = bias != nullptr ? *reinterpret_cast<const float4*>(bias + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ky = 0; ky < g.kh; ++ky) {
            const int ty = oy + g.pt - ky;
            if (ty < 0 || ty % g.sh != 0) continue;
            const int h = ty / g.sh;
            if (h >= g.H) continue;
            for (int kx = 0; kx < g.kw; ++kx) {
                const int tx = ox + g.pl - kx;
                if (tx < 0 || tx % g.sw != 0) continue;
                const int w_ = tx / g.sw;
                if (w_ >= g.W) continue;
                const float4 v = *reinterpret_cast<const float4*>(ycol + (((long long)n * g.H + h) * g.W + w_) * NC +
                                                                  (long long)(ky * g.kw + kx) * g.F + f);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        acc = make_float4(act_apply(act, acc.x, alpha), act_apply(act, acc.y, alpha), act_apply(act, acc.z, alpha), act_apply(act, acc.w, alpha));
        *reinterpret_cast<float4*>(y + i * 4) = acc;
    }
}

__global__ void __launch_bounds__(256) convt_window_gather_kernel(b200nn_convt_geom g, const float* __restrict__ err,
                                                                  const double* __restrict__ ss, uint64_t chain,
                                                                  float* __restrict__ ecol, long long total4) {
    const float scale = chain_scale(ss, chain);
    const int F4 = g.F >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F4) * 4; long long t = i / F4;
        const int kx = (int)(t % g.kw); t /= g.kw;
        const int ky = (int)(t % g.kh); t /= g.kh;
        const int w_ = (int)(t % g.W); t /= g.W;
        const int h = (int)(t % g.H); const int n = (int)(t / g.H);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool whole = h * g.sh + g.kh <= g.OH && w_ * g.sw + g.kw <= g.OW;          // layers.py:2651: truncated windows are skipped
        if (whole) {
            v = *reinterpret_cast<const float4*>(err + (((long long)n * g.OH + h * g.sh + ky) * g.OW + w_ * g.sw + kx) * g.F + f);
            v = make_float4(scale * v.x, scale * v.y, scale * v.z, scale * v.w);
        }
        *reinterpret_cast<float4*>(ecol + i * 4) = v;
    }
}

static bool convt_tc_ok(const b200nn_convt_geom* g, const void* a, const void* b, const void* c) {
    if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) != 0) return false;
    return g->C % 4 == 0 && g->F % 4 == 0 && g->C >= 32 && (long long)g->kh * g->kw * g->F >= 64 &&
           (long long)g->N * g->H * g->W * g->C * g->kh * g->kw * g->F >= (1LL << 22);
}
static size_t convt_col_floats(const b200nn_convt_geom* g) { return (size_t)g->N * g->H * g->W * g->kh * g->kw * g->F; }

static int check_convt_geom(const char* who, const b200nn_convt_geom* g) {
    B200NN_REQUIRE(g != nullptr, "%s: null geometry", who);
    B200NN_REQUIRE(g->N > 0 && g->H > 0 && g->W > 0 && g->C > 0 && g->F > 0 && g->kh > 0 && g->kw > 0 && g->sh > 0 &&
                       g->sw > 0 && g->pt >= 0 && g->pl >= 0 && g->OH > 0 && g->OW > 0,
                   "%s: bad geometry", who);
    B200NN_REQUIRE((long long)g->N * g->OH * g->OW < (1LL << 31), "%s: pixel count overflows int32", who);
    return B200NN_OK;
}

size_t b200nn_convt_ws_bytes(const b200nn_convt_geom* g) {
    const long long Mo = (long long)g->N * g->OH * g->OW, Mi = (long long)g->N * g->H * g->W;
    const long long Kf = (long long)g->kh * g->kw * g->C, Kb = (long long)g->kh * g->kw * g->F;
    size_t a = ws_bytes_for(Mo, g->F, Kf);
    size_t b = ws_bytes_for(Kb, g->C, Mi);
    size_t c = ws_bytes_for(Mi, g->C, Kb);
    size_t d = (size_t)kNumSMs * 2 * g->F * sizeof(float);  // colsum partials for db
    size_t r = a > b ? a : b;
    r = r > c ? r : c;
    return r > d ? r : d;
}

size_t b200nn_convt_ws_bytes_math(int math, const b200nn_convt_geom* g) {
    size_t r = b200nn_convt_ws_bytes(g);
    if (math != B200NN_MATH_FP32 && g != nullptr && convt_tc_ok(g, nullptr, nullptr, nullptr)) {
        const bool x3 = math == B200NN_MATH_TF32X3;
        const long long M = (long long)g->N * g->H * g->W, NC = (long long)g->kh * g->kw * g->F;
        size_t t = tcapi::ws2(M, NC, g->C, x3);                       // forward  Ycol = x . W^T
        const size_t d = tcapi::ws2(M, g->C, NC, x3), e = tcapi::ws2(NC, g->C, M, x3);
        t = t > d ? t : d; t = t > e ? t : e;
        const size_t db = (size_t)kNumSMs * 2 * g->F * sizeof(float);
        t = t > db ? t : db;
        const size_t tc = convt_col_floats(g) * sizeof(float) + 256 + t;   // column buffer | GEMM split-K / db partials
        r = r > tc ? r : tc;
    }
    return r;
}

int b200nn_convt_fwd(int math, const b200nn_convt_geom* g, const float* x, const float* w, const float* b,
                     float* y, int act, float alpha, float* ws, void* stream) {
    B200NN_REQUIRE(math >= B200NN_MATH_FP32 && math <= B200NN_MATH_TF32X3, "convt_fwd: unknown math mode %d", math);
    if (int rc = check_convt_geom("convt_fwd", g)) return rc;
    if (int rc = check_act("convt_fwd", act, alpha)) return rc;
    if (math != B200NN_MATH_FP32 && ws != nullptr && tcapi::tc2_enabled() && convt_tc_ok(g, x, w, y)) {
        cudaStream_t st = as_stream(stream);
        const int M = g->N * g->H * g->W, NC = g->kh * g->kw * g->F;
        float* ycol = ws;
        float* gws = ws + ((convt_col_floats(g) + 63) / 64) * 64;
        Epilogue e = plain_epilogue(ycol, NC, M);
        if (int rc = tc_gemm_any(x, g->C, false, w, g->C, false, e, M, NC, g->C, gws, st, math == B200NN_MATH_TF32X3)) return rc;
        const long long total4 = (long long)g->N * g->OH * g->OW * (g->F / 4);
        long long blocks = (total4 + 255) / 256;
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        convt_overlap_add_kernel<<<(int)blocks, 256, 0, st>>>(*g, ycol, b, y, act, alpha, total4);
        B200NN_LAUNCH_CHECK("convt_overlap_add");
        return B200NN_OK;
    }
    StridedGatherA af;
    af.src = x; af.SH = g->H; af.SW = g->W; af.SC = g->C; af.DH = g->OH; af.DW = g->OW;
    af.kh = g->kh; af.kw = g->kw; af.sh = g->sh; af.sw = g->sw; af.py = g->pt; af.px = g->pl;
    af.n_pix = g->N * g->OH * g->OW; af.n_tap = g->kh * g->kw * g->C;
    ConvTWFwdB bf{w, g->kh, g->kw, g->C, g->F};
    Epilogue e = plain_epilogue(y, g->F, af.n_pix);
    e.bias = b; e.act = act; e.alpha = alpha;
    return launch_gemm("convt_fwd", af, bf, e, af.n_pix, g->F, af.n_tap, ws, as_stream(stream));
}

int b200nn_convt_bwd(int math, const b200nn_convt_geom* g, const float* x, const float* w, const float* err,
                     const double* ss, uint64_t chain, float* dx, float* dw, float* db, int prev_act,
                     float prev_alpha, double* ss_out0, double* ss_out1, float* ws, void* stream) {
    B200NN_REQUIRE(math >= B200NN_MATH_FP32 && math <= B200NN_MATH_TF32X3, "convt_bwd: unknown math mode %d", math);
    if (int rc = check_convt_geom("convt_bwd", g)) return rc;
    if (int rc = check_act("convt_bwd", prev_act, prev_alpha)) return rc;
    cudaStream_t st = as_stream(stream);
    const int Mi = g->N * g->H * g->W;
    const int Kb = g->kh * g->kw * g->F;
    if (math != B200NN_MATH_FP32 && ws != nullptr && tcapi::tc2_enabled() && convt_tc_ok(g, x, w, err) &&
        ((reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0) {
        const bool x3 = math == B200NN_MATH_TF32X3;
        float* ecol = ws;
        float* gws = ws + ((convt_col_floats(g) + 63) / 64) * 64;
        if (db != nullptr) {  // layers.py:2641: sum over ALL error pixels
            const long long R = (long long)g->N * g->OH * g->OW;
            int nb = (int)((R + 63) / 64);
            if (nb > kNumSMs * 2) nb = kNumSMs * 2;
            if (colsum4r_ok(err, gws, g->F)) colsum_partial4r_kernel<<<nb, 256, 0, st>>>(err, R, g->F, gws);
            else colsum_partial_kernel<<<nb, 256, 0, st>>>(err, R, g->F, gws);
            B200NN_LAUNCH_CHECK("convt_db_partial");
            colsum_final_kernel<<<(g->F + 31) / 32, 256, 0, st>>>(gws, nb, g->F, ss, chain, db);
            B200NN_LAUNCH_CHECK("convt_db_final");
        }
        const long long total4 = (long long)Mi * g->kh * g->kw * (g->F / 4);
        long long blocks = (total4 + 255) / 256;
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        convt_window_gather_kernel<<<(int)blocks, 256, 0, st>>>(*g, err, ss, chain, ecol, total4);   // pending clips folded in here
        B200NN_LAUNCH_CHECK("convt_window_gather");
        if (dw != nullptr) {  // dw[(ky,kx,f), c] = Ecol^T . x : both operands MN-major
            Epilogue e = plain_epilogue(dw, g->C, Kb);
            if (int rc = tc_gemm_any(ecol, Kb, true, x, g->C, true, e, Kb, g->C, Mi, gws, st, x3)) return rc;
        }
        if (dx != nullptr) {  // dx[pix, c] = Ecol . w[(ky,kx,f), c]
            Epilogue e = plain_epilogue(dx, g->C, Mi);
            dx_epilogue(e, nullptr, 0, x, prev_act, prev_alpha, ss_out0, ss_out1);
            if (int rc = tc_gemm_any(ecol, Kb, false, w, g->C, true, e, Mi, g->C, Kb, gws, st, x3)) return rc;
        }
        return B200NN_OK;
    }
    // windows over the (cropped) error, stride s, no pad offset, truncated windows skipped (layers.py:2645-2651)
    PatchGather p;
    p.src = err; p.SH = g->OH; p.SW = g->OW; p.SC = g->F; p.OH = g->H; p.OW = g->W;
    p.kh = g->kh; p.kw = g->kw; p.sh = g->sh; p.sw = g->sw; p.ph = 0; p.pw = 0;
    p.n_pix = Mi; p.n_tap = Kb; p.skip_truncated = 1;
    if (db != nullptr) {  // layers.py:2641: sum over ALL error pixels
        B200NN_REQUIRE(ws != nullptr, "convt_bwd: workspace required");
        const long long R = (long long)g->N * g->OH * g->OW;
        int nb = (int)((R + 63) / 64);
        if (nb > kNumSMs * 2) nb = kNumSMs * 2;
        if (colsum4r_ok(err, ws, g->F)) colsum_partial4r_kernel<<<nb, 256, 0, st>>>(err, R, g->F, ws);
        else colsum_partial_kernel<<<nb, 256, 0, st>>>(err, R, g->F, ws);
        B200NN_LAUNCH_CHECK("convt_db_partial");
        colsum_final_kernel<<<(g->F + 31) / 32, 256, 0, st>>>(ws, nb, g->F, ss, chain, db);
        B200NN_LAUNCH_CHECK("convt_db_final");
    }
    if (dw != nullptr) {  // dw[(ky,kx,f), c] = sum_pix window(pix, tap) * x[pix, c]
        PatchTOnesA af{p, 0};
        RowMajorB bf{x, g->C, Mi, g->C};
        Epilogue e = plain_epilogue(dw, g->C, Kb);
        e.ss = ss; e.chain = chain;
        if (int rc = launch_gemm("convt_dw", af, bf, e, Kb, g->C, Mi, ws, st)) return rc;
    }
    if (dx != nullptr) {  // dx[pix, c] = sum_tap window(pix, tap) * w[tap, c]
        PatchA af{p};
        RowMajorB bf{w, g->C, Kb, g->C};
        Epilogue e = plain_epilogue(dx, g->C, Mi);
        dx_epilogue(e, ss, chain, x, prev_act, prev_alpha, ss_out0, ss_out1);
        if (int rc = launch_gemm("convt_dx", af, bf, e, Mi, g->C, Kb, ws, st)) return rc;
    }
    return B200NN_OK;
}

}  // extern "C"
