// f16_split.cu — operand preparation of the fp16-plane tensor-core mode (gemm_tc2.cuh F16; math mode TF32X3 on large Dense
// contractions, layers.py:173 / :189 / :196). A fp32 matrix X is rewritten ONCE per contraction as two K-major fp16 planes
//     s X = HI + 2^-11 LO,   HI = rn_fp16(s X),  LO = rn_fp16(2^11 (s X - HI)),   s = 2^(14 - floor(log2 max|X|))
// so that the three products HI.HI + 2^-11 (LO.HI + HI.LO) carry 22 mantissa bits per operand (the precision of the 3xTF32
// scheme) while the tensor core runs kind::f16 at twice the TF32 rate. The power-of-two scale keeps every |s X| below 2^15
// (fp16 range) and is undone in the GEMM epilogue; elements more than 2^29 below the tensor's maximum lose bits gradually —
// an absolute error below 2^-40 max|X|, irrelevant at the norm-relative 1e-5 bar. HBM-bound: 12 bytes per element
// (max pass 4 read; split pass 4 read + 4 written), with the transposing variant staged through shared memory so that both sides
// move full 128-byte lines.
#include <cuda_fp16.h>

#include "gemm_tc_api.cuh"

namespace b200nn {
namespace f16 {

__global__ void __launch_bounds__(256) maxabs_kernel(const float* __restrict__ x, long long ld, int R, int C, unsigned* __restrict__ out) {
    __shared__ unsigned sm[8];
    const int C4 = C >> 2;
    const long long total = (long long)R * C4;
    float m = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / C4;
        const int c = (int)(i - r * C4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = __float_as_uint(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned r = 0;
        for (int i = 0; i < 8; ++i) r = max(r, sm[i]);       // non-negative floats order like their bit patterns
        if (r != 0 && r < 0x7f800000u) atomicMax(out, r);     // NaN / Inf never become the scale
    }
}

// scale s (and 1 / s) from the bit pattern of max|X|
__device__ __forceinline__ float plane_scale(unsigned maxbits, float& inv) {
    int E = (int)(maxbits >> 23) - 127;
    if (maxbits == 0) E = 14;
    E = max(-100, min(100, E));
    inv = __uint_as_float((unsigned)(127 + E - 14) << 23);
    return __uint_as_float((unsigned)(127 + 14 - E) << 23);
}
__device__ __forceinline__ void split1(float x, float s, __half& hi, __half& lo) {
    const float v = x * s;
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * 2048.0f);
}

// planes[(p * rows + r) * pitch + c] for src[r][c] (row pitch ld): K-major as stored
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, long long ld, int R, int C, __half* __restrict__ planes,
                                                         long long pitch, const unsigned* __restrict__ maxbits, float* __restrict__ inv_out) {
    float inv;
    const float s = plane_scale(*maxbits, inv);
    if (blockIdx.x == 0 && threadIdx.x == 0) *inv_out = inv;
    const int C4 = C >> 2;
    const long long total = (long long)R * C4;
    __half* lo_plane = planes + (long long)R * pitch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / C4;
        const int c = (int)(i - r * C4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
        __half h[4], l[4];
        split1(v.x, s, h[0], l[0]); split1(v.y, s, h[1], l[1]); split1(v.z, s, h[2], l[2]); split1(v.w, s, h[3], l[3]);
        *reinterpret_cast<uint2*>(planes + r * pitch + c) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo_plane + r * pitch + c) = *reinterpret_cast<const uint2*>(l);
    }
}

// planes[(p * C + c) * pitch + r] for src[r][c]: the transposed (K-major over the source's rows) planes, 64 x 64 tiles
__global__ void __launch_bounds__(256) split_cols_kernel(const float* __restrict__ x, long long ld, int R, int C, __half* __restrict__ planes,
                                                         long long pitch, const unsigned* __restrict__ maxbits, float* __restrict__ inv_out) {
    __shared__ float tile[64][65];
    float inv;
    const float s = plane_scale(*maxbits, inv);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *inv_out = inv;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16: a float4 of columns, rows ty + 16 j
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + ty + 16 * j, c = c0 + tx * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < R && c < C) v = *reinterpret_cast<const float4*>(x + (long long)r * ld + c);     // C % 4 == 0
        tile[ty + 16 * j][tx * 4 + 0] = v.x; tile[ty + 16 * j][tx * 4 + 1] = v.y;
        tile[ty + 16 * j][tx * 4 + 2] = v.z; tile[ty + 16 * j][tx * 4 + 3] = v.w;
    }
    __syncthreads();
    __half* lo_plane = planes + (long long)C * pitch;
    const int px = threadIdx.x & 31, py = threadIdx.x >> 5;       // 32 x 8: a pair of source rows, columns py + 8 j
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + py + 8 * j, r = r0 + px * 2;
        if (c < C && r < R) {                                     // R % 2 == 0 is required by the caller (pitch % 8 == 0 anyway)
            __half h[2], l[2];
            split1(tile[px * 2][py + 8 * j], s, h[0], l[0]);
            split1(r + 1 < R ? tile[px * 2 + 1][py + 8 * j] : 0.0f, s, h[1], l[1]);
            *reinterpret_cast<unsigned*>(planes + (long long)c * pitch + r) = *reinterpret_cast<const unsigned*>(h);
            *reinterpret_cast<unsigned*>(lo_plane + (long long)c * pitch + r) = *reinterpret_cast<const unsigned*>(l);
        }
    }
}

// forward-convolution weights w[(c, tap), F] (the reference's K order, layers.py:470) -> K-major planes [F][(tap, c)]: the k order
// of the implicit-GEMM producer (one tap's channels are consecutive); one thread per plane element, a few hundred KB at most
__global__ void __launch_bounds__(256) split_conv_w_fwd_kernel(const float* __restrict__ w, int C, int taps, int F, __half* __restrict__ planes,
                                                               const unsigned* __restrict__ maxbits, float* __restrict__ inv_out) {
    float inv;
    const float s = plane_scale(*maxbits, inv);
    if (blockIdx.x == 0 && threadIdx.x == 0) *inv_out = inv;
    const long long K = (long long)taps * C, total = K * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / K), k = (int)(i - (long long)f * K), tap = k / C, c = k - tap * C;
        __half h, l;
        split1(__ldg(w + ((long long)c * taps + tap) * F + f), s, h, l);
        planes[i] = h;
        planes[total + i] = l;
    }
}

static inline long long pitch_of(long long k) { return (k + 7) / 8 * 8; }

}  // namespace f16

namespace tcapi {

// workspace floats needed beside the split-K partials: both operands' planes + the scale words
size_t f16_extra_bytes(long long M, long long N, long long K) {
    const long long p = f16::pitch_of(K);
    return (size_t)(2 * (M + N) * p) * sizeof(__half) + 1024;
}

bool f16_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200NN_F16X3");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1 && tc2_enabled();
}

// worth it where the halved tensor time outweighs the 12 B / element of operand preparation
bool f16_worth(long long M, long long N, long long K) {
    return M >= 1024 && N >= 1024 && K >= 1024 && M * N * K >= (1LL << 33) && (M % 4 == 0) && (N % 4 == 0) && (K % 4 == 0);
}

// one operand: fp32 [rows_src][cols_src] (pitch ld) -> planes with `mn` rows and K = the other dimension
// premax (optional): a word that already holds the bit pattern of max |src| (the max pass is skipped)
static int prepare(const float* src, long long ld, int mn, int K, bool transposed, __half* planes, unsigned* maxbits_own, float* inv,
                   cudaStream_t st, const unsigned* premax = nullptr) {
    const int R = transposed ? K : mn, C = transposed ? mn : K;      // source shape
    long long blocks = ((long long)R * (C / 4) + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    const unsigned* maxbits = premax != nullptr ? premax : maxbits_own;
    if (premax == nullptr) {
        B200NN_CUDA(cudaMemsetAsync(maxbits_own, 0, sizeof(unsigned), st));
        f16::maxabs_kernel<<<(int)blocks, 256, 0, st>>>(src, ld, R, C, maxbits_own);
        B200NN_LAUNCH_CHECK("f16_maxabs");
    }
    const long long pitch = f16::pitch_of(K);
    if (!transposed) {
        f16::split_rows_kernel<<<(int)blocks, 256, 0, st>>>(src, ld, R, C, planes, pitch, maxbits, inv);
        B200NN_LAUNCH_CHECK("f16_split_rows");
    } else {
        const dim3 grid((C + 63) / 64, (R + 63) / 64);
        f16::split_cols_kernel<<<grid, 256, 0, st>>>(src, ld, R, C, planes, pitch, maxbits, inv);
        B200NN_LAUNCH_CHECK("f16_split_cols");
    }
    return B200NN_OK;
}

// D[M,N] = A.B with fp32-grade products on the fp16 tensor pipe. a_mn: A stored [K][M] (ld = lda) else [M][K]; b_mn: B stored
// [K][N] else [N][K]. `ws`: ws2(M, N, ceil(K/2), x3) bytes of split-K partials followed by f16_extra_bytes(M, N, K).
int gemm2_f16x3(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi, int M, int N,
                int K, float* ws, cudaStream_t st, const unsigned* premax_a, const unsigned* premax_b) {
    B200NN_REQUIRE(ws != nullptr, "gemm2_f16x3: needs a workspace");
    const size_t part = (ws2(M, N, (K + 1) / 2, true) + 255) / 256 * 256;
    char* base = reinterpret_cast<char*>(ws) + part;
    const long long pitch = f16::pitch_of(K);
    __half* pa = reinterpret_cast<__half*>(base);
    __half* pb = pa + 2LL * M * pitch;
    unsigned* words = reinterpret_cast<unsigned*>(pb + 2LL * N * pitch);       // [0] max bits of A, [1] of B, [2..3] inverse scales
    float* inv = reinterpret_cast<float*>(words + 2);
    if (int rc = prepare(A, lda, M, K, a_mn, pa, words, inv, st, premax_a)) return rc;
    if (int rc = prepare(B, ldb, N, K, b_mn, pb, words + 1, inv + 1, st, premax_b)) return rc;
    return gemm2_f16_planes(pa, pitch, M, pb, pitch, N, inv, epi, M, N, K, ws, st);
}

// ---- implicit-GEMM convolution (cfg3's conv2 / conv3 in TF32X3 mode: layers.py:471-477 forward, :513-528 data gradient) ------------
// 0: off, 1 (default): where it pays, 2: wherever the geometry allows (tests)
static int conv_f16_mode() {
    const char* e = getenv("B200NN_F16CONV");
    return e != nullptr ? atoi(e) : 1;
}

bool conv_f16_worth(int which, const b200nn_conv_geom* g) {
    const int mode = conv_f16_mode();
    if (mode == 0 || !f16_enabled() || (which != 1 && which != 2)) return false;
    const long long Ck = which == 1 ? g->C : g->F, Nn = which == 1 ? g->F : g->C;
    if (Ck % 64 != 0) return false;
    const long long P = (long long)g->N * g->H * g->W, K = (long long)g->kh * g->kw * Ck;
    return mode >= 2 || P * Nn * K >= (1LL << 32);
}

// both operands' planes (the larger of the forward and the data-gradient pair) + the scale words
size_t conv_f16_extra_bytes(const b200nn_conv_geom* g) {
    const long long P = (long long)g->N * g->H * g->W, taps = (long long)g->kh * g->kw;
    const long long cmax = g->C > g->F ? g->C : g->F;
    return (size_t)(2 * P * cmax + 2 * taps * g->C * g->F) * sizeof(__half) + 1024;
}

int conv2_f16x3(int which, const b200nn_conv_geom* g, const float* a, const float* w, const Epilogue& epi, float* ws, cudaStream_t st,
                const unsigned* premax_a) {
    B200NN_REQUIRE(ws != nullptr, "conv2_f16x3: needs a workspace");
    const int taps = g->kh * g->kw, Ck = which == 1 ? g->C : g->F, Nn = which == 1 ? g->F : g->C;
    const long long P = (long long)g->N * g->H * g->W, K = (long long)taps * Ck;
    const size_t part = (ws2(P, Nn, K / 2, true) + 255) / 256 * 256;
    char* base = reinterpret_cast<char*>(ws) + part;
    __half* pa = reinterpret_cast<__half*>(base);                              // [2 N][H][W][Ck]
    __half* pb = pa + 2 * P * Ck;                                              // [2 Nn][K]
    unsigned* words = reinterpret_cast<unsigned*>(pb + 2 * (long long)Nn * K);
    float* inv = reinterpret_cast<float*>(words + 2);
    if (int rc = prepare(a, Ck, (int)P, Ck, false, pa, words, inv, st, premax_a)) return rc;
    if (which == 2) {
        // err . W^T contracts over (tap, f): the weights as stored, seen as the matrix [C][(tap, f)]
        if (int rc = prepare(w, K, Nn, (int)K, false, pb, words + 1, inv + 1, st)) return rc;
    } else {
        B200NN_CUDA(cudaMemsetAsync(words + 1, 0, sizeof(unsigned), st));
        const long long total = K * Nn;
        int blocks = (int)((total / 4 + 255) / 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        f16::maxabs_kernel<<<blocks, 256, 0, st>>>(w, g->F, (int)K, g->F, words + 1);
        B200NN_LAUNCH_CHECK("f16_maxabs");
        blocks = (int)((total + 255) / 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        f16::split_conv_w_fwd_kernel<<<blocks, 256, 0, st>>>(w, g->C, taps, g->F, pb, words + 1, inv + 1);
        B200NN_LAUNCH_CHECK("f16_split_conv_w");
    }
    return conv2_f16_planes(which, g, pa, pb, K, inv, epi, ws, st);
}

}  // namespace tcapi
}  // namespace b200nn
