// dense_skinny.cuh — Dense layers whose reduction dimension dwarfs the other two (K >> M, N; config 2's Dense(64) on the
// 6272 pooled features at batch 256): y[M,N] = x[M,K] . w[K,N],  dw[K,N] = x^T . e,  dx[M,K] = e . w^T   with M <= 256, N <= 64.
//
// A 64x64 output tile walked in 16-deep k-slabs (gemm_simt) pays one global-load latency per slab and, for the forward,
// needs 49-way split-K; here the work is cut along K instead: a CTA owns a SLAB of 32 (backward) or 64 (forward) input
// features, fetches everything it needs for that slab with ONE burst of 16-byte cp.async (all loads in flight at once), and
//   backward: produces the slab's 32 rows of dw AND its 32 columns of dx in the same kernel (they share the staged error
//             and input tiles), with the lazy clip multiplier, the preceding activation's backward and both clip slots fused;
//             no partials, no reduce pass;
//   forward : produces the slab's partial of the whole [M, N] output (reduced in fixed order by splitk_reduce_*).
// fp32 FFMA throughout (parity path). Included by gemm_simt.cu (uses Epilogue / launch_splitk_reduce).
#pragma once

namespace b200nn {
namespace skinny {

constexpr int MAXM = 256, MAXN = 64;
constexpr int BK_BWD = 24, BK_FWD = 44;   // slab widths (multiples of 4): 6272 features -> 262 / 143 CTAs on 148 SMs
constexpr int WS_STRIDE = MAXN + 4;   // padded weight-slab rows: conflict-free 128-bit reads of 8 consecutive rows

__device__ __forceinline__ void cp16(void* smem_dst, const void* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

static bool applicable(const float* x, const float* w, const float* other, int M, int N, int K) {
    return M <= MAXM && N <= MAXN && N % 4 == 0 && K % 4 == 0 && K >= 1024 &&
           ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(other)) & 15) == 0;
}

// ---- backward -----------------------------------------------------------------------------------------------------
// smem: Es[MAXM][MAXN] | Xs[MAXM][BK_BWD] | Ws[BK_BWD][WS_STRIDE]
constexpr size_t BWD_SMEM = (size_t)(MAXM * MAXN + MAXM * BK_BWD + BK_BWD * WS_STRIDE) * sizeof(float);

__global__ void __launch_bounds__(256) dense_skinny_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ err, const double* __restrict__ ss,
                                                               uint64_t chain, float* __restrict__ dx, float* __restrict__ dw,
                                                               float* __restrict__ db, int M, int N, int K, int prev_act,
                                                               float prev_alpha, double* ss0, double* ss1) {
    extern __shared__ __align__(16) float sk_smem[];
    __shared__ double scratch[32];
    float* Es = sk_smem;
    float* Xs = Es + MAXM * MAXN;
    float* Ws = Xs + MAXM * BK_BWD;
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * BK_BWD;
    const int kw = min(BK_BWD, K - k0);                 // multiple of 4
    const float scale = chain_scale(ss, chain);
    // ---- one burst of loads: error tile, input slab, weight slab
    const int n4 = N >> 2;
    for (int i = tid; i < M * n4; i += 256) cp16(Es + (i / n4) * MAXN + (i % n4) * 4, err + (long long)(i / n4) * N + (i % n4) * 4);
    for (int i = tid; i < M * (BK_BWD / 4); i += 256) {
        const int m = i / (BK_BWD / 4), c = (i % (BK_BWD / 4)) * 4;
        if (c < kw) cp16(Xs + m * BK_BWD + c, x + (long long)m * K + k0 + c);
        else *reinterpret_cast<float4*>(Xs + m * BK_BWD + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < BK_BWD * n4; i += 256) {
        const int r = i / n4, c = (i % n4) * 4;
        if (r < kw) cp16(Ws + r * WS_STRIDE + c, w + (long long)(k0 + r) * N + c);
        else *reinterpret_cast<float4*>(Ws + r * WS_STRIDE + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // zero the unused parts of the error tile (rows >= M never read; columns >= N are read by the 4-wide loops)
    if (N < MAXN)
        for (int i = tid; i < M * ((MAXN - N) >> 2); i += 256) {
            const int m = i / ((MAXN - N) >> 2), c = N + (i % ((MAXN - N) >> 2)) * 4;
            *reinterpret_cast<float4*>(Es + m * MAXN + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    cp_wait_all();
    __syncthreads();

    // ---- dw[k0 + r, n] = scale * sum_m x[m, k0 + r] * e[m, n]   (layers.py:196): thread = 2 rows x 4 columns
    if (dw != nullptr && tid < (BK_BWD / 2) * 16) {
        const int tk = tid >> 4, tn = tid & 15;          // rows 2tk, 2tk+1 ; columns 4tn .. 4tn+3
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
        for (int m = 0; m < M; ++m) {
            const float2 a = *reinterpret_cast<const float2*>(Xs + m * BK_BWD + 2 * tk);
            const float4 b = *reinterpret_cast<const float4*>(Es + m * MAXN + 4 * tn);
            acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
            acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
            acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
            acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
        }
        if (4 * tn < N) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = 2 * tk + i;
                if (r < kw)
                    *reinterpret_cast<float4*>(dw + (long long)(k0 + r) * N + 4 * tn) =
                        make_float4(scale * acc[i][0], scale * acc[i][1], scale * acc[i][2], scale * acc[i][3]);
            }
        }
    }
    if (dw != nullptr && blockIdx.x == 0 && db != nullptr && tid < N) {   // layers.py:197
        float s = 0.0f;
        for (int m = 0; m < M; ++m) s += Es[m * MAXN + tid];
        db[tid] = scale * s;
    }
    // ---- dx[m, k0 + c] = scale * sum_n e[m, n] * w[k0 + c, n]   (layers.py:189): thread = 8 rows x 4 columns
    if (dx != nullptr) {
        constexpr int NJ = BK_BWD / 8;
        const int tm = tid >> 3, tc = tid & 7;           // rows tm + 32 i ; columns tc + 8 j
        float acc[8][NJ];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j] = 0.0f;
        for (int n = 0; n < N; n += 4) {
            float4 bw[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bw[j] = *reinterpret_cast<const float4*>(Ws + (tc + 8 * j) * WS_STRIDE + n);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(Es + (tm + 32 * i) * MAXN + n);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    acc[i][j] = fmaf(a.x, bw[j].x, acc[i][j]); acc[i][j] = fmaf(a.y, bw[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a.z, bw[j].z, acc[i][j]); acc[i][j] = fmaf(a.w, bw[j].w, acc[i][j]);
                }
            }
        }
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = tm + 32 * i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int c = tc + 8 * j;
                if (c >= kw) continue;
                float v = scale * acc[i][j];
                a0 = fmaf(v, v, a0);
                if (prev_act != B200NN_ACT_NONE) {
                    v *= act_grad_from_y(prev_act, Xs[m * BK_BWD + c], prev_alpha);   // layers.py:237 of the preceding Activation
                    a1 = fmaf(v, v, a1);
                }
                dx[(long long)m * K + k0 + c] = v;
            }
        }
        if (ss0 != nullptr) block_accumulate(a0, ss0, scratch);
        if (prev_act != B200NN_ACT_NONE && ss1 != nullptr) block_accumulate(a1, ss1, scratch);
    }
}

// ---- forward: partial[cta][M][N] = x[:, slab] . w[slab, :] -----------------------------------------------------
// smem: Xs[MAXM][BK_FWD] | Ws[BK_FWD][WS_STRIDE]
constexpr size_t FWD_SMEM = (size_t)(MAXM * BK_FWD + BK_FWD * WS_STRIDE) * sizeof(float);

__global__ void __launch_bounds__(256) dense_skinny_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               float* __restrict__ partial, int M, int N, int K) {
    extern __shared__ __align__(16) float sk_smem[];
    float* Xs = sk_smem;
    float* Ws = Xs + MAXM * BK_FWD;
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * BK_FWD;
    const int kw = min(BK_FWD, K - k0);
    const int n4 = N >> 2;
    for (int i = tid; i < MAXM * (BK_FWD / 4); i += 256) {
        const int m = i / (BK_FWD / 4), c = (i % (BK_FWD / 4)) * 4;
        if (m < M && c < kw) cp16(Xs + m * BK_FWD + c, x + (long long)m * K + k0 + c);
        else *reinterpret_cast<float4*>(Xs + m * BK_FWD + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < BK_FWD * (MAXN / 4); i += 256) {
        const int r = i / (MAXN / 4), c = (i % (MAXN / 4)) * 4;
        if (r < kw && c < N) cp16(Ws + r * WS_STRIDE + c, w + (long long)(k0 + r) * N + c);
        else *reinterpret_cast<float4*>(Ws + r * WS_STRIDE + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_wait_all();
    __syncthreads();
    // thread = 16 rows (tm + 16 i) x 4 columns (4 tn ..)
    const int tm = tid >> 4, tn = tid & 15;
    float acc[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int k = 0; k < BK_FWD; k += 4) {
        float4 b[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4*>(Ws + (k + kk) * WS_STRIDE + 4 * tn);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(Xs + (tm + 16 * i) * BK_FWD + k);
            acc[i][0] = fmaf(a.x, b[0].x, acc[i][0]); acc[i][1] = fmaf(a.x, b[0].y, acc[i][1]);
            acc[i][2] = fmaf(a.x, b[0].z, acc[i][2]); acc[i][3] = fmaf(a.x, b[0].w, acc[i][3]);
            acc[i][0] = fmaf(a.y, b[1].x, acc[i][0]); acc[i][1] = fmaf(a.y, b[1].y, acc[i][1]);
            acc[i][2] = fmaf(a.y, b[1].z, acc[i][2]); acc[i][3] = fmaf(a.y, b[1].w, acc[i][3]);
            acc[i][0] = fmaf(a.z, b[2].x, acc[i][0]); acc[i][1] = fmaf(a.z, b[2].y, acc[i][1]);
            acc[i][2] = fmaf(a.z, b[2].z, acc[i][2]); acc[i][3] = fmaf(a.z, b[2].w, acc[i][3]);
            acc[i][0] = fmaf(a.w, b[3].x, acc[i][0]); acc[i][1] = fmaf(a.w, b[3].y, acc[i][1]);
            acc[i][2] = fmaf(a.w, b[3].z, acc[i][2]); acc[i][3] = fmaf(a.w, b[3].w, acc[i][3]);
        }
    }
    float* dst = partial + (long long)blockIdx.x * M * N;
    if (4 * tn < N) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int m = tm + 16 * i;
            if (m < M) *reinterpret_cast<float4*>(dst + (long long)m * N + 4 * tn) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
}

static int fwd_slabs(int K) { return (K + BK_FWD - 1) / BK_FWD; }
static size_t fwd_ws_bytes(int M, int N, int K) { return (size_t)fwd_slabs(K) * M * N * sizeof(float); }

static int launch_fwd(const float* x, const float* w, const Epilogue& epi, int M, int N, int K, float* ws, cudaStream_t st) {
    static thread_local bool configured = false;
    if (!configured) {
        B200NN_CUDA(cudaFuncSetAttribute(dense_skinny_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
        configured = true;
    }
    const int slabs = fwd_slabs(K);
    dense_skinny_fwd_kernel<<<slabs, 256, FWD_SMEM, st>>>(x, w, ws, M, N, K);
    B200NN_LAUNCH_CHECK("dense_skinny_fwd");
    return launch_splitk_reduce(ws, slabs, M, N, epi, st);
}

static int launch_bwd(const float* x, const float* w, const float* err, const double* ss, uint64_t chain, float* dx, float* dw,
                      float* db, int M, int N, int K, int prev_act, float prev_alpha, double* ss0, double* ss1, cudaStream_t st) {
    static thread_local bool configured = false;
    if (!configured) {
        B200NN_CUDA(cudaFuncSetAttribute(dense_skinny_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
        configured = true;
    }
    dense_skinny_bwd_kernel<<<(K + BK_BWD - 1) / BK_BWD, 256, BWD_SMEM, st>>>(x, w, err, ss, chain, dx, dw, db, M, N, K, prev_act,
                                                                             prev_alpha, ss0, ss1);
    B200NN_LAUNCH_CHECK("dense_skinny_bwd");
    return B200NN_OK;
}

}  // namespace skinny
}  // namespace b200nn
