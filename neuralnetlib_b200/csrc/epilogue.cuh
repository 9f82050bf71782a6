// epilogue.cuh — the epilogue shared by the FFMA contraction engine (gemm_simt.cu) and the tcgen05 kernels (gemm_tcgen05.cuh,
// gemm_tc2.cuh): bias + activation (forward), lazy clip multiplier, the preceding activation's backward and the fp64
// sum-of-squares slots (backward), conv weight-gradient row permutation.
#pragma once
#include "common.cuh"

namespace b200nn {

// ------------------------------------------------------------------------------------------------
// epilogue
// ------------------------------------------------------------------------------------------------
struct Epilogue {
    float* out;                // out[row * ldo + n]
    long long ldo;
    const float* bias;         // per n, nullable
    int act;                   // forward activation applied after bias
    float alpha;
    const double* ss;          // lazy clip multiplier applied to the accumulator (linear ops)
    uint64_t chain;
    const float* mask_y;       // preceding activation's OUTPUT (same indexing as out), nullable
    int mask_act;
    float mask_alpha;
    double* ss0;               // += sum(v^2) before the mask
    double* ss1;               // += sum(v^2) after the mask
    // weight-gradient row handling: GEMM rows are taps in (ky,kx,c) order (+ an optional trailing
    // "ones" row that carries db); row_mode 1 maps tap rows to the reference's (c,ky,kx) flat order.
    int row_mode;              // 0 plain, 1 conv-weight permute
    int kh, kw, C;
    int n_rows;                // number of real rows (rows >= n_rows go to db)
    float* db;                 // nullable
};

struct EpiState {
    float scale;
    float a0, a1;
};

__device__ __forceinline__ void epi_begin(const Epilogue& e, EpiState& st) {
    st.scale = chain_scale(e.ss, e.chain);
    st.a0 = 0.0f;
    st.a1 = 0.0f;
}

__device__ __forceinline__ void epi_store(const Epilogue& e, EpiState& st, int m, int n, float v) {
    v *= st.scale;
    if (m >= e.n_rows) {  // ones row -> db
        if (e.db != nullptr) e.db[n] = v;
        return;
    }
    long long row = m;
    if (e.row_mode == 1) {
        const int c = m % e.C;
        const int t = m / e.C;
        const int kx = t % e.kw, ky = t / e.kw;
        row = ((long long)c * e.kh + ky) * e.kw + kx;
    }
    if (e.bias != nullptr) v += e.bias[n];
    v = act_apply(e.act, v, e.alpha);
    const long long idx = row * e.ldo + n;
    st.a0 += v * v;
    if (e.mask_y != nullptr) {
        v *= act_grad_from_y(e.mask_act, e.mask_y[idx], e.mask_alpha);
        st.a1 += v * v;
    }
    e.out[idx] = v;
}

// same as epi_store for the plain row mode, with the activation-backward mask value already loaded by the caller
// (lets the caller batch the mask loads of several rows ahead of their use)
__device__ __forceinline__ void epi_store_premask(const Epilogue& e, EpiState& st, int m, int n, float v, float mask_val) {
    v *= st.scale;
    if (e.bias != nullptr) v += e.bias[n];
    v = act_apply(e.act, v, e.alpha);
    st.a0 += v * v;
    if (e.mask_y != nullptr) {
        v *= act_grad_from_y(e.mask_act, mask_val, e.mask_alpha);
        st.a1 += v * v;
    }
    e.out[(long long)m * e.ldo + n] = v;
}

__device__ __forceinline__ void epi_finish(const Epilogue& e, EpiState& st, double* scratch) {
    if (e.ss0 != nullptr) block_accumulate(st.a0, e.ss0, scratch);
    if (e.ss1 != nullptr) block_accumulate(st.a1, e.ss1, scratch);
}


// out = reduce over `splits` partial [M, N] matrices in `ws` (fixed order), then the epilogue (gemm_simt.cu)
int launch_splitk_reduce(const float* ws, int splits, int M, int N, const Epilogue& epi, cudaStream_t st);

static inline Epilogue plain_epilogue(float* out, long long ldo, int n_rows) {
    Epilogue e;
    e.out = out; e.ldo = ldo; e.bias = nullptr; e.act = B200NN_ACT_NONE; e.alpha = 0.f;
    e.ss = nullptr; e.chain = 0; e.mask_y = nullptr; e.mask_act = B200NN_ACT_NONE; e.mask_alpha = 0.f;
    e.ss0 = nullptr; e.ss1 = nullptr; e.row_mode = 0; e.kh = e.kw = e.C = 1; e.n_rows = n_rows; e.db = nullptr;
    return e;
}

}  // namespace b200nn
