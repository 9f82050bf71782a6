// bn_pool.cu — BatchNormalization (per-POSITION statistics over the batch axis, layers.py:1230-1276)
// and MaxPooling2D (layers.py:570-643). HBM-bound column reductions over x viewed as [B, P]:
// a CTA owns a strip of 32 adjacent positions (one 128-byte line per batch row), 8 row-groups of 32 lanes
// walk the batch, so every global access is a fully coalesced 128 B line and the strip (B x 128 B) stays
// in L1/L2 for the second and third sweep (exact two-pass variance, no E[x^2]-E[x]^2 cancellation).
#include "common.cuh"

namespace b200nn {

constexpr int BN_COLS = 32;
constexpr int BN_ROWG = 8;  // 256 threads

__device__ __forceinline__ float clip10(float v) { return fminf(fmaxf(v, -10.0f), 10.0f); }

// reduce `v` across the 8 row-groups of a column; result broadcast to every row-group
__device__ __forceinline__ float colgroup_sum(float v, float (*sm)[BN_COLS], int cx, int ry) {
    __syncthreads();
    sm[ry][cx] = v;
    __syncthreads();
    float s = 0.0f;
#pragma unroll
    for (int r = 0; r < BN_ROWG; ++r) s += sm[r][cx];
    return s;
}

__global__ void __launch_bounds__(256) bn_fwd_train_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ rmean,
                                                           float* __restrict__ rvar, float* __restrict__ save_mean,
                                                           float* __restrict__ save_invstd, float* __restrict__ y,
                                                           int B, long long P, float momentum, float eps) {
    __shared__ float sm[BN_ROWG][BN_COLS];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const long long col = (long long)blockIdx.x * BN_COLS + cx;
    const bool ok = col < P;
    const float* xc = x + col;
    float s = 0.0f;
    if (ok) {
#pragma unroll 4
        for (int r = ry; r < B; r += BN_ROWG) s += clip10(xc[(long long)r * P]);
    }
    const float mean = colgroup_sum(s, sm, cx, ry) / (float)B;  // layers.py:1237
    float q = 0.0f;
    if (ok) {
#pragma unroll 4
        for (int r = ry; r < B; r += BN_ROWG) {
            const float d = clip10(xc[(long long)r * P]) - mean;
            q += d * d;
        }
    }
    const float batch_var = colgroup_sum(q, sm, cx, ry) / (float)B + eps;  // layers.py:1238 (var + eps)
    const float invstd = rsqrtf(batch_var + eps);                           // layers.py:1252 (second eps)
    if (!ok) return;
    if (ry == 0) {
        rmean[col] = momentum * rmean[col] + (1.0f - momentum) * mean;      // layers.py:1240-1243
        rvar[col] = momentum * rvar[col] + (1.0f - momentum) * batch_var;
        save_mean[col] = mean;
        save_invstd[col] = invstd;
    }
    const float g = gamma[col] * invstd, b = beta[col];
    float* yc = y + col;
#pragma unroll 4
    for (int r = ry; r < B; r += BN_ROWG) yc[(long long)r * P] = g * (clip10(xc[(long long)r * P]) - mean) + b;
}

__global__ void __launch_bounds__(256) bn_fwd_infer_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const float* __restrict__ rmean,
                                                           const float* __restrict__ rvar, float* __restrict__ y,
                                                           long long total, long long P, float eps) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long c = i % P;
        y[i] = gamma[c] * ((clip10(x[i]) - rmean[c]) / sqrtf(rvar[c] + eps)) + beta[c];  // layers.py:1248-1256
    }
}

__global__ void __launch_bounds__(256) bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ err,
                                                     const double* __restrict__ ss, uint64_t chain,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ save_mean,
                                                     const float* __restrict__ save_invstd, float* __restrict__ dx,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int B,
                                                     long long P, int prev_act, float prev_alpha, double* ss_out0,
                                                     double* ss_out1) {
    __shared__ float sm[BN_ROWG][BN_COLS];
    __shared__ double scratch[32];
    const float scale = chain_scale(ss, chain);
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const long long col = (long long)blockIdx.x * BN_COLS + cx;
    const bool ok = col < P;
    const float mean = ok ? save_mean[col] : 0.0f, invstd = ok ? save_invstd[col] : 0.0f;
    const float* xc = x + col;
    const float* ec = err + col;
    float se = 0.0f, sex = 0.0f;
    if (ok) {
#pragma unroll 4
        for (int r = ry; r < B; r += BN_ROWG) {
            const float e = scale * ec[(long long)r * P];
            const float xh = (clip10(xc[(long long)r * P]) - mean) * invstd;
            se += e;
            sex += e * xh;
        }
    }
    const float sum_e = colgroup_sum(se, sm, cx, ry);       // d_beta   (layers.py:1262)
    const float sum_ex = colgroup_sum(sex, sm, cx, ry);     // d_gamma  (layers.py:1261)
    float a0 = 0.0f, a1 = 0.0f;
    if (ok) {
        if (ry == 0) {
            if (dgamma != nullptr) dgamma[col] = sum_ex;
            if (dbeta != nullptr) dbeta[col] = sum_e;
        }
        // layers.py:1264-1274 in closed form: dx = gamma*invstd*(e - xhat*sum(e*xhat)/B - sum(e)/B)
        const float gi = gamma[col] * invstd, invB = 1.0f / (float)B;
        float* dc = dx + col;
#pragma unroll 4
        for (int r = ry; r < B; r += BN_ROWG) {
            const float xv = xc[(long long)r * P];
            const float e = scale * ec[(long long)r * P];
            const float xh = (clip10(xv) - mean) * invstd;
            float d = gi * (e - xh * sum_ex * invB - sum_e * invB);
            a0 += d * d;
            if (prev_act != B200NN_ACT_NONE) {
                d *= act_grad_from_y(prev_act, xv, prev_alpha);
                a1 += d * d;
            }
            dc[(long long)r * P] = d;
        }
    }
    if (ss_out0 != nullptr) block_accumulate(a0, ss_out0, scratch);
    if (prev_act != B200NN_ACT_NONE && ss_out1 != nullptr) block_accumulate(a1, ss_out1, scratch);
}

// ------------------------------------------------------------------------------------------------ MaxPool
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(b200nn_pool_geom g, const float* __restrict__ x,
                                                          float* __restrict__ y, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % g.C); long long t = i / g.C;
        const int ox = (int)(t % g.OW); t /= g.OW;
        const int oy = (int)(t % g.OH); const int n = (int)(t / g.OH);
        const float* xi = x + (long long)n * g.H * g.W * g.C + c;
        float m = -INFINITY;
        for (int dy = 0; dy < g.ph; ++dy) {
            const int iy = oy * g.sh + dy - g.pad_h;
            for (int dxx = 0; dxx < g.pw; ++dxx) {
                const int ix = ox * g.sw + dxx - g.pad_w;
                const float v = (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) ? xi[((long long)iy * g.W + ix) * g.C] : 0.0f;
                m = fmaxf(m, v);  // zero padding takes part in the max (layers.py:582-585)
            }
        }
        y[i] = m;
    }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(b200nn_pool_geom g, const float* __restrict__ x,
                                                          const float* __restrict__ y, const float* __restrict__ err,
                                                          const double* __restrict__ ss, uint64_t chain,
                                                          float* __restrict__ dx, double* ss_out, long long total) {
    __shared__ double scratch[32];
    const float scale = chain_scale(ss, chain);
    float acc = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % g.C); long long t = i / g.C;
        const int w = (int)(t % g.W); t /= g.W;
        const int h = (int)(t % g.H); const int n = (int)(t / g.H);
        const float xv = x[i];
        const int hp = h + g.pad_h, wp = w + g.pad_w;
        // windows (oy, ox) with oy*sh <= hp < oy*sh + ph
        int oy_lo = hp - g.ph + 1; oy_lo = oy_lo <= 0 ? 0 : (oy_lo + g.sh - 1) / g.sh;
        int oy_hi = hp / g.sh; if (oy_hi > g.OH - 1) oy_hi = g.OH - 1;
        int ox_lo = wp - g.pw + 1; ox_lo = ox_lo <= 0 ? 0 : (ox_lo + g.sw - 1) / g.sw;
        int ox_hi = wp / g.sw; if (ox_hi > g.OW - 1) ox_hi = g.OW - 1;
        float d = 0.0f;
        for (int oy = oy_lo; oy <= oy_hi; ++oy)
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                const long long o = (((long long)n * g.OH + oy) * g.OW + ox) * g.C + c;
                if (xv == y[o]) d += err[o];  // every tied maximum gets the gradient (layers.py:631-637)
            }
        d *= scale;
        acc += d * d;
        dx[i] = d;
    }
    if (ss_out != nullptr) block_accumulate(acc, ss_out, scratch);
}

}  // namespace b200nn

using namespace b200nn;

extern "C" {

int b200nn_bn_fwd_train(const float* x, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, float* save_mean, float* save_invstd, float* y, int B, long long P,
                        float momentum, float eps, void* stream) {
    B200NN_REQUIRE(B > 0 && P > 0, "bn_fwd_train: bad shape B=%d P=%lld", B, P);
    const long long blocks = (P + BN_COLS - 1) / BN_COLS;
    B200NN_REQUIRE(blocks < (1LL << 31), "bn_fwd_train: too many positions");
    bn_fwd_train_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, gamma, beta, running_mean, running_var, save_mean,
                                                                  save_invstd, y, B, P, momentum, eps);
    B200NN_LAUNCH_CHECK("bn_fwd_train");
    return B200NN_OK;
}

int b200nn_bn_fwd_infer(const float* x, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float* y, int B, long long P, float eps, void* stream) {
    B200NN_REQUIRE(B > 0 && P > 0, "bn_fwd_infer: bad shape B=%d P=%lld", B, P);
    const long long total = (long long)B * P;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    bn_fwd_infer_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, gamma, beta, running_mean, running_var, y, total, P, eps);
    B200NN_LAUNCH_CHECK("bn_fwd_infer");
    return B200NN_OK;
}

int b200nn_bn_bwd(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                  const float* save_mean, const float* save_invstd, float* dx, float* dgamma, float* dbeta, int B,
                  long long P, int prev_act, float prev_alpha, double* ss_out0, double* ss_out1, void* stream) {
    B200NN_REQUIRE(B > 0 && P > 0, "bn_bwd: bad shape B=%d P=%lld", B, P);
    B200NN_REQUIRE(prev_act >= 0 && prev_act <= B200NN_ACT_LEAKYRELU, "bn_bwd: unknown activation %d", prev_act);
    const long long blocks = (P + BN_COLS - 1) / BN_COLS;
    bn_bwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, err, ss, chain, gamma, save_mean, save_invstd, dx, dgamma,
                                                            dbeta, B, P, prev_act, prev_alpha, ss_out0, ss_out1);
    B200NN_LAUNCH_CHECK("bn_bwd");
    return B200NN_OK;
}

static int check_pool_geom(const char* who, const b200nn_pool_geom* g) {
    B200NN_REQUIRE(g != nullptr, "%s: null geometry", who);
    B200NN_REQUIRE(g->N > 0 && g->H > 0 && g->W > 0 && g->C > 0 && g->ph > 0 && g->pw > 0 && g->sh > 0 && g->sw > 0 &&
                       g->pad_h >= 0 && g->pad_w >= 0,
                   "%s: bad geometry", who);
    B200NN_REQUIRE(g->OH == (g->H + 2 * g->pad_h - g->ph) / g->sh + 1 && g->OW == (g->W + 2 * g->pad_w - g->pw) / g->sw + 1 &&
                       g->OH > 0 && g->OW > 0,
                   "%s: output size inconsistent with geometry", who);
    return B200NN_OK;
}

int b200nn_maxpool_fwd(const b200nn_pool_geom* g, const float* x, float* y, void* stream) {
    if (int rc = check_pool_geom("maxpool_fwd", g)) return rc;
    const long long total = (long long)g->N * g->OH * g->OW * g->C;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    maxpool_fwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(*g, x, y, total);
    B200NN_LAUNCH_CHECK("maxpool_fwd");
    return B200NN_OK;
}

int b200nn_maxpool_bwd(const b200nn_pool_geom* g, const float* x, const float* y, const float* err,
                       const double* ss, uint64_t chain, float* dx, double* ss_out, void* stream) {
    if (int rc = check_pool_geom("maxpool_bwd", g)) return rc;
    const long long total = (long long)g->N * g->H * g->W * g->C;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    maxpool_bwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(*g, x, y, err, ss, chain, dx, ss_out, total);
    B200NN_LAUNCH_CHECK("maxpool_bwd");
    return B200NN_OK;
}

}  // extern "C"
