// bn_pool.cu — BatchNormalization (per-POSITION statistics over the batch axis, layers.py:1230-1276)
// and MaxPooling2D (layers.py:570-643). HBM-bound column reductions over x viewed as [B, P]:
// a CTA owns a strip of 32 adjacent positions (one 128-byte line per batch row), 8 row-groups of 32 lanes
// walk the batch, so every global access is a fully coalesced 128 B line and the strip (B x 128 B) stays
// in L1/L2 for the second and third sweep (exact two-pass variance, no E[x^2]-E[x]^2 cancellation).
#include "common.cuh"

namespace b200nn {

constexpr int BN_COLS = 32;
constexpr int BN_ROWG = 8;  // 256 threads


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

// ---- data-parallel split of the forward (comm.cu): local partial sums -> sum over ranks -> apply -------------------
// stats[col] = sum_rows clip(x), stats[P + col] = sum_rows clip(x)^2, both fp64. The second moment is formed from the
// exact two-pass local M2 (M2 + S^2/B in fp64), so the only cancellation left is the fp64 one in the combine.
__global__ void __launch_bounds__(256) bn_stats_partial_kernel(const float* __restrict__ x, double* __restrict__ stats, int B,
                                                               long long P) {
    __shared__ float sm[BN_ROWG][BN_COLS];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const long long col = (long long)blockIdx.x * BN_COLS + cx;
    const bool ok = col < P;
    const float* xc = x + col;
    float s = 0.0f;
    if (ok) {
#pragma unroll 8
        for (int r = ry; r < B; r += BN_ROWG) s += clip10(xc[(long long)r * P]);
    }
    const float sum = colgroup_sum(s, sm, cx, ry);
    const float mean = sum / (float)B;
    float q = 0.0f;
    if (ok) {
#pragma unroll 8
        for (int r = ry; r < B; r += BN_ROWG) {
            const float d = clip10(xc[(long long)r * P]) - mean;
            q += d * d;
        }
    }
    const float m2 = colgroup_sum(q, sm, cx, ry);
    if (ok && ry == 0) {
        const double S = (double)mean * (double)B;
        stats[col] = S;
        stats[P + col] = (double)m2 + S * S / (double)B;
    }
}

__global__ void __launch_bounds__(256) bn_fwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ rmean,
                                                           float* __restrict__ rvar, float* __restrict__ save_mean,
                                                           float* __restrict__ save_invstd, const double* __restrict__ stats,
                                                           float* __restrict__ y, int B, long long P, double n_total,
                                                           float momentum, float eps) {
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const long long col = (long long)blockIdx.x * BN_COLS + cx;
    if (col >= P) return;
    float mean, batch_var, invstd;
    bn_finalize(stats[col], stats[P + col], n_total, eps, mean, batch_var, invstd);
    if (ry == 0) {
        rmean[col] = momentum * rmean[col] + (1.0f - momentum) * mean;
        rvar[col] = momentum * rvar[col] + (1.0f - momentum) * batch_var;
        save_mean[col] = mean;
        save_invstd[col] = invstd;
    }
    const float g = gamma[col] * invstd, b = beta[col];
    const float* xc = x + col;
    float* yc = y + col;
#pragma unroll 4
    for (int r = ry; r < B; r += BN_ROWG) yc[(long long)r * P] = g * (clip10(xc[(long long)r * P]) - mean) + b;
}

// BatchNorm apply + MaxPool 2x2/s2 from rank-summed moments: a thread owns (window, channel quad) and `rows_per_chunk` batch rows
// (8, or 32 at large batch: the prologue finalises 16 positions' moments in fp64 and is worth amortising)
constexpr int FB_APPLY_ROWS = 8;
__global__ void __launch_bounds__(256) bn_pool_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ rmean,
                                                            float* __restrict__ rvar, float* __restrict__ save_mean,
                                                            float* __restrict__ save_invstd, const double* __restrict__ stats,
                                                            float* __restrict__ y, int B, int H, int W, int C, double n_total,
                                                            float momentum, float eps, long long total, int rows_per_chunk) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int CQ = C >> 2, OW = W >> 1, OH = H >> 1;
    const int cq = (int)(idx % CQ); long long t = idx / CQ;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH); const int chunk = (int)(t / OH);
    const int c = cq * 4;
    const long long P = (long long)H * W * C;
    float4 mean[4], gq[4], bq[4];
    long long pos[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        pos[q] = ((long long)(2 * oy + (q >> 1)) * W + 2 * ox + (q & 1)) * C + c;
        const float4 gm = *reinterpret_cast<const float4*>(gamma + pos[q]);
        bq[q] = *reinterpret_cast<const float4*>(beta + pos[q]);
        float mv[4], bv[4], is[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) bn_finalize(stats[pos[q] + k], stats[P + pos[q] + k], n_total, eps, mv[k], bv[k], is[k]);
        mean[q] = make_float4(mv[0], mv[1], mv[2], mv[3]);
        gq[q] = make_float4(gm.x * is[0], gm.y * is[1], gm.z * is[2], gm.w * is[3]);
        if (chunk == 0) {
            float4 rm = *reinterpret_cast<const float4*>(rmean + pos[q]), rv = *reinterpret_cast<const float4*>(rvar + pos[q]);
            rm = make_float4(momentum * rm.x + (1.0f - momentum) * mv[0], momentum * rm.y + (1.0f - momentum) * mv[1],
                             momentum * rm.z + (1.0f - momentum) * mv[2], momentum * rm.w + (1.0f - momentum) * mv[3]);
            rv = make_float4(momentum * rv.x + (1.0f - momentum) * bv[0], momentum * rv.y + (1.0f - momentum) * bv[1],
                             momentum * rv.z + (1.0f - momentum) * bv[2], momentum * rv.w + (1.0f - momentum) * bv[3]);
            *reinterpret_cast<float4*>(rmean + pos[q]) = rm;
            *reinterpret_cast<float4*>(rvar + pos[q]) = rv;
            *reinterpret_cast<float4*>(save_mean + pos[q]) = mean[q];
            *reinterpret_cast<float4*>(save_invstd + pos[q]) = make_float4(is[0], is[1], is[2], is[3]);
        }
    }
    const long long OP = (long long)OH * OW * C, obase = ((long long)oy * OW + ox) * C + c;
    const int r0 = chunk * rows_per_chunk, r1 = min(B, r0 + rows_per_chunk);
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
        const float* xr = x + (long long)r * P;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 v = *reinterpret_cast<const float4*>(xr + pos[q]);
            v = make_float4(clip10(v.x), clip10(v.y), clip10(v.z), clip10(v.w));
            m.x = fmaxf(m.x, fmaf(gq[q].x, v.x - mean[q].x, bq[q].x)); m.y = fmaxf(m.y, fmaf(gq[q].y, v.y - mean[q].y, bq[q].y));
            m.z = fmaxf(m.z, fmaf(gq[q].z, v.z - mean[q].z, bq[q].z)); m.w = fmaxf(m.w, fmaf(gq[q].w, v.w - mean[q].w, bq[q].w));
        }
        *reinterpret_cast<float4*>(y + (long long)r * OP + obase) = m;
    }
}

// MODE 0: single GPU. MODE 1 (data parallel, first half): only the two batch reductions, written as LOCAL partial sums to
// dbeta / dgamma. MODE 2 (second half): dbeta / dgamma hold the rank-summed reductions; B_total is the global batch.
template <int MODE>
__global__ void __launch_bounds__(256) bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ err,
                                                     const double* __restrict__ ss, uint64_t chain,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ save_mean,
                                                     const float* __restrict__ save_invstd, float* __restrict__ dx,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int B,
                                                     long long P, int B_total, int prev_act, float prev_alpha,
                                                     double* ss_out0, double* ss_out1) {
    __shared__ float sm[BN_ROWG][BN_COLS];
    __shared__ double scratch[32];
    const float scale = chain_scale(ss, chain);
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const long long col = (long long)blockIdx.x * BN_COLS + cx;
    const bool ok = col < P;
    const float mean = ok ? save_mean[col] : 0.0f, invstd = ok ? save_invstd[col] : 0.0f;
    const float* xc = x + col;
    const float* ec = err + col;
    float sum_e, sum_ex;
    if (MODE != 2) {
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
        sum_e = colgroup_sum(se, sm, cx, ry);       // d_beta   (layers.py:1262)
        sum_ex = colgroup_sum(sex, sm, cx, ry);     // d_gamma  (layers.py:1261)
        if (ok && ry == 0) {
            if (dgamma != nullptr) dgamma[col] = sum_ex;
            if (dbeta != nullptr) dbeta[col] = sum_e;
        }
        if (MODE == 1) return;
    } else {
        sum_e = ok ? dbeta[col] : 0.0f;
        sum_ex = ok ? dgamma[col] : 0.0f;
    }
    float a0 = 0.0f, a1 = 0.0f;
    if (ok) {
        // layers.py:1264-1274 in closed form: dx = gamma*invstd*(e - xhat*sum(e*xhat)/B - sum(e)/B)
        const float gi = gamma[col] * invstd, invB = 1.0f / (float)B_total;
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


// ------------------------------------------------------------------------------------------------
// Fused conv-block kernels (BASELINE configs 2/3: Conv+ReLU -> BatchNorm -> MaxPool 2x2/s2):
//   forward : BatchNorm(train) + MaxPool in ONE pass over the conv output (read S once, write P + stats)
//   backward: MaxPool-bwd + BatchNorm-bwd + Activation-bwd in ONE pass (read S and the pooled error once, write R)
// Tiling: one pool-window position x 32 channels x the WHOLE batch belongs to one thread-block CLUSTER; each CTA of
// the cluster stages <= 128 batch rows of the window (4 x 128-byte NHWC lines per row, cp.async 16 B) in shared
// memory, reduces its rows, and the per-position batch statistics are completed across the cluster through
// distributed shared memory (one 128-float exchange per reduction). Cluster size = ceil(B / 128) <= 8, so batches
// up to 1024 stay single-pass. All global traffic is full 128-byte lines; the 2x2 max is thread-local.
// ------------------------------------------------------------------------------------------------
constexpr int FB_ROWS = 128;     // batch rows staged per CTA
constexpr int FB_THREADS = 256;  // 32 channel lanes x 8 row groups
constexpr int FB_RG = FB_THREADS / 32;
constexpr int FB_MAX_CLUSTER = 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local_smem_ptr, unsigned rank) {
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(local_smem_ptr));
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(a), "r"(rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];\n" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

// Thread mapping of the fused kernels: thread = (channel quad cq = tid % 8 -> channels 4cq..4cq+3 as one float4,
// row group rg = tid / 8 in 0..31). Every shared-memory access is a conflict-free 128-bit LDS, every statistic is
// kept per channel quad, and the 2x2 max is thread-local.
constexpr int FB_RGV = FB_THREADS / 8;   // 32 row groups

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_clip10(float4 a) { return make_float4(clip10(a.x), clip10(a.y), clip10(a.z), clip10(a.w)); }

// Sum v[NQ] (float4 per window position / statistic; partial over this thread's rows) over the 32 row groups of the
// CTA and over the CTAs of the cluster (distributed shared memory); every thread gets the totals.
// `red` is [FB_RGV][NQ][8] float4, `part` is [2][NQ][8] float4 (double-buffered by `phase`).
template <int NQ>
__device__ __forceinline__ void cluster_colsum4(float4 (&v)[NQ], float4* red, float4* part, int phase, int cq, int rg,
                                                unsigned csize) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) red[(rg * NQ + q) * 8 + cq] = v[q];
    __syncthreads();
    // 8 * NQ float4 totals per CTA: thread t < 8*NQ sums column t over the 32 row groups
    if (threadIdx.x < 8 * NQ) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int r = 0; r < FB_RGV; ++r) s = f4_add(s, red[r * NQ * 8 + threadIdx.x]);
        part[phase * NQ * 8 + threadIdx.x] = s;
    }
    if (csize > 1) cluster_sync_all(); else __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const float* src = reinterpret_cast<const float*>(&part[phase * NQ * 8 + q * 8 + cq]);
        float4 s;
        if (csize > 1) {
            s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (unsigned r = 0; r < csize; ++r)
                s = f4_add(s, make_float4(ld_dsmem(src, r), ld_dsmem(src + 1, r), ld_dsmem(src + 2, r), ld_dsmem(src + 3, r)));
        } else {
            s = part[phase * NQ * 8 + q * 8 + cq];
        }
        v[q] = s;
    }
}

struct FbGeom {
    int B, H, W, C;
};

// stage rows [row0, row0+nrows) of the 2x2 window at (oy, ox), channels [c0, c0+32) into tile[row][q][32]
__device__ __forceinline__ void fb_stage_window(float* tile, const float* __restrict__ x, const FbGeom& g, int row0,
                                                int nrows, int oy, int ox, int c0) {
    const long long HWC = (long long)g.H * g.W * g.C;
    const int n4 = nrows * 32;   // float4 chunks: nrows x 4 positions x 8
    for (int i = threadIdx.x; i < n4; i += FB_THREADS) {
        const int j = i & 7, q = (i >> 3) & 3, r = i >> 5;
        const long long src = (long long)(row0 + r) * HWC + ((long long)(2 * oy + (q >> 1)) * g.W + 2 * ox + (q & 1)) * g.C + c0 + j * 4;
        cp_async16(tile + (r * 4 + q) * 32 + j * 4, x + src);
    }
}

#define FB_FOR4(...) _Pragma("unroll") for (int k_ = 0; k_ < 4; ++k_) { __VA_ARGS__ }
__device__ __forceinline__ float& f4_at(float4& a, int k) { return reinterpret_cast<float*>(&a)[k]; }

__global__ void __launch_bounds__(FB_THREADS) bn_pool_fwd_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 float* __restrict__ rmean, float* __restrict__ rvar,
                                                                 float* __restrict__ save_mean,
                                                                 float* __restrict__ save_invstd,
                                                                 float* __restrict__ y, FbGeom g, float momentum,
                                                                 float eps) {
    extern __shared__ __align__(16) float tile[];   // [FB_ROWS][4][32]
    __shared__ float4 red[FB_RGV * 4 * 8];
    __shared__ float4 part[2 * 4 * 8];
    const unsigned csize = cluster_size(), crank = cluster_rank();
    const int cq = threadIdx.x & 7, rg = threadIdx.x >> 3;
    const int ox = blockIdx.x / csize, oy = blockIdx.y, c0 = blockIdx.z * 32, c = c0 + cq * 4;
    const int rows_per = (g.B + csize - 1) / csize;
    const int row0 = crank * rows_per;
    const int nrows = max(0, min(rows_per, g.B - row0));
    fb_stage_window(tile, x, g, row0, nrows, oy, ox, c0);
    cp_async_wait_all();
    __syncthreads();
    const float4* t4 = reinterpret_cast<const float4*>(tile);
    float4 acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = rg; r < nrows; r += FB_RGV) {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = f4_add(acc[q], f4_clip10(t4[(r * 4 + q) * 8 + cq]));   // layers.py:1234
    }
    cluster_colsum4<4>(acc, red, part, 0, cq, rg, csize);
    float4 mean[4];
    const float invB = 1.0f / (float)g.B;
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // layers.py:1237
        mean[q] = make_float4(acc[q].x * invB, acc[q].y * invB, acc[q].z * invB, acc[q].w * invB);
        acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int r = rg; r < nrows; r += FB_RGV) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = f4_clip10(t4[(r * 4 + q) * 8 + cq]);
            const float dx_ = v.x - mean[q].x, dy_ = v.y - mean[q].y, dz_ = v.z - mean[q].z, dw_ = v.w - mean[q].w;
            acc[q].x = fmaf(dx_, dx_, acc[q].x); acc[q].y = fmaf(dy_, dy_, acc[q].y);
            acc[q].z = fmaf(dz_, dz_, acc[q].z); acc[q].w = fmaf(dw_, dw_, acc[q].w);
        }
    }
    cluster_colsum4<4>(acc, red, part, 1, cq, rg, csize);
    float4 gq[4], bq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const long long p = ((long long)(2 * oy + (q >> 1)) * g.W + 2 * ox + (q & 1)) * g.C + c;
        const float4 gm = *reinterpret_cast<const float4*>(gamma + p);
        bq[q] = *reinterpret_cast<const float4*>(beta + p);
        float4 bv, is;
        FB_FOR4(f4_at(bv, k_) = f4_at(acc[q], k_) * invB + eps;            /* layers.py:1238 */
                f4_at(is, k_) = rsqrtf(f4_at(bv, k_) + eps);               /* layers.py:1252 */
                f4_at(gq[q], k_) = f4_at(const_cast<float4&>(gm), k_) * f4_at(is, k_);)
        if (crank == 0 && rg == 0) {
            float4 rm = *reinterpret_cast<const float4*>(rmean + p), rv = *reinterpret_cast<const float4*>(rvar + p);
            FB_FOR4(f4_at(rm, k_) = momentum * f4_at(rm, k_) + (1.0f - momentum) * f4_at(mean[q], k_);   /* layers.py:1240-1243 */
                    f4_at(rv, k_) = momentum * f4_at(rv, k_) + (1.0f - momentum) * f4_at(bv, k_);)
            *reinterpret_cast<float4*>(rmean + p) = rm;
            *reinterpret_cast<float4*>(rvar + p) = rv;
            *reinterpret_cast<float4*>(save_mean + p) = mean[q];
            *reinterpret_cast<float4*>(save_invstd + p) = is;
        }
    }
    const int OW = g.W >> 1;
    const long long OHWC = (long long)(g.H >> 1) * OW * g.C, obase = ((long long)oy * OW + ox) * g.C + c;
    for (int r = rg; r < nrows; r += FB_RGV) {
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = f4_clip10(t4[(r * 4 + q) * 8 + cq]);
            m.x = fmaxf(m.x, fmaf(gq[q].x, v.x - mean[q].x, bq[q].x)); m.y = fmaxf(m.y, fmaf(gq[q].y, v.y - mean[q].y, bq[q].y));
            m.z = fmaxf(m.z, fmaf(gq[q].z, v.z - mean[q].z, bq[q].z)); m.w = fmaxf(m.w, fmaf(gq[q].w, v.w - mean[q].w, bq[q].w));
        }
        *reinterpret_cast<float4*>(y + (long long)(row0 + r) * OHWC + obase) = m;   // layers.py:1256 then :599-600
    }
    if (csize > 1) cluster_sync_all();   // peers may still be reading this CTA's `part`
}

// MODE 0: single GPU. MODE 1 (data parallel, first half): pass 1 only -- LOCAL partial sums to dbeta / dgamma and ss0.
// MODE 2 (second half): dbeta / dgamma hold the rank-summed reductions; pass 2 only (ss1, ss2). B_total = global batch.
template <int MODE>
__global__ void __launch_bounds__(FB_THREADS, 2) pool_bn_act_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ err, const double* __restrict__ ss, uint64_t chain,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
    const float* __restrict__ save_invstd, float* __restrict__ r0, float* __restrict__ dgamma,
    float* __restrict__ dbeta, FbGeom g, int B_total, int prev_act, float prev_alpha, double* ss0, double* ss1, double* ss2) {
    extern __shared__ __align__(16) float tile[];   // [FB_ROWS][4][32] window values, then [FB_ROWS][32] pooled error
    __shared__ float4 red[FB_RGV * 8 * 8];
    __shared__ float4 part[2 * 8 * 8];
    __shared__ double scratch[32];
    const float scale = chain_scale(ss, chain);
    const unsigned csize = cluster_size(), crank = cluster_rank();
    const int cq = threadIdx.x & 7, rg = threadIdx.x >> 3;
    const int ox = blockIdx.x / csize, oy = blockIdx.y, c0 = blockIdx.z * 32, c = c0 + cq * 4;
    const int rows_per = (g.B + csize - 1) / csize;
    const int row0 = crank * rows_per;
    const int nrows = max(0, min(rows_per, g.B - row0));
    float* etile = tile + FB_ROWS * 128;
    const int OW = g.W >> 1;
    const long long OHWC = (long long)(g.H >> 1) * OW * g.C, obase = ((long long)oy * OW + ox) * g.C + c0;
    fb_stage_window(tile, x, g, row0, nrows, oy, ox, c0);
    for (int i = threadIdx.x; i < nrows * 8; i += FB_THREADS)
        cp_async16(etile + (i >> 3) * 32 + (i & 7) * 4, err + (long long)(row0 + (i >> 3)) * OHWC + obase + (i & 7) * 4);
    float4 mean[4], invstd[4], gq[4], bq[4];
    long long pos[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        pos[q] = ((long long)(2 * oy + (q >> 1)) * g.W + 2 * ox + (q & 1)) * g.C + c;
        mean[q] = *reinterpret_cast<const float4*>(save_mean + pos[q]);
        invstd[q] = *reinterpret_cast<const float4*>(save_invstd + pos[q]);
        const float4 gm = *reinterpret_cast<const float4*>(gamma + pos[q]);
        gq[q] = make_float4(gm.x * invstd[q].x, gm.y * invstd[q].y, gm.z * invstd[q].z, gm.w * invstd[q].w);
        bq[q] = *reinterpret_cast<const float4*>(beta + pos[q]);
    }
    cp_async_wait_all();
    __syncthreads();
    const float4* t4 = reinterpret_cast<const float4*>(tile);
    const float4* e4 = reinterpret_cast<const float4*>(etile);
    // pass 1: U = pool-bwd(e): EVERY tied maximum of the BN output receives e (layers.py:631-637); column sums
    float4 sums[8];   // [0..3] sum U (d_beta), [4..7] sum U*xhat (d_gamma)
#pragma unroll
    for (int q = 0; q < 8; ++q) sums[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    float a0 = 0.0f;
    if (MODE != 2) {
        for (int r = rg; r < nrows; r += FB_RGV) {
            float4 e = e4[r * 8 + cq];
            e = make_float4(scale * e.x, scale * e.y, scale * e.z, scale * e.w);
            float4 xv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) xv[q] = f4_clip10(t4[(r * 4 + q) * 8 + cq]);
            FB_FOR4(
                float yq[4], xh[4];
                _Pragma("unroll") for (int q = 0; q < 4; ++q) {
                    const float d = f4_at(xv[q], k_) - f4_at(mean[q], k_);
                    yq[q] = fmaf(f4_at(gq[q], k_), d, f4_at(bq[q], k_));
                    xh[q] = d * f4_at(invstd[q], k_);
                }
                const float m = fmaxf(fmaxf(yq[0], yq[1]), fmaxf(yq[2], yq[3]));
                _Pragma("unroll") for (int q = 0; q < 4; ++q) {
                    const float u = (yq[q] == m) ? f4_at(e, k_) : 0.0f;
                    f4_at(sums[q], k_) += u;
                    f4_at(sums[4 + q], k_) = fmaf(u, xh[q], f4_at(sums[4 + q], k_));
                    a0 = fmaf(u, u, a0);
                })
        }
        cluster_colsum4<8>(sums, red, part, 0, cq, rg, csize);
        if (crank == 0 && rg == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (dbeta != nullptr) *reinterpret_cast<float4*>(dbeta + pos[q]) = sums[q];          // layers.py:1262
                if (dgamma != nullptr) *reinterpret_cast<float4*>(dgamma + pos[q]) = sums[4 + q];    // layers.py:1261
            }
        }
        if (MODE == 1) {
            if (ss0 != nullptr) block_accumulate(a0, ss0, scratch);
            if (csize > 1) cluster_sync_all();
            return;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sums[q] = *reinterpret_cast<const float4*>(dbeta + pos[q]);
            sums[4 + q] = *reinterpret_cast<const float4*>(dgamma + pos[q]);
        }
    }
    const float invB = 1.0f / (float)B_total;
#pragma unroll
    for (int q = 0; q < 8; ++q) sums[q] = make_float4(sums[q].x * invB, sums[q].y * invB, sums[q].z * invB, sums[q].w * invB);
    float a1 = 0.0f, a2 = 0.0f;
    const long long HWC = (long long)g.H * g.W * g.C;
    for (int r = rg; r < nrows; r += FB_RGV) {
        float4 e = e4[r * 8 + cq];
        e = make_float4(scale * e.x, scale * e.y, scale * e.z, scale * e.w);
        float4 xraw[4], out[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) xraw[q] = t4[(r * 4 + q) * 8 + cq];
        FB_FOR4(
            float yq[4], xh[4];
            _Pragma("unroll") for (int q = 0; q < 4; ++q) {
                const float d = clip10(f4_at(xraw[q], k_)) - f4_at(mean[q], k_);
                yq[q] = fmaf(f4_at(gq[q], k_), d, f4_at(bq[q], k_));
                xh[q] = d * f4_at(invstd[q], k_);
            }
            const float m = fmaxf(fmaxf(yq[0], yq[1]), fmaxf(yq[2], yq[3]));
            _Pragma("unroll") for (int q = 0; q < 4; ++q) {
                const float u = (yq[q] == m) ? f4_at(e, k_) : 0.0f;
                float d = f4_at(gq[q], k_) * (u - xh[q] * f4_at(sums[4 + q], k_) - f4_at(sums[q], k_));   /* layers.py:1264-1274 */
                a1 = fmaf(d, d, a1);
                if (prev_act != B200NN_ACT_NONE) {
                    d *= act_grad_from_y(prev_act, f4_at(xraw[q], k_), prev_alpha);                    /* layers.py:237 */
                    a2 = fmaf(d, d, a2);
                }
                f4_at(out[q], k_) = d;
            })
        float* rb = r0 + (long long)(row0 + r) * HWC;
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(rb + pos[q]) = out[q];
    }
    if (MODE == 0 && ss0 != nullptr) block_accumulate(a0, ss0, scratch);
    if (ss1 != nullptr) block_accumulate(a1, ss1, scratch);
    if (prev_act != B200NN_ACT_NONE && ss2 != nullptr) block_accumulate(a2, ss2, scratch);
    if (MODE == 0 && csize > 1) cluster_sync_all();
}


// ------------------------------------------------------------------------------------------------
// STREAMING form of the fused block kernels for large batches (config 3 at batch 1024: the cluster kernels above stage a
// window's whole batch through 128-byte lines 256 KB apart and run at ~0.8 TB/s there). Mapping: thread = (pool window,
// channel quad) as in bn_pool_apply_kernel — a CTA of 256 threads covers a run of adjacent windows, i.e. two contiguous
// multi-KB segments of every batch row — and blockIdx.y = a chunk of SB_ROWS batch rows, so every load instruction of a warp
// covers whole 128-byte lines and each thread keeps several rows in flight. Backward = three launches:
//   partial : per (chunk, position) sums of U and U * (x - mean)  (U = pool-bwd of the incoming error)   -> ws[chunk][2][P]
//   reduce  : fixed-order sum over the chunks -> d_beta, d_gamma (deterministic; no float atomics)
//   apply   : BN-bwd + activation-bwd from the (rank-summed) reductions -> r0, clip slots
// x and err are read twice (2 * (S + P) + S bytes against the single-pass 2S + P), at several TB/s instead of 0.8.
constexpr int SB_ROWS = 32;

struct SbPos {
    long long pos[4];
    long long obase;
};
__device__ __forceinline__ SbPos sb_positions(long long idx, int H, int W, int C) {
    const int CQ = C >> 2, OW = W >> 1;
    const int cq = (int)(idx % CQ); long long t = idx / CQ;
    const int ox = (int)(t % OW); const int oy = (int)(t / OW);
    const int c = cq * 4;
    SbPos r;
#pragma unroll
    for (int q = 0; q < 4; ++q) r.pos[q] = ((long long)(2 * oy + (q >> 1)) * W + 2 * ox + (q & 1)) * C + c;
    r.obase = ((long long)oy * OW + ox) * C + c;
    return r;
}

__global__ void __launch_bounds__(256) pool_bn_bwd_stream_partial_kernel(
    const float* __restrict__ x, const float* __restrict__ err, const double* __restrict__ ss, uint64_t chain,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
    const float* __restrict__ save_invstd, float* __restrict__ part, double* ss0, int B, int H, int W, int C, long long items) {
    __shared__ double scratch[32];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = idx < items;
    const float scale = chain_scale(ss, chain);
    const long long P = (long long)H * W * C, OP = P >> 2;
    float a0 = 0.0f;
    if (ok) {
        const SbPos sp = sb_positions(idx, H, W, C);
        float4 mean[4], gq[4], bq[4], is[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mean[q] = *reinterpret_cast<const float4*>(save_mean + sp.pos[q]);
            is[q] = *reinterpret_cast<const float4*>(save_invstd + sp.pos[q]);
            const float4 gm = *reinterpret_cast<const float4*>(gamma + sp.pos[q]);
            gq[q] = make_float4(gm.x * is[q].x, gm.y * is[q].y, gm.z * is[q].z, gm.w * is[q].w);
            bq[q] = *reinterpret_cast<const float4*>(beta + sp.pos[q]);
        }
        float4 sb[4], sg[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) sb[q] = sg[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int r0 = blockIdx.y * SB_ROWS, r1 = min(B, r0 + SB_ROWS);
#pragma unroll 2
        for (int r = r0; r < r1; ++r) {
            const float* xr = x + (long long)r * P;
            float4 e = *reinterpret_cast<const float4*>(err + (long long)r * OP + sp.obase);
            float4 xv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) xv[q] = *reinterpret_cast<const float4*>(xr + sp.pos[q]);
            e = make_float4(scale * e.x, scale * e.y, scale * e.z, scale * e.w);
            FB_FOR4(
                float yq[4], dq[4];
                _Pragma("unroll") for (int q = 0; q < 4; ++q) {
                    dq[q] = clip10(f4_at(xv[q], k_)) - f4_at(mean[q], k_);
                    yq[q] = fmaf(f4_at(gq[q], k_), dq[q], f4_at(bq[q], k_));
                }
                const float m = fmaxf(fmaxf(yq[0], yq[1]), fmaxf(yq[2], yq[3]));
                _Pragma("unroll") for (int q = 0; q < 4; ++q) {
                    const float u = (yq[q] == m) ? f4_at(e, k_) : 0.0f;      /* every tied maximum, layers.py:631-637 */
                    f4_at(sb[q], k_) += u;
                    f4_at(sg[q], k_) = fmaf(u, dq[q], f4_at(sg[q], k_));
                    a0 = fmaf(u, u, a0);
                })
        }
        float* pb = part + (long long)blockIdx.y * 2 * P;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<float4*>(pb + sp.pos[q]) = sb[q];                                                   // sum U
            *reinterpret_cast<float4*>(pb + P + sp.pos[q]) = make_float4(sg[q].x * is[q].x, sg[q].y * is[q].y,  // sum U * xhat
                                                                         sg[q].z * is[q].z, sg[q].w * is[q].w);
        }
    }
    if (ss0 != nullptr) block_accumulate(a0, ss0, scratch);
}

// out0[i] = sum_c part[c][i], out1[i] = sum_c part[c][P + i]  (fixed chunk order)
__global__ void __launch_bounds__(256) stream_reduce2_kernel(const float* __restrict__ part, int chunks, long long P,
                                                             float* __restrict__ out0, float* __restrict__ out1) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= 2 * P) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16   // the chunk loads are independent: 16 in flight per thread (a 4-deep loop spent 8 HBM round trips on 32 chunks)
    for (int c = 0; c < chunks; ++c) acc = f4_add(acc, *reinterpret_cast<const float4*>(part + (long long)c * 2 * P + i4));
    float* o = i4 < P ? out0 + i4 : out1 + (i4 - P);
    *reinterpret_cast<float4*>(o) = acc;
}

__global__ void __launch_bounds__(256) pool_bn_act_bwd_stream_apply_kernel(
    const float* __restrict__ x, const float* __restrict__ err, const double* __restrict__ ss, uint64_t chain,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
    const float* __restrict__ save_invstd, float* __restrict__ r0out, const float* __restrict__ dgamma,
    const float* __restrict__ dbeta, int B, int H, int W, int C, float inv_total, int prev_act, float prev_alpha, double* ss1,
    double* ss2, long long items) {
    __shared__ double scratch[32];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = idx < items;
    const float scale = chain_scale(ss, chain);
    const long long P = (long long)H * W * C, OP = P >> 2;
    float a1 = 0.0f, a2 = 0.0f;
    if (ok) {
        const SbPos sp = sb_positions(idx, H, W, C);
        float4 mean[4], gq[4], bq[4], ka[4], kb[4], isv[4];   // dx = g * (U - xhat * kb - ka), ka = dbeta / N, kb = dgamma / N
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mean[q] = *reinterpret_cast<const float4*>(save_mean + sp.pos[q]);
            const float4 is = *reinterpret_cast<const float4*>(save_invstd + sp.pos[q]);
            const float4 gm = *reinterpret_cast<const float4*>(gamma + sp.pos[q]);
            gq[q] = make_float4(gm.x * is.x, gm.y * is.y, gm.z * is.z, gm.w * is.w);
            bq[q] = *reinterpret_cast<const float4*>(beta + sp.pos[q]);
            const float4 db = *reinterpret_cast<const float4*>(dbeta + sp.pos[q]), dg = *reinterpret_cast<const float4*>(dgamma + sp.pos[q]);
            ka[q] = make_float4(db.x * inv_total, db.y * inv_total, db.z * inv_total, db.w * inv_total);
            kb[q] = make_float4(dg.x * inv_total, dg.y * inv_total, dg.z * inv_total, dg.w * inv_total);
            isv[q] = is;
        }
        const int r0 = blockIdx.y * SB_ROWS, r1 = min(B, r0 + SB_ROWS);
#pragma unroll 2
        for (int r = r0; r < r1; ++r) {
            const float* xr = x + (long long)r * P;
            float4 e = *reinterpret_cast<const float4*>(err + (long long)r * OP + sp.obase);
            float4 xraw[4], out[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) xraw[q] = *reinterpret_cast<const float4*>(xr + sp.pos[q]);
            e = make_float4(scale * e.x, scale * e.y, scale * e.z, scale * e.w);
            FB_FOR4(
                float yq[4], xh[4];
                _Pragma("unroll") for (int q = 0; q < 4; ++q) {
                    const float d = clip10(f4_at(xraw[q], k_)) - f4_at(mean[q], k_);
                    yq[q] = fmaf(f4_at(gq[q], k_), d, f4_at(bq[q], k_));
                    xh[q] = d * f4_at(isv[q], k_);
                }
                const float m = fmaxf(fmaxf(yq[0], yq[1]), fmaxf(yq[2], yq[3]));
                _Pragma("unroll") for (int q = 0; q < 4; ++q) {
                    const float u = (yq[q] == m) ? f4_at(e, k_) : 0.0f;
                    float d = f4_at(gq[q], k_) * (u - xh[q] * f4_at(kb[q], k_) - f4_at(ka[q], k_));   /* layers.py:1264-1274 */
                    a1 = fmaf(d, d, a1);
                    if (prev_act != B200NN_ACT_NONE) {
                        d *= act_grad_from_y(prev_act, f4_at(xraw[q], k_), prev_alpha);               /* layers.py:237 */
                        a2 = fmaf(d, d, a2);
                    }
                    f4_at(out[q], k_) = d;
                })
            float* rb = r0out + (long long)r * P;
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(rb + sp.pos[q]) = out[q];
        }
    }
    if (ss1 != nullptr) block_accumulate(a1, ss1, scratch);
    if (prev_act != B200NN_ACT_NONE && ss2 != nullptr) block_accumulate(a2, ss2, scratch);
}

// ---- one-channel-per-thread forms of the two streaming backward kernels (round 2). The channel-quad kernels above keep six float4
// constants per pooled position in registers (apply: 221 registers = ONE 256-thread CTA per SM, ~41 KB of loads in flight per SM,
// 2.2 TB/s; ncu: profiles/r2_ncu_cfg3_kernels.txt). With one channel per thread the constants are 24 scalars, a warp still covers whole
// 128-byte lines (32 consecutive channels), four batch rows are in flight per thread (20 loads) and three CTAs fit an SM. The
// arithmetic per element is unchanged.
constexpr int SB1_UNROLL = 4;

struct SbPos1 {
    long long pos[4];
    long long obase;
};
__device__ __forceinline__ SbPos1 sb_positions1(long long idx, int W, int C) {
    const int OW = W >> 1;
    const int c = (int)(idx % C); long long t = idx / C;
    const int ox = (int)(t % OW); const int oy = (int)(t / OW);
    SbPos1 r;
#pragma unroll
    for (int q = 0; q < 4; ++q) r.pos[q] = ((long long)(2 * oy + (q >> 1)) * W + 2 * ox + (q & 1)) * C + c;
    r.obase = ((long long)oy * OW + ox) * C + c;
    return r;
}

__global__ void __launch_bounds__(256, 3) pool_bn_bwd_stream_partial1_kernel(
    const float* __restrict__ x, const float* __restrict__ err, const double* __restrict__ ss, uint64_t chain,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
    const float* __restrict__ save_invstd, float* __restrict__ part, double* ss0, int B, int H, int W, int C, long long items) {
    __shared__ double scratch[32];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = idx < items;
    const float scale = chain_scale(ss, chain);
    const long long P = (long long)H * W * C, OP = P >> 2;
    float a0 = 0.0f;
    if (ok) {
        const SbPos1 sp = sb_positions1(idx, W, C);
        float mean[4], gq[4], bq[4], is[4], sb[4], sg[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mean[q] = save_mean[sp.pos[q]];
            is[q] = save_invstd[sp.pos[q]];
            gq[q] = gamma[sp.pos[q]] * is[q];
            bq[q] = beta[sp.pos[q]];
            sb[q] = sg[q] = 0.0f;
        }
        const int r0 = blockIdx.y * SB_ROWS, r1 = min(B, r0 + SB_ROWS);
        for (int rb = r0; rb < r1; rb += SB1_UNROLL) {
            float e[SB1_UNROLL], xv[SB1_UNROLL][4];
#pragma unroll
            for (int j = 0; j < SB1_UNROLL; ++j) {
                const int r = min(rb + j, r1 - 1);
                e[j] = err[(long long)r * OP + sp.obase];
#pragma unroll
                for (int q = 0; q < 4; ++q) xv[j][q] = x[(long long)r * P + sp.pos[q]];
            }
#pragma unroll
            for (int j = 0; j < SB1_UNROLL; ++j) {
                if (rb + j < r1) {
                    const float ev = scale * e[j];
                    float yq[4], dq[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        dq[q] = clip10(xv[j][q]) - mean[q];
                        yq[q] = fmaf(gq[q], dq[q], bq[q]);
                    }
                    const float m = fmaxf(fmaxf(yq[0], yq[1]), fmaxf(yq[2], yq[3]));
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float u = (yq[q] == m) ? ev : 0.0f;      /* every tied maximum, layers.py:631-637 */
                        sb[q] += u;
                        sg[q] = fmaf(u, dq[q], sg[q]);
                        a0 = fmaf(u, u, a0);
                    }
                }
            }
        }
        float* pb = part + (long long)blockIdx.y * 2 * P;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pb[sp.pos[q]] = sb[q];                       // sum U
            pb[P + sp.pos[q]] = sg[q] * is[q];           // sum U * xhat
        }
    }
    if (ss0 != nullptr) block_accumulate(a0, ss0, scratch);
}

__global__ void __launch_bounds__(256, 3) pool_bn_act_bwd_stream_apply1_kernel(
    const float* __restrict__ x, const float* __restrict__ err, const double* __restrict__ ss, uint64_t chain,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
    const float* __restrict__ save_invstd, float* __restrict__ r0out, const float* __restrict__ dgamma,
    const float* __restrict__ dbeta, int B, int H, int W, int C, float inv_total, int prev_act, float prev_alpha, double* ss1,
    double* ss2, long long items) {
    __shared__ double scratch[32];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = idx < items;
    const float scale = chain_scale(ss, chain);
    const long long P = (long long)H * W * C, OP = P >> 2;
    float a1 = 0.0f, a2 = 0.0f;
    if (ok) {
        const SbPos1 sp = sb_positions1(idx, W, C);
        float mean[4], gq[4], bq[4], ka[4], kb[4], isv[4];   // dx = g * (U - xhat * kb - ka), ka = dbeta / N, kb = dgamma / N
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mean[q] = save_mean[sp.pos[q]];
            isv[q] = save_invstd[sp.pos[q]];
            gq[q] = gamma[sp.pos[q]] * isv[q];
            bq[q] = beta[sp.pos[q]];
            ka[q] = dbeta[sp.pos[q]] * inv_total;
            kb[q] = dgamma[sp.pos[q]] * inv_total;
        }
        const int r0 = blockIdx.y * SB_ROWS, r1 = min(B, r0 + SB_ROWS);
        for (int rb = r0; rb < r1; rb += SB1_UNROLL) {
            float e[SB1_UNROLL], xraw[SB1_UNROLL][4];
#pragma unroll
            for (int j = 0; j < SB1_UNROLL; ++j) {
                const int r = min(rb + j, r1 - 1);
                e[j] = err[(long long)r * OP + sp.obase];
#pragma unroll
                for (int q = 0; q < 4; ++q) xraw[j][q] = x[(long long)r * P + sp.pos[q]];
            }
#pragma unroll
            for (int j = 0; j < SB1_UNROLL; ++j) {
                if (rb + j < r1) {
                    const float ev = scale * e[j];
                    float yq[4], xh[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float d = clip10(xraw[j][q]) - mean[q];
                        yq[q] = fmaf(gq[q], d, bq[q]);
                        xh[q] = d * isv[q];
                    }
                    const float m = fmaxf(fmaxf(yq[0], yq[1]), fmaxf(yq[2], yq[3]));
                    float* rb_out = r0out + (long long)(rb + j) * P;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float u = (yq[q] == m) ? ev : 0.0f;
                        float d = gq[q] * (u - xh[q] * kb[q] - ka[q]);                       /* layers.py:1264-1274 */
                        a1 = fmaf(d, d, a1);
                        if (prev_act != B200NN_ACT_NONE) {
                            d *= act_grad_from_y(prev_act, xraw[j][q], prev_alpha);           /* layers.py:237 */
                            a2 = fmaf(d, d, a2);
                        }
                        rb_out[sp.pos[q]] = d;
                    }
                }
            }
        }
    }
    if (ss1 != nullptr) block_accumulate(a1, ss1, scratch);
    if (prev_act != B200NN_ACT_NONE && ss2 != nullptr) block_accumulate(a2, ss2, scratch);
}

// Forward statistics, streaming form: ONE sweep over x. Thread = 4 adjacent positions (float4), blockIdx.y = a chunk of
// SS_ROWS batch rows; per chunk the sums are taken about a per-position shift K = clip(x[first row of the chunk]) so that
// sum((x - K)^2) - sum(x - K)^2 / n loses nothing to cancellation (a one-pass variance about zero would); the chunks are then
// merged pairwise-exactly (Chan) in fp64 by stream_stats_combine_kernel into the [sum | sum of squares] format of
// bn_stats_partial. part = [chunk][3][P]: K, sum(x - K), sum((x - K)^2).
constexpr int SS_ROWS = 64;
__global__ void __launch_bounds__(256) bn_stats_stream_kernel(const float* __restrict__ x, float* __restrict__ part, int B, long long P) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= P) return;
    const int r0 = blockIdx.y * SS_ROWS, r1 = min(B, r0 + SS_ROWS);
    const float* xc = x + i4;
    const float4 K = f4_clip10(*reinterpret_cast<const float4*>(xc + (long long)r0 * P));
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll 8
    for (int r = r0; r < r1; ++r) {
        const float4 v = f4_clip10(*reinterpret_cast<const float4*>(xc + (long long)r * P));      // layers.py:1234
        const float4 d = make_float4(v.x - K.x, v.y - K.y, v.z - K.z, v.w - K.w);
        s1 = f4_add(s1, d);
        s2 = make_float4(fmaf(d.x, d.x, s2.x), fmaf(d.y, d.y, s2.y), fmaf(d.z, d.z, s2.z), fmaf(d.w, d.w, s2.w));
    }
    float* pb = part + (long long)blockIdx.y * 3 * P + i4;
    *reinterpret_cast<float4*>(pb) = K;
    *reinterpret_cast<float4*>(pb + P) = s1;
    *reinterpret_cast<float4*>(pb + 2 * P) = s2;
}

// Chan's pairwise merge of two (count, mean, M2) summaries; an empty side leaves the other unchanged
struct Moments {
    double n, mean, M2;
};
__device__ __forceinline__ Moments merge_moments(const Moments& a, const Moments& b) {
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    Moments r;
    r.n = a.n + b.n;
    const double delta = b.mean - a.mean, f = b.n / r.n;
    r.mean = a.mean + delta * f;
    r.M2 = a.M2 + b.M2 + delta * delta * a.n * f;
    return r;
}

__global__ void __launch_bounds__(256) stream_stats_combine_kernel(const float* __restrict__ part, int chunks, int B, long long P,
                                                                   double* __restrict__ stats) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    Moments tot = {0.0, 0.0, 0.0};
    for (int c0 = 0; c0 < chunks; c0 += 8) {
        // eight chunks at a time: their loads in flight together, their summaries formed independently, then merged as a
        // balanced tree (depth 3 instead of a serial chain of eight fp64 divide chains); fixed order, so every run and every
        // replica forms the same bits
        float kv[8], s1v[8], s2v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float* pb = part + (long long)min(c0 + j, chunks - 1) * 3 * P + i;
            kv[j] = pb[0]; s1v[j] = pb[P]; s2v[j] = pb[2 * P];
        }
        Moments m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const double nc = c < chunks ? (double)min(SS_ROWS, B - c * SS_ROWS) : 0.0;
            const double K = (double)kv[j], S1 = (double)s1v[j], S2 = (double)s2v[j];
            const double inv = nc > 0.0 ? 1.0 / nc : 0.0;
            m[j].n = nc;
            m[j].mean = K + S1 * inv;
            m[j].M2 = fmax(S2 - S1 * S1 * inv, 0.0);
        }
#pragma unroll
        for (int step = 1; step < 8; step *= 2)
#pragma unroll
            for (int j = 0; j + step < 8; j += 2 * step) m[j] = merge_moments(m[j], m[j + step]);
        tot = merge_moments(tot, m[0]);
    }
    const double S = tot.mean * tot.n;
    stats[i] = S;                               // sum clip(x)
    stats[P + i] = tot.M2 + S * S / tot.n;      // sum clip(x)^2, formed from the exact M2 (see bn_stats_partial_kernel)
}

static int check_fused_block(const char* who, int B, int H, int W, int C) {
    B200NN_REQUIRE(B > 0 && B <= FB_ROWS * FB_MAX_CLUSTER, "%s: batch %d outside the single-pass range (1..%d)", who, B,
                   FB_ROWS * FB_MAX_CLUSTER);
    B200NN_REQUIRE(H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0 && C % 32 == 0,
                   "%s: needs even H, W and C %% 32 == 0 (got %dx%dx%d)", who, H, W, C);
    B200NN_REQUIRE((long long)(W / 2) * FB_MAX_CLUSTER < (1LL << 31) && H / 2 <= 65535 && C / 32 <= 65535, "%s: grid too large", who);
    return B200NN_OK;
}

template <class Kern, class... Args>
static int launch_fused_block(const char* name, Kern kern, size_t smem, int B, int H, int W, int C, cudaStream_t st,
                              Args... args) {
    const int cl = (B + FB_ROWS - 1) / FB_ROWS;
    static thread_local const void* configured[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool seen = false;
    for (auto p : configured) seen = seen || p == (const void*)kern;
    if (!seen) {
        B200NN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (auto& p : configured) if (p == nullptr) { p = (const void*)kern; break; }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((W / 2) * cl, H / 2, C / 32);
    cfg.blockDim = dim3(FB_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200NN_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    B200NN_LAUNCH_CHECK(name);
    return B200NN_OK;
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
    B200NN_REQUIRE(prev_act >= 0 && prev_act <= B200NN_ACT_MAX_FUSED, "bn_bwd: unknown activation %d", prev_act);
    const long long blocks = (P + BN_COLS - 1) / BN_COLS;
    bn_bwd_kernel<0><<<(int)blocks, 256, 0, as_stream(stream)>>>(x, err, ss, chain, gamma, save_mean, save_invstd, dx, dgamma,
                                                               dbeta, B, P, B, prev_act, prev_alpha, ss_out0, ss_out1);
    B200NN_LAUNCH_CHECK("bn_bwd");
    return B200NN_OK;
}

// ---- data-parallel halves (see b200nn.h) ------------------------------------------------------------------------
int b200nn_bn_stats_partial(const float* x, double* stats, int B, long long P, void* stream) {
    B200NN_REQUIRE(B > 0 && P > 0 && stats != nullptr, "bn_stats_partial: bad argument B=%d P=%lld", B, P);
    const long long blocks = (P + BN_COLS - 1) / BN_COLS;
    B200NN_REQUIRE(blocks < (1LL << 31), "bn_stats_partial: too many positions");
    bn_stats_partial_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, stats, B, P);
    B200NN_LAUNCH_CHECK("bn_stats_partial");
    return B200NN_OK;
}

int b200nn_bn_fwd_apply(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                        float* save_mean, float* save_invstd, const double* stats, float* y, int B, long long P,
                        long long B_total, float momentum, float eps, void* stream) {
    B200NN_REQUIRE(B > 0 && P > 0 && B_total >= B && stats != nullptr, "bn_fwd_apply: bad argument");
    const long long blocks = (P + BN_COLS - 1) / BN_COLS;
    bn_fwd_apply_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, gamma, beta, running_mean, running_var, save_mean,
                                                                  save_invstd, stats, y, B, P, (double)B_total, momentum, eps);
    B200NN_LAUNCH_CHECK("bn_fwd_apply");
    return B200NN_OK;
}

int b200nn_bn_bwd_partial(const float* x, const float* err, const double* ss, uint64_t chain, const float* save_mean,
                          const float* save_invstd, float* dgamma, float* dbeta, int B, long long P, void* stream) {
    B200NN_REQUIRE(B > 0 && P > 0 && dgamma != nullptr && dbeta != nullptr, "bn_bwd_partial: bad argument");
    const long long blocks = (P + BN_COLS - 1) / BN_COLS;
    bn_bwd_kernel<1><<<(int)blocks, 256, 0, as_stream(stream)>>>(x, err, ss, chain, nullptr, save_mean, save_invstd, nullptr, dgamma,
                                                               dbeta, B, P, B, 0, 0.f, nullptr, nullptr);
    B200NN_LAUNCH_CHECK("bn_bwd_partial");
    return B200NN_OK;
}

int b200nn_bn_bwd_apply(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                        const float* save_mean, const float* save_invstd, float* dx, const float* dgamma, const float* dbeta,
                        int B, long long P, long long B_total, int prev_act, float prev_alpha, double* ss_out0,
                        double* ss_out1, void* stream) {
    B200NN_REQUIRE(B > 0 && P > 0 && B_total >= B && dgamma != nullptr && dbeta != nullptr, "bn_bwd_apply: bad argument");
    B200NN_REQUIRE(prev_act >= 0 && prev_act <= B200NN_ACT_MAX_FUSED, "bn_bwd_apply: unknown activation %d", prev_act);
    const long long blocks = (P + BN_COLS - 1) / BN_COLS;
    bn_bwd_kernel<2><<<(int)blocks, 256, 0, as_stream(stream)>>>(x, err, ss, chain, gamma, save_mean, save_invstd, dx,
                                                               const_cast<float*>(dgamma), const_cast<float*>(dbeta), B, P,
                                                               (int)B_total, prev_act, prev_alpha, ss_out0, ss_out1);
    B200NN_LAUNCH_CHECK("bn_bwd_apply");
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


int b200nn_fused_block_max_batch(void) { return FB_ROWS * FB_MAX_CLUSTER; }

int b200nn_bn_pool_fwd_train(const float* x, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, float* save_mean, float* save_invstd, float* y_pool, int B, int H,
                             int W, int C, float momentum, float eps, void* stream) {
    if (int rc = check_fused_block("bn_pool_fwd_train", B, H, W, C)) return rc;
    B200NN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "bn_pool_fwd_train: x must be 16-byte aligned");
    const FbGeom g{B, H, W, C};
    return launch_fused_block("bn_pool_fwd_train", bn_pool_fwd_kernel, (size_t)FB_ROWS * 128 * sizeof(float), B, H, W, C,
                              as_stream(stream), x, gamma, beta, running_mean, running_var, save_mean, save_invstd, y_pool, g,
                              momentum, eps);
}

int b200nn_pool_bn_act_bwd(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                           const float* beta, const float* save_mean, const float* save_invstd, float* r0,
                           float* dgamma, float* dbeta, int B, int H, int W, int C, int prev_act, float prev_alpha,
                           double* ss_out0, double* ss_out1, double* ss_out2, void* stream) {
    if (int rc = check_fused_block("pool_bn_act_bwd", B, H, W, C)) return rc;
    B200NN_REQUIRE(prev_act >= 0 && prev_act <= B200NN_ACT_MAX_FUSED, "pool_bn_act_bwd: unknown activation %d", prev_act);
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(err)) & 15) == 0,
                   "pool_bn_act_bwd: x and err must be 16-byte aligned");
    const FbGeom g{B, H, W, C};
    return launch_fused_block("pool_bn_act_bwd", pool_bn_act_bwd_kernel<0>, (size_t)FB_ROWS * 160 * sizeof(float), B, H, W, C,
                              as_stream(stream), x, err, ss, chain, gamma, beta, save_mean, save_invstd, r0, dgamma, dbeta, g,
                              B, prev_act, prev_alpha, ss_out0, ss_out1, ss_out2);
}

int b200nn_bn_pool_fwd_apply(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                             float* save_mean, float* save_invstd, const double* stats, float* y_pool, int B, int H, int W,
                             int C, long long B_total, float momentum, float eps, void* stream) {
    B200NN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0 && B_total >= B && stats != nullptr,
                   "bn_pool_fwd_apply: needs even H, W and C %% 4 == 0 (got %dx%dx%d)", H, W, C);
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y_pool)) & 15) == 0,
                   "bn_pool_fwd_apply: x and y must be 16-byte aligned");
    // rows per thread: 32 or 16 while that still leaves >= 2 CTAs per SM, else 8
    const long long windows = (long long)(H / 2) * (W / 2) * (C / 4);
    int rpc = FB_APPLY_ROWS;
    for (int cand = 32; cand > FB_APPLY_ROWS; cand >>= 1)
        if ((long long)((B + cand - 1) / cand) * windows >= 2LL * 256 * kNumSMs) { rpc = cand; break; }
    const long long chunks = (B + rpc - 1) / rpc;
    const long long total = chunks * windows;
    const long long blocks = (total + 255) / 256;
    B200NN_REQUIRE(blocks < (1LL << 31), "bn_pool_fwd_apply: grid too large");
    bn_pool_apply_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, gamma, beta, running_mean, running_var, save_mean,
                                                                   save_invstd, stats, y_pool, B, H, W, C, (double)B_total,
                                                                   momentum, eps, total, rpc);
    B200NN_LAUNCH_CHECK("bn_pool_fwd_apply");
    return B200NN_OK;
}

int b200nn_pool_bn_bwd_partial(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                               const float* beta, const float* save_mean, const float* save_invstd, float* dgamma,
                               float* dbeta, int B, int H, int W, int C, double* ss_out0, void* stream) {
    if (int rc = check_fused_block("pool_bn_bwd_partial", B, H, W, C)) return rc;
    B200NN_REQUIRE(dgamma != nullptr && dbeta != nullptr, "pool_bn_bwd_partial: dgamma / dbeta are required");
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(err)) & 15) == 0,
                   "pool_bn_bwd_partial: x and err must be 16-byte aligned");
    const FbGeom g{B, H, W, C};
    return launch_fused_block("pool_bn_bwd_partial", pool_bn_act_bwd_kernel<1>, (size_t)FB_ROWS * 160 * sizeof(float), B, H, W, C,
                              as_stream(stream), x, err, ss, chain, gamma, beta, save_mean, save_invstd, (float*)nullptr, dgamma,
                              dbeta, g, B, 0, 0.f, ss_out0, (double*)nullptr, (double*)nullptr);
}

int b200nn_pool_bn_act_bwd_apply(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                                 const float* beta, const float* save_mean, const float* save_invstd, float* r0,
                                 const float* dgamma, const float* dbeta, int B, int H, int W, int C, long long B_total,
                                 int prev_act, float prev_alpha, double* ss_out1, double* ss_out2, void* stream) {
    if (int rc = check_fused_block("pool_bn_act_bwd_apply", B, H, W, C)) return rc;
    B200NN_REQUIRE(prev_act >= 0 && prev_act <= B200NN_ACT_MAX_FUSED, "pool_bn_act_bwd_apply: unknown activation %d", prev_act);
    B200NN_REQUIRE(dgamma != nullptr && dbeta != nullptr && B_total >= B, "pool_bn_act_bwd_apply: bad argument");
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(err)) & 15) == 0,
                   "pool_bn_act_bwd_apply: x and err must be 16-byte aligned");
    const FbGeom g{B, H, W, C};
    return launch_fused_block("pool_bn_act_bwd_apply", pool_bn_act_bwd_kernel<2>, (size_t)FB_ROWS * 160 * sizeof(float), B, H, W,
                              C, as_stream(stream), x, err, ss, chain, gamma, beta, save_mean, save_invstd, r0,
                              const_cast<float*>(dgamma), const_cast<float*>(dbeta), g, (int)B_total, prev_act, prev_alpha,
                              (double*)nullptr, ss_out1, ss_out2);
}

// ---- streaming forms (large batches; any batch size, C % 4 == 0) -----------------------------------------------------------
// one channel per thread when a warp then covers whole 128-byte lines (C % 32 == 0); B200NN_STREAM_SCALAR=0 keeps the quad kernels
static bool stream_scalar_form(int C) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200NN_STREAM_SCALAR");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1 && C % 32 == 0;
}

static int check_stream_block(const char* who, int B, int H, int W, int C, const void* a, const void* b, const void* c) {
    B200NN_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0,
                   "%s: needs even H, W and C %% 4 == 0 (got %dx%dx%d)", who, H, W, C);
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0,
                   "%s: tensors must be 16-byte aligned", who);
    return B200NN_OK;
}

size_t b200nn_pool_bn_bwd_stream_ws_bytes(int B, int H, int W, int C) {
    return (size_t)((B + SB_ROWS - 1) / SB_ROWS) * 2 * H * W * C * sizeof(float);
}

int b200nn_pool_bn_bwd_stream_partial(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                                      const float* beta, const float* save_mean, const float* save_invstd, float* dgamma,
                                      float* dbeta, float* ws, int B, int H, int W, int C, double* ss_out0, void* stream) {
    if (int rc = check_stream_block("pool_bn_bwd_stream_partial", B, H, W, C, x, err, ws)) return rc;
    B200NN_REQUIRE(dgamma != nullptr && dbeta != nullptr && ws != nullptr, "pool_bn_bwd_stream_partial: dgamma / dbeta / ws are required");
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta)) & 15) == 0,
                   "pool_bn_bwd_stream_partial: dgamma / dbeta must be 16-byte aligned");
    const long long items = (long long)(H / 2) * (W / 2) * (C / 4), P = (long long)H * W * C;
    const int chunks = (B + SB_ROWS - 1) / SB_ROWS;
    const dim3 grid((unsigned)((items + 255) / 256), (unsigned)chunks);
    cudaStream_t st = as_stream(stream);
    if (stream_scalar_form(C)) {
        const dim3 grid1((unsigned)((items * 4 + 255) / 256), (unsigned)chunks);
        pool_bn_bwd_stream_partial1_kernel<<<grid1, 256, 0, st>>>(x, err, ss, chain, gamma, beta, save_mean, save_invstd, ws, ss_out0, B, H,
                                                                W, C, items * 4);
    } else {
        pool_bn_bwd_stream_partial_kernel<<<grid, 256, 0, st>>>(x, err, ss, chain, gamma, beta, save_mean, save_invstd, ws, ss_out0, B, H,
                                                              W, C, items);
    }
    B200NN_LAUNCH_CHECK("pool_bn_bwd_stream_partial");
    stream_reduce2_kernel<<<(unsigned)((2 * P / 4 + 255) / 256), 256, 0, st>>>(ws, chunks, P, dbeta, dgamma);
    B200NN_LAUNCH_CHECK("stream_reduce2");
    return B200NN_OK;
}

int b200nn_pool_bn_act_bwd_stream_apply(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                                        const float* beta, const float* save_mean, const float* save_invstd, float* r0,
                                        const float* dgamma, const float* dbeta, int B, int H, int W, int C, long long B_total,
                                        int prev_act, float prev_alpha, double* ss_out1, double* ss_out2, void* stream) {
    if (int rc = check_stream_block("pool_bn_act_bwd_stream_apply", B, H, W, C, x, err, r0)) return rc;
    B200NN_REQUIRE(prev_act >= 0 && prev_act <= B200NN_ACT_MAX_FUSED, "pool_bn_act_bwd_stream_apply: unknown activation %d", prev_act);
    B200NN_REQUIRE(dgamma != nullptr && dbeta != nullptr && B_total >= B, "pool_bn_act_bwd_stream_apply: bad argument");
    const long long items = (long long)(H / 2) * (W / 2) * (C / 4);
    const dim3 grid((unsigned)((items + 255) / 256), (unsigned)((B + SB_ROWS - 1) / SB_ROWS));
    if (stream_scalar_form(C)) {
        const dim3 grid1((unsigned)((items * 4 + 255) / 256), (unsigned)((B + SB_ROWS - 1) / SB_ROWS));
        pool_bn_act_bwd_stream_apply1_kernel<<<grid1, 256, 0, as_stream(stream)>>>(x, err, ss, chain, gamma, beta, save_mean, save_invstd,
                                                                                r0, dgamma, dbeta, B, H, W, C, 1.0f / (float)B_total,
                                                                                prev_act, prev_alpha, ss_out1, ss_out2, items * 4);
    } else
    pool_bn_act_bwd_stream_apply_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, err, ss, chain, gamma, beta, save_mean, save_invstd, r0,
                                                                           dgamma, dbeta, B, H, W, C, 1.0f / (float)B_total, prev_act,
                                                                           prev_alpha, ss_out1, ss_out2, items);
    B200NN_LAUNCH_CHECK("pool_bn_act_bwd_stream_apply");
    return B200NN_OK;
}

size_t b200nn_bn_stats_stream_ws_bytes(int B, long long P) { return (size_t)((B + SS_ROWS - 1) / SS_ROWS) * 3 * P * sizeof(float); }

/* one-sweep form of b200nn_bn_stats_partial for large batches (same `stats` format); P % 4 == 0 */
int b200nn_bn_stats_stream(const float* x, double* stats, float* ws, int B, long long P, void* stream) {
    B200NN_REQUIRE(B > 0 && P > 0 && P % 4 == 0 && x != nullptr && stats != nullptr && ws != nullptr, "bn_stats_stream: bad argument");
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0, "bn_stats_stream: x / ws must be 16-byte aligned");
    const int chunks = (B + SS_ROWS - 1) / SS_ROWS;
    const dim3 grid((unsigned)((P / 4 + 255) / 256), (unsigned)chunks);
    cudaStream_t st = as_stream(stream);
    bn_stats_stream_kernel<<<grid, 256, 0, st>>>(x, ws, B, P);
    B200NN_LAUNCH_CHECK("bn_stats_stream");
    stream_stats_combine_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(ws, chunks, B, P, stats);
    B200NN_LAUNCH_CHECK("stream_stats_combine");
    return B200NN_OK;
}

}  // extern "C"
