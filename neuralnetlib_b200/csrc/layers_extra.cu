// layers_extra.cu — the sibling layers of the hot-path kernel families: Dropout (layers.py:264-342),
// AveragePooling2D (:908-1011), GlobalAveragePooling2D (:1419-1442), UpSampling2D nearest (:2695-2815), and the im2col / col2im
// pair that puts Conv1D (:667-816, preprocessing.py:128-186) on the Dense kernels. All are HBM-bound streaming kernels; every
// backward follows the lazy-clip protocol of the other layers: dx = chain(ss) * (...), ss_out += sum(dx^2).
#include "common.cuh"

namespace b200nn {

// out = chain(ss) * a * b   (Dropout: b = the host-generated mask / (1 - rate))
__global__ void __launch_bounds__(256) mul_kernel(const float* __restrict__ a, const float* __restrict__ b, const double* __restrict__ ss,
                                                  uint64_t chain, float* __restrict__ out, double* ss_out, long long n) {
    __shared__ double scratch[32];
    const float s = chain_scale(ss, chain);
    float acc = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = s * a[i] * b[i];
        acc = fmaf(v, v, acc);
        out[i] = v;
    }
    if (ss_out != nullptr) block_accumulate(acc, ss_out, scratch);
}

// out = a + s * b  (Autoencoder skip connections, models.py:1014-1019)
__global__ void __launch_bounds__(256) axpy_kernel(const float* __restrict__ a, const float* __restrict__ b, float s, float* __restrict__ out,
                                                   long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fmaf(s, b[i], a[i]);
}

// Autoencoder.backward_pass's own clip rule (models.py:1033-1048), applied to every error tensor and every gradient:
//   g *= min(1, threshold / ||g||);  g = clip(g, -10, 10);  g /= std(g) + 1e-8          (np.std: population, over all elements)
// scratch[0] = sum g^2 (pass 1), scratch[1] / [2] = sum / sum of squares of the clamped tensor (pass 2), pass 3 applies.
__global__ void __launch_bounds__(256) ae_clip_sumsq_kernel(const float* __restrict__ g, long long n, double* scratch) {
    __shared__ double sc[32];
    float acc = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc = fmaf(g[i], g[i], acc);
    block_accumulate(acc, &scratch[0], sc);
}
__device__ __forceinline__ float ae_norm_scale(const double* scratch, float threshold) {
    if (threshold <= 0.0f) return 1.0f;
    const double nrm = sqrt(scratch[0]);
    return nrm > (double)threshold ? (float)((double)threshold / nrm) : 1.0f;
}
__global__ void __launch_bounds__(256) ae_clip_stats_kernel(const float* __restrict__ g, long long n, float threshold, double* scratch) {
    const float s = ae_norm_scale(scratch, threshold);
    double a1 = 0.0, a2 = 0.0;       // fp64 per thread: the variance is a difference of two nearly equal sums when |mean| >> std
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)clip10(g[i] * s);
        a1 += v; a2 += v * v;
    }
    // block reduction in fp64
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    a1 = warp_sum(a1); a2 = warp_sum(a2);
    __shared__ double s1[8], s2[8];
    if (lane == 0) { s1[wid] = a1; s2[wid] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < 8; ++i) { t1 += s1[i]; t2 += s2[i]; }
        atomicAdd(&scratch[1], t1);
        atomicAdd(&scratch[2], t2);
    }
}
__global__ void __launch_bounds__(256) ae_clip_apply_kernel(const float* __restrict__ g, float* __restrict__ out, long long n, float threshold,
                                                            const double* __restrict__ scratch) {
    const float s = ae_norm_scale(scratch, threshold);
    const double mean = scratch[1] / (double)n;
    double var = scratch[2] / (double)n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = threshold > 0.0f ? (float)(1.0 / (sqrt(var) + 1e-8)) : 1.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = threshold > 0.0f ? clip10(g[i] * s) * inv : g[i];
}

// mean over every (ph x pw) window of the zero-padded input (np.mean of the padded slice: padding counts in the divisor)
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(b200nn_pool_geom g, const float* __restrict__ x, float* __restrict__ y,
                                                          long long total) {
    const float inv = 1.0f / (float)(g.ph * g.pw);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % g.C); long long t = i / g.C;
        const int ox = (int)(t % g.OW); t /= g.OW;
        const int oy = (int)(t % g.OH); const int n = (int)(t / g.OH);
        float acc = 0.0f;
        for (int ky = 0; ky < g.ph; ++ky) {
            const int iy = oy * g.sh - g.pad_h + ky;
            if (iy < 0 || iy >= g.H) continue;
            for (int kx = 0; kx < g.pw; ++kx) {
                const int ix = ox * g.sw - g.pad_w + kx;
                if (ix < 0 || ix >= g.W) continue;
                acc += x[(((long long)n * g.H + iy) * g.W + ix) * g.C + c];
            }
        }
        y[i] = acc * inv;
    }
}

// dx[n,iy,ix,c] = chain * sum over the windows that contain (iy,ix) of err[n,oy,ox,c] / (ph * pw)
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(b200nn_pool_geom g, const float* __restrict__ err, const double* __restrict__ ss,
                                                          uint64_t chain, float* __restrict__ dx, double* ss_out, long long total) {
    __shared__ double scratch[32];
    const float s = chain_scale(ss, chain) / (float)(g.ph * g.pw);
    float acc2 = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % g.C); long long t = i / g.C;
        const int ix = (int)(t % g.W); t /= g.W;
        const int iy = (int)(t % g.H); const int n = (int)(t / g.H);
        float acc = 0.0f;
        for (int ky = 0; ky < g.ph; ++ky) {
            const int ty = iy + g.pad_h - ky;
            if (ty < 0 || ty % g.sh != 0) continue;
            const int oy = ty / g.sh;
            if (oy >= g.OH) continue;
            for (int kx = 0; kx < g.pw; ++kx) {
                const int tx = ix + g.pad_w - kx;
                if (tx < 0 || tx % g.sw != 0) continue;
                const int ox = tx / g.sw;
                if (ox >= g.OW) continue;
                acc += err[(((long long)n * g.OH + oy) * g.OW + ox) * g.C + c];
            }
        }
        const float v = acc * s;
        acc2 = fmaf(v, v, acc2);
        dx[i] = v;
    }
    if (ss_out != nullptr) block_accumulate(acc2, ss_out, scratch);
}

// y[n, c] = mean over the HW positions of x[n, :, c]; one warp per (n, 32 channels), lanes = channels
__global__ void __launch_bounds__(256) gap_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int HW, int C) {
    const long long total = (long long)N * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const int n = (int)(i / C);
        const float* xp = x + (long long)n * HW * C + c;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        int p = 0;
        for (; p + 4 <= HW; p += 4) {
            a0 += xp[(long long)p * C]; a1 += xp[(long long)(p + 1) * C]; a2 += xp[(long long)(p + 2) * C]; a3 += xp[(long long)(p + 3) * C];
        }
        for (; p < HW; ++p) a0 += xp[(long long)p * C];
        y[i] = ((a0 + a1) + (a2 + a3)) / (float)HW;
    }
}

__global__ void __launch_bounds__(256) gap_bwd_kernel(const float* __restrict__ err, const double* __restrict__ ss, uint64_t chain,
                                                      float* __restrict__ dx, double* ss_out, int N, int HW, int C) {
    __shared__ double scratch[32];
    const float s = chain_scale(ss, chain) / (float)HW;
    const long long total = (long long)N * HW * C;
    float acc = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const int n = (int)(i / ((long long)HW * C));
        const float v = s * err[(long long)n * C + c];
        acc = fmaf(v, v, acc);
        dx[i] = v;
    }
    if (ss_out != nullptr) block_accumulate(acc, ss_out, scratch);
}

// nearest-neighbour upsampling by (fh, fw): np.repeat along H then W
__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                                           int fh, int fw) {
    const long long total = (long long)N * H * fh * W * fw * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); long long t = i / C;
        const int ox = (int)(t % (W * fw)); t /= (W * fw);
        const int oy = (int)(t % (H * fh)); const int n = (int)(t / (H * fh));
        y[i] = x[(((long long)n * H + oy / fh) * W + ox / fw) * C + c];
    }
}

__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ err, const double* __restrict__ ss, uint64_t chain,
                                                           float* __restrict__ dx, double* ss_out, int N, int H, int W, int C, int fh,
                                                           int fw) {
    __shared__ double scratch[32];
    const float s = chain_scale(ss, chain);
    const long long total = (long long)N * H * W * C;
    float acc2 = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); long long t = i / C;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H); const int n = (int)(t / H);
        float acc = 0.0f;
        for (int a = 0; a < fh; ++a)
            for (int b = 0; b < fw; ++b)
                acc += err[(((long long)n * H * fh + iy * fh + a) * W * fw + ix * fw + b) * C + c];
        const float v = s * acc;
        acc2 = fmaf(v, v, acc2);
        dx[i] = v;
    }
    if (ss_out != nullptr) block_accumulate(acc2, ss_out, scratch);
}

// im2col with columns ordered (ky, kx, c) — any channel count (scalar accesses): the 1-D convolution's im2col_1d order
// (preprocessing.py:128-156) when kh == 1
__global__ void __launch_bounds__(256) im2col_any_kernel(b200nn_conv_geom g, const float* __restrict__ x, float* __restrict__ col,
                                                         long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % g.C); long long t = i / g.C;
        const int kx = (int)(t % g.kw); t /= g.kw;
        const int ky = (int)(t % g.kh); t /= g.kh;
        const int ox = (int)(t % g.OW); t /= g.OW;
        const int oy = (int)(t % g.OH); const int n = (int)(t / g.OH);
        const int iy = oy * g.sh - g.ph + ky, ix = ox * g.sw - g.pw + kx;
        col[i] = (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) ? x[(((long long)n * g.H + iy) * g.W + ix) * g.C + c] : 0.0f;
    }
}

__global__ void __launch_bounds__(256) col2im_any_kernel(b200nn_conv_geom g, const float* __restrict__ dcol, const double* __restrict__ ss,
                                                         uint64_t chain, float* __restrict__ dx, double* ss_out, long long total) {
    __shared__ double scratch[32];
    const float s = chain_scale(ss, chain);
    const int K = g.kh * g.kw * g.C;
    float acc2 = 0.0f;
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
                acc += dcol[(((long long)n * g.OH + oy) * g.OW + ox) * K + (long long)(ky * g.kw + kx) * g.C + c];
            }
        }
        const float v = s * acc;
        acc2 = fmaf(v, v, acc2);
        dx[i] = v;
    }
    if (ss_out != nullptr) block_accumulate(acc2, ss_out, scratch);
}

static int grid_for(long long total) {
    long long b = (total + 255) / 256;
    if (b > kNumSMs * 32) b = kNumSMs * 32;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace b200nn

using namespace b200nn;

extern "C" {

int b200nn_mul(const float* a, const float* b, const double* ss, uint64_t chain, float* out, double* ss_out, long long n, void* stream) {
    B200NN_REQUIRE(a != nullptr && b != nullptr && out != nullptr && n >= 0, "mul: bad argument");
    if (n == 0) return B200NN_OK;
    mul_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(a, b, ss, chain, out, ss_out, n);
    B200NN_LAUNCH_CHECK("mul");
    return B200NN_OK;
}

int b200nn_axpy(const float* a, const float* b, float s, float* out, long long n, void* stream) {
    B200NN_REQUIRE(a != nullptr && b != nullptr && out != nullptr && n >= 0, "axpy: bad argument");
    if (n == 0) return B200NN_OK;
    axpy_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(a, b, s, out, n);
    B200NN_LAUNCH_CHECK("axpy");
    return B200NN_OK;
}

int b200nn_ae_clip(const float* g, float* out, long long n, float threshold, double* scratch3, void* stream) {
    B200NN_REQUIRE(g != nullptr && out != nullptr && scratch3 != nullptr && n >= 0, "ae_clip: bad argument");
    if (n == 0) return B200NN_OK;
    cudaStream_t st = as_stream(stream);
    B200NN_CUDA(cudaMemsetAsync(scratch3, 0, 3 * sizeof(double), st));
    ae_clip_sumsq_kernel<<<grid_for(n), 256, 0, st>>>(g, n, scratch3);
    B200NN_LAUNCH_CHECK("ae_clip_sumsq");
    ae_clip_stats_kernel<<<grid_for(n), 256, 0, st>>>(g, n, threshold, scratch3);
    B200NN_LAUNCH_CHECK("ae_clip_stats");
    ae_clip_apply_kernel<<<grid_for(n), 256, 0, st>>>(g, out, n, threshold, scratch3);
    B200NN_LAUNCH_CHECK("ae_clip_apply");
    return B200NN_OK;
}

static int check_pool(const char* who, const b200nn_pool_geom* g) {
    B200NN_REQUIRE(g != nullptr && g->N > 0 && g->H > 0 && g->W > 0 && g->C > 0 && g->ph > 0 && g->pw > 0 && g->sh > 0 && g->sw > 0 &&
                   g->OH > 0 && g->OW > 0 && g->pad_h >= 0 && g->pad_w >= 0, "%s: bad geometry", who);
    return B200NN_OK;
}

int b200nn_avgpool2d_fwd(const b200nn_pool_geom* g, const float* x, float* y, void* stream) {
    if (int rc = check_pool("avgpool2d_fwd", g)) return rc;
    const long long total = (long long)g->N * g->OH * g->OW * g->C;
    avgpool_fwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(*g, x, y, total);
    B200NN_LAUNCH_CHECK("avgpool2d_fwd");
    return B200NN_OK;
}

int b200nn_avgpool2d_bwd(const b200nn_pool_geom* g, const float* err, const double* ss, uint64_t chain, float* dx, double* ss_out,
                         void* stream) {
    if (int rc = check_pool("avgpool2d_bwd", g)) return rc;
    const long long total = (long long)g->N * g->H * g->W * g->C;
    avgpool_bwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(*g, err, ss, chain, dx, ss_out, total);
    B200NN_LAUNCH_CHECK("avgpool2d_bwd");
    return B200NN_OK;
}

int b200nn_gap2d_fwd(const float* x, float* y, int N, int HW, int C, void* stream) {
    B200NN_REQUIRE(N > 0 && HW > 0 && C > 0, "gap2d_fwd: bad shape");
    gap_fwd_kernel<<<grid_for((long long)N * C), 256, 0, as_stream(stream)>>>(x, y, N, HW, C);
    B200NN_LAUNCH_CHECK("gap2d_fwd");
    return B200NN_OK;
}

int b200nn_gap2d_bwd(const float* err, const double* ss, uint64_t chain, float* dx, double* ss_out, int N, int HW, int C, void* stream) {
    B200NN_REQUIRE(N > 0 && HW > 0 && C > 0, "gap2d_bwd: bad shape");
    gap_bwd_kernel<<<grid_for((long long)N * HW * C), 256, 0, as_stream(stream)>>>(err, ss, chain, dx, ss_out, N, HW, C);
    B200NN_LAUNCH_CHECK("gap2d_bwd");
    return B200NN_OK;
}

int b200nn_upsample2d_fwd(const float* x, float* y, int N, int H, int W, int C, int fh, int fw, void* stream) {
    B200NN_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && fh > 0 && fw > 0, "upsample2d_fwd: bad shape");
    upsample_fwd_kernel<<<grid_for((long long)N * H * fh * W * fw * C), 256, 0, as_stream(stream)>>>(x, y, N, H, W, C, fh, fw);
    B200NN_LAUNCH_CHECK("upsample2d_fwd");
    return B200NN_OK;
}

int b200nn_upsample2d_bwd(const float* err, const double* ss, uint64_t chain, float* dx, double* ss_out, int N, int H, int W, int C,
                          int fh, int fw, void* stream) {
    B200NN_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && fh > 0 && fw > 0, "upsample2d_bwd: bad shape");
    upsample_bwd_kernel<<<grid_for((long long)N * H * W * C), 256, 0, as_stream(stream)>>>(err, ss, chain, dx, ss_out, N, H, W, C, fh, fw);
    B200NN_LAUNCH_CHECK("upsample2d_bwd");
    return B200NN_OK;
}

int b200nn_im2col(const b200nn_conv_geom* g, const float* x, float* col, void* stream) {
    B200NN_REQUIRE(g != nullptr && x != nullptr && col != nullptr, "im2col: null argument");
    const long long total = (long long)g->N * g->OH * g->OW * g->kh * g->kw * g->C;
    im2col_any_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(*g, x, col, total);
    B200NN_LAUNCH_CHECK("im2col");
    return B200NN_OK;
}

int b200nn_col2im(const b200nn_conv_geom* g, const float* dcol, const double* ss, uint64_t chain, float* dx, double* ss_out, void* stream) {
    B200NN_REQUIRE(g != nullptr && dcol != nullptr && dx != nullptr, "col2im: null argument");
    const long long total = (long long)g->N * g->H * g->W * g->C;
    col2im_any_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(*g, dcol, ss, chain, dx, ss_out, total);
    B200NN_LAUNCH_CHECK("col2im");
    return B200NN_OK;
}

}  // extern "C"
