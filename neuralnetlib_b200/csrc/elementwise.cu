// elementwise.cu — API basics, step bookkeeping, CUDA-graph helpers, activations, clip helpers,
// softmax / losses, and the fused multi-tensor optimizers. All HBM-bound: 128-bit vector access,
// grid-stride loops sized in multiples of the SM count, one fp64 atomic per CTA for norms.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "opt_common.cuh"

namespace b200nn {

static thread_local char g_last_error[512] = {0};
static long long g_launches = 0;
static std::mutex g_graph_mu;
static std::unordered_map<void*, long long> g_graph_kernels;   // graph exec -> kernel nodes (see b200nn_graph_end)

void count_launch() { __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

static inline int grid_for(long long work_items, int threads, int items_per_thread = 4, int max_waves = 8) {
    long long blocks = (work_items + (long long)threads * items_per_thread - 1) / ((long long)threads * items_per_thread);
    if (blocks < 1) blocks = 1;
    const long long cap = (long long)kNumSMs * max_waves;
    return (int)(blocks < cap ? blocks : cap);
}

// ------------------------------------------------------------------------------------------------
__global__ void step_begin_kernel(double* ss, int n_slots, double* acc, int n_acc, long long* counter, long long inc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) ss[i] = 0.0;
    if (i < n_acc) acc[i] = 0.0;
    if (i == 0 && counter != nullptr) *counter += inc;
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx,
                                                          float* __restrict__ dst, long long rows, long long row_elems) {
    // one warp per 128-float chunk of a row; rows are contiguous so every access is a coalesced line
    const long long chunks_per_row = (row_elems + 127) / 128;
    const long long total = rows * chunks_per_row;
    const int lane = threadIdx.x & 31;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
         w += ((long long)gridDim.x * blockDim.x) >> 5) {
        const long long r = w / chunks_per_row, c = w % chunks_per_row;
        const float* s = src + idx[r] * row_elems + c * 128;
        float* d = dst + r * row_elems + c * 128;
        const long long left = row_elems - c * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = lane + j * 32;
            if (e < left) d[e] = s[e];
        }
    }
}

__global__ void zero_f64_kernel(double* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) act_fwd_kernel(int act, float alpha, const float* __restrict__ x,
                                                      float* __restrict__ y, long long n) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = x4[i];
        v.x = act_apply(act, v.x, alpha); v.y = act_apply(act, v.y, alpha);
        v.z = act_apply(act, v.z, alpha); v.w = act_apply(act, v.w, alpha);
        y4[i] = v;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = act_apply(act, x[i], alpha);
}

// dx = s * err * f'(y); ss_out += sum(dx^2).  has_y == false => f' = 1 (pure scale / sumsq of scaled err)
template <bool kWrite>
__global__ void __launch_bounds__(256) act_bwd_kernel(int act, float alpha, const float* __restrict__ y,
                                                      const float* __restrict__ err, const double* __restrict__ ss,
                                                      uint64_t chain, float* __restrict__ dx, double* ss_out,
                                                      long long n) {
    __shared__ double scratch[32];
    const float s = chain_scale(ss, chain);
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4* e4 = reinterpret_cast<const float4*>(err);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    float4* d4 = reinterpret_cast<float4*>(dx);
    float acc = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 e = e4[i];
        float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
        if (y != nullptr) {
            const float4 yy = y4[i];
            g.x = act_grad_from_y(act, yy.x, alpha); g.y = act_grad_from_y(act, yy.y, alpha);
            g.z = act_grad_from_y(act, yy.z, alpha); g.w = act_grad_from_y(act, yy.w, alpha);
        }
        e.x = s * e.x * g.x; e.y = s * e.y * g.y; e.z = s * e.z * g.z; e.w = s * e.w * g.w;
        acc += e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w;
        if (kWrite) d4[i] = e;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float g = y != nullptr ? act_grad_from_y(act, y[i], alpha) : 1.0f;
        const float e = s * err[i] * g;
        acc += e * e;
        if (kWrite) dx[i] = e;
    }
    if (ss_out != nullptr) block_accumulate(acc, ss_out, scratch);
}

// ------------------------------------------------------------------------------------------------
// one warp per row; cols arbitrary
__global__ void __launch_bounds__(256) softmax_kernel(const float* __restrict__ x, float* __restrict__ p, int rows,
                                                      int cols) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float* xr = x + (long long)warp * cols;
    float* pr = p + (long long)warp * cols;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, xr[c]);
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int c = lane; c < cols; c += 32) sum += expf(xr[c] - mx);
    sum = warp_sum(sum);
    for (int c = lane; c < cols; c += 32) pr[c] = expf(xr[c] - mx) / sum;
}

// `param`: Huber's delta (losses.py:158-176)
__device__ __forceinline__ float loss_term(int loss, float y, float p, float param) {
    if (loss == B200NN_LOSS_MSE) return (y - p) * (y - p);
    if (loss == B200NN_LOSS_MAE) return fabsf(y - p);                                   // losses.py:148
    if (loss == B200NN_LOSS_HUBER) {                                                    // losses.py:163-168
        const float e = fabsf(y - p);
        return e <= param ? 0.5f * e * e : param * (e - 0.5f * param);
    }
    const float pc = fminf(fmaxf(p, 1e-15f), 1.0f);  // 1 - 1e-15 rounds to 1 in fp32
    if (loss == B200NN_LOSS_CCE) return -y * logf(pc);
    const float q = fmaxf(1.0f - pc, 1.1102230246251565e-15f);  // fp64 value of 1-(1-1e-15)
    return -(y * logf(pc) + (1.0f - y) * logf(q));
}
__device__ __forceinline__ float loss_deriv(int loss, float y, float p, float inv_size, float param) {
    if (loss == B200NN_LOSS_MSE) return 2.0f * (p - y) * inv_size;
    if (loss == B200NN_LOSS_MAE) return p > y ? 1.0f : -1.0f;                           // losses.py:151 (not divided by the size)
    if (loss == B200NN_LOSS_HUBER) {                                                    // losses.py:170-172: sign as the reference has it
        const float e = y - p;
        return fabsf(e) <= param ? e : (e > 0.0f ? param : (e < 0.0f ? -param : 0.0f));
    }
    const float pc = fminf(fmaxf(p, 1e-15f), 1.0f);
    if (loss == B200NN_LOSS_CCE) return -y / pc;
    const float q = fmaxf(1.0f - pc, 1.1102230246251565e-15f);
    return (pc - y) / (pc * q);
}

// warp per row (cols > 1) or thread per element (cols == 1)
__global__ void __launch_bounds__(256) loss_kernel(int loss, int fused_pair, const float* __restrict__ p,
                                                   const float* __restrict__ y, float* __restrict__ err,
                                                   double* loss_acc, double* ss_out, int rows, int cols, double global_elems,
                                                   float param) {
    __shared__ double scratch[32];
    float lsum = 0.0f, esq = 0.0f, correct = 0.0f;
    // MeanSquaredError.derivative divides by y_true.size of the WHOLE batch (losses.py:84): under data parallelism that is
    // the global element count, not this rank's shard
    const float inv_size = (float)(1.0 / global_elems);
    if (cols == 1) {
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
            const float pv = p[r], yv = y[r];
            lsum += loss_term(loss, yv, pv, param);
            const float e = fused_pair ? (pv - yv) : loss_deriv(loss, yv, pv, inv_size, param);
            if (err != nullptr) err[r] = e;
            esq += e * e;
            correct += ((pv >= 0.5f ? 1.0f : 0.0f) == yv) ? 1.0f : 0.0f;
        }
    } else {
        const int lane = threadIdx.x & 31;
        const int nwarps = (gridDim.x * blockDim.x) >> 5;
        for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
            const float* pr = p + (long long)r * cols;
            const float* yr = y + (long long)r * cols;
            float bp = -INFINITY, by = -INFINITY;
            int ip = 0x7fffffff, iy = 0x7fffffff;
            for (int c = lane; c < cols; c += 32) {
                const float pv = pr[c], yv = yr[c];
                lsum += loss_term(loss, yv, pv, param);
                const float e = fused_pair ? (pv - yv) : loss_deriv(loss, yv, pv, inv_size, param);
                if (err != nullptr) err[(long long)r * cols + c] = e;
                esq += e * e;
                if (pv > bp) { bp = pv; ip = c; }  // strict > keeps the first index within a lane
                if (yv > by) { by = yv; iy = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {  // first-index tie break across lanes (np.argmax)
                const float obp = __shfl_xor_sync(0xffffffffu, bp, o);
                const int oip = __shfl_xor_sync(0xffffffffu, ip, o);
                if (obp > bp || (obp == bp && oip < ip)) { bp = obp; ip = oip; }
                const float oby = __shfl_xor_sync(0xffffffffu, by, o);
                const int oiy = __shfl_xor_sync(0xffffffffu, iy, o);
                if (oby > by || (oby == by && oiy < iy)) { by = oby; iy = oiy; }
            }
            if (lane == 0 && ip == iy) correct += 1.0f;
        }
    }
    const double l = block_sum_to_double(lsum, scratch);
    if (threadIdx.x == 0 && loss_acc != nullptr) atomicAdd(&loss_acc[0], l);
    const double c = block_sum_to_double(correct, scratch);
    if (threadIdx.x == 0 && loss_acc != nullptr) atomicAdd(&loss_acc[1], c);
    if (ss_out != nullptr) block_accumulate(esq, ss_out, scratch);
}

// ------------------------------------------------------------------------------------------------
// multi-tensor optimizers. block_map holds (tensor, chunk) pairs; one CTA per 4096-element chunk.
__global__ void __launch_bounds__(256) sumsq_multi_kernel(const b200nn_param* __restrict__ params,
                                                          const int* __restrict__ block_map, double* gss) {
    __shared__ double scratch[32];
    const int ti = block_map[2 * blockIdx.x], chunk = block_map[2 * blockIdx.x + 1];
    const b200nn_param P = params[ti];
    const long long base = (long long)chunk * B200NN_OPT_CHUNK;
    const long long end = min(P.n, base + B200NN_OPT_CHUNK);
    float acc = 0.0f;
    if ((reinterpret_cast<uintptr_t>(P.g) & 15) == 0) {
        const long long e4 = base + ((end - base) & ~3LL);
        for (long long i = base + threadIdx.x * 4; i < e4; i += 256 * 4) {
            const float4 g = *reinterpret_cast<const float4*>(P.g + i);
            acc += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
        }
        for (long long i = e4 + threadIdx.x; i < end; i += 256) acc += P.g[i] * P.g[i];
    } else {
        for (long long i = base + threadIdx.x; i < end; i += 256) acc += P.g[i] * P.g[i];
    }
    block_accumulate(acc, &gss[ti], scratch);
}

// FUSED: gradient norm + update in ONE launch for parameter sets of at most one CTA per SM (every CTA is resident, so a
// grid-wide barrier on a device counter is safe): each CTA adds its chunk's sum of squares to gss[tensor], arrives, waits
// for all, then updates. The last CTA to leave re-zeroes gss and the two counters (stored after gss[n_params]) for the next
// launch -- no memset node, no separate norm kernel on the critical path of small models.
template <bool FUSED>
__global__ void __launch_bounds__(256) adam_multi_kernel(const b200nn_param* __restrict__ params, int n_layers, int n_params,
                                                         double* __restrict__ gss,
                                                         const double* __restrict__ hyper,
                                                         const long long* __restrict__ step,
                                                         const int* __restrict__ block_map, const double* __restrict__ ss) {
    __shared__ OptScalars sh;
    __shared__ double scratch[32];
    const int ti = block_map[2 * blockIdx.x], chunk = block_map[2 * blockIdx.x + 1];
    const b200nn_param P = params[ti];
    unsigned* sync = reinterpret_cast<unsigned*>(gss + n_params);
    if (FUSED) {
        const long long b0 = (long long)chunk * B200NN_OPT_CHUNK, e0 = min(P.n, b0 + B200NN_OPT_CHUNK);
        float part = 0.0f;
        for (long long i = b0 + threadIdx.x; i < e0; i += 256) part = fmaf(P.g[i], P.g[i], part);
        const double r = block_sum_to_double(part, scratch);
        if (threadIdx.x == 0) {
            atomicAdd(&gss[ti], r);
            __threadfence();
            atomicAdd(&sync[0], 1u);
            while (*reinterpret_cast<volatile unsigned*>(&sync[0]) < gridDim.x) {
            }
            __threadfence();
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
        const double pre = (double)chain_scale(ss, P.chain);                 // deferred error clips (models.py:214) owed by g
        const double n2 = *reinterpret_cast<volatile double*>(&gss[ti]) * pre * pre;
        const long long t = *step - (long long)n_layers + P.t_index;         // optimizers.py:236 (Q3)
        sh.clip = (float)(pre * ((P.clip && n2 > 1.0) ? rsqrt(n2) : 1.0));          // models.py:240-242
        sh.lr = (float)lr; sh.b1 = (float)b1; sh.b2 = (float)b2;
        sh.omb1 = (float)(1.0 - b1); sh.omb2 = (float)(1.0 - b2);
        sh.inv_bc1 = (float)(1.0 / (1.0 - powi(b1, t)));
        sh.inv_bc2 = (float)(1.0 / (1.0 - powi(b2, t)));
        sh.eps = (float)eps;
    }
    const long long base = (long long)chunk * B200NN_OPT_CHUNK;
    const long long end = min(P.n, base + B200NN_OPT_CHUNK);
    const bool aligned = ((reinterpret_cast<uintptr_t>(P.p) | reinterpret_cast<uintptr_t>(P.g) |
                           reinterpret_cast<uintptr_t>(P.m) | reinterpret_cast<uintptr_t>(P.v)) & 15) == 0;
    long long tail = base;
    if (aligned) {
        // the chunk's 16 loads per thread are issued up front, BEFORE waiting for thread 0's scalars: one memory round trip
        constexpr int IT = B200NN_OPT_CHUNK / (256 * 4);
        const long long e4 = base + ((end - base) & ~3LL);
        float4 p[IT], g[IT], m[IT], v[IT];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const long long i = base + ((long long)it * 256 + threadIdx.x) * 4;
            if (i < e4) {
                p[it] = *reinterpret_cast<const float4*>(P.p + i);
                g[it] = *reinterpret_cast<const float4*>(P.g + i);
                m[it] = *reinterpret_cast<const float4*>(P.m + i);
                v[it] = *reinterpret_cast<const float4*>(P.v + i);
            }
        }
        __syncthreads();
        const OptScalars s = sh;
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const long long i = base + ((long long)it * 256 + threadIdx.x) * 4;
            if (i < e4) {
                adam_elem(p[it].x, g[it].x, m[it].x, v[it].x, s); adam_elem(p[it].y, g[it].y, m[it].y, v[it].y, s);
                adam_elem(p[it].z, g[it].z, m[it].z, v[it].z, s); adam_elem(p[it].w, g[it].w, m[it].w, v[it].w, s);
                *reinterpret_cast<float4*>(P.p + i) = p[it];
                *reinterpret_cast<float4*>(P.m + i) = m[it];
                *reinterpret_cast<float4*>(P.v + i) = v[it];
            }
        }
        tail = e4;
    } else {
        __syncthreads();
    }
    const OptScalars s = sh;
    for (long long i = tail + threadIdx.x; i < end; i += 256) {
        float p = P.p[i], m = P.m[i], v = P.v[i];
        adam_elem(p, P.g[i], m, v, s);
        P.p[i] = p; P.m[i] = m; P.v[i] = v;
    }
    if (FUSED) {
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&sync[1], 1u) == gridDim.x - 1) {   // last CTA out: everyone has read gss
            for (int i = 0; i < n_params; ++i) gss[i] = 0.0;
            sync[0] = 0u;
            sync[1] = 0u;
        }
    }
}

// Single-launch Adam for tiny parameter sets (<= 16 chunks: the last layer updated in a step): every load the kernel needs is
// issued before the grid-wide barrier that completes the gradient norms, so the critical path is one memory round trip, the
// barrier, and the update. gss and its two trailing counters are zero on entry and re-zeroed by the last CTA to leave.
__global__ void __launch_bounds__(256) adam_small_fused_kernel(const b200nn_param* __restrict__ params, int n_layers, int n_params,
                                                               double* __restrict__ gss, const double* __restrict__ hyper,
                                                               const long long* __restrict__ step,
                                                               const int* __restrict__ block_map, const double* __restrict__ ss) {
    __shared__ OptScalars sh;
    __shared__ double scratch[32];
    constexpr int IT = B200NN_OPT_CHUNK / 256;
    const int ti = block_map[2 * blockIdx.x], chunk = block_map[2 * blockIdx.x + 1];
    const b200nn_param P = params[ti];
    unsigned* sync = reinterpret_cast<unsigned*>(gss + n_params);
    const long long base = (long long)chunk * B200NN_OPT_CHUNK, end = min(P.n, base + B200NN_OPT_CHUNK);
    float p[IT], g[IT], m[IT], v[IT];
    float part = 0.0f;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const long long i = base + it * 256 + threadIdx.x;
        p[it] = g[it] = m[it] = v[it] = 0.0f;
        if (i < end) { p[it] = P.p[i]; g[it] = P.g[i]; m[it] = P.m[i]; v[it] = P.v[i]; }
    }
    double lr = 0, b1 = 0, b2 = 0, eps = 0, pre = 1;
    long long t = 1;
    if (threadIdx.x == 0) {
        lr = hyper[0]; b1 = hyper[1]; b2 = hyper[2]; eps = hyper[3];
        pre = (double)chain_scale(ss, P.chain);
        t = *step - (long long)n_layers + P.t_index;                          // optimizers.py:236 (Q3)
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) part = fmaf(g[it], g[it], part);
    const double r = block_sum_to_double(part, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(&gss[ti], r);
        __threadfence();
        atomicAdd(&sync[0], 1u);
        sh.lr = (float)lr; sh.b1 = (float)b1; sh.b2 = (float)b2;              // everything that does not need the norm
        sh.omb1 = (float)(1.0 - b1); sh.omb2 = (float)(1.0 - b2);
        sh.inv_bc1 = (float)(1.0 / (1.0 - powi(b1, t)));
        sh.inv_bc2 = (float)(1.0 / (1.0 - powi(b2, t)));
        sh.eps = (float)eps;
        while (*reinterpret_cast<volatile unsigned*>(&sync[0]) < gridDim.x) {
        }
        __threadfence();
        const double n2 = *reinterpret_cast<volatile double*>(&gss[ti]) * pre * pre;
        sh.clip = (float)(pre * ((P.clip && n2 > 1.0) ? rsqrt(n2) : 1.0));          // models.py:214 (deferred) and :240-242
    }
    __syncthreads();
    const OptScalars s = sh;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const long long i = base + it * 256 + threadIdx.x;
        if (i < end) {
            adam_elem(p[it], g[it], m[it], v[it], s);
            P.p[i] = p[it]; P.m[i] = m[it]; P.v[i] = v[it];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&sync[1], 1u) == gridDim.x - 1) {       // last CTA out: everyone has read gss
        for (int i = 0; i < n_params; ++i) gss[i] = 0.0;
        sync[0] = 0u;
        sync[1] = 0u;
    }
}

__global__ void __launch_bounds__(256) sgd_multi_kernel(const b200nn_param* __restrict__ params,
                                                        const double* __restrict__ gss,
                                                        const double* __restrict__ hyper,
                                                        const int* __restrict__ block_map, const double* __restrict__ ss) {
    const int ti = block_map[2 * blockIdx.x], chunk = block_map[2 * blockIdx.x + 1];
    const b200nn_param P = params[ti];
    const double pre = (double)chain_scale(ss, P.chain);
    const double n2 = gss[ti] * pre * pre;
    const float scale = (float)(hyper[0] * pre * ((P.clip && n2 > 1.0) ? rsqrt(n2) : 1.0));
    const long long base = (long long)chunk * B200NN_OPT_CHUNK;
    const long long end = min(P.n, base + B200NN_OPT_CHUNK);
    for (long long i = base + threadIdx.x; i < end; i += 256) P.p[i] -= scale * P.g[i];  // optimizers.py:48
}


// ------------------------------------------------------------------------------------------------
// The remaining update rules of optimizers.py on the same multi-tensor skeleton (one CTA per 4096-element chunk of one tensor;
// gss[tensor] = ||g||^2 from sumsq_multi): Momentum (:63-111), RMSprop (:114-164), Adam with its own clip_norm / clip_value
// (:190-202), AdaBelief (:286-402), RAdam (:405-523). hyper = {lr, beta1 | momentum | rho, beta2, eps, clip_norm, clip_value}
// (clip_norm / clip_value <= 0: off).
struct OptScalarsX {
    float clip, clip_value, lr, b1, b2, omb1, omb2, inv_bc1, inv_bc2, eps, rt;
    int rect;
};

template <int KIND>
__device__ __forceinline__ void opt_elem(float& p, float g, float& m, float& v, const OptScalarsX& s) {
    g *= s.clip;
    if (s.clip_value > 0.0f) g = fminf(fmaxf(g, -s.clip_value), s.clip_value);     // np.clip(grad, -clip_value, clip_value)
    if (KIND == B200NN_OPT_MOMENTUM) {            // optimizers.py:79-82: v = momentum * v - lr * g ; w += v
        m = s.b1 * m - s.lr * g;
        p += m;
        return;
    }
    if (KIND == B200NN_OPT_RMSPROP) {             // optimizers.py:130-136
        v = s.b1 * v + s.omb1 * g * g;
        p -= s.lr * g / (sqrtf(v) + s.eps);
        return;
    }
    float upd;
    m = s.b1 * m + s.omb1 * g;
    if (KIND == B200NN_OPT_ADABELIEF) {           // optimizers.py:322-341: second moment of the residual g - m
        const float r = g - m;
        v = s.b2 * v + s.omb2 * r * r;
        const float denom = sqrtf(v * s.inv_bc2 + s.eps);
        upd = s.lr * (m * s.inv_bc1) / fmaxf(denom, 1e-16f);
    } else {
        v = s.b2 * v + s.omb2 * g * g;
        if (KIND == B200NN_OPT_RADAM) {           // optimizers.py:446-466
            if (s.rect) upd = s.rt * s.lr * (m * s.inv_bc1) / fmaxf(sqrtf(v * s.inv_bc2) + s.eps, 1e-16f);
            else upd = s.lr * (m * s.inv_bc1);
        } else {                                  // Adam, optimizers.py:204-225
            upd = s.lr * (m * s.inv_bc1) / fmaxf(sqrtf(v * s.inv_bc2) + s.eps, 1e-16f);
        }
    }
    if (!isfinite(upd)) upd = 0.0f;               // np.nan_to_num(update, nan=0, posinf=0, neginf=0)
    p -= upd;
}

template <int KIND>
__global__ void __launch_bounds__(256) opt_multi_kernel(const b200nn_param* __restrict__ params, int n_layers,
                                                        const double* __restrict__ gss, const double* __restrict__ hyper,
                                                        const long long* __restrict__ step, const int* __restrict__ block_map,
                                                        const double* __restrict__ ss) {
    __shared__ OptScalarsX sh;
    const int ti = block_map[2 * blockIdx.x], chunk = block_map[2 * blockIdx.x + 1];
    const b200nn_param P = params[ti];
    if (threadIdx.x == 0) {
        const double lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], cn = hyper[4], cv = hyper[5];
        const double pre = (double)chain_scale(ss, P.chain);                 // deferred error clips (models.py:214) owed by g
        const double n2 = gss[ti] * pre * pre;
        double clip = pre * ((P.clip && n2 > 1.0) ? rsqrt(n2) : 1.0);        // models.py:240-242
        if (cn > 0.0) {                                                       // the optimizer's own clip_norm on the clipped gradient
            const double norm = sqrt(gss[ti]) * clip;
            if (norm > cn) clip *= cn / (norm + 1e-16);
        }
        sh.clip = (float)clip;
        sh.clip_value = (float)cv;
        sh.lr = (float)lr; sh.b1 = (float)b1; sh.b2 = (float)b2;
        sh.omb1 = (float)(1.0 - b1); sh.omb2 = (float)(1.0 - b2); sh.eps = (float)eps;
        sh.inv_bc1 = sh.inv_bc2 = 1.0f; sh.rt = 1.0f; sh.rect = 0;
        if (KIND == B200NN_OPT_ADAM || KIND == B200NN_OPT_ADABELIEF || KIND == B200NN_OPT_RADAM) {
            const long long t = *step - (long long)n_layers + P.t_index;     // t counts update() calls = layers (Q3)
            const double b1t = powi(b1, t), b2t = powi(b2, t);
            sh.inv_bc1 = (float)(1.0 / (1.0 - b1t));
            sh.inv_bc2 = (float)(1.0 / (1.0 - b2t));
            if (KIND == B200NN_OPT_RADAM) {
                const double rho_inf = 2.0 / (1.0 - b2) - 1.0;
                const double rho_t = rho_inf - 2.0 * (double)t * b2t / (1.0 - b2t);
                sh.rect = rho_t > 4.0 ? 1 : 0;
                if (sh.rect) sh.rt = (float)sqrt(((rho_t - 4.0) * (rho_t - 2.0) * rho_inf) / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t));
            }
        }
    }
    __syncthreads();
    const OptScalarsX s = sh;
    const long long base = (long long)chunk * B200NN_OPT_CHUNK;
    const long long end = min(P.n, base + B200NN_OPT_CHUNK);
    for (long long i = base + threadIdx.x; i < end; i += 256) {
        float p = P.p[i], m = P.m != nullptr ? P.m[i] : 0.0f, v = P.v != nullptr ? P.v[i] : 0.0f;
        opt_elem<KIND>(p, P.g[i], m, v, s);
        P.p[i] = p;
        if (P.m != nullptr) P.m[i] = m;
        if (P.v != nullptr) P.v[i] = v;
    }
}

}  // namespace b200nn

using namespace b200nn;

extern "C" {

int b200nn_version(void) { return 100; }

#ifndef B200NN_HAS_TCGEN05
#define B200NN_HAS_TCGEN05 1   /* gemm_tcgen05.cuh: kind::tf32 GEMM behind B200NN_MATH_TF32 */
#endif
int b200nn_has_tcgen05(void) { return B200NN_HAS_TCGEN05; }

long long b200nn_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

size_t b200nn_last_error(char* buf, size_t cap) {
    const size_t n = strlen(g_last_error);
    if (buf != nullptr && cap > 0) {
        const size_t c = n < cap - 1 ? n : cap - 1;
        memcpy(buf, g_last_error, c);
        buf[c] = 0;
    }
    return n;
}

int b200nn_gather_rows(const float* src, const long long* idx, float* dst, long long rows, long long row_elems,
                       void* stream) {
    if (rows <= 0 || row_elems <= 0) return B200NN_OK;
    const long long warps = rows * ((row_elems + 127) / 128);
    long long blocks = (warps * 32 + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    gather_rows_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(src, idx, dst, rows, row_elems);
    B200NN_LAUNCH_CHECK("gather_rows");
    return B200NN_OK;
}

int b200nn_step_begin(double* ss, int n_slots, double* acc, int n_acc, long long* counter, long long inc, void* stream) {
    int n = n_slots > n_acc ? n_slots : n_acc;
    if (n < 1) n = 1;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    step_begin_kernel<<<blocks, threads, 0, as_stream(stream)>>>(ss, n_slots, acc, n_acc, counter, inc);
    B200NN_LAUNCH_CHECK("step_begin");
    return B200NN_OK;
}

int b200nn_zero_f64(double* p, int n, void* stream) {
    if (n <= 0) return B200NN_OK;
    zero_f64_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(p, n);
    B200NN_LAUNCH_CHECK("zero_f64");
    return B200NN_OK;
}

// Host <-> device plumbing of the public entry points (Sequential.train_on_batch: batch upload, loss read-back) as single
// calls, so that the host side of a step is three ctypes calls: upload x, upload y, launch graph + read loss.
int b200nn_upload(void* dst_dev, const void* src_host, size_t bytes, void* stream) {
    B200NN_REQUIRE(dst_dev != nullptr && src_host != nullptr, "upload: null pointer");
    if (bytes == 0) return B200NN_OK;
    B200NN_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return B200NN_OK;
}

// copies `bytes` from the device to (preferably pinned) host memory and waits for the stream: the value is valid on return
int b200nn_read_sync(void* dst_host, const void* src_dev, size_t bytes, void* stream) {
    B200NN_REQUIRE(dst_host != nullptr && src_dev != nullptr, "read_sync: null pointer");
    B200NN_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    B200NN_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return B200NN_OK;
}

int b200nn_graph_begin(void* stream) {
    B200NN_CUDA(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
    return B200NN_OK;
}

int b200nn_graph_end(void* stream, void** graph_exec_out) {
    cudaGraph_t graph = nullptr;
    B200NN_CUDA(cudaStreamEndCapture(as_stream(stream), &graph));
    // kernel nodes of the captured step: b200nn_graph_launch adds them to the launch count (b200nn_launch_count counts KERNELS)
    size_t n_nodes = 0, n_kernels = 0;
    if (cudaGraphGetNodes(graph, nullptr, &n_nodes) == cudaSuccess && n_nodes > 0) {
        std::vector<cudaGraphNode_t> nodes(n_nodes);
        if (cudaGraphGetNodes(graph, nodes.data(), &n_nodes) == cudaSuccess)
            for (size_t i = 0; i < n_nodes; ++i) {
                cudaGraphNodeType t;
                if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++n_kernels;
            }
    }
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        return B200NN_ECUDA;
    }
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        g_graph_kernels[exec] = (long long)n_kernels;
    }
    *graph_exec_out = exec;
    return B200NN_OK;
}

int b200nn_graph_launch(void* graph_exec, void* stream) {
    B200NN_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), as_stream(stream)));
    long long n = 0;
    {
        std::lock_guard<std::mutex> lk(g_graph_mu);
        auto it = g_graph_kernels.find(graph_exec);
        if (it != g_graph_kernels.end()) n = it->second;
    }
    __atomic_fetch_add(&g_launches, n, __ATOMIC_RELAXED);
    return B200NN_OK;
}

int b200nn_graph_destroy(void* graph_exec) {
    if (graph_exec != nullptr) {
        {
            std::lock_guard<std::mutex> lk(g_graph_mu);
            g_graph_kernels.erase(graph_exec);
        }
        B200NN_CUDA(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)));
    }
    return B200NN_OK;
}

int b200nn_act_fwd(int act, float alpha, const float* x, float* y, long long n, void* stream) {
    B200NN_REQUIRE(act >= 0 && act <= B200NN_ACT_GELU, "act_fwd: unknown activation %d", act);
    B200NN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                   "act_fwd: buffers must be 16-byte aligned");
    if (n <= 0) return B200NN_OK;
    act_fwd_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(act, alpha, x, y, n);
    B200NN_LAUNCH_CHECK("act_fwd");
    return B200NN_OK;
}

int b200nn_act_bwd(int act, float alpha, const float* y, const float* err, const double* ss, uint64_t chain,
                   float* dx, double* ss_out, long long n, void* stream) {
    B200NN_REQUIRE(act >= 0 && act <= B200NN_ACT_GELU, "act_bwd: unknown activation %d", act);
    B200NN_REQUIRE(act != B200NN_ACT_LEAKYRELU || alpha > 0.0f, "act_bwd: LeakyReLU needs alpha > 0 (got %g)", alpha);
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(err) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
                   "act_bwd: buffers must be 16-byte aligned");
    if (n <= 0) return B200NN_OK;
    act_bwd_kernel<true><<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(act, alpha, y, err, ss, chain, dx, ss_out, n);
    B200NN_LAUNCH_CHECK("act_bwd");
    return B200NN_OK;
}

int b200nn_scale_apply(const float* in, const double* ss, uint64_t chain, float* out, long long n, void* stream) {
    if (n <= 0) return B200NN_OK;
    B200NN_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                   "scale_apply: buffers must be 16-byte aligned");
    act_bwd_kernel<true><<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(B200NN_ACT_NONE, 0.f, nullptr, in, ss, chain, out, nullptr, n);
    B200NN_LAUNCH_CHECK("scale_apply");
    return B200NN_OK;
}

int b200nn_sumsq(const float* x, long long n, double* ss_out, void* stream) {
    if (n <= 0) return B200NN_OK;
    B200NN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sumsq: buffer must be 16-byte aligned");
    act_bwd_kernel<false><<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(B200NN_ACT_NONE, 0.f, nullptr, x, nullptr, 0, nullptr, ss_out, n);
    B200NN_LAUNCH_CHECK("sumsq");
    return B200NN_OK;
}

int b200nn_softmax_fwd(const float* x, float* p, int rows, int cols, void* stream) {
    B200NN_REQUIRE(rows >= 0 && cols > 0, "softmax: bad shape %d x %d", rows, cols);
    if (rows == 0) return B200NN_OK;
    const int blocks = (rows * 32 + 255) / 256;
    softmax_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, p, rows, cols);
    B200NN_LAUNCH_CHECK("softmax");
    return B200NN_OK;
}

int b200nn_loss_fwd_bwd_ex(int loss, int fused_pair, const float* p, const float* y, float* err, double* loss_acc,
                           double* ss_out, int rows, int cols, long long global_rows, float param, void* stream) {
    B200NN_REQUIRE(loss >= 0 && loss <= B200NN_LOSS_HUBER, "loss: unknown kind %d", loss);
    B200NN_REQUIRE(rows > 0 && cols > 0 && global_rows >= rows, "loss: bad shape %d x %d (global rows %lld)", rows, cols, global_rows);
    B200NN_REQUIRE(loss != B200NN_LOSS_HUBER || param > 0.0f, "loss: Huber needs delta > 0");
    long long threads_needed = cols == 1 ? rows : (long long)rows * 32;
    int blocks = (int)((threads_needed + 255) / 256);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    loss_kernel<<<blocks, 256, 0, as_stream(stream)>>>(loss, fused_pair, p, y, err, loss_acc, ss_out, rows, cols,
                                                       (double)global_rows * (double)cols, param);
    B200NN_LAUNCH_CHECK("loss");
    return B200NN_OK;
}

int b200nn_loss_fwd_bwd(int loss, int fused_pair, const float* p, const float* y, float* err, double* loss_acc,
                        double* ss_out, int rows, int cols, void* stream) {
    return b200nn_loss_fwd_bwd_ex(loss, fused_pair, p, y, err, loss_acc, ss_out, rows, cols, rows, 1.0f, stream);
}

int b200nn_adam_multi(const b200nn_param* params, int n_params, int n_layers, double* gss, const double* hyper,
                      const long long* step, const int* block_map, int total_blocks, const double* ss, void* stream) {
    B200NN_REQUIRE(n_params > 0 && total_blocks > 0 && n_layers > 0, "adam_multi: empty parameter list");
    if (total_blocks <= 16) {   // tiny sets only (CTAs spinning at the barrier would steal SM slots from concurrent kernels); gss + counters are zero on entry, restored by the kernel
        adam_small_fused_kernel<<<total_blocks, 256, 0, as_stream(stream)>>>(params, n_layers, n_params, gss, hyper, step, block_map, ss);
        B200NN_LAUNCH_CHECK("adam_small_fused");
        return B200NN_OK;
    }
    B200NN_CUDA(cudaMemsetAsync(gss, 0, sizeof(double) * n_params, as_stream(stream)));
    sumsq_multi_kernel<<<total_blocks, 256, 0, as_stream(stream)>>>(params, block_map, gss);
    B200NN_LAUNCH_CHECK("sumsq_multi");
    adam_multi_kernel<false><<<total_blocks, 256, 0, as_stream(stream)>>>(params, n_layers, n_params, gss, hyper, step, block_map, ss);
    B200NN_LAUNCH_CHECK("adam_multi");
    return B200NN_OK;
}

int b200nn_opt_multi(int kind, const b200nn_param* params, int n_params, int n_layers, double* gss, const double* hyper,
                     const long long* step, const int* block_map, int total_blocks, const double* ss, void* stream) {
    B200NN_REQUIRE(kind >= B200NN_OPT_ADAM && kind <= B200NN_OPT_RADAM, "opt_multi: unknown update rule %d", kind);
    B200NN_REQUIRE(n_params > 0 && total_blocks > 0 && n_layers > 0, "opt_multi: empty parameter list");
    B200NN_REQUIRE(step != nullptr || kind == B200NN_OPT_MOMENTUM || kind == B200NN_OPT_RMSPROP, "opt_multi: this rule needs the t counter");
    cudaStream_t st = as_stream(stream);
    B200NN_CUDA(cudaMemsetAsync(gss, 0, sizeof(double) * n_params, st));
    sumsq_multi_kernel<<<total_blocks, 256, 0, st>>>(params, block_map, gss);
    B200NN_LAUNCH_CHECK("sumsq_multi");
    switch (kind) {
        case B200NN_OPT_ADAM: opt_multi_kernel<B200NN_OPT_ADAM><<<total_blocks, 256, 0, st>>>(params, n_layers, gss, hyper, step, block_map, ss); break;
        case B200NN_OPT_MOMENTUM: opt_multi_kernel<B200NN_OPT_MOMENTUM><<<total_blocks, 256, 0, st>>>(params, n_layers, gss, hyper, step, block_map, ss); break;
        case B200NN_OPT_RMSPROP: opt_multi_kernel<B200NN_OPT_RMSPROP><<<total_blocks, 256, 0, st>>>(params, n_layers, gss, hyper, step, block_map, ss); break;
        case B200NN_OPT_ADABELIEF: opt_multi_kernel<B200NN_OPT_ADABELIEF><<<total_blocks, 256, 0, st>>>(params, n_layers, gss, hyper, step, block_map, ss); break;
        default: opt_multi_kernel<B200NN_OPT_RADAM><<<total_blocks, 256, 0, st>>>(params, n_layers, gss, hyper, step, block_map, ss); break;
    }
    B200NN_LAUNCH_CHECK("opt_multi");
    return B200NN_OK;
}

int b200nn_sgd_multi(const b200nn_param* params, int n_params, double* gss, const double* hyper,
                     const int* block_map, int total_blocks, const double* ss, void* stream) {
    B200NN_REQUIRE(n_params > 0 && total_blocks > 0, "sgd_multi: empty parameter list");
    B200NN_CUDA(cudaMemsetAsync(gss, 0, sizeof(double) * n_params, as_stream(stream)));
    sumsq_multi_kernel<<<total_blocks, 256, 0, as_stream(stream)>>>(params, block_map, gss);
    B200NN_LAUNCH_CHECK("sumsq_multi");
    sgd_multi_kernel<<<total_blocks, 256, 0, as_stream(stream)>>>(params, gss, hyper, block_map, ss);
    B200NN_LAUNCH_CHECK("sgd_multi");
    return B200NN_OK;
}

}  // extern "C"
