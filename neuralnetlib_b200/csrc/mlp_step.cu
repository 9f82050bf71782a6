// mlp_step.cu — the whole train step of a small MLP (config 1: 784-128-64-10, batch 32, 14.5 MFLOP) as ONE persistent
// cooperative kernel. Layer by layer the step is 12-16 graph nodes of ~10 us dependency latency each for a few hundred
// nanoseconds of arithmetic; here the phases
//     [zero slots, t += L] -> Dense+act forward x (L-1) -> head (Dense + softmax + CCE + p - y) ->
//     per layer, last to first: { dW, db (+ their squared norms), dx + activation backward (+ the two clip slots) } -> Adam
// are separated by grid-wide barriers (one atomic counter, every CTA resident: the launch is cooperative) instead of kernel
// boundaries. Semantics are exactly those of the layer kernels (models.py:159-264): errors are stored already multiplied by the
// clip chain of their producer's input, every clip is a float64 sum-of-squares slot read through chain_scale, the per-tensor
// gradient clip and Adam's per-layer t (optimizers.py:236) come from the same b200nn_param table as b200nn_adam_multi.
//
// Work decomposition (256 threads = 8 warps per CTA, items dealt round-robin to CTAs):
//   forward : item = 8 rows x 32 columns; lane <-> column (coalesced weight rows), warps split k, the input rows sit in shared
//             memory (128-bit broadcast reads), partial sums folded through shared memory; bias + activation in the fold.
//   dW      : item = 32 k x 32 n; lane <-> n, warp <-> 4 k; a[m][k] and e[m][n] tiles staged once, loop over the batch.
//   dx      : item = 8 rows x 32 k; lane <-> k; W rows staged TRANSPOSED-readable (pitch + 1), warps split n, fold, then
//             clip multiplier, ||dx||^2, activation backward, ||masked dx||^2.
#include "common.cuh"
#include "opt_common.cuh"

namespace b200nn {
namespace mlp {

constexpr int THREADS = 256, WARPS = 8, RB = 8, CB = 32;
constexpr int FWD_KC = 512;                    // input features staged per pass (forward)
constexpr int DX_NC = 128;                     // output features staged per pass (dx)
constexpr int MAXB = B200NN_MLP_MAX_BATCH;
constexpr int HEAD_W_FLOATS = 8192;            // the head's weights are staged in shared memory when K * N fits
constexpr int HEAD_K_MAX = 1024;               // ... and its input rows when K does
// shared-memory layouts (floats): forward As[RB][FWD_KC] | Ws[FWD_KC][CB] | red[WARPS][RB][CB];  dx Ws[32][DX_NC + 1] | Es[RB][DX_NC] | red;
// dW At[MAXB][32] | Es[MAXB][32];  head Wh[HEAD_W_FLOATS] | xs[WARPS][HEAD_K_MAX]
constexpr int FWD_FLOATS = RB * FWD_KC + FWD_KC * CB + WARPS * RB * CB;
constexpr int HEAD_FLOATS = HEAD_W_FLOATS + WARPS * HEAD_K_MAX;
constexpr int SMEM_FLOATS = FWD_FLOATS > HEAD_FLOATS ? FWD_FLOATS : HEAD_FLOATS;
constexpr size_t SMEM_BYTES = (size_t)SMEM_FLOATS * sizeof(float);
constexpr unsigned long long BARRIER_TIMEOUT_NS = 2000000000ull;   // a barrier that never completes must not hang the GPU

struct Shared {
    float* f;              // dynamic shared memory, SMEM_FLOATS floats, 16-byte aligned
    double* scratch;       // 32 doubles
    OptScalars* opt;
};

__device__ __forceinline__ void cp16_zfill(void* smem_dst, const void* gmem_src, bool ok) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(ok ? 16 : 0));
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// global -> shared staging with every load of a batch in flight before the first store (latency, not bandwidth, bounds this kernel)
template <typename LoadF, typename StoreF>
__device__ __forceinline__ void stage_batched(int total, LoadF ld, StoreF st) {
    constexpr int PER = 8;
    for (int base = threadIdx.x; base < total; base += THREADS * PER) {
        float v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = base + j * THREADS;
            v[j] = i < total ? ld(i) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = base + j * THREADS;
            if (i < total) st(i, v[j]);
        }
    }
}

__device__ __forceinline__ double ldcg_d(const double* p) { return __ldcg(p); }

// chain_scale (common.cuh) over slots produced INSIDE this kernel by other CTAs: L2 loads
__device__ __forceinline__ float chain_scale_cg(const double* ss, uint64_t chain) {
    const int n = static_cast<int>(chain & 0xff);
    double acc = 1.0;
    for (int i = 0; i < n; ++i) {
        const int slot = static_cast<int>((chain >> (8 + 12 * i)) & 0xfff);
        const double n2 = acc * acc * ldcg_d(ss + slot);
        if (n2 > 1.0) acc *= rsqrt(n2);
    }
    return static_cast<float>(acc);
}

// phase time stamps for scripts/mlp_step_phases.py: CTA 0 records %globaltimer when it enters and leaves every barrier
// (stamps 2e + 1, 2e + 2, 8 bytes each, behind the two barrier words)
__device__ __forceinline__ void stamp(unsigned* sync, int idx) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && idx < 15) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        reinterpret_cast<unsigned long long*>(sync + 2)[idx] = now;
    }
}

__device__ __forceinline__ void grid_barrier(unsigned* sync, unsigned& epoch) {
    __syncthreads();
    stamp(sync, 2 * (int)epoch + 1);
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&sync[0], 1u);
        const unsigned target = (epoch + 1u) * gridDim.x;
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (*reinterpret_cast<volatile unsigned*>(&sync[0]) < target) {
            if ((++spins & 0xfffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > BARRIER_TIMEOUT_NS) break;
            }
        }
        __threadfence();
    }
    ++epoch;
    __syncthreads();
    stamp(sync, 2 * (int)epoch);
}

// out[r0.., n0..] = act(in . w + b) for one 8 x 32 tile
__device__ void fwd_item(const b200nn_mlp_step& a, int l, int item, Shared& sh) {
    const b200nn_mlp_layer& Ly = a.layer[l];
    const int K = Ly.K, N = Ly.N, B = a.B;
    const int ncb = (N + CB - 1) / CB;
    const int r0 = (item / ncb) * RB, n0 = (item % ncb) * CB;
    const float* in = l == 0 ? a.x : a.layer[l - 1].out;
    float* As = sh.f;                                  // [RB][FWD_KC]
    float* Ws = sh.f + RB * FWD_KC;                    // [FWD_KC][CB]
    float* red = Ws + FWD_KC * CB;                     // [WARPS][RB][CB]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = n0 + lane;
    const bool vec = (N & 3) == 0 && (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(Ly.w) | reinterpret_cast<uintptr_t>(in)) & 15) == 0;
    float acc[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = 0.0f;
    for (int kc0 = 0; kc0 < K; kc0 += FWD_KC) {
        const int kc = min(FWD_KC, K - kc0);
        const int kc4 = (kc + 3) & ~3;
        __syncthreads();
        if (vec) {      // 16-byte cp.async (L2 path: coherent with what other CTAs wrote before the last barrier), zero-filled tails
            for (int i = threadIdx.x; i < RB * (kc4 >> 2); i += THREADS) {
                const int r = i / (kc4 >> 2), q = i - r * (kc4 >> 2), m = r0 + r;
                const bool ok = m < B;                                   // K % 4 == 0: a quad is entirely inside or outside the row
                cp16_zfill(&As[r * FWD_KC + q * 4], ok ? in + (long long)m * K + kc0 + q * 4 : in, ok);
            }
            for (int i = threadIdx.x; i < kc * (CB / 4); i += THREADS) {
                const int k = i >> 3, q = i & 7;
                const bool ok = n0 + q * 4 < N;
                cp16_zfill(&Ws[k * CB + q * 4], ok ? Ly.w + (long long)(kc0 + k) * N + n0 + q * 4 : Ly.w, ok);
            }
            cp_wait_all();
        } else {
            stage_batched(RB * kc4,
                          [&](int i) { const int r = i / kc4, k = i - r * kc4, m = r0 + r;
                                       return (m < B && k < kc) ? __ldcg(in + (long long)m * K + kc0 + k) : 0.0f; },
                          [&](int i, float v) { const int r = i / kc4, k = i - r * kc4; As[r * FWD_KC + k] = v; });
            stage_batched(kc * CB,
                          [&](int i) { const int k = i >> 5, c = i & 31;
                                       return n0 + c < N ? __ldcg(Ly.w + (long long)(kc0 + k) * N + n0 + c) : 0.0f; },
                          [&](int i, float v) { Ws[i] = v; });
        }
        for (int i = kc * CB + threadIdx.x; i < kc4 * CB; i += THREADS) Ws[i] = 0.0f;     // k tail of the last quad
        __syncthreads();
        const int per = (kc4 / 4 + WARPS - 1) / WARPS * 4;          // k per warp, a multiple of 4
        const int kb = w * per, ke = min(kc4, kb + per);
#pragma unroll 2
        for (int k = kb; k < ke; k += 4) {
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = Ws[(k + j) * CB + lane];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const float4 av = *reinterpret_cast<const float4*>(&As[r * FWD_KC + k]);
                acc[r] = fmaf(av.x, wv[0], acc[r]); acc[r] = fmaf(av.y, wv[1], acc[r]);
                acc[r] = fmaf(av.z, wv[2], acc[r]); acc[r] = fmaf(av.w, wv[3], acc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) red[(w * RB + r) * CB + lane] = acc[r];
    __syncthreads();
    {
        const int r = threadIdx.x >> 5, m = r0 + r;
        float s = 0.0f;
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) s += red[(ww * RB + r) * CB + lane];
        if (m < B && n < N) {
            s += Ly.b != nullptr ? __ldcg(Ly.b + n) : 0.0f;
            Ly.out[(long long)m * N + n] = act_apply(Ly.act, s, Ly.alpha);
        }
    }
}

// classifier head: one warp per row (layers.py:173, activations.py:69-71, losses.py:106-109, models.py:200-205, metrics.py:84-89)
__device__ void head_rows(const b200nn_mlp_step& a, Shared& sh) {
    const b200nn_mlp_layer& H = a.layer[a.L - 1];
    const int K = H.K, N = H.N, B = a.B;
    const float* in = a.layer[a.L - 2].out;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gw = blockIdx.x * WARPS + w, nw = gridDim.x * WARPS;
    if ((int)blockIdx.x * WARPS >= B) return;            // no row for this CTA (uniform per CTA: no barrier inside is skipped)
    const bool staged = K * N <= HEAD_W_FLOATS && K <= HEAD_K_MAX;
    float* Wh = sh.f;                                    // [K][N]
    float* xs = sh.f + HEAD_W_FLOATS + w * HEAD_K_MAX;   // this warp's input row
    if (staged) {
        __syncthreads();
        stage_batched(K * N, [&](int i) { return __ldcg(H.w + i); }, [&](int i, float v) { Wh[i] = v; });
        __syncthreads();
    }
    for (int m = gw; m < B; m += nw) {
        float logit = 0.0f;
        if (staged) {
            __syncwarp();
            for (int k0 = lane; k0 < K; k0 += 32 * 8) {          // 8 loads in flight per lane
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = k0 + 32 * j < K ? __ldcg(in + (long long)m * K + k0 + 32 * j) : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) if (k0 + 32 * j < K) xs[k0 + 32 * j] = v[j];
            }
            __syncwarp();
            // lane <-> (class n, k part): every lane runs one serial dot product over its share of k, the parts are folded with
            // log2(parts) shuffles (instead of N warp-wide reductions of 5 dependent shuffles each)
            const int NP = N <= 8 ? 8 : (N <= 16 ? 16 : 32), parts = 32 / NP;
            const int n = lane % NP, part = lane / NP;
            float t = 0.0f;
            if (n < N)
                for (int k = part; k < K; k += parts) t = fmaf(xs[k], Wh[k * N + n], t);
            for (int o = NP; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            logit = t;                                   // lanes < N hold the logit of class `lane` (n == lane there)
        } else {
            for (int n = 0; n < N; ++n) {
                float t = 0.0f;
                for (int k = lane; k < K; k += 32) t = fmaf(__ldcg(in + (long long)m * K + k), __ldcg(H.w + (long long)k * N + n), t);
                t = warp_sum(t);
                if (lane == n) logit = t;
            }
        }
        const bool on = lane < N;
        logit = on ? logit + (H.b != nullptr ? __ldcg(H.b + lane) : 0.0f) : -INFINITY;
        const float mx = warp_max(logit);
        const float ex = on ? expf(logit - mx) : 0.0f;
        const float pv = ex / warp_sum(ex);
        const float yv = on ? __ldg(a.y + (long long)m * N + lane) : 0.0f;
        float lsum = 0.0f, esq = 0.0f;
        if (on) {
            H.out[(long long)m * N + lane] = pv;
            const float pc = fminf(fmaxf(pv, 1e-15f), 1.0f);
            lsum = -yv * logf(pc);
            const float e = pv - yv;
            H.err[(long long)m * N + lane] = e;
            esq = e * e;
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
        const float l = warp_sum(lsum), q = warp_sum(esq);      // <= 32 terms each
        if (lane == 0) {
            atomicAdd(&a.acc[0], (double)l);
            if (ip == iy) atomicAdd(&a.acc[1], 1.0);
            atomicAdd(&a.ss[a.s0], (double)q);
        }
    }
}

// sum of per-thread fp32 partials -> one atomicAdd(double) per CTA: fp32 shuffles inside a warp (<= 32 terms), fp64 across warps
__device__ __forceinline__ void block_accumulate_fast(float v, double* out, double* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = (double)v;
    __syncthreads();
    if (threadIdx.x == 0 && out != nullptr) {
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < WARPS; ++i) r += scratch[i];
        atomicAdd(out, r);
    }
}

// dw[k0.., n0..] = scale * a_prev^T . err for one 32 x 32 tile (layers.py:196), + its share of ||dw||^2
__device__ void dw_item(const b200nn_mlp_step& a, int l, int item, float scale, Shared& sh) {
    const b200nn_mlp_layer& Ly = a.layer[l];
    const int K = Ly.K, N = Ly.N, B = a.B;
    const int ncb = (N + CB - 1) / CB;
    const int k0 = (item / ncb) * 32, n0 = (item % ncb) * CB;
    const float* in = l == 0 ? a.x : a.layer[l - 1].out;
    float* At = sh.f;                 // [MAXB][32]
    float* Es = sh.f + MAXB * 32;     // [MAXB][32]
    __syncthreads();
    stage_batched(B * 32,
                  [&](int i) { const int m = i >> 5, j = i & 31;
                               return k0 + j < K ? __ldcg(in + (long long)m * K + k0 + j) : 0.0f; },
                  [&](int i, float v) { At[i] = v; });
    stage_batched(B * 32,
                  [&](int i) { const int m = i >> 5, j = i & 31;
                               return n0 + j < N ? __ldcg(Ly.err + (long long)m * N + n0 + j) : 0.0f; },
                  [&](int i, float v) { Es[i] = v; });
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
    for (int m = 0; m < B; ++m) {
        const float ev = Es[m * 32 + lane];
        const float4 av = *reinterpret_cast<const float4*>(&At[m * 32 + w * 4]);
        acc[0] = fmaf(av.x, ev, acc[0]); acc[1] = fmaf(av.y, ev, acc[1]);
        acc[2] = fmaf(av.z, ev, acc[2]); acc[3] = fmaf(av.w, ev, acc[3]);
    }
    float g2 = 0.0f;
    const int n = n0 + lane;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int k = k0 + w * 4 + r;
        if (k < K && n < N) {
            const float v = acc[r] * scale;
            Ly.dw[(long long)k * N + n] = v;
            g2 = fmaf(v, v, g2);
        }
    }
    block_accumulate_fast(g2, &a.gss[Ly.gi_w], sh.scratch);
}

// db[n] = scale * sum_m err[m][n] (layers.py:197) for 256 columns
__device__ void db_item(const b200nn_mlp_step& a, int l, int item, float scale, Shared& sh) {
    const b200nn_mlp_layer& Ly = a.layer[l];
    const int n = item * THREADS + threadIdx.x;
    float g2 = 0.0f;
    if (n < Ly.N) {
        float s = 0.0f;
#pragma unroll 8
        for (int m = 0; m < a.B; ++m) s += __ldcg(Ly.err + (long long)m * Ly.N + n);
        const float v = s * scale;
        Ly.db[n] = v;
        g2 = v * v;
    }
    block_accumulate_fast(g2, &a.gss[Ly.gi_b], sh.scratch);
}

// err_{l-1}[r0.., k0..] = act'(out_{l-1}) * scale * err_l . w_l^T for one 8 x 32 tile (layers.py:189 + the Activation backward), both slots
__device__ void dx_item(const b200nn_mlp_step& a, int l, int item, float scale, Shared& sh) {
    const b200nn_mlp_layer& Ly = a.layer[l];
    const b200nn_mlp_layer& Pv = a.layer[l - 1];
    const int K = Ly.K, N = Ly.N, B = a.B;
    const int nkb = (K + 31) / 32;
    const int r0 = (item / nkb) * RB, k0 = (item % nkb) * 32;
    constexpr int WP = DX_NC + 1;
    float* Ws = sh.f;                         // [32 k][DX_NC + 1]
    float* Es = sh.f + 32 * WP;               // [RB][DX_NC]
    float* red = Es + RB * DX_NC;             // [WARPS][RB][32]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = 0.0f;
    for (int nc0 = 0; nc0 < N; nc0 += DX_NC) {
        const int nc = min(DX_NC, N - nc0);
        __syncthreads();
        stage_batched(32 * nc,
                      [&](int i) { const int j = i / nc, n = i - j * nc;
                                   return k0 + j < K ? __ldcg(Ly.w + (long long)(k0 + j) * N + nc0 + n) : 0.0f; },
                      [&](int i, float v) { const int j = i / nc, n = i - j * nc; Ws[j * WP + n] = v; });
        stage_batched(RB * nc,
                      [&](int i) { const int r = i / nc, n = i - r * nc, m = r0 + r;
                                   return m < B ? __ldcg(Ly.err + (long long)m * N + nc0 + n) : 0.0f; },
                      [&](int i, float v) { const int r = i / nc, n = i - r * nc; Es[r * DX_NC + n] = v; });
        __syncthreads();
        const int per = (nc + WARPS - 1) / WARPS;
        const int nb = w * per, ne = min(nc, nb + per);
#pragma unroll 4
        for (int n = nb; n < ne; ++n) {
            const float wv = Ws[lane * WP + n];
#pragma unroll
            for (int r = 0; r < RB; ++r) acc[r] = fmaf(Es[r * DX_NC + n], wv, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) red[(w * RB + r) * 32 + lane] = acc[r];
    __syncthreads();
    float a0 = 0.0f, a1 = 0.0f;
    {
        const int r = threadIdx.x >> 5, m = r0 + r, k = k0 + lane;
        float s = 0.0f;
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) s += red[(ww * RB + r) * 32 + lane];
        if (m < B && k < K) {
            float v = s * scale;
            a0 = v * v;
            if (Pv.act != B200NN_ACT_NONE) {
                v *= act_grad_from_y(Pv.act, __ldcg(Pv.out + (long long)m * K + k), Pv.alpha);
                a1 = v * v;
            }
            Pv.err[(long long)m * K + k] = v;
        }
    }
    block_accumulate_fast(a0, Ly.s_out0 >= 0 ? &a.ss[Ly.s_out0] : nullptr, sh.scratch);
    if (Ly.s_out1 >= 0) block_accumulate_fast(a1, &a.ss[Ly.s_out1], sh.scratch);
}

__global__ void __launch_bounds__(THREADS, 1) mlp_train_step_kernel(const __grid_constant__ b200nn_mlp_step a) {
    extern __shared__ __align__(16) float mlp_smem[];
    __shared__ double scratch[32];
    __shared__ OptScalars opt_scalars;
    Shared sh{mlp_smem, scratch, &opt_scalars};
    unsigned epoch = 0;
    const int G = gridDim.x;
    stamp(a.sync, 0);
    // ---- step_begin: slots, accumulators, Adam's t (visible to everyone after the first barrier; nothing reads them before)
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < a.n_slots; i += THREADS) a.ss[i] = 0.0;
        if (threadIdx.x < 4) a.acc[threadIdx.x] = 0.0;
        if (threadIdx.x == 0 && a.t_counter != nullptr) *a.t_counter += a.n_layers_opt;
    }
    // ---- warm L2: parameters and Adam moments (the only large operands: 28 B per parameter), one 128-byte line per thread and
    // tensor; the step's dependent phases then wait on L2 hits instead of HBM round trips
    {
        // warp <-> tensor (round robin), so that a thread reads ONE descriptor instead of walking the table serially
        const int gw = blockIdx.x * WARPS + (threadIdx.x >> 5), nw = G * WARPS, lane = threadIdx.x & 31;
        if (gw < nw - nw % a.n_params || nw < a.n_params) {
            const int ti = gw % a.n_params;
            const b200nn_param P = a.params[ti];
            const long long lines = (P.n * 4 + 127) >> 7;
            const int per = nw >= a.n_params ? nw / a.n_params : 1;          // warps that share this tensor
            for (long long i = (long long)(gw / a.n_params) * 32 + lane; i < lines; i += (long long)per * 32) {
                prefetch_l2(reinterpret_cast<const char*>(P.p) + (i << 7));
                prefetch_l2(reinterpret_cast<const char*>(P.m) + (i << 7));
                prefetch_l2(reinterpret_cast<const char*>(P.v) + (i << 7));
            }
        }
    }
    // ---- forward: hidden layers
    for (int l = 0; l + 1 < a.L; ++l) {
        const b200nn_mlp_layer& Ly = a.layer[l];
        const int items = ((a.B + RB - 1) / RB) * ((Ly.N + CB - 1) / CB);
        for (int it = blockIdx.x; it < items; it += G) fwd_item(a, l, it, sh);
        grid_barrier(a.sync, epoch);
    }
    // ---- head + loss
    head_rows(a, sh);
    grid_barrier(a.sync, epoch);
    // ---- backward, last layer first: dW / db / dx items of a layer are independent and share one phase
    for (int l = a.L - 1; l >= 0; --l) {
        const b200nn_mlp_layer& Ly = a.layer[l];
        const float scale = chain_scale_cg(a.ss, Ly.chain);
        const int n_dw = ((Ly.K + 31) / 32) * ((Ly.N + CB - 1) / CB);
        const int n_db = (Ly.N + THREADS - 1) / THREADS;
        const int n_dx = l > 0 ? ((a.B + RB - 1) / RB) * ((Ly.K + 31) / 32) : 0;
        // the dx items gate the next phase: deal them first
        for (int it = blockIdx.x; it < n_dx + n_dw + n_db; it += G) {
            if (it < n_dx) dx_item(a, l, it, scale, sh);
            else if (it < n_dx + n_dw) dw_item(a, l, it - n_dx, scale, sh);
            else db_item(a, l, it - n_dx - n_dw, scale, sh);
        }
        grid_barrier(a.sync, epoch);
    }
    // ---- per-tensor gradient clip (models.py:240-242) + Adam (optimizers.py:204-245), one 4096-element chunk per item
    for (int blk = blockIdx.x; blk < a.total_blocks; blk += G) {
        const int ti = a.block_map[2 * blk], chunk = a.block_map[2 * blk + 1];
        const b200nn_param P = a.params[ti];
        constexpr int IT = B200NN_OPT_CHUNK / THREADS;
        const long long base = (long long)chunk * B200NN_OPT_CHUNK, end = min(P.n, base + B200NN_OPT_CHUNK);
        float p[IT], g[IT], m[IT], v[IT];
#pragma unroll
        for (int it = 0; it < IT; ++it) {                      // every load of the chunk in flight before the scalars are known
            const long long i = base + it * THREADS + threadIdx.x;
            p[it] = g[it] = m[it] = v[it] = 0.0f;
            if (i < end) { p[it] = P.p[i]; g[it] = __ldcg(P.g + i); m[it] = P.m[i]; v[it] = P.v[i]; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double lr = a.hyper[0], b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3];
            const double pre = (double)chain_scale_cg(a.ss, P.chain);
            const double n2 = ldcg_d(a.gss + ti) * pre * pre;
            const long long t = __ldcg(a.t_counter) - (long long)a.n_layers_opt + P.t_index;   // optimizers.py:236 (Q3)
            OptScalars& s = *sh.opt;
            s.clip = (float)(pre * ((P.clip && n2 > 1.0) ? rsqrt(n2) : 1.0));
            s.lr = (float)lr; s.b1 = (float)b1; s.b2 = (float)b2;
            s.omb1 = (float)(1.0 - b1); s.omb2 = (float)(1.0 - b2);
            s.inv_bc1 = (float)(1.0 / (1.0 - powi(b1, t)));
            s.inv_bc2 = (float)(1.0 / (1.0 - powi(b2, t)));
            s.eps = (float)eps;
        }
        __syncthreads();
        const OptScalars s = *sh.opt;
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const long long i = base + it * THREADS + threadIdx.x;
            if (i < end) {
                adam_elem(p[it], g[it], m[it], v[it], s);
                P.p[i] = p[it]; P.m[i] = m[it]; P.v[i] = v[it];
            }
        }
    }
    // ---- last CTA out re-arms the barrier and the norm slots for the next launch
    __syncthreads();
    stamp(a.sync, 2 * (int)epoch + 1);
    if (threadIdx.x == 0 && atomicAdd(&a.sync[1], 1u) == (unsigned)G - 1u) {
        for (int i = 0; i < a.n_params; ++i) a.gss[i] = 0.0;
        a.sync[0] = 0u;
        a.sync[1] = 0u;
    }
}

}  // namespace mlp
}  // namespace b200nn

using namespace b200nn;

extern "C" int b200nn_mlp_train_step(const b200nn_mlp_step* step, void* stream) {
    B200NN_REQUIRE(step != nullptr, "mlp_train_step: null descriptor");
    const b200nn_mlp_step& a = *step;
    B200NN_REQUIRE(a.L >= 2 && a.L <= B200NN_MLP_MAX_LAYERS && a.B >= 1 && a.B <= B200NN_MLP_MAX_BATCH,
                   "mlp_train_step: unsupported shape (L=%d, batch=%d)", a.L, a.B);
    B200NN_REQUIRE(a.layer[a.L - 1].N <= 32, "mlp_train_step: the classifier head needs N <= 32");
    B200NN_REQUIRE(a.x && a.y && a.ss && a.acc && a.gss && a.hyper && a.t_counter && a.params && a.block_map && a.sync,
                   "mlp_train_step: null buffer");
    // grid: as many CTAs as the widest phase has items, at most one per SM (the barrier needs every CTA resident)
    int items = a.total_blocks;
    for (int l = 0; l < a.L; ++l) {
        const b200nn_mlp_layer& Ly = a.layer[l];
        const int rb = (a.B + mlp::RB - 1) / mlp::RB, nb = (Ly.N + 31) / 32, kb = (Ly.K + 31) / 32;
        const int f = rb * nb, b = kb * nb + (Ly.N + 255) / 256 + (l > 0 ? rb * kb : 0);
        items = items > f ? items : f;
        items = items > b ? items : b;
    }
    int grid = items < kNumSMs ? items : kNumSMs;
    if (grid < 1) grid = 1;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        B200NN_CUDA(cudaFuncSetAttribute(mlp::mlp_train_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp::SMEM_BYTES));
        configured[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(mlp::THREADS);
    cfg.dynamicSmemBytes = mlp::SMEM_BYTES;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    B200NN_CUDA(cudaLaunchKernelEx(&cfg, mlp::mlp_train_step_kernel, a));
    B200NN_LAUNCH_CHECK("mlp_train_step");
    return B200NN_OK;
}
