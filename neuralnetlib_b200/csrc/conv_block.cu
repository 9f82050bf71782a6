// conv_block.cu — the first block of a CNN as TWO kernels:  Conv2D(3x3, stride 1, 'same', C in {1..3}) + Activation ->
// BatchNormalization(train, per-position statistics) -> MaxPooling2D(2x2 / 2)     (BASELINE configs 2 and 3, block 1)
//
// With K = 9*C <= 27 taps the convolution is ~5 FLOP per output byte: the block is bound by the HBM traffic of the conv
// OUTPUT S (B*H*W*F floats: written by the conv, read by BatchNorm, read again by the backward, and the error w.r.t. it
// written and read once more -> 5 passes over the largest tensor of the model). These kernels never materialise S or its
// error: the conv is RECOMPUTED from the input image wherever it is needed (36*C FMAs per sample and filter, from a 4x4xC
// input patch held in shared memory), so the block's HBM traffic drops from ~5*S to x + the pooled tensors:
//
//   forward : x -> [conv + act + clip +-10 -> batch mean (pass 1) -> exact two-pass variance (pass 2) -> normalise, 2x2 max
//             (pass 3)] -> pooled output, running statistics, save_mean / save_invstd
//   backward: x, pooled error -> [recompute -> pool-bwd (all tied maxima, layers.py:631-637) -> the two BatchNorm batch
//             reductions (pass 1) -> BN-bwd, act', conv weight / bias gradient (pass 2)] -> per-window dW partials,
//             d_gamma / d_beta, the three clip slots the reference applies between those layers (models.py:214)
//
// Work decomposition: CTA = (pool window, 8 filters); warp = one filter; lane = batch sample (mod 32). Statistics over the batch
// are warp-shuffle reductions, no shared-memory or cross-CTA reduction exists, all conv outputs of the window stay in registers (batch <= 256 per GPU;
// larger batches use the unfused kernels), and the 784 CTAs of config 2 balance over 148 SMs. Replaces layers.py:442-489 (+
// preprocessing.py:34-79), :231-237, :1230-1276, :570-643 and the weight-gradient half of :491-530 for this pattern.
#include <stdlib.h>
#include <string.h>

#include "comm.cuh"

namespace b200nn {

constexpr int CB_FPC = 8;        // filters (= warps) per CTA
constexpr int CB_THREADS = CB_FPC * 32;
constexpr int CB_MAXB = 256;     // batch rows per CTA: sample s = lane + 32 j, j < CB_NS, all conv outputs stay in registers
constexpr int CB_NS = CB_MAXB / 32;

template <int C>
struct CbCfg {
    static constexpr int K = 9 * C;                 // taps, reference order (c, ky, kx)  (preprocessing.py:77, Q1)
    static constexpr int EPS = 16 * C;              // patch elements per sample
    static constexpr int STRIDE = 16 * C + 4;       // floats per staged patch (+4: conflict-free 128-bit rows across lanes)
};

struct CbArgs {
    int B, H, W, F;                 // conv input H x W (even), F filters; conv output = H x W x F, pooled = H/2 x W/2 x F
    const float* x;                 // [B, H, W, C]
    const float* w;                 // [3, 3, C, F] buffer read as (C, 3, 3, F)
    const float* bias;              // [F]
    int act;
    float alpha;
    const float* gamma;             // [H, W, F]
    const float* beta;
    float* save_mean;
    float* save_invstd;
    double n_total;                 // global batch (data parallel) or B
};

// ACT > 0: compile-time activation kind; ACT < 0: run-time `act`
template <int ACT>
__device__ __forceinline__ float cb_act(float v, int act, float alpha) {
    if (ACT == B200NN_ACT_RELU) return fmaxf(v, 0.0f);
    return act_apply(act, v, alpha);
}
// backward keeps act(conv) per sample; for ReLU the +-10 clip is applied once when the value is produced (the clip keeps
// the sign, so ReLU' can be read from the clipped value), for the other activations at every use
template <int ACT>
__device__ __forceinline__ float cb_keep(float a) { return ACT == B200NN_ACT_RELU ? fminf(a, 10.0f) : a; }
template <int ACT>
__device__ __forceinline__ float cb_clipped(float kept) { return ACT == B200NN_ACT_RELU ? kept : clip10(kept); }
template <int ACT>
__device__ __forceinline__ float cb_act_grad(float y, int act, float alpha) {
    if (ACT == B200NN_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
    return act_grad_from_y(act, y, alpha);
}

// stage the 4x4xC input patches of all B samples around window (oy, ox): patch[s][c*16 + py*4 + px].
// Loads are issued 8 at a time into registers before any of them is stored, so their latencies overlap.
template <int C>
__device__ __forceinline__ void cb_stage(float* patch, const CbArgs& a, int oy, int ox) {
    constexpr int EPS = CbCfg<C>::EPS, ROW = 4 * C, ITER = CB_MAXB * EPS / CB_THREADS, U = 8;
    const int iy0 = 2 * oy - 1, ix0 = 2 * ox - 1;
    const long long img = (long long)a.H * a.W * C;
    const int total = a.B * EPS;
#pragma unroll
    for (int k0 = 0; k0 < ITER; k0 += U) {
        if (k0 * CB_THREADS >= total) break;
        float v[U];
        int dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = (int)threadIdx.x + (k0 + u) * CB_THREADS;
            const int sidx = i / EPS, e = i % EPS;             // memory order within a sample: py, px, c
            const int py = e / ROW, r = e % ROW, px = r / C, c = r % C;
            const int iy = iy0 + py, ix = ix0 + px;
            dst[u] = i < total ? sidx * CbCfg<C>::STRIDE + c * 16 + py * 4 + px : -1;
            v[u] = 0.0f;                                        // zero padding of 'same' (layers.py:450-465)
            if (i < total && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
                v[u] = a.x[(long long)sidx * img + ((long long)iy * a.W + ix) * C + c];
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (dst[u] >= 0) patch[dst[u]] = v[u];
    }
}

// the same staging with cp.async (4-byte copies, zero-filled when the source pixel is padding): returns immediately, so a
// persistent CTA can fetch its NEXT item while it computes the current one (double-buffered shared memory)
template <int C>
__device__ __forceinline__ void cb_stage_async(float* patch, const CbArgs& a, int oy, int ox) {
    constexpr int EPS = CbCfg<C>::EPS, ROW = 4 * C;
    const int iy0 = 2 * oy - 1, ix0 = 2 * ox - 1;
    const long long img = (long long)a.H * a.W * C;
    const int total = a.B * EPS;
    for (int i = threadIdx.x; i < total; i += CB_THREADS) {
        const int sidx = i / EPS, e = i % EPS;
        const int py = e / ROW, r = e % ROW, px = r / C, c = r % C;
        const int iy = iy0 + py, ix = ix0 + px;
        const bool ok = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
        const float* src = ok ? a.x + (long long)sidx * img + ((long long)iy * a.W + ix) * C + c : a.x;
        const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(patch + sidx * CbCfg<C>::STRIDE + c * 16 + py * 4 + px));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(ok ? 4 : 0));
    }
}
__device__ __forceinline__ void cb_cp16(void* smem_dst, const void* gmem_src) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cb_async_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
template <int C>
struct CbBuf {   // per staging buffer: patch[CB_MAXB][STRIDE] | aux[CB_FPC * CB_MAXB]; two buffers when they fit 3 CTAs per SM
    static constexpr int FLOATS = CB_MAXB * CbCfg<C>::STRIDE + CB_FPC * CB_MAXB;
    static constexpr int N = C == 1 ? 2 : 1;
};

// conv + activation of one sample at the window's 4 positions (q = 2*qy + qx) for one filter
template <int C, int ACT>
__device__ __forceinline__ void cb_conv(const float* __restrict__ p, const float (&wr)[9 * C], float bias, int act, float alpha,
                                        float (&S)[4]) {
    // packed FFMA2 (fma.rn.f32x2, sm_100): the two outputs of a window row share the tap weight, so one instruction does both
    // fused multiply-adds — the same roundings as two scalar FFMAs, half the FMA-pipe slots (scripts/ffma2_probe.cu: 46 -> 66 TFLOP/s)
    float2 S01 = make_float2(bias, bias), S23 = S01;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float rows[4][4];
#pragma unroll
        for (int py = 0; py < 4; ++py) {
            const float4 r = *reinterpret_cast<const float4*>(p + c * 16 + py * 4);
            rows[py][0] = r.x; rows[py][1] = r.y; rows[py][2] = r.z; rows[py][3] = r.w;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float wv = wr[(c * 3 + ky) * 3 + kx];
                const float2 w2 = make_float2(wv, wv);
                S01 = __ffma2_rn(make_float2(rows[ky][kx], rows[ky][kx + 1]), w2, S01);
                S23 = __ffma2_rn(make_float2(rows[ky + 1][kx], rows[ky + 1][kx + 1]), w2, S23);
            }
    }
    S[0] = S01.x; S[1] = S01.y; S[2] = S23.x; S[3] = S23.y;
#pragma unroll
    for (int q = 0; q < 4; ++q) S[q] = cb_act<ACT>(S[q], act, alpha);     // layers.py:231-234 (auto-appended Activation)
}

template <int C>
__device__ __forceinline__ void cb_load_filter(const CbArgs& a, int f, float (&wr)[9 * C], float& bias) {
#pragma unroll
    for (int k = 0; k < 9 * C; ++k) wr[k] = a.w[(long long)k * a.F + f];
    bias = a.bias != nullptr ? a.bias[f] : 0.0f;
}

// ---- peer-fused mode (MODE 3): the batch reductions of BatchNorm completed ACROSS GPUs inside the kernel ------------------
// Every warp contributes 8 floats per item (lane l < 8 sends element l). Low-latency protocol: each element travels as ONE
// 8-byte store {value, round number} straight into every peer's inbox over NVLink (a naturally aligned 8-byte store is a
// single transaction, so the value is valid whenever its round number is), and the receiving lane polls its own word:
// no fence, no separate flag, no block-wide barrier -- warps proceed independently and the other CTAs of the SM compute
// while one waits. Region (b200nn_comm_region_create): word[2 parity][world][items][64] of 8 bytes, double-buffered on the
// parity of the device-resident round counter. The launch is persistent (grid <= co-resident CTAs, items strided), so the
// CTA a rank waits for is always resident on the peer.
__host__ __device__ inline size_t cb_region_bytes(int items, int world) { return (size_t)2 * world * items * 64 * 8; }

__device__ __forceinline__ void cb_peer_gather(const PeerRegion& pr, unsigned seq, int item, int items, float mine,
                                               float (&in)[MAX_WORLD]) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned par = seq & 1u;
    SymHeader* me = reinterpret_cast<SymHeader*>(pr.peers.p[pr.rank]);
    const size_t slot = (size_t)item * 64 + wid * 8 + lane;
    const size_t per_rank = (size_t)items * 64;
#pragma unroll
    for (int r = 0; r < MAX_WORLD; ++r) in[r] = 0.0f;
    if (lane < 8) {
        const unsigned bits = __float_as_uint(mine);
#pragma unroll 1
        for (int p = 0; p < MAX_WORLD; ++p) {
            if (p < pr.world) {
                uint2* dst = reinterpret_cast<uint2*>(pr.peers.p[p] + pr.off) + ((size_t)par * pr.world + pr.rank) * per_rank + slot;
                asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(bits), "r"(seq) : "memory");
            }
        }
        // poll ALL ranks' words together: the loads of one sweep are in flight at once, so a sweep costs one memory round trip
        // instead of one per rank (sequential polling made the wait grow with the world size even when every word had arrived:
        // convblock_bwd_peer 41 -> 60 -> 99 us at 1 / 2 / 8 GPUs)
        const uint2* base = reinterpret_cast<const uint2*>(pr.peers.p[pr.rank] + pr.off) + (size_t)par * pr.world * per_rank + slot;
        unsigned got = 0, spins = 0;
        const unsigned full = (1u << pr.world) - 1u;
        unsigned long long t0 = 0;
        if (!pr.poll_all) {                                  // rank by rank (B200NN_PEER_POLL_ALL=0): the A/B baseline
#pragma unroll 1
            for (int r = 0; r < MAX_WORLD; ++r) {
                if (r < pr.world) {
                    unsigned vx, vy;
                    for (;;) {
                        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(vx), "=r"(vy) : "l"(base + (size_t)r * per_rank) : "memory");
                        if (vy == seq) break;
                        if ((++spins & 0x3ffu) == 0) {
                            unsigned long long now;
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                            if (t0 == 0) t0 = now;
                            else if (now - t0 > WAIT_TIMEOUT_NS) { atomicAdd(&me->timeouts, 1u); break; }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < MAX_WORLD; ++q) if (q == r) in[q] = __uint_as_float(vx);
                }
            }
            got = full;
        }
        for (; got != full;) {
            unsigned vx[MAX_WORLD], vy[MAX_WORLD];
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) {
                vx[r] = 0; vy[r] = 0;
                if (r < pr.world && !((got >> r) & 1u))
                    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(vx[r]), "=r"(vy[r]) : "l"(base + (size_t)r * per_rank) : "memory");
            }
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r)
                if (r < pr.world && !((got >> r) & 1u) && vy[r] == seq) { got |= 1u << r; in[r] = __uint_as_float(vx[r]); }
            if (got == full) break;
            if ((++spins & 0x3ffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > WAIT_TIMEOUT_NS) { atomicAdd(&me->timeouts, 1u); break; }
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void cb_round_done(const PeerRegion& pr, unsigned seq) {
    SymHeader* me = reinterpret_cast<SymHeader*>(pr.peers.p[pr.rank]);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&me->done[pr.site], 1u) == gridDim.x - 1) {
            me->done[pr.site] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned*>(&me->seq[pr.site]) = seq;
        }
    }
}

// MODE 0: single GPU. MODE 1 (data parallel): local moments only -> stats (fp64 [2][P]: sum, sum of squares).
// MODE 2: normalise + pool from rank-summed `stats`. MODE 3: one kernel, moments merged across GPUs over peer memory.
// 1-D grid over items = (pool window, group of 8 filters), strided (persistent in MODE 3).
template <int C, int MODE, int ACT>
__global__ void __launch_bounds__(CB_THREADS, C == 1 ? 3 : 2) convblock_fwd_kernel(CbArgs a, float* __restrict__ rmean, float* __restrict__ rvar,
                                                                   double* __restrict__ stats, float* __restrict__ y_pool,
                                                                   float momentum, float eps, PeerRegion pr, int items) {
    using Cfg = CbCfg<C>;
    extern __shared__ __align__(16) float cb_smem[];                         // CbBuf<C>::N x { patch | outs[CB_FPC][CB_MAXB] }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int OW = a.W >> 1, OH = a.H >> 1;
    const long long P = (long long)a.H * a.W * a.F;
    auto stage = [&](int it, int b) {
        cb_stage_async<C>(cb_smem + b * CbBuf<C>::FLOATS, a, (it / OW) % OH, it % OW);
    };
    int buf = 0;
    if (CbBuf<C>::N == 2 && (int)blockIdx.x < items) stage(blockIdx.x, 0);
    unsigned seq = 0;
    if (MODE == 3) {
        seq = *reinterpret_cast<volatile unsigned*>(&reinterpret_cast<SymHeader*>(pr.peers.p[pr.rank])->seq[pr.site]) + 1u;
        // CTAs that share an SM (block ids one SM count apart: residues 0, 1, 2 mod 3 since 148 = 1 mod 3) start a third of an
        // item period apart, so that one computes while another waits on NVLink; every rank applies the same offsets
        if (pr.stagger_ns > 0) __nanosleep((blockIdx.x % 3) * pr.stagger_ns);
    }
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int ox = item % OW, oy = (item / OW) % OH, f0 = (item / (OW * OH)) * CB_FPC, f = f0 + wid;
    float wr[9 * C], bias;
    cb_load_filter<C>(a, f, wr, bias);
    long long pos[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) pos[q] = ((long long)(2 * oy + (q >> 1)) * a.W + 2 * ox + (q & 1)) * a.F + f;
    float* patch = cb_smem + buf * CbBuf<C>::FLOATS;
    float (*outs)[CB_MAXB] = reinterpret_cast<float (*)[CB_MAXB]>(patch + CB_MAXB * Cfg::STRIDE);
    if (CbBuf<C>::N == 2) {
        cb_async_wait();
        __syncthreads();                                                      // this item's patches have landed; the previous item is done
        if (item + (int)gridDim.x < items) stage(item + gridDim.x, buf ^ 1);   // fetch the next item while computing this one
    } else {
        __syncthreads();                                                      // the previous item is done with the shared tiles
        cb_stage<C>(patch, a, oy, ox);
        __syncthreads();
    }
    float S[CB_NS][4];                                                        // clip(act(conv)) of this lane's samples (layers.py:1234)
#pragma unroll
    for (int j = 0; j < CB_NS; ++j) {
        const int s = lane + 32 * j;
        if (s < a.B) {
            cb_conv<C, ACT>(patch + s * Cfg::STRIDE, wr, bias, a.act, a.alpha, S[j]);
#pragma unroll
            for (int q = 0; q < 4; ++q) S[j][q] = clip10(S[j][q]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) S[j][q] = 0.0f;
        }
    }
    float mean[4], invstd[4], bvar[4];
    if (MODE != 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                         // batch mean (layers.py:1237)
            float t = 0.0f;
#pragma unroll
            for (int j = 0; j < CB_NS; ++j) t += S[j][q];
            mean[q] = warp_sum(t) / (float)a.B;
        }
        float m2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                         // exact two-pass variance
            float t = 0.0f;
#pragma unroll
            for (int j = 0; j < CB_NS; ++j) {
                const float d = S[j][q] - mean[q];
                t = (lane + 32 * j < a.B) ? fmaf(d, d, t) : t;
            }
            m2[q] = warp_sum(t);
        }
        if (MODE == 1) {
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double S1 = (double)mean[q] * (double)a.B;
                    stats[pos[q]] = S1;
                    stats[P + pos[q]] = (double)m2[q] + S1 * S1 / (double)a.B;
                }
            }
            if (CbBuf<C>::N == 2) buf ^= 1;
            continue;
        }
        float ntot = (float)a.B;
        if (MODE == 3) {
            // merge the per-rank (mean, M2) pairs (equal counts) without cancellation: mean_g = avg(mean_r),
            // M2_g = sum_r [ M2_r + B * (mean_r - mean_g)^2 ]; every rank sums in rank order => identical statistics
            float mine = mean[0];                                             // lane l < 4: mean[l]; lane 4 + l: m2[l]
            mine = lane == 1 ? mean[1] : mine; mine = lane == 2 ? mean[2] : mine; mine = lane == 3 ? mean[3] : mine;
            mine = lane == 4 ? m2[0] : mine; mine = lane == 5 ? m2[1] : mine; mine = lane == 6 ? m2[2] : mine; mine = lane == 7 ? m2[3] : mine;
            float mr[MAX_WORLD], qr[MAX_WORLD];
            cb_peer_gather(pr, seq, item, items, mine, mr);
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) qr[r] = __shfl_sync(0xffffffffu, mr[r], (lane + 4) & 31);   // lane q <- M2_r[q]
            float mg = 0.0f;
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) mg += r < pr.world ? mr[r] : 0.0f;
            mg /= (float)pr.world;
            float M2 = 0.0f;
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) {
                const float d = mr[r] - mg;
                M2 += r < pr.world ? fmaf((float)a.B * d, d, qr[r]) : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                mean[q] = __shfl_sync(0xffffffffu, mg, q);
                m2[q] = __shfl_sync(0xffffffffu, M2, q);
            }
            ntot = (float)a.n_total;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bvar[q] = m2[q] / ntot + eps;                                     // layers.py:1238
            invstd[q] = rsqrtf(bvar[q] + eps);                                // layers.py:1252
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) bn_finalize(stats[pos[q]], stats[P + pos[q]], a.n_total, eps, mean[q], bvar[q], invstd[q]);
    }
    float g[4], bt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        g[q] = a.gamma[pos[q]] * invstd[q];
        bt[q] = a.beta[pos[q]];
        if (lane == 0) {
            rmean[pos[q]] = momentum * rmean[pos[q]] + (1.0f - momentum) * mean[q];      // layers.py:1240-1243
            rvar[pos[q]] = momentum * rvar[pos[q]] + (1.0f - momentum) * bvar[q];
            a.save_mean[pos[q]] = mean[q];
            a.save_invstd[pos[q]] = invstd[q];
        }
    }
#pragma unroll
    for (int j = 0; j < CB_NS; ++j) {                                         // normalise, 2x2 max (layers.py:1256, :599-600)
        float m = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) m = fmaxf(m, fmaf(g[q], S[j][q] - mean[q], bt[q]));
        outs[wid][lane + 32 * j] = m;
    }
    __syncthreads();
    const long long OP = (long long)OH * OW * a.F, obase = ((long long)oy * OW + ox) * a.F + f0;
    for (int s = threadIdx.x; s < a.B; s += CB_THREADS) {                     // one 32-byte run (8 filters) per sample
        const float4 v0 = make_float4(outs[0][s], outs[1][s], outs[2][s], outs[3][s]);
        const float4 v1 = make_float4(outs[4][s], outs[5][s], outs[6][s], outs[7][s]);
        float4* dst = reinterpret_cast<float4*>(y_pool + (long long)s * OP + obase);
        dst[0] = v0;
        dst[1] = v1;
    }
    if (CbBuf<C>::N == 2) buf ^= 1;
  }
    if (MODE == 3) cb_round_done(pr, seq);
}

// Backward. MODE 0: single GPU. MODE 1: pass 1 only (local sum(u), sum(u*xhat) -> dbeta / dgamma, ss0). MODE 2: pass 2 from
// rank-summed dbeta / dgamma (ss1, ss2, dW partials).  partial: [windows][9C + 1][F] (row 9C = bias gradient).
template <int C, int MODE, int ACT>
__global__ void __launch_bounds__(CB_THREADS, C == 1 ? 2 : 1) convblock_bwd_kernel(CbArgs a, const float* __restrict__ err,
                                                                      const double* __restrict__ ss, uint64_t chain,
                                                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                      float* __restrict__ partial, float* __restrict__ r_out,
                                                                      double* ss0, double* ss1, double* ss2, PeerRegion pr,
                                                                      int items) {
    using Cfg = CbCfg<C>;
    extern __shared__ __align__(16) float cb_smem[];                         // CbBuf<C>::N x { patch | errs[CB_MAXB][CB_FPC] }
    __shared__ double scratch[32];
    const float scale = chain_scale(ss, chain);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int OW = a.W >> 1, OH = a.H >> 1;
    const long long OP = (long long)OH * OW * a.F;
    auto stage = [&](int it, int b) {                                         // patches + the pooled error (one 32-byte run per sample)
        float* base = cb_smem + b * CbBuf<C>::FLOATS;
        const int sox = it % OW, soy = (it / OW) % OH, sf0 = (it / (OW * OH)) * CB_FPC;
        cb_stage_async<C>(base, a, soy, sox);
        float* eb = base + CB_MAXB * Cfg::STRIDE;
        const float* src = err + ((long long)soy * OW + sox) * a.F + sf0;
        for (int i = threadIdx.x; i < 2 * a.B; i += CB_THREADS) cb_cp16(eb + i * 4, src + (long long)(i >> 1) * OP + (i & 1) * 4);
    };
    int buf = 0;
    if (CbBuf<C>::N == 2 && (int)blockIdx.x < items) stage(blockIdx.x, 0);
    unsigned seq = 0;
    if (MODE == 3) {
        seq = *reinterpret_cast<volatile unsigned*>(&reinterpret_cast<SymHeader*>(pr.peers.p[pr.rank])->seq[pr.site]) + 1u;
        if (pr.stagger_ns > 0) __nanosleep((blockIdx.x % 3) * pr.stagger_ns);
    }
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;                                   // clip-slot partials, accumulated over this CTA's items
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int ox = item % OW, oy = (item / OW) % OH, f0 = (item / (OW * OH)) * CB_FPC, f = f0 + wid;
    float wr[9 * C], bias;
    cb_load_filter<C>(a, f, wr, bias);
    long long pos[4];
    float mean[4], invstd[4], g[4], bt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        pos[q] = ((long long)(2 * oy + (q >> 1)) * a.W + 2 * ox + (q & 1)) * a.F + f;
        mean[q] = a.save_mean[pos[q]];
        invstd[q] = a.save_invstd[pos[q]];
        g[q] = a.gamma[pos[q]] * invstd[q];
        bt[q] = a.beta[pos[q]] - g[q] * mean[q];                              // y = g * c + bt  with c = clip(act(conv))
        mean[q] = -mean[q] * invstd[q];                                       // xhat = c * invstd + mean   (folded constants)
    }
    float* patch = cb_smem + buf * CbBuf<C>::FLOATS;
    const float* errs = patch + CB_MAXB * Cfg::STRIDE;                        // [sample][8 filters]
    if (CbBuf<C>::N == 2) {
        cb_async_wait();
        __syncthreads();
        if (item + (int)gridDim.x < items) stage(item + gridDim.x, buf ^ 1);
    } else {
        __syncthreads();                                                      // the previous item is done with the shared tiles
        stage(item, 0);
        cb_async_wait();
        __syncthreads();
    }
    float S[CB_NS][4];                                                        // act(conv) (ReLU: already clipped, see cb_keep)
#pragma unroll
    for (int j = 0; j < CB_NS; ++j) {
        const int s = lane + 32 * j;
        if (s < a.B) {
            cb_conv<C, ACT>(patch + s * Cfg::STRIDE, wr, bias, a.act, a.alpha, S[j]);
#pragma unroll
            for (int q = 0; q < 4; ++q) S[j][q] = cb_keep<ACT>(S[j][q]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) S[j][q] = 0.0f;
        }
    }
    float su[4] = {0.f, 0.f, 0.f, 0.f}, sux[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned tie = 0u;                                                        // 4 tie bits per sample, kept from pass 1 for pass 2
    if (MODE != 2) {
#pragma unroll
        for (int j = 0; j < CB_NS; ++j) {                    // pass 1: U = pool-bwd(e); d_beta = sum U, d_gamma = sum U*xhat
            const int s = lane + 32 * j;
            const float e = s < a.B ? scale * errs[s * CB_FPC + wid] : 0.0f;
            float y[4], xh[4], m = -INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float c = cb_clipped<ACT>(S[j][q]);
                y[q] = fmaf(g[q], c, bt[q]);
                xh[q] = fmaf(c, invstd[q], mean[q]);
                m = fmaxf(m, y[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool hit = y[q] == m;                                 // every tied maximum (layers.py:631-637)
                tie |= hit ? (1u << (4 * j + q)) : 0u;
                const float u = hit ? e : 0.0f;
                su[q] += u;
                sux[q] = fmaf(u, xh[q], sux[q]);
                a0 = fmaf(u, u, a0);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            su[q] = warp_sum(su[q]);
            sux[q] = warp_sum(sux[q]);
        }
        if (MODE == 3) {                                     // sum the two reductions over the ranks, in rank order
            float mine = su[0];                                               // lane l < 4: su[l]; lane 4 + l: sux[l]
            mine = lane == 1 ? su[1] : mine; mine = lane == 2 ? su[2] : mine; mine = lane == 3 ? su[3] : mine;
            mine = lane == 4 ? sux[0] : mine; mine = lane == 5 ? sux[1] : mine; mine = lane == 6 ? sux[2] : mine; mine = lane == 7 ? sux[3] : mine;
            float in[MAX_WORLD];
            cb_peer_gather(pr, seq, item, items, mine, in);
            float tot = 0.0f;
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) tot += r < pr.world ? in[r] : 0.0f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                su[q] = __shfl_sync(0xffffffffu, tot, q);
                sux[q] = __shfl_sync(0xffffffffu, tot, 4 + q);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (dbeta != nullptr) dbeta[pos[q]] = su[q];                 // layers.py:1262
                if (dgamma != nullptr) dgamma[pos[q]] = sux[q];              // layers.py:1261
            }
        }
        if (MODE == 1) {
            if (CbBuf<C>::N == 2) buf ^= 1;
            continue;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            su[q] = dbeta[pos[q]];
            sux[q] = dgamma[pos[q]];
        }
    }
    const float invB = (float)(1.0 / a.n_total);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        su[q] *= -invB * g[q];                                                // d = g*u - xhat * g*sum(u xhat)/B - g*sum(u)/B
        sux[q] *= -invB * g[q];
    }
    float gw[9 * C], gb = 0.0f;
#pragma unroll
    for (int k = 0; k < 9 * C; ++k) gw[k] = 0.0f;
#pragma unroll
    for (int j = 0; j < CB_NS; ++j) {                        // pass 2: BN-bwd, act', weight gradient
        const int s = lane + 32 * j;
        if (s >= a.B) continue;
        const float e = scale * errs[s * CB_FPC + wid];
        float r[4];
        unsigned bits;
        if (MODE == 2) {                                                      // pass 1 ran in another launch: recompute the ties
            float y[4], m = -INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                y[q] = fmaf(g[q], cb_clipped<ACT>(S[j][q]), bt[q]);
                m = fmaxf(m, y[q]);
            }
            bits = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) bits |= (y[q] == m) ? (1u << q) : 0u;
        } else {
            bits = tie >> (4 * j);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float u = (bits >> q) & 1u ? e : 0.0f;
            const float xh = fmaf(cb_clipped<ACT>(S[j][q]), invstd[q], mean[q]);
            float d = fmaf(g[q], u, fmaf(xh, sux[q], su[q]));                 // layers.py:1264-1274 in closed form
            a1 = fmaf(d, d, a1);
            d *= cb_act_grad<ACT>(S[j][q], a.act, a.alpha);                   // layers.py:237
            a2 = fmaf(d, d, a2);
            r[q] = d;
            gb += d;
        }
        if (r_out != nullptr) {
            const long long nb = (long long)s * a.H * a.W * a.F;
#pragma unroll
            for (int q = 0; q < 4; ++q) r_out[nb + pos[q]] = r[q];
        }
        const float* p = patch + s * Cfg::STRIDE;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float rows[4][4];
#pragma unroll
            for (int py = 0; py < 4; ++py) {
                const float4 rr = *reinterpret_cast<const float4*>(p + c * 16 + py * 4);
                rows[py][0] = rr.x; rows[py][1] = rr.y; rows[py][2] = rr.z; rows[py][3] = rr.w;
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {                              // layers.py:522-524 (col^T . dY), K order (c,ky,kx)
                    float t = gw[(c * 3 + ky) * 3 + kx];
                    t = fmaf(rows[ky][kx], r[0], t);
                    t = fmaf(rows[ky][kx + 1], r[1], t);
                    t = fmaf(rows[ky + 1][kx], r[2], t);
                    t = fmaf(rows[ky + 1][kx + 1], r[3], t);
                    gw[(c * 3 + ky) * 3 + kx] = t;
                }
        }
    }
    if (partial != nullptr) {
        float* dst = partial + ((long long)(oy * OW + ox) * (Cfg::K + 1)) * a.F + f;
#pragma unroll
        for (int k = 0; k < 9 * C; ++k) {
            const float t = warp_sum(gw[k]);
            if (lane == 0) dst[(long long)k * a.F] = t;
        }
        const float tb = warp_sum(gb);
        if (lane == 0) dst[(long long)Cfg::K * a.F] = tb;                     // layers.py:521
    }
    if (CbBuf<C>::N == 2) buf ^= 1;
  }
    if (MODE != 2 && ss0 != nullptr) block_accumulate(a0, ss0, scratch);
    if (MODE != 1 && ss1 != nullptr) block_accumulate(a1, ss1, scratch);
    if (MODE != 1 && ss2 != nullptr) block_accumulate(a2, ss2, scratch);
    if (MODE == 3) cb_round_done(pr, seq);
}

// partial[s][m][n] -> out, fixed order, with the lazy clip multiplier; rows >= n_rows go to db (same contract as the split-K
// reduction of gemm_simt.cu, kept local so that this file is self-contained)
__global__ void __launch_bounds__(1024) convblock_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                                                const double* __restrict__ ss, uint64_t chain,
                                                                float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float sm[32][32];
    const float scale = chain_scale(ss, chain);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long total = (long long)M * N;
    const long long i = (long long)blockIdx.x * 32 + lane;
    float v = 0.0f;
    if (i < total) {
#pragma unroll 4
        for (int s = w; s < splits; s += 32) v += ws[(long long)s * total + i];
    }
    sm[w][lane] = v;
    __syncthreads();
    if (w == 0 && i < total) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += sm[k][lane];
        t *= scale;
        const int m = (int)(i / N), n = (int)(i % N);
        if (m == M - 1) { if (db != nullptr) db[n] = t; }
        else dw[(long long)m * N + n] = t;
    }
}

static int cb_check(const char* who, const b200nn_conv_geom* g, int act, float alpha) {
    B200NN_REQUIRE(g != nullptr, "%s: null geometry", who);
    B200NN_REQUIRE(g->kh == 3 && g->kw == 3 && g->sh == 1 && g->sw == 1 && g->ph == 1 && g->pw == 1 && g->OH == g->H && g->OW == g->W,
                   "%s: needs a 3x3 stride-1 'same' convolution", who);
    B200NN_REQUIRE(g->C >= 1 && g->C <= 3 && g->F % CB_FPC == 0 && g->H % 2 == 0 && g->W % 2 == 0 && g->N > 0 && g->N <= CB_MAXB,
                   "%s: needs C <= 3, F %% 8 == 0, even H and W, batch <= %d (got C=%d F=%d %dx%d batch %d)", who, CB_MAXB, g->C, g->F,
                   g->H, g->W, g->N);
    B200NN_REQUIRE(act >= 0 && act <= B200NN_ACT_LEAKYRELU && (act != B200NN_ACT_LEAKYRELU || alpha > 0.0f), "%s: bad activation", who);
    B200NN_REQUIRE(g->W / 2 <= 65535 && g->H / 2 <= 65535 && g->F / CB_FPC <= 65535, "%s: grid too large", who);
    return B200NN_OK;
}

// grid: one CTA per item, or (persistent) as many CTAs as are co-resident on the device, whichever is smaller
template <class Kern, class... Args>
static int cb_launch(Kern kern, int items, int persistent /* 0 no, 1 yes, 2 yes with head-room for NCCL */, size_t smem,
                     cudaStream_t st, Args... args) {
    struct Entry { const void* k; int resident; };
    static thread_local Entry table[64] = {};
    Entry* e = nullptr;
    for (auto& t : table) {
        if (t.k == (const void*)kern) { e = &t; break; }
        if (t.k == nullptr) {
            B200NN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            B200NN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CB_THREADS, smem));
            t.k = (const void*)kern;
            t.resident = (per_sm > 0 ? per_sm : 1) * kNumSMs;
            e = &t;
            break;
        }
    }
    B200NN_REQUIRE(e != nullptr, "convblock: kernel table full");
    // persistent launches leave a few CTA slots free: an NCCL gradient allreduce may be running on a side stream, and every
    // CTA of the grid must be able to become resident while its peers on the other GPUs wait for it
    const int cap = persistent == 2 && e->resident > 64 ? e->resident - 24 : e->resident;
    const int grid = persistent && items > cap ? cap : items;
    kern<<<grid, CB_THREADS, smem, st>>>(args...);
    return B200NN_OK;
}

}  // namespace b200nn

using namespace b200nn;

extern "C" {

int b200nn_convblock_supported(const b200nn_conv_geom* g, int act, float alpha) {
    if (g == nullptr) return 0;
    return g->kh == 3 && g->kw == 3 && g->sh == 1 && g->sw == 1 && g->ph == 1 && g->pw == 1 && g->OH == g->H && g->OW == g->W &&
           g->C >= 1 && g->C <= 3 && g->F % CB_FPC == 0 && g->H % 2 == 0 && g->W % 2 == 0 && g->N > 0 && g->N <= CB_MAXB && act >= 0 &&
           act <= B200NN_ACT_LEAKYRELU && (act != B200NN_ACT_LEAKYRELU || alpha > 0.0f);
}

size_t b200nn_convblock_ws_bytes(const b200nn_conv_geom* g) {
    return (size_t)(g->H / 2) * (g->W / 2) * (9 * g->C + 1) * g->F * sizeof(float);
}

#define CB_SMEM(C) ((size_t)CbBuf<C>::N * CbBuf<C>::FLOATS * sizeof(float))
// PERS: launch policy of the single-GPU / split modes (1 = persistent grid with next-item prefetch, 0 = one CTA per item);
// the peer-fused mode 3 is always persistent with head-room
#define CB_DISPATCH_C(KERN, MODE, ACT, PERS, ...)                                                                            \
    do {                                                                                                                     \
        int rc_;                                                                                                             \
        if (g->C == 1) rc_ = cb_launch(KERN<1, MODE, ACT>, items, MODE == 3 ? 2 : PERS, CB_SMEM(1), st, __VA_ARGS__);        \
        else if (g->C == 2) rc_ = cb_launch(KERN<2, MODE, ACT>, items, MODE == 3 ? 2 : 0, CB_SMEM(2), st, __VA_ARGS__);      \
        else rc_ = cb_launch(KERN<3, MODE, ACT>, items, MODE == 3 ? 2 : 0, CB_SMEM(3), st, __VA_ARGS__);                     \
        if (rc_) return rc_;                                                                                                 \
    } while (0)
#define CB_DISPATCH(KERN, MODE, PERS, ...)                                                             \
    do {                                                                                               \
        if (act == B200NN_ACT_RELU) CB_DISPATCH_C(KERN, MODE, B200NN_ACT_RELU, PERS, __VA_ARGS__);     \
        else CB_DISPATCH_C(KERN, MODE, -1, PERS, __VA_ARGS__);                                         \
    } while (0)

static int cb_peer_region(const char* who, int mode, b200nn_comm* comm, int site, size_t region_off, int items, PeerRegion* pr) {
    memset(pr, 0, sizeof(*pr));
    if (mode != 3) return B200NN_OK;
    B200NN_REQUIRE(comm != nullptr && comm->peer_ready && site >= 0 && site < comm->n_sites, "%s: mode 3 needs a peer-mapped communicator", who);
    B200NN_REQUIRE(region_off == comm->site_off[site] && region_off + cb_region_bytes(items, comm->world) <= comm->sym_bytes,
                   "%s: exchange region too small (needs %zu bytes)", who, cb_region_bytes(items, comm->world));
    for (int p = 0; p < MAX_WORLD; ++p) pr->peers.p[p] = p < comm->world ? comm->peer[p] : nullptr;
    pr->rank = comm->rank; pr->world = comm->world; pr->site = site; pr->off = region_off;
    static int stagger = -1;
    if (stagger < 0) {
        const char* e = getenv("B200NN_PEER_STAGGER_NS");
        stagger = e != nullptr ? atoi(e) : 0;
    }
    pr->stagger_ns = comm->world > 1 ? (unsigned)stagger : 0u;
    static int poll_all = -1;
    if (poll_all < 0) {
        const char* e = getenv("B200NN_PEER_POLL_ALL");
        poll_all = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    pr->poll_all = poll_all;
    return B200NN_OK;
}

size_t b200nn_convblock_region_bytes(const b200nn_conv_geom* g, int world) {
    return cb_region_bytes((g->H / 2) * (g->W / 2) * (g->F / CB_FPC), world);
}

int b200nn_convblock_fwd(const b200nn_conv_geom* g, const float* x, const float* w, const float* b, int act, float alpha,
                         const float* gamma, const float* beta, float* running_mean, float* running_var, float* save_mean,
                         float* save_invstd, double* stats, int mode, long long B_total, float* y_pool, float momentum,
                         float eps, b200nn_comm* comm, int site, size_t region_off, void* stream) {
    if (int rc = cb_check("convblock_fwd", g, act, alpha)) return rc;
    B200NN_REQUIRE(mode >= 0 && mode <= 3 && (mode == 0 || mode == 3 || stats != nullptr), "convblock_fwd: mode %d needs stats", mode);
    B200NN_REQUIRE(mode == 1 || (reinterpret_cast<uintptr_t>(y_pool) & 15) == 0, "convblock_fwd: y_pool must be 16-byte aligned");
    CbArgs a{g->N, g->H, g->W, g->F, x, w, b, act, alpha, gamma, beta, save_mean, save_invstd, (double)(mode == 0 ? g->N : B_total)};
    const int items = (g->W / 2) * (g->H / 2) * (g->F / CB_FPC);
    PeerRegion pr;
    if (int rc = cb_peer_region("convblock_fwd", mode, comm, site, region_off, items, &pr)) return rc;
    cudaStream_t st = as_stream(stream);
    if (mode == 0) CB_DISPATCH(convblock_fwd_kernel, 0, 1, a, running_mean, running_var, stats, y_pool, momentum, eps, pr, items);
    else if (mode == 1) CB_DISPATCH(convblock_fwd_kernel, 1, 1, a, running_mean, running_var, stats, y_pool, momentum, eps, pr, items);
    else if (mode == 2) CB_DISPATCH(convblock_fwd_kernel, 2, 1, a, running_mean, running_var, stats, y_pool, momentum, eps, pr, items);
    else CB_DISPATCH(convblock_fwd_kernel, 3, 1, a, running_mean, running_var, stats, y_pool, momentum, eps, pr, items);
    B200NN_LAUNCH_CHECK("convblock_fwd");
    return B200NN_OK;
}

int b200nn_convblock_bwd(const b200nn_conv_geom* g, const float* x, const float* w, const float* b, int act, float alpha,
                         const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                         const float* err, const double* ss, uint64_t chain, float* dgamma, float* dbeta, float* dw, float* db,
                         float* r_out, int mode, long long B_total, double* ss_out0, double* ss_out1, double* ss_out2,
                         uint64_t chain_out, float* ws, b200nn_comm* comm, int site, size_t region_off, void* stream) {
    if (int rc = cb_check("convblock_bwd", g, act, alpha)) return rc;
    B200NN_REQUIRE(mode >= 0 && mode <= 3, "convblock_bwd: mode %d", mode);
    B200NN_REQUIRE((mode == 0 || mode == 3 || (dgamma != nullptr && dbeta != nullptr)) && (mode == 1 || ws != nullptr),
                   "convblock_bwd: missing buffer for mode %d", mode);
    B200NN_REQUIRE((reinterpret_cast<uintptr_t>(err) & 15) == 0, "convblock_bwd: err must be 16-byte aligned");
    CbArgs a{g->N, g->H, g->W, g->F, x, w, b, act, alpha, gamma, beta, const_cast<float*>(save_mean), const_cast<float*>(save_invstd),
             (double)(mode == 0 ? g->N : B_total)};
    const int items = (g->W / 2) * (g->H / 2) * (g->F / CB_FPC);
    PeerRegion pr;
    if (int rc = cb_peer_region("convblock_bwd", mode, comm, site, region_off, items, &pr)) return rc;
    cudaStream_t st = as_stream(stream);
    if (mode == 0) CB_DISPATCH(convblock_bwd_kernel, 0, 1, a, err, ss, chain, dgamma, dbeta, ws, r_out, ss_out0, ss_out1, ss_out2, pr, items);
    else if (mode == 1) CB_DISPATCH(convblock_bwd_kernel, 1, 1, a, err, ss, chain, dgamma, dbeta, ws, r_out, ss_out0, ss_out1, ss_out2, pr, items);
    else if (mode == 2) CB_DISPATCH(convblock_bwd_kernel, 2, 1, a, err, ss, chain, dgamma, dbeta, ws, r_out, ss_out0, ss_out1, ss_out2, pr, items);
    else CB_DISPATCH(convblock_bwd_kernel, 3, 1, a, err, ss, chain, dgamma, dbeta, ws, r_out, ss_out0, ss_out1, ss_out2, pr, items);
    B200NN_LAUNCH_CHECK("convblock_bwd");
    if (mode != 1 && dw != nullptr) {
        // the weight gradient still owes the three clips this kernel produced (pool-bwd / BN-bwd / act-bwd outputs): chain_out
        const int M = 9 * g->C + 1, splits = (g->H / 2) * (g->W / 2);
        const long long total = (long long)M * g->F;
        convblock_reduce_kernel<<<(int)((total + 31) / 32), 1024, 0, st>>>(ws, splits, M, g->F, ss, chain_out, dw, db);
        B200NN_LAUNCH_CHECK("convblock_reduce");
    }
    return B200NN_OK;
}

// The fixed-order reduction of the per-window weight-gradient partials alone (data parallel: it must run AFTER the clip slots
// in chain_out have been summed over the ranks, i.e. as a separate launch from b200nn_convblock_bwd(mode 2, dw = NULL)).
int b200nn_convblock_wgrad_reduce(const b200nn_conv_geom* g, const float* ws, const double* ss, uint64_t chain_out, float* dw,
                                  float* db, void* stream) {
    B200NN_REQUIRE(g != nullptr && ws != nullptr && dw != nullptr, "convblock_wgrad_reduce: null argument");
    const int M = 9 * g->C + 1, splits = (g->H / 2) * (g->W / 2);
    const long long total = (long long)M * g->F;
    convblock_reduce_kernel<<<(int)((total + 31) / 32), 1024, 0, as_stream(stream)>>>(ws, splits, M, g->F, ss, chain_out, dw, db);
    B200NN_LAUNCH_CHECK("convblock_reduce");
    return B200NN_OK;
}

}  // extern "C"
