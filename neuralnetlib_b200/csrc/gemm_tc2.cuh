// gemm_tc2.cuh — second-generation Blackwell tensor-core contraction (math modes TF32 / TF32x3), replacing the one-tile-per-CTA
// kernel of gemm_tcgen05.cuh wherever it applies. Same operands (fp32 read as TF32 by tcgen05.mma.kind::tf32, both majors, the
// implicit-GEMM convolution producers of ConvTma), different machine mapping:
//   * PERSISTENT: one CTA pair (or CTA) per SM pair, static round-robin over (tile, k-split) work items in an L2-friendly
//     grouped order; barriers, TMEM and tensor-map prefetch are paid once per kernel, not once per tile;
//   * CTA PAIRS: tcgen05.mma.cta_group::2 — a 256 x BN tile per pair, each CTA stages its own 128 rows of A and its own HALF of
//     the B columns (per-SM shared-memory traffic per MMA falls from 12 KB to 8 KB), TMA loads signal the leader's mbarrier,
//     tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs;
//   * TWO ACCUMULATORS in TMEM (2 x BN columns): the epilogue of work item i overlaps the main loop of item i + 1;
//   * TMA-STORE EPILOGUE: tcgen05.ld (thread = accumulator row) -> bias / activation / clip multiplier / activation-backward
//     mask / sums of squares in registers -> 128-byte-swizzled shared staging (conflict-free st.shared.v4) ->
//     cp.async.bulk.tensor store of 32 x 32 boxes (full lines, tails clipped by the tensor map). Outputs TMA cannot describe
//     (conv weight gradient row permutation, unaligned pitches, split-K partials) take a direct float4 path.
// X3 (error-compensated 3xTF32, see gemm_tcgen05.cuh) keeps its in-kernel splitter: eight extra warps write the lo tiles of
// each staged operand tile to a mirror ring.
#pragma once

namespace b200nn {
namespace tc2 {

using tc::BLOCK_K;
using tc::BLOCK_M;
using tc::ConvTma;
using tc::SLAB_BYTES;
using tc::UMMA_K;

constexpr int EPI_WARPS = 4;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;        // TMA warp, MMA warp, four epilogue warps
constexpr int SPLIT_WARPS = 8;
constexpr int NUM_THREADS_X3 = NUM_THREADS + 32 * SPLIT_WARPS;
constexpr uint32_t EPI_BUF_BYTES = 32 * 128;            // one 32 x 32 fp32 box
constexpr uint32_t EPI_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default .release.cta semantics (what CUTLASS' umma_arrive_2x1SM_sm0 issues): the .cluster-scope form costs a MEMBAR.ALL +
    // ERRBAR per arrival; the data the arrival orders (TMEM reads, smem writes behind fence.proxy.async) is fenced explicitly
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the data lands in the executing CTA, the bytes are counted on the LEADER's mbarrier
template <int CG>
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if constexpr (CG == 2)
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
    else
        tc::tma_load_2d(dst, map, bar, c0, c1);
}
template <int CG>
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    if constexpr (CG == 2)
        asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
        tc::tma_load_3d(dst, map, bar, c0, c1, c2);
}
template <int CG>
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    if constexpr (CG == 2)
        asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
        tc::tma_load_4d(dst, map, bar, c0, c1, c2, c3);
}
template <int CG>
__device__ __forceinline__ void tc2_commit(uint32_t bar) {
    if constexpr (CG == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3) : "memory");
    else
        tc::tc_commit(bar);
}
template <int CG>
__device__ __forceinline__ void tc2_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int CG>
__device__ __forceinline__ void tc2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        tc::tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// (tile, split) of work item w: tiles in groups of 8 row blocks, column blocks fastest inside a group, so that the ~74 pairs
// running at one time share a band of A rows and a band of B columns in L2
struct Work {
    int m_blk, n_blk, split;
};
__device__ __forceinline__ Work decode_work(int w, int num_m, int num_n) {
    const int tiles = num_m * num_n;
    Work r;
    r.split = w / tiles;
    const int t = w - r.split * tiles;
    constexpr int GROUP = 8;
    const int gsz = GROUP * num_n;
    const int g = t / gsz;
    const int first = g * GROUP;
    const int gm = min(num_m - first, GROUP);
    const int in = t - g * gsz;
    r.m_blk = first + in % gm;
    r.n_blk = in / gm;
    return r;
}

// epilogue arithmetic on one accumulator row segment of 32 columns (compile-time activation kinds, see tc::tc_epi_rows)
template <int AM, int MM>
__device__ __forceinline__ void epi_math(const Epilogue& e, float scale, float (&v)[32], const float* __restrict__ bias_n,
                                         const float* __restrict__ mask_row, int ncols, bool row_ok, float& a0, float& a1) {
    const int act = AM == 0 ? B200NN_ACT_NONE : (AM == 1 ? B200NN_ACT_RELU : e.act);
    const int mact = MM == 1 ? B200NN_ACT_RELU : e.mask_act;
    float b[32], mk[32];
    if (bias_n != nullptr) {
        if (ncols == 32 && (reinterpret_cast<uintptr_t>(bias_n) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(bias_n) + j);
                b[4 * j] = t.x; b[4 * j + 1] = t.y; b[4 * j + 2] = t.z; b[4 * j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) b[i] = i < ncols ? __ldg(bias_n + i) : 0.0f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) b[i] = 0.0f;
    }
    if (MM != 0) {
        if (row_ok && ncols == 32 && (reinterpret_cast<uintptr_t>(mask_row) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(mask_row) + j);
                mk[4 * j] = t.x; mk[4 * j + 1] = t.y; mk[4 * j + 2] = t.z; mk[4 * j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mk[i] = (row_ok && i < ncols) ? __ldg(mask_row + i) : 0.0f;
        }
    }
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float x = act_apply(act, fmaf(v[i], scale, b[i]), e.alpha);
        if (!row_ok || i >= ncols) x = 0.0f;
        s0 = fmaf(x, x, s0);
        if (MM != 0) {
            x *= act_grad_from_y(mact, mk[i], e.mask_alpha);
            s1 = fmaf(x, x, s1);
        }
        v[i] = x;
    }
    a0 += s0;
    a1 += s1;
}

// F16: both operands arrive as fp16 hi / lo PLANES (K-major, split once per call by f16_split.cu: x * s = hi + 2^-11 lo with a
// power-of-two range scale s per tensor); 3 MMAs of kind::f16 per k-step (hi.hi into the main accumulator, lo.hi + hi.lo into the
// second one) at TWICE the TF32 rate -- the same 22-bit products as X3 for half the tensor-pipe time, and no splitter warps.
template <int BN, int STAGES, bool A_MN, bool B_MN, int CG, bool X3, bool F16 = false>
__global__ void __launch_bounds__(X3 ? NUM_THREADS_X3 : NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmOut, Epilogue epi, int M, int N, int K, int kb_per_split, int splits,
                float* __restrict__ ws, const ConvTma cv, int tma_out) {
    constexpr int BNH = BN / CG;                                   // B columns staged by one CTA
    constexpr uint32_t A_BYTES = BLOCK_M * 128, B_BYTES = BNH * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    static_assert(!(X3 && F16) && !(F16 && (A_MN || B_MN)), "F16 planes are K-major and pre-split");
    constexpr bool DUAL = X3 || F16;                               // hi and lo operand tiles, two accumulators
    constexpr uint32_t LO_OFFSET = STAGES * STAGE_BYTES;           // DUAL: the lo tiles mirror the hi / raw ring behind it
    constexpr int ACC_COLS = BN * (DUAL ? 2 : 1);                  // DUAL: main + lo-product accumulators
    constexpr int NUM_ACC = ACC_COLS * 2 <= 512 ? 2 : 1;
    constexpr int TMEM_COLS = ACC_COLS * NUM_ACC < 32 ? 32 : ACC_COLS * NUM_ACC;
    constexpr uint32_t RING_BYTES = STAGES * STAGE_BYTES * (DUAL ? 2 : 1);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], split_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-byte alignment
    const uint32_t epi_base = smem_base + RING_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;
    const int num_m = (M + BLOCK_M * CG - 1) / (BLOCK_M * CG), num_n = (N + BN - 1) / BN;
    const int total_work = num_m * num_n * splits;
    const int kb_total = (K + BLOCK_K - 1) / BLOCK_K;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        if (tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmOut)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
                tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
                tc::mbar_init(tc::smem_u32(&split_bar[s]), SPLIT_WARPS * CG);   // X3: the splitter warps of BOTH CTAs
            }
            for (int a = 0; a < 2; ++a) {
                tc::mbar_init(tc::smem_u32(&tmem_full_bar[a]), 1);
                tc::mbar_init(tc::smem_u32(&tmem_empty_bar[a]), EPI_WARPS * CG);   // the epilogue warps of BOTH CTAs
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base_slot)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base_slot)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc::tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer (one elected lane per CTA; every CTA stages its own A rows and its own half of the B columns) =====
        if (tc::elect_one()) {
            uint32_t it = 0;   // k-blocks issued so far (ring position)
            for (int w = unit; w < total_work; w += num_units) {
                const Work wk = decode_work(w, num_m, num_n);
                const int m0 = (wk.m_blk * CG + (int)rank) * BLOCK_M;
                const int n0 = wk.n_blk * BN + (int)rank * BNH;
                const int kb_begin = wk.split * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
                // convolution producers: every division is hoisted out of the k loop (the loop below is the serial path that
                // feeds the whole SM); tap / channel / pixel positions advance incrementally
                int tap = 0, c0 = 0, ky = 0, kx = 0, img = 0, y0 = 0, x0 = 0;
                if (cv.mode == 1 || cv.mode == 2) {
                    const int k0 = kb_begin * BLOCK_K * (F16 ? 2 : 1), HW = cv.H * cv.W;
                    tap = k0 / cv.C; c0 = k0 - tap * cv.C; ky = tap / cv.kw; kx = tap - ky * cv.kw;
                    img = m0 / HW; y0 = (m0 - img * HW) / cv.W;
                } else if (cv.mode == 3) {
                    const int k0 = kb_begin * BLOCK_K, HW = cv.H * cv.W;
                    img = k0 / HW;
                    const int rem = k0 - img * HW;
                    y0 = rem / cv.W; x0 = rem - y0 * cv.W;
                }
                int a_tap[BLOCK_M / 32], a_c0[BLOCK_M / 32], a_ky[BLOCK_M / 32], a_kx[BLOCK_M / 32];
                if (cv.mode == 3) {
#pragma unroll
                    for (int j = 0; j < BLOCK_M / 32; ++j) {
                        const int r = m0 + j * 32;
                        a_tap[j] = r / cv.C; a_c0[j] = r - a_tap[j] * cv.C;
                        a_ky[j] = a_tap[j] / cv.kw; a_kx[j] = a_tap[j] - a_ky[j] * cv.kw;
                    }
                }
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t round = it / STAGES;
                    tc::mbar_wait(tc::smem_u32(&empty_bar[s]), (round & 1) ^ 1);   // first pass over the ring returns immediately
                    const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
                    const uint32_t own = tc::smem_u32(&full_bar[s]);
                    constexpr int LG = X3 ? 1 : CG;     // X3: loads are counted per CTA (see the splitter), else on the leader's barrier
                    if (X3 || rank == 0) tc::mbar_expect_tx(own, STAGE_BYTES * LG * (F16 ? 2 : 1));
                    const uint32_t bar = LG == 2 ? mapa(own, 0) : own;
                    const int k0 = kb * BLOCK_K * (F16 ? 2 : 1);          // F16: a k-block is 64 two-byte elements
                    if (cv.mode == 1 || cv.mode == 2) {
                        // A tile = 128 consecutive pixels (whole image rows / whole images), one tap, 32 channels
                        const int dy = cv.mode == 1 ? ky - cv.pad_y : cv.pad_y - ky;
                        const int dx = cv.mode == 1 ? kx - cv.pad_x : cv.pad_x - kx;
                        tma2_load_4d<LG>(sa, &tmA, bar, c0, dx, y0 + dy, img);
                        if constexpr (F16) {
                            // fp16 planes: 64 channels of one tap per k-block; the lo images follow the hi images in the same
                            // map; the weights are a plain K-major plane matrix [filters | channels][(tap, channel)]
                            tma2_load_4d<LG>(sa + LO_OFFSET, &tmA, bar, c0, dx, y0 + dy, img + cv.lo_a);
                            tma2_load_2d<LG>(sb, &tmB, bar, k0, n0);
                            tma2_load_2d<LG>(sb + LO_OFFSET, &tmB, bar, k0, n0 + cv.lo_b);
                        } else if constexpr (B_MN) {
#pragma unroll
                            for (int j = 0; j < BNH / 32; ++j) tma2_load_3d<LG>(sb + j * SLAB_BYTES, &tmB, bar, n0 + j * 32, tap, c0);
                        } else {
                            tma2_load_3d<LG>(sb, &tmB, bar, c0, tap, n0);
                        }
                        c0 += BLOCK_K * (F16 ? 2 : 1);
                        if (c0 >= cv.C) {
                            c0 = 0; ++tap;
                            if (++kx == cv.kw) { kx = 0; ++ky; }
                        }
                    } else if (cv.mode == 3) {
                        // A slab j = 32 (tap, c) rows x 32 consecutive pixels; rows past the last tap are fetched far out of bounds (zeros)
                        if constexpr (A_MN) {
#pragma unroll
                            for (int j = 0; j < BLOCK_M / 32; ++j) {
                                const int yy = a_tap[j] < cv.taps ? y0 + a_ky[j] - cv.pad_y : -(1 << 20);
                                tma2_load_4d<LG>(sa + j * SLAB_BYTES, &tmA, bar, a_c0[j], x0 + a_kx[j] - cv.pad_x, yy, img);
                            }
                        }
                        if constexpr (B_MN) {
#pragma unroll
                            for (int j = 0; j < BNH / 32; ++j) tma2_load_2d<LG>(sb + j * SLAB_BYTES, &tmB, bar, n0 + j * 32, k0);
                        }
                        x0 += BLOCK_K;                       // 32 pixels further (W divides 32 or 32 divides W: conv_tc_ok)
                        while (x0 >= cv.W) { x0 -= cv.W; ++y0; }
                        while (y0 >= cv.H) { y0 -= cv.H; ++img; }
                    } else {
                        if constexpr (A_MN) {
#pragma unroll
                            for (int j = 0; j < BLOCK_M / 32; ++j) tma2_load_2d<LG>(sa + j * SLAB_BYTES, &tmA, bar, m0 + j * 32, k0);
                        } else {
                            tma2_load_2d<LG>(sa, &tmA, bar, k0, m0);
                            if constexpr (F16) tma2_load_2d<LG>(sa + LO_OFFSET, &tmA, bar, k0, m0 + cv.lo_a);
                        }
                        if constexpr (B_MN) {
#pragma unroll
                            for (int j = 0; j < BNH / 32; ++j) tma2_load_2d<LG>(sb + j * SLAB_BYTES, &tmB, bar, n0 + j * 32, k0);
                        } else {
                            tma2_load_2d<LG>(sb, &tmB, bar, k0, n0);
                            if constexpr (F16) tma2_load_2d<LG>(sb + LO_OFFSET, &tmB, bar, k0, n0 + cv.lo_b);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread of the LEADER CTA issues for the pair) =====
        if (rank == 0 && tc::elect_one()) {
            constexpr uint32_t idesc = F16 ? tc::make_idesc_f16(BLOCK_M * CG, BN) : tc::make_idesc(BLOCK_M * CG, BN, A_MN, B_MN);
            auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t accum) {
                if constexpr (F16) tc2_mma_f16<CG>(d, ad, bd, id, accum); else tc2_mma<CG>(d, ad, bd, id, accum);
            };
            uint32_t it = 0, item = 0;
            for (int w = unit; w < total_work; w += num_units, ++item) {
                const Work wk = decode_work(w, num_m, num_n);
                const int kb_begin = wk.split * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
                const uint32_t acc = NUM_ACC == 2 ? (item & 1) : 0;
                const uint32_t use = NUM_ACC == 2 ? (item >> 1) : item;        // how often this accumulator has been used before
                tc::mbar_wait(tc::smem_u32(&tmem_empty_bar[acc]), (use & 1) ^ 1);   // the epilogues of both CTAs have drained it
                tc::tc_fence_after();
                const uint32_t d_main = tmem_base + acc * ACC_COLS;
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t round = it / STAGES;
                    tc::mbar_wait(tc::smem_u32(X3 ? &split_bar[s] : &full_bar[s]), round & 1);
                    tc::tc_fence_after();
                    const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
                    const bool first = kb == kb_begin;
#pragma unroll
                    for (int j = 0; j < BLOCK_K / UMMA_K; ++j) {
                        const uint64_t ad = A_MN ? tc::desc_mnmajor(sa, j) : tc::desc_kmajor(sa, j);
                        const uint64_t bd = B_MN ? tc::desc_mnmajor(sb, j) : tc::desc_kmajor(sb, j);
                        mma(d_main, ad, bd, idesc, (!first || j > 0) ? 1u : 0u);
                        if constexpr (DUAL) {
                            const uint64_t adl = A_MN ? tc::desc_mnmajor(sa + LO_OFFSET, j) : tc::desc_kmajor(sa + LO_OFFSET, j);
                            const uint64_t bdl = B_MN ? tc::desc_mnmajor(sb + LO_OFFSET, j) : tc::desc_kmajor(sb + LO_OFFSET, j);
                            mma(d_main + BN, adl, bd, idesc, (!first || j > 0) ? 1u : 0u);   // A_lo . B_hi
                            mma(d_main + BN, ad, bdl, idesc, 1u);                           // A_hi . B_lo
                        }
                    }
                    tc2_commit<CG>(tc::smem_u32(&empty_bar[s]));        // frees the stage (in both CTAs) once these MMAs have read it
                }
                tc2_commit<CG>(tc::smem_u32(&tmem_full_bar[acc]));       // accumulator complete (signalled in both CTAs)
            }
        }
        __syncwarp();
    } else if (X3 && warp >= 2 + EPI_WARPS) {
        // ===== X3 operand splitter (eight warps): lo = x - trunc_tf32(x) of every staged tile into the mirror ring =====
        constexpr int SPLIT_THREADS = 32 * SPLIT_WARPS;
        constexpr int PER = (STAGE_BYTES / 16 + SPLIT_THREADS - 1) / SPLIT_THREADS;   // float4 per thread and stage
        const int t = threadIdx.x - 32 * (2 + EPI_WARPS);
        uint32_t it = 0;
        for (int w = unit; w < total_work; w += num_units) {
            const Work wk = decode_work(w, num_m, num_n);
            const int kb_begin = wk.split * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
            for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t round = it / STAGES;
                // X3: every CTA's loads complete on its OWN full barrier (an mbarrier can only be waited on locally); the MMA
                // issuer waits for the splitter warps of both CTAs instead (split_bar in the leader)
                tc::mbar_wait(tc::smem_u32(&full_bar[s]), round & 1);
                const uint32_t off = (smem_base - tc::smem_u32(smem_raw)) + s * STAGE_BYTES;
                const float4* src = reinterpret_cast<const float4*>(smem_raw + off);
                float4* dst = reinterpret_cast<float4*>(smem_raw + off + LO_OFFSET);
                float4 x[PER];
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const uint32_t v = t + u * SPLIT_THREADS;
                    if (v < STAGE_BYTES / 16) x[u] = src[v];
                }
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const uint32_t v = t + u * SPLIT_THREADS;
                    if (v < STAGE_BYTES / 16)
                        dst[v] = make_float4(tc::tf32_lo(x[u].x), tc::tf32_lo(x[u].y), tc::tf32_lo(x[u].z), tc::tf32_lo(x[u].w));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_cluster(mapa(tc::smem_u32(&split_bar[s]), 0));
                    else tc::mbar_arrive(tc::smem_u32(&split_bar[s]));
                }
            }
        }
    } else if (warp >= 2 && warp < 2 + EPI_WARPS) {
        // ===== epilogue (four warps; TMEM lane quarter = warp % 4; thread = one accumulator row) =====
        const int q = warp & 3;
        const int ew = warp - 2;
        const float scale = chain_scale(epi.ss, epi.chain);
        float f16_unscale = 1.0f;
        if constexpr (F16) f16_unscale = __ldg(cv.inv_scale) * __ldg(cv.inv_scale + 1);
        const int am = epi.act == B200NN_ACT_NONE ? 0 : (epi.act == B200NN_ACT_RELU ? 1 : 2);
        const int mm = epi.mask_y == nullptr ? 0 : (epi.mask_act == B200NN_ACT_RELU ? 1 : 2);
        double d0 = 0.0, d1 = 0.0;
        uint32_t item = 0, sbuf = 0;
        for (int w = unit; w < total_work; w += num_units, ++item) {
            const Work wk = decode_work(w, num_m, num_n);
            const int m0 = (wk.m_blk * CG + (int)rank) * BLOCK_M + q * 32;   // first row of this warp
            const int n0 = wk.n_blk * BN;
            const uint32_t acc = NUM_ACC == 2 ? (item & 1) : 0;
            const uint32_t use = NUM_ACC == 2 ? (item >> 1) : item;
            tc::mbar_wait(tc::smem_u32(&tmem_full_bar[acc]), use & 1);
            tc::tc_fence_after();
            const int m = m0 + lane;
            const bool row_ok = m < M;
            long long orow = m;
            if (epi.row_mode == 1 && row_ok) {      // conv weight gradient: GEMM rows (tap, c) -> the reference's (c, ky, kx) order
                const int c = m % epi.C, t = m / epi.C;
                const int kx = t % epi.kw, ky = t / epi.kw;
                orow = ((long long)c * epi.kh + ky) * epi.kw + kx;
            }
            // First-order compensation of the tensor core's truncating accumulate. Every MMA adds into the fp32 accumulator with
            // truncation; measured (scripts/probe_trunc_bias.py, profiles/r2_trunc_bias.txt) the result is the exact sum times
            // (1 - 1.64e-8 * MMAs chained on the main accumulator), independent of shape and of the operands' sign pattern
            // (-4.2e-6 at 256 MMAs, -2.1e-6 at 128). Undoing that factor halves the error of the compensated modes and removes
            // the part that compounds from layer to layer (a same-signed shrink of every activation).
            float trunc_comp = 1.0f;
            if constexpr (DUAL) {
                const int kb_begin = wk.split * kb_per_split, kb_end = min(kb_total, kb_begin + kb_per_split);
                trunc_comp = 1.0f + 1.64e-8f * (float)((kb_end - kb_begin) * (BLOCK_K / UMMA_K));
            }
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                const int n = n0 + c;
                if (n >= N) break;
                const int ncols = min(32, N - n);
                float v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + (uint32_t)c;
                tc::tc_ld32(taddr, v);
                if constexpr (X3) {
                    float lo[32];
                    tc::tc_ld32(taddr + BN, lo);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = (v[i] + lo[i]) * trunc_comp;
                }
                if constexpr (F16) {      // (hi.hi + 2^-11 (lo.hi + hi.lo)) / (s_a s_b)
                    float lo[32];
                    tc::tc_ld32(taddr + BN, lo);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaf(lo[i], 0.00048828125f, v[i]) * (f16_unscale * trunc_comp);
                }
                if (splits > 1) {                 // split-K partial: the raw accumulator, reduced by splitk_reduce_*
                    if (row_ok) {
                        float* p = ws + ((long long)wk.split * M + m) * N + n;
                        if (ncols == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)      // unrolled + predicated: v[] must stay in registers
                                if (i < ncols) p[i] = v[i];
                        }
                    }
                    continue;
                }
                const float* bias_n = epi.bias != nullptr ? epi.bias + n : nullptr;
                const float* mask_row = epi.mask_y != nullptr ? epi.mask_y + orow * epi.ldo + n : nullptr;
                switch (am * 3 + mm) {
                    case 0: epi_math<0, 0>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                    case 1: epi_math<0, 1>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                    case 2: epi_math<0, 2>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                    case 3: epi_math<1, 0>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                    case 4: epi_math<1, 1>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                    case 5: epi_math<1, 2>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                    case 6: epi_math<2, 0>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                    case 7: epi_math<2, 1>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                    default: epi_math<2, 2>(epi, scale, v, bias_n, mask_row, ncols, row_ok, a0, a1); break;
                }
                if (tma_out) {
                    // 32 x 32 box in the SWIZZLE_128B layout: row = lane, 16-byte chunk j at (j ^ (lane & 7)) -- conflict-free v4 stores
                    const uint32_t buf = epi_base + (ew * 2 + (sbuf & 1)) * EPI_BUF_BYTES;
                    if (lane == 0) bulk_wait_read<1>();          // the store that last read this buffer (two stores ago) is done
                    __syncwarp();
                    const uint32_t rowaddr = buf + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((j ^ (lane & 7)) << 4)), "f"(v[4 * j]),
                                     "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut, buf, n, m0);      // rows / columns beyond the tensor are clipped by the map
                        bulk_commit();
                    }
                    ++sbuf;
                } else if (row_ok) {
                    float* o = epi.out + orow * epi.ldo + n;
                    if (ncols == 32 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < ncols) o[i] = v[i];
                    }
                }
            }
            d0 += (double)a0;
            d1 += (double)a1;
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) {       // this warp has drained its quarter of the accumulator
                const uint32_t own = tc::smem_u32(&tmem_empty_bar[acc]);
                if constexpr (CG == 2) mbar_arrive_cluster(mapa(own, 0));
                else tc::mbar_arrive(own);
            }
        }
        if (tma_out && lane == 0) bulk_wait_all();
        if (splits == 1) {
            if (epi.ss0 != nullptr) {
                const double r = warp_sum(d0);
                if (lane == 0) atomicAdd(epi.ss0, r);
            }
            if (epi.ss1 != nullptr) {
                const double r = warp_sum(d1);
                if (lane == 0) atomicAdd(epi.ss1, r);
            }
        }
    }
    // teardown: nobody may leave while the peer can still signal into this CTA's shared memory / use the pair's TMEM
    tc::tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        if constexpr (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct Plan2 {
    int bn, cg, stages, splits, kb_per_split, grid;
};

static bool enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200NN_TC2");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
static int forced_cg() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200NN_TC2_CG");
        v = e != nullptr ? atoi(e) : 0;
    }
    return v;
}

// X3: chains of more than tc::X3_CHAIN_KB k-blocks are cut through split-K (see gemm_tcgen05.cuh)
static Plan2 plan2(long long M, long long N, long long K, bool x3) {
    Plan2 p;
    p.bn = N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
    p.cg = (M > BLOCK_M && p.bn >= 64) ? 2 : 1;
    if (forced_cg() == 1) p.cg = 1;
    const long long kb_total = (K + BLOCK_K - 1) / BLOCK_K;
    const long long tiles = ((M + BLOCK_M * p.cg - 1) / (BLOCK_M * p.cg)) * ((N + p.bn - 1) / p.bn);
    const long long units = kNumSMs / p.cg;
    long long splits = 1;
    if (x3) {
        long long chain = tc::X3_CHAIN_KB;
        splits = (kb_total + chain - 1) / chain;
        while (splits > 1 && splits * M * N * (long long)sizeof(float) > tc::X3_WS_CAP) {
            chain *= 2;
            splits = (kb_total + chain - 1) / chain;
        }
    }
    if (tiles * splits < units && kb_total >= 8) {   // short of one wave: split the reduction, tiles kept wide
        long long more = units / tiles;
        const long long maxs = kb_total / 4;           // at least 4 k-blocks (128 k) per split
        if (more > maxs) more = maxs;
        if (more > splits) splits = more;
    }
    const long long per = (kb_total + splits - 1) / splits;
    p.splits = (int)((kb_total + per - 1) / per);
    p.kb_per_split = (int)per;
    const long long work = tiles * p.splits;
    p.grid = (int)((work < units ? work : units) * p.cg);
    // ring depth: what fits beside the 32 KB of epilogue staging in 227 KB
    const long long stage = (long long)(BLOCK_M * 128 + (p.bn / p.cg) * 128) * (x3 ? 2 : 1);
    long long st = (227 * 1024 - 1024 - (long long)EPI_BYTES - 512) / stage;
    p.stages = (int)(st > 8 ? 8 : st);
    return p;
}

static size_t ws_bytes2(long long M, long long N, long long K, bool x3) {
    const Plan2 p = plan2(M, N, K, x3);
    return p.splits > 1 ? (size_t)p.splits * M * N * sizeof(float) : 0;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int CG, bool X3, bool F16 = false>
static int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const Epilogue& epi, int M, int N, int K,
                   const Plan2& p, float* ws, cudaStream_t st, const ConvTma& cv, int tma_out) {
    constexpr size_t smem = (size_t)STAGES * (BLOCK_M * 128 + (BN / CG) * 128) * ((X3 || F16) ? 2 : 1) + EPI_BYTES + 1024;
    static_assert(smem <= 227 * 1024, "ring + staging exceed the shared memory of an SM");
    auto kern = gemm_tc2_kernel<BN, STAGES, A_MN, B_MN, CG, X3, F16>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        B200NN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(X3 ? NUM_THREADS_X3 : NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    B200NN_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, epi, M, N, K, p.kb_per_split, p.splits, p.splits > 1 ? ws : (float*)nullptr, cv, tma_out));
    B200NN_LAUNCH_CHECK("gemm_tc2");
    if (p.splits > 1) return launch_splitk_reduce(ws, p.splits, M, N, epi, st);
    return B200NN_OK;
}

template <int BN, bool A_MN, bool B_MN, int CG, bool X3, bool F16 = false>
static int launch2_stages(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const Epilogue& epi, int M, int N,
                          int K, const Plan2& p, float* ws, cudaStream_t st, const ConvTma& cv, int tma_out) {
    constexpr long long stage = (long long)(BLOCK_M * 128 + (BN / CG) * 128) * ((X3 || F16) ? 2 : 1);
    constexpr long long fit = (227 * 1024 - 1024 - (long long)EPI_BYTES - 512) / stage;
    constexpr int S = fit > 8 ? 8 : (int)fit;
    static_assert(S >= 2, "at least a double-buffered ring");
    return launch2<BN, S, A_MN, B_MN, CG, X3, F16>(ta, tb, to, epi, M, N, K, p, ws, st, cv, tma_out);
}

template <bool X3>
static int dispatch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, bool a_mn, bool b_mn, const Epilogue& epi,
                     int M, int N, int K, const Plan2& p, float* ws, cudaStream_t st, const ConvTma& cv, int tma_out) {
    B200NN_REQUIRE(!(a_mn && !b_mn), "gemm_tc2: (A MN-major, B K-major) is not instantiated");
#define TC2_MAJ(BN, CG)                                                                                              \
    if (a_mn) return launch2_stages<BN, true, true, CG, X3>(ta, tb, to, epi, M, N, K, p, ws, st, cv, tma_out);       \
    if (b_mn) return launch2_stages<BN, false, true, CG, X3>(ta, tb, to, epi, M, N, K, p, ws, st, cv, tma_out);      \
    return launch2_stages<BN, false, false, CG, X3>(ta, tb, to, epi, M, N, K, p, ws, st, cv, tma_out);
    if (p.cg == 2) {
        switch (p.bn) {
            case 64: TC2_MAJ(64, 2)
            case 128: TC2_MAJ(128, 2)
            default: TC2_MAJ(256, 2)
        }
    }
    switch (p.bn) {
        case 32: TC2_MAJ(32, 1)
        case 64: TC2_MAJ(64, 1)
        case 128: TC2_MAJ(128, 1)
        default: TC2_MAJ(256, 1)
    }
#undef TC2_MAJ
}

// store map over out[M, N] (row pitch ldo): boxes of 32 x 32, 128-byte swizzle; 0 when TMA cannot describe the output
static int make_out_map(CUtensorMap* map, const Epilogue& epi, int M, int N, int splits, int* tma_out) {
    *tma_out = 0;
    memset(map, 0, sizeof(*map));
    static int allow = -1;
    if (allow < 0) {
        const char* e = getenv("B200NN_TC2_TMAOUT");
        allow = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!allow || splits > 1 || epi.row_mode != 0 || epi.n_rows < M || epi.db != nullptr) return B200NN_OK;
    if (!tc::tma_ok(epi.out, epi.ldo) || N < 32) return B200NN_OK;
    if (int rc = tc::make_map(map, epi.out, N, M, epi.ldo, 32, false)) return rc;
    *tma_out = 1;
    return B200NN_OK;
}

template <bool X3>
static int run2(const CUtensorMap& ta, const CUtensorMap& tb, bool a_mn, bool b_mn, const Epilogue& epi, int M, int N, int K,
                const Plan2& p, float* ws, cudaStream_t st, const ConvTma& cv) {
    B200NN_REQUIRE(p.splits == 1 || ws != nullptr, "gemm_tc2: split-K needs a workspace");
    CUtensorMap to;
    int tma_out = 0;
    if (int rc = make_out_map(&to, epi, M, N, p.splits, &tma_out)) return rc;
    return dispatch2<X3>(ta, tb, to, a_mn, b_mn, epi, M, N, K, p, ws, st, cv, tma_out);
}

// D[M,N] = A.B ; a_mn: A stored [K][M] (ld = lda) else [M][K]; b_mn: B stored [K][N] else [N][K]
template <bool X3>
static int gemm2(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi,
                 int M, int N, int K, float* ws, cudaStream_t st) {
    const Plan2 p = plan2(M, N, K, X3);
    CUtensorMap ta, tb;
    if (a_mn) { if (int rc = tc::make_map(&ta, A, M, K, lda, BLOCK_K, true)) return rc; }
    else      { if (int rc = tc::make_map(&ta, A, K, M, lda, BLOCK_M, false)) return rc; }
    if (b_mn) { if (int rc = tc::make_map(&tb, B, N, K, ldb, BLOCK_K, true)) return rc; }
    else      { if (int rc = tc::make_map(&tb, B, K, N, ldb, p.bn / p.cg, false)) return rc; }
    ConvTma cv{};
    return run2<X3>(ta, tb, a_mn, b_mn, epi, M, N, K, p, ws, st, cv);
}

template <int UNUSED = 0>   // a template so that only gemm_tc2_f16.cu instantiates the F16 kernels
static int dispatch2_f16(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const Epilogue& epi, int M, int N, int K4,
                         const Plan2& p, float* ws, cudaStream_t st, const ConvTma& cv, int tma_out) {
#define TC2_F16(BN, CG) return launch2_stages<BN, false, false, CG, false, true>(ta, tb, to, epi, M, N, K4, p, ws, st, cv, tma_out);
    if (p.cg == 2) {
        switch (p.bn) {
            case 64: TC2_F16(64, 2)
            case 128: TC2_F16(128, 2)
            default: TC2_F16(256, 2)
        }
    }
    switch (p.bn) {
        case 32: TC2_F16(32, 1)
        case 64: TC2_F16(64, 1)
        case 128: TC2_F16(128, 1)
        default: TC2_F16(256, 1)
    }
#undef TC2_F16
}

// D[M,N] = A.B from fp16 hi / lo planes: Ap = [hi plane (M rows) | lo plane] with the lo plane lo_a rows below, row pitch lda
// (elements), K two-byte elements per row; Bp likewise with N rows. K4 = ceil(K / 2) is the reduction length in the 4-byte units the
// planner counts in (a k-block is 128 bytes either way).
template <int UNUSED = 0>
static int gemm2_f16(const void* Ap, long long lda, int lo_a, const void* Bp, long long ldb, int lo_b, const float* inv_scale,
                     const Epilogue& epi, int M, int N, int K, float* ws, cudaStream_t st) {
    const int K4 = (K + 1) / 2;
    const Plan2 p = plan2(M, N, K4, true);
    CUtensorMap ta, tb, to;
    if (int rc = tc::make_map_f16(&ta, Ap, K, (long long)lo_a + M, lda, BLOCK_M)) return rc;
    if (int rc = tc::make_map_f16(&tb, Bp, K, (long long)lo_b + N, ldb, p.bn / p.cg)) return rc;
    ConvTma cv{};
    cv.lo_a = lo_a; cv.lo_b = lo_b; cv.inv_scale = inv_scale;
    B200NN_REQUIRE(p.splits == 1 || ws != nullptr, "gemm_tc2 (f16 planes): split-K needs a workspace");
    int tma_out = 0;
    if (int rc = make_out_map(&to, epi, M, N, p.splits, &tma_out)) return rc;
    return dispatch2_f16(ta, tb, to, epi, M, N, K4, p, ws, st, cv, tma_out);
}

// implicit-GEMM convolution from fp16 planes (forward: which = 1, Ap = planes of x; data gradient: which = 2, Ap = planes of err).
// Ap = [hi images (g->N) | lo images][H][W][Ck] two-byte elements, Ck = channels along K (C_in resp. F), a multiple of 64;
// Bp = K-major weight planes [hi rows (Nn) | lo rows][K = (tap, channel)] with row pitch ldb, Nn = F resp. C_in.
template <int UNUSED = 0>
static int conv2_f16(int which, const b200nn_conv_geom* g, const void* Ap, const void* Bp, long long ldb, const float* inv_scale,
                     const Epilogue& epi, float* ws, cudaStream_t st) {
    const int Ck = which == 1 ? g->C : g->F, Nn = which == 1 ? g->F : g->C;
    const int M = g->N * g->H * g->W, K = g->kh * g->kw * Ck, K4 = K / 2;
    B200NN_REQUIRE((which == 1 || which == 2) && Ck % 64 == 0, "conv2_f16: forward / data gradient with 64 | channels only");
    const Plan2 p = plan2(M, Nn, K4, true);
    CUtensorMap ta, tb, to;
    if (int rc = tc::make_image_map_f16(&ta, Ap, 2 * g->N, g->H, g->W, Ck, BLOCK_M)) return rc;
    if (int rc = tc::make_map_f16(&tb, Bp, K, 2LL * Nn, ldb, p.bn / p.cg)) return rc;
    ConvTma cv = tc::conv_tma(which, g, Ck);
    cv.lo_a = g->N; cv.lo_b = Nn; cv.inv_scale = inv_scale;
    B200NN_REQUIRE(p.splits == 1 || ws != nullptr, "conv2_f16: split-K needs a workspace");
    int tma_out = 0;
    if (int rc = make_out_map(&to, epi, M, Nn, p.splits, &tma_out)) return rc;
    return dispatch2_f16(ta, tb, to, epi, M, Nn, K4, p, ws, st, cv, tma_out);
}

template <bool X3>
static int conv_fwd2(const b200nn_conv_geom* g, const float* x, const float* w, const Epilogue& epi, float* ws, cudaStream_t st) {
    const int M = g->N * g->H * g->W, N = g->F, K = g->kh * g->kw * g->C;
    const Plan2 p = plan2(M, N, K, X3);
    CUtensorMap ta, tb;
    if (int rc = tc::make_image_map(&ta, x, g->N, g->H, g->W, g->C, BLOCK_M, false)) return rc;
    if (int rc = tc::make_weight_map(&tb, w, g->kh * g->kw, g->C, g->F, 32, true)) return rc;
    return run2<X3>(ta, tb, false, true, epi, M, N, K, p, ws, st, tc::conv_tma(1, g, g->C));
}

template <bool X3>
static int conv_dgrad2(const b200nn_conv_geom* g, const float* err, const float* w, const Epilogue& epi, float* ws, cudaStream_t st) {
    const int M = g->N * g->H * g->W, N = g->C, K = g->kh * g->kw * g->F;
    const Plan2 p = plan2(M, N, K, X3);
    CUtensorMap ta, tb;
    if (int rc = tc::make_image_map(&ta, err, g->N, g->H, g->W, g->F, BLOCK_M, false)) return rc;
    if (int rc = tc::make_weight_map(&tb, w, g->kh * g->kw, g->C, g->F, p.bn / p.cg, false)) return rc;
    return run2<X3>(ta, tb, false, false, epi, M, N, K, p, ws, st, tc::conv_tma(2, g, g->F));
}

template <bool X3>
static int conv_wgrad2(const b200nn_conv_geom* g, const float* x, const float* err, const Epilogue& epi, float* ws, cudaStream_t st) {
    const int M = g->kh * g->kw * g->C, N = g->F, K = g->N * g->H * g->W;
    const Plan2 p = plan2(M, N, K, X3);
    CUtensorMap ta, tb;
    if (int rc = tc::make_image_map(&ta, x, g->N, g->H, g->W, g->C, BLOCK_K, true)) return rc;
    if (int rc = tc::make_map(&tb, err, N, K, N, BLOCK_K, true)) return rc;
    return run2<X3>(ta, tb, true, true, epi, M, N, K, p, ws, st, tc::conv_tma(3, g, g->C));
}

static size_t conv_ws_bytes2(const b200nn_conv_geom* g, bool x3) {
    const long long P = (long long)g->N * g->H * g->W, taps = (long long)g->kh * g->kw;
    size_t r = ws_bytes2(P, g->F, taps * g->C, x3);
    const size_t b = ws_bytes2(P, g->C, taps * g->F, x3), c = ws_bytes2(taps * g->C, g->F, P, x3);
    r = r > b ? r : b;
    return r > c ? r : c;
}

}  // namespace tc2
}  // namespace b200nn
