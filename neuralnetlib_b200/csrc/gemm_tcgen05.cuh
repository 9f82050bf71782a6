// gemm_tcgen05.cuh — Blackwell tensor-core contraction for Dense layers (math mode TF32, rel 2e-3):
//   D[M,N] (fp32) = A[M,K] . B[K,N]  with fp32 operands read as TF32 by tcgen05.mma.kind::tf32,
//   fp32 accumulation in TMEM, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a
//   multi-stage mbarrier pipeline. Warp-specialised: warp 0 = TMA producer (one elected lane), warp 1 = TMEM
//   allocator + single-thread MMA issuer, warps 2..5 = epilogue (tcgen05.ld 32x32b -> registers -> the shared
//   Epilogue: bias / activation / lazy clip multiplier / fused activation backward / fp64 sum-of-squares slots).
//
// Both operand majors are supported WITHOUT transposed copies (the three GEMMs of a Dense layer need all of them):
//   K-major  : element (mn, k) at  base[mn * ld + k]   -> TMA box {32 k, rows mn}, smem tile [mn][128 B], SW128,
//              UMMA descriptor SBO = 1024 B, one MMA k-step (8 tf32) advances the start address by 32 B;
//   MN-major : element (mn, k) at  base[k * ld + mn]   -> BLOCK_MN/32 TMA boxes {32 mn, BLOCK_K k}, each a
//              [k][128 B] slab of 4 KB. 32-bit MN-major operands must use the 32-byte-atom swizzle
//              (UMMA LayoutType SWIZZLE_128B_BASE32B = Swizzle<2,5,2>, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
//              plain SWIZZLE_128B silently yields zeros, see scripts/tc_probe.cu): atom = 128 B x 4 k rows, so
//              LBO = 4096 B (next 32 mn), SBO = 512 B (next 4 k), one MMA k-step advances the start by 1024 B.
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
// Tails in M, N and K are handled by TMA zero fill + guarded stores. Skinny shapes use split-K over blockIdx.z
// (partials to a workspace, reduced in fixed order by splitk_reduce_*).
#pragma once
#include <cuda.h>

namespace b200nn {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;          // tf32 elements = 128 bytes = one swizzle span
constexpr int UMMA_K = 8;            // tf32 elements per tcgen05.mma
constexpr int NUM_THREADS = 192;      // TMA warp, MMA warp, four epilogue warps
constexpr int NUM_THREADS_X3 = 320;   // + four more operand-splitter warps (the split of a stage must not outlast its MMAs)
constexpr uint32_t SLAB_BYTES = BLOCK_K * 128;   // one 32-wide MN slab of an MN-major operand

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline that has not advanced for 4 s of wall time is a protocol bug or a lost peer; trapping turns a
// device hang into a CUDA error the host sees at its next synchronisation.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// Implicit-GEMM convolution operands, fetched by TMA straight from the NHWC tensors (stride 1, 'same', odd kernel, power-of-two
// image, channel counts multiples of 32). The patched tensor is a 4-D map (c, x, y, image); a box of {32 channels, a run of
// pixels} at (c0, x0 + kx - pad, y0 + ky - pad, image) IS a k-block of im2col rows in the canonical swizzled layout, and the
// out-of-bounds pixels TMA zero-fills are the padding -- no im2col buffer, no gather functor (replaces preprocessing.py:34-126
// on the tensor-core path). The weights [(c, ky, kx), F] (the reference's K order, layers.py:470) are a 3-D map (f, tap, c).
//   mode 1  forward : A rows = output pixels, K = (tap, c)   ; B = W  MN-major, slabs {32 f, tap, 32 c}
//   mode 2  dgrad   : A rows = input pixels of err (mirrored taps), K = (tap, f) ; B = W K-major, box {32 f, tap, BLOCK_N c}
//   mode 3  wgrad   : A rows = (tap, c) MN-major, K = pixels ; B = err [pixels, F] MN-major (plain 2-D)
struct ConvTma {
    int mode;          // 0: plain 2-D operands (Dense)
    int C;             // channels per tap: along K (modes 1, 2: C_in resp. F) or along M (mode 3: C_in)
    int kw, taps;
    int W, H;          // image size (input = output)
    int pad_y, pad_x;
    // fp16 hi/lo operand planes (gemm_tc2.cuh F16 mode): the lo plane of A / B starts this many rows below the hi plane in the
    // same tensor map; inv_scale = {1 / s_a, 1 / s_b} (device) undoes the power-of-two range scaling of the planes
    int lo_a, lo_b;
    const float* inv_scale;
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"   // same asm block: the registers are only defined after the wait
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | layout [61,64): SWIZZLE_128B = 2 (K-major), SWIZZLE_128B_BASE32B = 1 (MN-major 32-bit)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kstep) { return make_desc(tile + kstep * 32, 16, 1024, 2); }
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int kstep) { return make_desc(tile + kstep * 1024, SLAB_BYTES, 512, 1); }
// instruction descriptor (InstrDescriptor): c_format F32 = 1 [4,6) | a/b format TF32 = 2 [7,10) [10,13) | a_major [15] |
// b_major [16] | N>>3 [17,23) | M>>4 [24,29)
// kind::f16 with fp16 operands (a / b format 0), fp32 accumulate, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Epilogue inner loop of one warp: `rows` rows x 32 columns (one column per lane) out of the transposed tile. The activation
// kinds are compile-time (AM / MM: 0 none, 1 ReLU, 2 decided at run time) so that the common cases are ~8 instructions per
// element with 8 independent elements in flight; the generic per-element epi_store costs ~47 and, with only four epilogue
// warps per SM, took 45 us per 128x256 tile (ncu: gpurun r1 tc_dw, 4.5 % tensor-pipe activity at M = 256).
template <int AM, int MM>
__device__ __forceinline__ void tc_epi_rows(const Epilogue& e, EpiState& st, const float (*tile)[33], int lane, int rows,
                                            float* __restrict__ o, const float* __restrict__ my, float bias) {
    const int act = AM == 0 ? B200NN_ACT_NONE : (AM == 1 ? B200NN_ACT_RELU : e.act);
    const int mact = MM == 1 ? B200NN_ACT_RELU : e.mask_act;
    const long long ld = e.ldo;
    const float scale = st.scale;
    float a0 = 0.0f, a1 = 0.0f;
    auto one = [&](float acc, float mk, long long off) {
        float x = act_apply(act, fmaf(acc, scale, bias), e.alpha);
        a0 = fmaf(x, x, a0);
        if (MM != 0) {
            x *= act_grad_from_y(mact, mk, e.mask_alpha);
            a1 = fmaf(x, x, a1);
        }
        o[off] = x;
    };
    if (rows == 32) {
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
            float v[8], mk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {   // 8 independent smem (+ global mask) loads in flight before any use
                v[j] = tile[r0 + j][lane];
                mk[j] = MM != 0 ? my[(r0 + j) * ld] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) one(v[j], mk[j], (r0 + j) * ld);
        }
    } else {
        for (int r = 0; r < rows; ++r) one(tile[r][lane], MM != 0 ? my[r * ld] : 0.0f, r * ld);
    }
    st.a0 += a0;
    st.a1 += a1;
}

// fp32 -> nearest TF32 (low 13 mantissa bits zero), so that the tensor core's own truncation of the operand is exact
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// low part of x with respect to the TF32 the tensor core reads from the raw fp32 word (it truncates the 13 low mantissa
// bits): hi = trunc(x) is what the MMA sees of the untouched tile, lo = x - hi exactly, rounded to TF32 itself
__device__ __forceinline__ float tf32_lo(float x) { return tf32_rn(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)); }

// X3: error-compensated "3xTF32". Eight splitter warps (the four epilogue warps, idle during the main loop, plus four)
// compute lo = x - trunc_tf32(x) of every staged operand tile into a mirror ring; the raw tile is left untouched, the tensor
// core's own truncation of it IS hi. The MMA warp issues A.B into the main TMEM accumulator and A_lo.B + A.B_lo into a
// second one. The dropped lo.lo term is a same-signed ~2^-22 fraction of every product, i.e. a 2e-7 relative scaling of the
// result -- harmless; what does matter is the accumulate (see X3_CHAIN_KB below).
template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, bool X3>
__global__ void __launch_bounds__(X3 ? NUM_THREADS_X3 : NUM_THREADS)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Epilogue epi, int M,
                 int N, int K, int kb_per_split, float* __restrict__ ws, const ConvTma cv) {
    constexpr uint32_t A_BYTES = BLOCK_M * 128, B_BYTES = BLOCK_N * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t LO_OFFSET = STAGES * STAGE_BYTES;   // X3: the lo tiles mirror the raw ring behind it
    // power of two >= 32 (BLOCK_N in {32,64,128,256}); X3 keeps the two lo products in a second accumulator (see below)
    constexpr int TMEM_COLS = (BLOCK_N < 32 ? 32 : BLOCK_N) * (X3 ? 2 : 1);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], split_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ double scratch[32];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-byte alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BLOCK_N;
    const int kb_total = (K + BLOCK_K - 1) / BLOCK_K;
    const int kb_begin = blockIdx.z * kb_per_split;
    const int kb_end = min(kb_total, kb_begin + kb_per_split);
    const int num_kb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(smem_u32(&full_bar[s]), 1);
                mbar_init(smem_u32(&empty_bar[s]), 1);
                mbar_init(smem_u32(&split_bar[s]), 8);     // one arrival per splitter warp
            }
            mbar_init(smem_u32(&tmem_full_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t round = i / STAGES;
                mbar_wait(smem_u32(&empty_bar[s]), (round & 1) ^ 1);   // first pass over the ring returns immediately
                const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
                const uint32_t bar = smem_u32(&full_bar[s]);
                mbar_expect_tx(bar, STAGE_BYTES);
                const int k0 = (kb_begin + i) * BLOCK_K;
                if (cv.mode == 1 || cv.mode == 2) {
                    // A tile = 128 consecutive pixels (whole image rows / whole images), one tap, 32 channels
                    const int tap = k0 / cv.C, c0 = k0 - tap * cv.C;
                    const int ky = tap / cv.kw, kx = tap - ky * cv.kw;
                    const int HW = cv.H * cv.W;
                    const int img = m0 / HW, y0 = (m0 - img * HW) / cv.W;
                    const int dy = cv.mode == 1 ? ky - cv.pad_y : cv.pad_y - ky;
                    const int dx = cv.mode == 1 ? kx - cv.pad_x : cv.pad_x - kx;
                    tma_load_4d(sa, &tmA, bar, c0, dx, y0 + dy, img);
                    if constexpr (B_MN) {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 32; ++j) tma_load_3d(sb + j * SLAB_BYTES, &tmB, bar, n0 + j * 32, tap, c0);
                    } else {
                        tma_load_3d(sb, &tmB, bar, c0, tap, n0);
                    }
                } else if (cv.mode == 3) {
                    // A slab j = 32 (tap, c) rows x 32 consecutive pixels; rows past the last tap are fetched far out of bounds (zeros)
                    const int HW = cv.H * cv.W;
                    const int img = k0 / HW, rem = k0 - img * HW;
                    const int y0 = rem / cv.W, x0 = rem - y0 * cv.W;
                    if constexpr (A_MN) {
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 32; ++j) {
                            const int r = m0 + j * 32;
                            const int tap = r / cv.C, c0 = r - tap * cv.C;
                            const int ky = tap / cv.kw, kx = tap - ky * cv.kw;
                            const int yy = tap < cv.taps ? y0 + ky - cv.pad_y : -(1 << 20);
                            tma_load_4d(sa + j * SLAB_BYTES, &tmA, bar, c0, x0 + kx - cv.pad_x, yy, img);
                        }
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 32; ++j) tma_load_2d(sb + j * SLAB_BYTES, &tmB, bar, n0 + j * 32, k0);
                    }
                } else {
                    if constexpr (A_MN) {
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 32; ++j) tma_load_2d(sa + j * SLAB_BYTES, &tmA, bar, m0 + j * 32, k0);
                    } else {
                        tma_load_2d(sa, &tmA, bar, k0, m0);
                    }
                    if constexpr (B_MN) {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 32; ++j) tma_load_2d(sb + j * SLAB_BYTES, &tmB, bar, n0 + j * 32, k0);
                    } else {
                        tma_load_2d(sb, &tmB, bar, k0, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single elected thread) =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, A_MN, B_MN);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t round = i / STAGES;
                mbar_wait(smem_u32(X3 ? &split_bar[s] : &full_bar[s]), round & 1);
                tc_fence_after();
                const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
                for (int j = 0; j < BLOCK_K / UMMA_K; ++j) {
                    const uint64_t ad = A_MN ? desc_mnmajor(sa, j) : desc_kmajor(sa, j);
                    const uint64_t bd = B_MN ? desc_mnmajor(sb, j) : desc_kmajor(sb, j);
                    tc_mma_tf32(tmem_base, ad, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);
                    if (X3) {
                        const uint64_t adl = A_MN ? desc_mnmajor(sa + LO_OFFSET, j) : desc_kmajor(sa + LO_OFFSET, j);
                        const uint64_t bdl = B_MN ? desc_mnmajor(sb + LO_OFFSET, j) : desc_kmajor(sb + LO_OFFSET, j);
                        // the lo products go to their own accumulator: 2^-11 of the main one, so their truncating adds
                        // cost nothing, and the main accumulator sees one add per k-step instead of three
                        tc_mma_tf32(tmem_base + BLOCK_N, adl, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);   // A_lo . B_hi
                        tc_mma_tf32(tmem_base + BLOCK_N, ad, bdl, idesc, 1u);                           // A_hi . B_lo
                    }
                }
                tc_commit(smem_u32(&empty_bar[s]));          // frees the smem stage once these MMAs have read it
            }
            tc_commit(smem_u32(&tmem_full_bar));             // accumulator complete
        }
    }

    // ===== epilogue (warps 2..5; TMEM lane quarter = warp % 4) =====
    // tcgen05.ld hands thread t the 32 columns of accumulator ROW t; a 32x32 shared-memory transpose turns that into
    // lane = column, so every global access of the epilogue (output, bias, activation-backward mask) is a coalesced
    // 128-byte line instead of 32 rows x 4 bytes.
    __shared__ float xpose[4][32][33];
    EpiState st;
    epi_begin(epi, st);
    if (X3 && warp >= 2) {
        // ===== operand splitter (eight warps: the four epilogue warps, idle during the main loop, and four more) =====
        constexpr int SPLIT_THREADS = NUM_THREADS_X3 - 64;
        constexpr int PER = (STAGE_BYTES / 16 + SPLIT_THREADS - 1) / SPLIT_THREADS;   // float4 per thread and stage
        const int t = threadIdx.x - 64;
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % STAGES;
            const uint32_t round = i / STAGES;
            mbar_wait(smem_u32(&full_bar[s]), round & 1);                 // the TMA tiles of this stage have landed
            const uint32_t off = (smem_base - smem_u32(smem_raw)) + s * STAGE_BYTES;
            const float4* src = reinterpret_cast<const float4*>(smem_raw + off);
            float4* dst = reinterpret_cast<float4*>(smem_raw + off + LO_OFFSET);
            float4 x[PER];
#pragma unroll
            for (int u = 0; u < PER; ++u) {                               // all loads in flight first
                const uint32_t v = t + u * SPLIT_THREADS;
                if (v < STAGE_BYTES / 16) x[u] = src[v];
            }
#pragma unroll
            for (int u = 0; u < PER; ++u) {                               // elementwise: the swizzled layout carries over unchanged
                const uint32_t v = t + u * SPLIT_THREADS;
                if (v < STAGE_BYTES / 16)
                    dst[v] = make_float4(tf32_lo(x[u].x), tf32_lo(x[u].y), tf32_lo(x[u].z), tf32_lo(x[u].w));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&split_bar[s]));
        }
    }
    if (warp >= 2 && warp < 6) {
        const int q = warp & 3;
        float (*tile)[33] = xpose[warp - 2];
        if (num_kb > 0) {
            mbar_wait(smem_u32(&tmem_full_bar), 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
            if (n0 + c >= N) break;
            float v[32];
            if (num_kb > 0) {
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
                if (X3) {
                    float lo[32];
                    tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BLOCK_N + c), lo);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += lo[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 32; ++i) tile[lane][i] = v[i];   // row = lane
            __syncwarp();
            const int n = n0 + c + lane;
            const int mrow = m0 + q * 32;
            const int rows = min(32, M - mrow);
            if (n < N && rows > 0) {
                if (ws != nullptr) {                     // split-K partial: raw accumulator
                    float* p = ws + ((long long)blockIdx.z * M + mrow) * N + n;
                    if (rows == 32) {
#pragma unroll 1
                        for (int r0 = 0; r0 < 32; r0 += 8) {
                            float v8[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v8[j] = tile[r0 + j][lane];
#pragma unroll
                            for (int j = 0; j < 8; ++j) p[(long long)(r0 + j) * N] = v8[j];
                        }
                    } else {
                        for (int r = 0; r < rows; ++r) p[(long long)r * N] = tile[r][lane];
                    }
                } else if (epi.row_mode == 0 && epi.n_rows >= M) {
                    const float bias = epi.bias != nullptr ? epi.bias[n] : 0.0f;
                    float* o = epi.out + (long long)mrow * epi.ldo + n;
                    const float* my = epi.mask_y != nullptr ? epi.mask_y + (long long)mrow * epi.ldo + n : nullptr;
                    const int am = epi.act == B200NN_ACT_NONE ? 0 : (epi.act == B200NN_ACT_RELU ? 1 : 2);
                    const int mm = my == nullptr ? 0 : (epi.mask_act == B200NN_ACT_RELU ? 1 : 2);
                    switch (am * 3 + mm) {
                        case 0: tc_epi_rows<0, 0>(epi, st, tile, lane, rows, o, my, bias); break;
                        case 1: tc_epi_rows<0, 1>(epi, st, tile, lane, rows, o, my, bias); break;
                        case 2: tc_epi_rows<0, 2>(epi, st, tile, lane, rows, o, my, bias); break;
                        case 3: tc_epi_rows<1, 0>(epi, st, tile, lane, rows, o, my, bias); break;
                        case 4: tc_epi_rows<1, 1>(epi, st, tile, lane, rows, o, my, bias); break;
                        case 5: tc_epi_rows<1, 2>(epi, st, tile, lane, rows, o, my, bias); break;
                        case 6: tc_epi_rows<2, 0>(epi, st, tile, lane, rows, o, my, bias); break;
                        case 7: tc_epi_rows<2, 1>(epi, st, tile, lane, rows, o, my, bias); break;
                        default: tc_epi_rows<2, 2>(epi, st, tile, lane, rows, o, my, bias); break;
                    }
                } else {
                    for (int r = 0; r < rows; ++r) epi_store(epi, st, mrow + r, n, tile[r][lane]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (ws == nullptr) epi_finish(epi, st, scratch);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows of `ld` elements; box = {32, box_rows};
// K-major operands use SWIZZLE_128B, MN-major operands the 32-byte-atom variant (see the header comment)
// K-major fp16 operand planes: element (row, k) at base[row * ld + k]; box {64 k = 128 bytes, box_rows}, 128-byte swizzle
static int make_map_f16(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    B200NN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200NN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp16) failed with %d (inner=%lld outer=%lld ld=%lld)", (int)r, inner, outer, ld);
    return B200NN_OK;
}

static int make_map(CUtensorMap* map, const float* base, long long inner, long long outer, long long ld, int box_rows,
                    bool mn_major) {
    EncodeTiledFn fn = encode_fn();
    B200NN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200NN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (inner=%lld outer=%lld ld=%lld)", (int)r, inner, outer, ld);
    return B200NN_OK;
}

// rank-3 / rank-4 fp32 tensor map with explicit byte strides (dims[0] is contiguous) and box
static int make_map_nd(CUtensorMap* map, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                       const cuuint32_t* box, bool mn_major) {
    EncodeTiledFn fn = encode_fn();
    B200NN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200NN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (rank %d) failed with %d", rank, (int)r);
    return B200NN_OK;
}

// can this operand be described by a TMA map? (16-byte aligned base and row pitch)
static bool tma_ok(const float* base, long long ld) {
    return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * sizeof(float)) % 16 == 0 && ld > 0;
}

static int tc_block_n(int N) { return N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32)); }

struct TcPlan {
    int block_n, splits, kb_per_split;
};

// X3: the tensor core adds into the fp32 accumulator with truncation, a same-signed ~2^-25 relative error per MMA that
// grows with the chain (2e-5 at 1536 chained MMAs, measured) and would dominate the compensated product error. Two
// measures: the lo products accumulate in a second TMEM accumulator (a third of the adds on the main one), and the chain
// is cut every X3_CHAIN_KB k-blocks (2048 k = 256 adds on the main accumulator) through split-K, the reduce kernel adding
// the partials in round-to-nearest fp32 (the idea of Ootomo & Yokota's error-corrected tensor-core GEMM, at tile-chunk
// granularity). The extra CTAs also let small-M problems keep wide tiles.
constexpr int X3_CHAIN_KB = 64;
constexpr long long X3_WS_CAP = 512LL << 20;

static TcPlan tc_plan(long long M, long long N, long long K, bool x3 = false) {
    TcPlan p;
    p.block_n = tc_block_n((int)N);
    const long long kb_total = (K + BLOCK_K - 1) / BLOCK_K;
    auto tiles_for = [&](int bn) { return ((M + BLOCK_M - 1) / BLOCK_M) * ((N + bn - 1) / bn); };
    if (x3) {
        long long chain = X3_CHAIN_KB;
        long long splits = (kb_total + chain - 1) / chain;
        while (splits > 1 && splits * M * N * (long long)sizeof(float) > X3_WS_CAP) {   // bound the partial-sum workspace
            chain *= 2;
            splits = (kb_total + chain - 1) / chain;
        }
        const long long can = splits > kb_total / 16 ? splits : kb_total / 16;   // splits a long reduction could take
        while (p.block_n > 32 && tiles_for(p.block_n) < kNumSMs && tiles_for(p.block_n) * can < (3 * kNumSMs) / 4) p.block_n >>= 1;
        const long long tiles = tiles_for(p.block_n);
        if (tiles * splits < kNumSMs && kb_total >= 8) {          // short of a wave: split further, tiles kept wide
            long long more = kNumSMs / tiles;
            const long long maxs = kb_total / 4;
            if (more > maxs) more = maxs;
            if (more > splits) splits = more;
        }
        const long long per = (kb_total + splits - 1) / splits;
        p.splits = (int)((kb_total + per - 1) / per);
        p.kb_per_split = (int)per;
        return p;
    }
    // If the tile grid cannot fill the 148 SMs: a long reduction is split over K with the tiles kept wide (the MMA rate grows
    // with N; the weight-gradient GEMMs of a convolution have a handful of tiles and thousands of k-blocks), a short one gets
    // narrower tiles that spread the epilogue-heavy work over more SMs.
    while (p.block_n > 32 && tiles_for(p.block_n) < kNumSMs &&
           tiles_for(p.block_n) * (kb_total / 16) < (3 * kNumSMs) / 4) p.block_n >>= 1;
    const long long tiles = tiles_for(p.block_n);
    long long splits = 1;
    if (tiles < kNumSMs && kb_total >= 8) {
        splits = kNumSMs / tiles;                  // one wave: a second, nearly empty wave would double the time
        const long long maxs = kb_total / 4;       // at least 4 k-blocks (128 k) per split
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    long long per = (kb_total + splits - 1) / splits;
    splits = (kb_total + per - 1) / per;           // every split non-empty
    p.splits = (int)splits;
    p.kb_per_split = (int)per;
    return p;
}

static size_t tc_ws_bytes(long long M, long long N, long long K, bool x3 = false) {
    const TcPlan p = tc_plan(M, N, K, x3);
    return p.splits > 1 ? (size_t)p.splits * M * N * sizeof(float) : 0;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, bool X3>
static int tc_launch_stages(const CUtensorMap& ta, const CUtensorMap& tb, const Epilogue& epi, int M, int N, int K,
                            const TcPlan& p, float* ws, cudaStream_t st, const ConvTma& cv) {
    constexpr size_t smem = (size_t)STAGES * (BLOCK_M * 128 + BN * 128) * (X3 ? 2 : 1) + 1024;
    auto kern = gemm_tf32_kernel<BN, STAGES, A_MN, B_MN, X3>;
    static bool configured[64] = {};   // per device (one template instance per kernel variant)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        B200NN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63] = true;
    }
    const dim3 grid((N + BN - 1) / BN, (M + BLOCK_M - 1) / BLOCK_M, p.splits);
    kern<<<grid, X3 ? NUM_THREADS_X3 : NUM_THREADS, smem, st>>>(ta, tb, epi, M, N, K, p.kb_per_split, p.splits > 1 ? ws : nullptr, cv);
    B200NN_LAUNCH_CHECK("gemm_tf32");
    if (p.splits > 1) return launch_splitk_reduce(ws, p.splits, M, N, epi, st);
    return B200NN_OK;
}

// short reductions (<= 2 k-blocks per CTA) are epilogue-bound: a 2-stage ring keeps the shared-memory footprint small
// so that several CTAs (and their epilogue warps) share an SM; long reductions get a deep ring
template <int BN, bool A_MN, bool B_MN>
static int tc_launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const Epilogue& epi, int M, int N, int K,
                             const TcPlan& p, float* ws, cudaStream_t st, bool x3, const ConvTma& cv) {
    if (x3)   // twice the shared memory per stage (raw + lo tiles): 2 stages at 256 columns, 3 below
        return tc_launch_stages<BN, (BN >= 256 ? 2 : 3), A_MN, B_MN, true>(ta, tb, epi, M, N, K, p, ws, st, cv);
    if (p.kb_per_split <= 2) return tc_launch_stages<BN, 2, A_MN, B_MN, false>(ta, tb, epi, M, N, K, p, ws, st, cv);
    return tc_launch_stages<BN, (BN >= 256 ? 4 : 6), A_MN, B_MN, false>(ta, tb, epi, M, N, K, p, ws, st, cv);
}

static int tc_dispatch(const CUtensorMap& ta, const CUtensorMap& tb, bool a_mn, bool b_mn, const Epilogue& epi, int M, int N, int K,
                       const TcPlan& p, float* ws, cudaStream_t st, bool x3, const ConvTma& cv) {
    B200NN_REQUIRE(!(a_mn && !b_mn), "gemm_tf32: (A MN-major, B K-major) is not instantiated");
#define TC_DISPATCH(BN)                                                                                  \
    if (a_mn) return tc_launch_variant<BN, true, true>(ta, tb, epi, M, N, K, p, ws, st, x3, cv);         \
    if (b_mn) return tc_launch_variant<BN, false, true>(ta, tb, epi, M, N, K, p, ws, st, x3, cv);        \
    return tc_launch_variant<BN, false, false>(ta, tb, epi, M, N, K, p, ws, st, x3, cv);
    switch (p.block_n) {
        case 32: TC_DISPATCH(32)
        case 64: TC_DISPATCH(64)
        case 128: TC_DISPATCH(128)
        default: TC_DISPATCH(256)
    }
#undef TC_DISPATCH
}

// D[M,N] = A.B ; a_mn: A stored [K][M] (ld = lda) else [M][K]; b_mn: B stored [K][N] else [N][K]
static int tc_gemm(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi,
                   int M, int N, int K, float* ws, cudaStream_t st, bool x3 = false) {
    const TcPlan p = tc_plan(M, N, K, x3);
    B200NN_REQUIRE(p.splits == 1 || ws != nullptr, "gemm_tf32: split-K needs a workspace");
    CUtensorMap ta, tb;
    if (a_mn) { if (int rc = make_map(&ta, A, M, K, lda, BLOCK_K, true)) return rc; }
    else      { if (int rc = make_map(&ta, A, K, M, lda, BLOCK_M, false)) return rc; }
    if (b_mn) { if (int rc = make_map(&tb, B, N, K, ldb, BLOCK_K, true)) return rc; }
    else      { if (int rc = make_map(&tb, B, K, N, ldb, p.block_n, false)) return rc; }
    ConvTma cv{};
    return tc_dispatch(ta, tb, a_mn, b_mn, epi, M, N, K, p, ws, st, x3, cv);
}

// ---- implicit-GEMM convolution on the same kernel (see ConvTma) ------------------------------------------------
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// stride 1, 'same' with an odd kernel, power-of-two image (W in 4..128), channel counts multiples of 32
static bool conv_tc_ok(const b200nn_conv_geom* g, const void* p0, const void* p1, const void* p2) {
    if (((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2)) & 15) != 0) return false;
    return g->sh == 1 && g->sw == 1 && g->OH == g->H && g->OW == g->W && (g->kh & 1) && (g->kw & 1) &&
           g->ph == (g->kh - 1) / 2 && g->pw == (g->kw - 1) / 2 && g->C % 32 == 0 && g->F % 32 == 0 && pow2(g->W) &&
           pow2(g->H) && g->W >= 4 && g->W <= 128 && g->H <= 256 &&
           (long long)g->N * g->H * g->W * g->F * g->kh * g->kw * g->C >= (1LL << 22);
}

// 4-D map over an NHWC tensor t[Nimg, H, W, Cn]; the box covers `pixels` consecutive pixels (whole rows / whole images)
static int make_image_map(CUtensorMap* map, const float* t, int Nimg, int H, int W, int Cn, int pixels, bool mn_major) {
    const cuuint64_t dims[4] = {(cuuint64_t)Cn, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
    const cuuint64_t strides[3] = {(cuuint64_t)Cn * 4, (cuuint64_t)W * Cn * 4, (cuuint64_t)H * W * Cn * 4};
    const int bx = W < pixels ? W : pixels;
    const int by = H < pixels / bx ? H : pixels / bx;
    const int bb = pixels / (bx * by);
    const cuuint32_t box[4] = {32, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bb};
    return make_map_nd(map, t, 4, dims, strides, box, mn_major);
}

// the same over fp16 operand planes t[Nimg, H, W, Cn] (gemm_tc2.cuh F16 mode): 64 channels = 128 bytes per box row
static int make_image_map_f16(CUtensorMap* map, const void* t, int Nimg, int H, int W, int Cn, int pixels) {
    EncodeTiledFn fn = encode_fn();
    B200NN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[4] = {(cuuint64_t)Cn, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
    const cuuint64_t strides[3] = {(cuuint64_t)Cn * 2, (cuuint64_t)W * Cn * 2, (cuuint64_t)H * W * Cn * 2};
    const int bx = W < pixels ? W : pixels;
    const int by = H < pixels / bx ? H : pixels / bx;
    const int bb = pixels / (bx * by);
    const cuuint32_t box[4] = {64, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bb};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(t), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200NN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp16 images) failed with %d", (int)r);
    return B200NN_OK;
}

// 3-D map over the weights w[(c, tap), F]: (f, tap, c)
static int make_weight_map(CUtensorMap* map, const float* w, int taps, int C, int F, int box_c, bool mn_major) {
    const cuuint64_t dims[3] = {(cuuint64_t)F, (cuuint64_t)taps, (cuuint64_t)C};
    const cuuint64_t strides[2] = {(cuuint64_t)F * 4, (cuuint64_t)taps * F * 4};
    const cuuint32_t box[3] = {32, 1, (cuuint32_t)box_c};
    return make_map_nd(map, w, 3, dims, strides, box, mn_major);
}

static ConvTma conv_tma(int mode, const b200nn_conv_geom* g, int C) {
    ConvTma cv;
    cv.mode = mode; cv.C = C; cv.kw = g->kw; cv.taps = g->kh * g->kw; cv.W = g->W; cv.H = g->H; cv.pad_y = g->ph; cv.pad_x = g->pw;
    return cv;
}

// y[pixels, F] = patches(x) . W   (epi: bias + activation)
static int tc_conv_fwd(const b200nn_conv_geom* g, const float* x, const float* w, const Epilogue& epi, float* ws,
                       cudaStream_t st, bool x3) {
    const int M = g->N * g->H * g->W, N = g->F, K = g->kh * g->kw * g->C;
    const TcPlan p = tc_plan(M, N, K, x3);
    B200NN_REQUIRE(p.splits == 1 || ws != nullptr, "conv_tf32: split-K needs a workspace");
    CUtensorMap ta, tb;
    if (int rc = make_image_map(&ta, x, g->N, g->H, g->W, g->C, BLOCK_M, false)) return rc;
    if (int rc = make_weight_map(&tb, w, g->kh * g->kw, g->C, g->F, 32, true)) return rc;
    return tc_dispatch(ta, tb, false, true, epi, M, N, K, p, ws, st, x3, conv_tma(1, g, g->C));
}

// dx[pixels, C] = patches'(err) . W^T   (mirrored taps; epi: clip multiplier, activation-backward mask, sums of squares)
static int tc_conv_dgrad(const b200nn_conv_geom* g, const float* err, const float* w, const Epilogue& epi, float* ws,
                         cudaStream_t st, bool x3) {
    const int M = g->N * g->H * g->W, N = g->C, K = g->kh * g->kw * g->F;
    const TcPlan p = tc_plan(M, N, K, x3);
    B200NN_REQUIRE(p.splits == 1 || ws != nullptr, "conv_tf32: split-K needs a workspace");
    CUtensorMap ta, tb;
    if (int rc = make_image_map(&ta, err, g->N, g->H, g->W, g->F, BLOCK_M, false)) return rc;
    if (int rc = make_weight_map(&tb, w, g->kh * g->kw, g->C, g->F, p.block_n, false)) return rc;
    return tc_dispatch(ta, tb, false, false, epi, M, N, K, p, ws, st, x3, conv_tma(2, g, g->F));
}

// dw[(tap, c), F] = patches(x)^T . err   (epi: row_mode 1 writes the reference's (c, ky, kx) row order)
static int tc_conv_wgrad(const b200nn_conv_geom* g, const float* x, const float* err, const Epilogue& epi, float* ws,
                         cudaStream_t st, bool x3) {
    const int M = g->kh * g->kw * g->C, N = g->F, K = g->N * g->H * g->W;
    const TcPlan p = tc_plan(M, N, K, x3);
    B200NN_REQUIRE(p.splits == 1 || ws != nullptr, "conv_tf32: split-K needs a workspace");
    CUtensorMap ta, tb;
    if (int rc = make_image_map(&ta, x, g->N, g->H, g->W, g->C, BLOCK_K, true)) return rc;
    if (int rc = make_map(&tb, err, N, K, N, BLOCK_K, true)) return rc;
    return tc_dispatch(ta, tb, true, true, epi, M, N, K, p, ws, st, x3, conv_tma(3, g, g->C));
}

static size_t tc_conv_ws_bytes(const b200nn_conv_geom* g, bool x3) {
    const long long P = (long long)g->N * g->H * g->W, taps = (long long)g->kh * g->kw;
    size_t r = tc_ws_bytes(P, g->F, taps * g->C, x3);
    const size_t b = tc_ws_bytes(P, g->C, taps * g->F, x3), c = tc_ws_bytes(taps * g->C, g->F, P, x3);
    r = r > b ? r : b;
    return r > c ? r : c;
}

}  // namespace tc
}  // namespace b200nn
