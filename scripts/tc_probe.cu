// standalone probe for the tcgen05 / TMA building blocks (debug aid; not part of the library)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include <math.h>
#define B200NN_REQUIRE(c, ...) do { if (!(c)) { printf(__VA_ARGS__); printf("\n"); return -1; } } while (0)
#define B200NN_OK 0
namespace b200nn { struct Epilogue {}; }
// minimal copies of the helpers under test
namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout = 2) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | (layout << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
}
using namespace tc;
constexpr int BM = 128, BN = 64, BK = 32;

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* dumpA, float* dumpB, float* dumpD, int b_mn) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar, done_bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sa = base, sb = base + BM * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&full_bar), 1); mbar_init(smem_u32(&done_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&full_bar), BM * 128 + BN * 128);
        tma_load_2d(sa, &tmA, smem_u32(&full_bar), 0, 0);
        if (b_mn) { tma_load_2d(sb, &tmB, smem_u32(&full_bar), 0, 0); tma_load_2d(sb + 4096, &tmB, smem_u32(&full_bar), 32, 0); }
        else tma_load_2d(sb, &tmB, smem_u32(&full_bar), 0, 0);
    }
    mbar_wait(smem_u32(&full_bar), 0);
    for (int i = threadIdx.x; i < BM * 32; i += 128) dumpA[i] = reinterpret_cast<float*>(gen)[i];
    for (int i = threadIdx.x; i < BN * 32; i += 128) dumpB[i] = reinterpret_cast<float*>(gen + BM * 128)[i];
    __syncthreads();
    if (threadIdx.x == 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc(BM, BN, false, b_mn != 0);
        for (int j = 0; j < 4; ++j) {
            const uint64_t ad = make_desc(sa + j * 32, 16, 1024);
            const uint64_t bd = b_mn ? make_desc(sb + j * 1024, 4096, 512, 1) : make_desc(sb + j * 32, 16, 1024);
            const uint32_t acc = j > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    }
    mbar_wait(smem_u32(&done_bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
        for (int i = 0; i < 32; ++i) dumpD[(warp * 32 + lane) * BN + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
}

static int make_map(CUtensorMap* map, const float* base, long long inner, long long outer, long long ld, int box_rows, bool mn = false) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUresult r = ((Fn)p)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    return (int)r;
}

int main() {
    std::vector<float> A(BM * BK), Bm(BK * BN), Bk(BN * BK);
    for (int m = 0; m < BM; ++m) for (int k = 0; k < BK; ++k) A[m * BK + k] = (float)(m * 100 + k);   // A[m][k]
    for (int k = 0; k < BK; ++k) for (int n = 0; n < BN; ++n) { Bm[k * BN + n] = (k == (n % 32)) ? 1.f : 0.f; Bk[n * BK + k] = Bm[k * BN + n]; }
    float *dA, *dB, *dumpA, *dumpB, *dumpD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bm.size() * 4); cudaMalloc(&dumpA, BM * 32 * 4); cudaMalloc(&dumpB, BN * 32 * 4); cudaMalloc(&dumpD, BM * BN * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    for (int b_mn = 0; b_mn < 2; ++b_mn) {
        cudaMemcpy(dB, b_mn ? Bm.data() : Bk.data(), Bm.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dumpD, 0xff, BM * BN * 4);
        CUtensorMap ta, tb;
        make_map(&ta, dA, BK, BM, BK, BM);
        if (b_mn) make_map(&tb, dB, BN, BK, BN, BK, true); else make_map(&tb, dB, BK, BN, BK, BN);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        probe<<<1, 128, 40 * 1024>>>(ta, tb, dumpA, dumpB, dumpD, b_mn);
        cudaError_t e = cudaDeviceSynchronize();
        printf("b_mn=%d sync: %s\n", b_mn, cudaGetErrorString(e));
        std::vector<float> hA(BM * 32), hB(BN * 32), hD(BM * BN);
        cudaMemcpy(hA.data(), dumpA, hA.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hB.data(), dumpB, hB.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hD.data(), dumpD, hD.size() * 4, cudaMemcpyDeviceToHost);
        printf("smem A row0: "); for (int i = 0; i < 12; ++i) printf("%g ", hA[i]); printf("\nsmem A row1: "); for (int i = 0; i < 12; ++i) printf("%g ", hA[32 + i]);
        printf("\nsmem B first: "); for (int i = 0; i < 12; ++i) printf("%g ", hB[i]);
        // expected D[m][n] = A[m][n % 32] = m*100 + n%32
        double maxerr = 0; for (int m = 0; m < BM; ++m) for (int n = 0; n < BN; ++n) maxerr = fmax(maxerr, fabs(hD[m * BN + n] - (m * 100 + n % 32)));
        printf("\nD row0: "); for (int i = 0; i < 10; ++i) printf("%g ", hD[i]); printf("\nD row3: "); for (int i = 0; i < 10; ++i) printf("%g ", hD[3 * BN + i]);
        printf("\nD row127 tail: "); for (int i = 54; i < 64; ++i) printf("%g ", hD[127 * BN + i]);
        printf("\nmax err %g\n", maxerr);
    }
    return 0;
}
