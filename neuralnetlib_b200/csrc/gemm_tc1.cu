// gemm_tc1.cu — translation unit of the first-generation tcgen05 contraction (gemm_tcgen05.cuh): one tile per CTA,
// cta_group::1. Kept as the B200NN_TC2=0 fallback and as the A/B baseline of the persistent CTA-pair kernel.
#include "gemm_tc_api.cuh"
#include "gemm_tcgen05.cuh"

namespace b200nn {
namespace tcapi {

bool tma_ok(const float* base, long long ld) { return tc::tma_ok(base, ld); }
bool conv_ok(const b200nn_conv_geom* g, const void* p0, const void* p1, const void* p2) { return tc::conv_tc_ok(g, p0, p1, p2); }

int gemm1(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi, int M, int N,
          int K, float* ws, cudaStream_t st, bool x3) {
    return tc::tc_gemm(A, lda, a_mn, B, ldb, b_mn, epi, M, N, K, ws, st, x3);
}
int conv1(int which, const b200nn_conv_geom* g, const float* a, const float* b, const Epilogue& epi, float* ws, cudaStream_t st, bool x3) {
    if (which == 1) return tc::tc_conv_fwd(g, a, b, epi, ws, st, x3);
    if (which == 2) return tc::tc_conv_dgrad(g, a, b, epi, ws, st, x3);
    return tc::tc_conv_wgrad(g, a, b, epi, ws, st, x3);
}
size_t ws1(long long M, long long N, long long K, bool x3) { return tc::tc_ws_bytes(M, N, K, x3); }
size_t conv_ws1(const b200nn_conv_geom* g, bool x3) { return tc::tc_conv_ws_bytes(g, x3); }

}  // namespace tcapi
}  // namespace b200nn
