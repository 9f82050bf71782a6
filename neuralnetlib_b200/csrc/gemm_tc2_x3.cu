// gemm_tc2_x3.cu — translation unit of the persistent CTA-pair tcgen05 contraction (gemm_tc2.cuh), error-compensated 3xTF32.
#include "gemm_tc_api.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tc2.cuh"

namespace b200nn {
namespace tcapi {

int gemm2_x3(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi, int M, int N,
             int K, float* ws, cudaStream_t st) {
    return tc2::gemm2<true>(A, lda, a_mn, B, ldb, b_mn, epi, M, N, K, ws, st);
}
int conv2_x3(int which, const b200nn_conv_geom* g, const float* a, const float* b, const Epilogue& epi, float* ws, cudaStream_t st) {
    if (which == 1) return tc2::conv_fwd2<true>(g, a, b, epi, ws, st);
    if (which == 2) return tc2::conv_dgrad2<true>(g, a, b, epi, ws, st);
    return tc2::conv_wgrad2<true>(g, a, b, epi, ws, st);
}

}  // namespace tcapi
}  // namespace b200nn
