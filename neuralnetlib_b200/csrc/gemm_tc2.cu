// gemm_tc2.cu — translation unit of the persistent CTA-pair tcgen05 contraction (gemm_tc2.cuh), single-pass TF32.
#include "gemm_tc_api.cuh"
#include "gemm_tcgen05.cuh"   // PTX wrappers, descriptors and tensor-map builders shared with the first generation
#include "gemm_tc2.cuh"

namespace b200nn {
namespace tcapi {

bool tc2_enabled() { return tc2::enabled(); }
int gemm2_plain(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi, int M,
                int N, int K, float* ws, cudaStream_t st) {
    return tc2::gemm2<false>(A, lda, a_mn, B, ldb, b_mn, epi, M, N, K, ws, st);
}
int conv2_plain(int which, const b200nn_conv_geom* g, const float* a, const float* b, const Epilogue& epi, float* ws, cudaStream_t st) {
    if (which == 1) return tc2::conv_fwd2<false>(g, a, b, epi, ws, st);
    if (which == 2) return tc2::conv_dgrad2<false>(g, a, b, epi, ws, st);
    return tc2::conv_wgrad2<false>(g, a, b, epi, ws, st);
}
size_t ws2(long long M, long long N, long long K, bool x3) { return tc2::ws_bytes2(M, N, K, x3); }
size_t conv_ws2(const b200nn_conv_geom* g, bool x3) { return tc2::conv_ws_bytes2(g, x3); }

}  // namespace tcapi
}  // namespace b200nn
