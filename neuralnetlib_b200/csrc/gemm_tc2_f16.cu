// gemm_tc2_f16.cu — translation unit of the persistent CTA-pair tcgen05 contraction (gemm_tc2.cuh) in its fp16-plane mode.
#include "gemm_tc_api.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tc2.cuh"

namespace b200nn {
namespace tcapi {

int gemm2_f16_planes(const void* Ap, long long lda, int lo_a, const void* Bp, long long ldb, int lo_b, const float* inv_scale,
                     const Epilogue& epi, int M, int N, int K, float* ws, cudaStream_t st) {
    return tc2::gemm2_f16(Ap, lda, lo_a, Bp, ldb, lo_b, inv_scale, epi, M, N, K, ws, st);
}

int conv2_f16_planes(int which, const b200nn_conv_geom* g, const void* Ap, const void* Bp, long long ldb, const float* inv_scale,
                     const Epilogue& epi, float* ws, cudaStream_t st) {
    return tc2::conv2_f16(which, g, Ap, Bp, ldb, inv_scale, epi, ws, st);
}

}  // namespace tcapi
}  // namespace b200nn
