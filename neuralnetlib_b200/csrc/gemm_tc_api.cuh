// gemm_tc_api.cuh — host entry points of the tcgen05 contraction kernels, each generation / math mode in its own translation
// unit (gemm_tc1.cu: first-generation one-tile-per-CTA kernel; gemm_tc2.cu / gemm_tc2_x3.cu: persistent CTA-pair kernel).
#pragma once
#include "epilogue.cuh"

namespace b200nn {
namespace tcapi {

bool tc2_enabled();   // B200NN_TC2=0 selects the first generation everywhere
// can this operand be described by a TMA map? (16-byte aligned base and row pitch)
bool tma_ok(const float* base, long long ld);
// implicit-GEMM convolution eligibility (stride 1, 'same', odd kernel, power-of-two image, channel counts multiples of 32)
bool conv_ok(const b200nn_conv_geom* g, const void* p0, const void* p1, const void* p2);

// D[M,N] = A.B ; a_mn: A stored [K][M] (ld = lda) else [M][K]; b_mn: B stored [K][N] else [N][K]
int gemm1(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi, int M, int N,
          int K, float* ws, cudaStream_t st, bool x3);
int gemm2_plain(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi, int M,
                int N, int K, float* ws, cudaStream_t st);
int gemm2_x3(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi, int M, int N,
             int K, float* ws, cudaStream_t st);
// TF32X3 accuracy at twice the rate for LARGE contractions: operands pre-split into fp16 hi / lo planes (f16_split.cu), three
// kind::f16 MMAs per k-step (gemm_tc2.cuh F16 mode). ws: ws2(M, N, ceil(K / 2), true) + f16_extra_bytes(M, N, K) bytes.
bool f16_enabled();
bool f16_worth(long long M, long long N, long long K);
size_t f16_extra_bytes(long long M, long long N, long long K);
// premax_a / premax_b (optional): device words already holding the bit pattern of max |A| / max |B| over finite elements (the max pass
// of that operand is skipped)
int gemm2_f16x3(const float* A, long long lda, bool a_mn, const float* B, long long ldb, bool b_mn, const Epilogue& epi, int M, int N,
                int K, float* ws, cudaStream_t st, const unsigned* premax_a = nullptr, const unsigned* premax_b = nullptr);
int gemm2_f16_planes(const void* Ap, long long lda, int lo_a, const void* Bp, long long ldb, int lo_b, const float* inv_scale,
                     const Epilogue& epi, int M, int N, int K, float* ws, cudaStream_t st);
// the same for the implicit-GEMM convolution (forward and data gradient with 64 | channels along K): the image operand is split
// into fp16 hi / lo IMAGES, the weights into a K-major plane matrix in the kernel's (tap, channel) order.
// ws: ws2(P, Nn, K / 2, true) + conv_f16_extra_bytes(g) bytes.
bool conv_f16_worth(int which, const b200nn_conv_geom* g);
size_t conv_f16_extra_bytes(const b200nn_conv_geom* g);
int conv2_f16x3(int which, const b200nn_conv_geom* g, const float* a, const float* w, const Epilogue& epi, float* ws, cudaStream_t st,
                const unsigned* premax_a = nullptr);
int conv2_f16_planes(int which, const b200nn_conv_geom* g, const void* Ap, const void* Bp, long long ldb, const float* inv_scale,
                     const Epilogue& epi, float* ws, cudaStream_t st);
// which: 1 forward (a = x, b = w), 2 data gradient (a = err, b = w), 3 weight gradient (a = x, b = err)
int conv1(int which, const b200nn_conv_geom* g, const float* a, const float* b, const Epilogue& epi, float* ws, cudaStream_t st, bool x3);
int conv2_plain(int which, const b200nn_conv_geom* g, const float* a, const float* b, const Epilogue& epi, float* ws, cudaStream_t st);
int conv2_x3(int which, const b200nn_conv_geom* g, const float* a, const float* b, const Epilogue& epi, float* ws, cudaStream_t st);
// split-K workspace bytes
size_t ws1(long long M, long long N, long long K, bool x3 = false);
size_t ws2(long long M, long long N, long long K, bool x3 = false);
size_t conv_ws1(const b200nn_conv_geom* g, bool x3);
size_t conv_ws2(const b200nn_conv_geom* g, bool x3);

}  // namespace tcapi
}  // namespace b200nn
