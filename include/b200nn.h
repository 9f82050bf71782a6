/*
 * b200nn.h — C ABI of libb200nn.so: the B200 (sm_100a) numeric core behind neuralnetlib's
 * Python Layer / Optimizer / Sequential API.
 *
 * The upstream reference (marcpinet/neuralnetlib, pure NumPy) has no FFI; each entry point below
 * replaces a NumPy code region of the reference's `Sequential.fit` hot path and cites it
 * (paths relative to /root/reference/neuralnetlib/). INTEGRATION.md shows the ctypes binding a
 * maintainer would add on the reference side.
 *
 * Conventions
 *  - plain pointers and sizes only; every tensor is float32, row-major, images are NHWC;
 *  - the C side never allocates or frees tensors: the caller owns every device buffer;
 *  - every function is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant,
 *    capturable into a CUDA graph, and returns 0 on success or a negative B200NN_E* code;
 *    the message of the last failure on the calling thread is available from b200nn_last_error;
 *  - "lazy clip": the reference clips every error tensor to L2 norm <= 1 before every layer's
 *    backward (models.py:214, preprocessing.py:220-224). All backward ops are linear in the
 *    error, so producers accumulate the sum of squares of what they emit into a float64 slot
 *    (`ss_out`, one atomicAdd per CTA) and consumers turn a *chain* of such slots into one
 *    multiplier applied on load.  A chain is (`ss`, `chain`): `ss` points at the slot array,
 *    `chain` packs n (bits 0-7) and up to four 12-bit slot indices (bits 8.., 20.., 32.., 44..).
 *        acc = 1; for i < n: n2 = acc*acc*ss[slot_i]; if (n2 > 1) acc *= rsqrt(n2);
 *    n == 0 (or ss == NULL) means "no clip". Slots must be zeroed before the producer runs
 *    (b200nn_step_begin, or b200nn_zero_f64).
 *    Range form (bit 63 set; lo in bits 0-15, count in bits 16-31): slots [lo, lo+count) in order. Used when the clips are
 *    DEFERRED (data parallelism): backward kernels then run without any multiplier, record the raw sums of squares, and the
 *    whole pending chain is applied once to each weight gradient -- so the slots need a single exchange per step.
 */
#ifndef B200NN_H_
#define B200NN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NN_OK 0
#define B200NN_EINVAL (-1)   /* bad argument / unsupported geometry */
#define B200NN_ECUDA (-2)    /* CUDA runtime error (launch, capture, ...) */
#define B200NN_ENOTSUP (-3)  /* valid request that this build cannot serve */

/* activation kinds (activations.py:33-109) */
#define B200NN_ACT_NONE 0      /* Linear                     activations.py:81-89  */
#define B200NN_ACT_RELU 1      /* max(0,x); f' = x>0         activations.py:46-54  */
#define B200NN_ACT_SIGMOID 2   /* 1/(1+exp(-clip(x,+-50)))   activations.py:33-43  */
#define B200NN_ACT_TANH 3      /*                            activations.py:57-65  */
#define B200NN_ACT_LEAKYRELU 4 /* max(x, a*x); f' = x>0?1:a  activations.py:92-109 */
#define B200NN_ACT_ELU 5       /* x>0 ? x : exp(x)-1         activations.py:112-123 */
#define B200NN_ACT_SELU 6      /* s*(x>0 ? x : a(exp(x)-1)), a = alpha, s = 1.0507009873554805   activations.py:126-147 */
#define B200NN_ACT_MAX_FUSED 6 /* kinds 0..6 are strictly monotone or piecewise so that f' is a function of the OUTPUT: they fuse into
                                  GEMM / BatchNorm epilogues; GELU (7) needs the pre-activation and runs through b200nn_act_fwd / _bwd only */
#define B200NN_ACT_GELU 7      /* tanh approximation         activations.py:150-166; b200nn_act_bwd takes the INPUT x in place of y */

/* loss kinds (losses.py:79-120) */
#define B200NN_LOSS_MSE 0
#define B200NN_LOSS_BCE 1
#define B200NN_LOSS_CCE 2
#define B200NN_LOSS_MAE 3   /* losses.py:146-155 MeanAbsoluteError */
#define B200NN_LOSS_HUBER 4 /* losses.py:158-180 Huber(delta) */

/* GEMM math modes */
#define B200NN_MATH_FP32 0 /* FFMA, fp32 accumulate: parity path (rel 1e-5)                    */
#define B200NN_MATH_TF32 1 /* tcgen05 kind::tf32, fp32 accumulate in TMEM: perf path (rel 2e-3 per contraction) */
#define B200NN_MATH_TF32X3 2 /* tcgen05, error-compensated 3xTF32 (hi/lo operand split in shared memory): fp32 accuracy on tensor cores */

int b200nn_version(void);
/* copies the calling thread's last error message (NUL terminated) into buf; returns its length */
size_t b200nn_last_error(char* buf, size_t cap);
/* 1 when tcgen05 GEMM kernels are compiled in */
int b200nn_has_tcgen05(void);
/* number of kernels this library has launched (or recorded into a graph being captured) in this process */
long long b200nn_launch_count(void);

/* ---- step bookkeeping --------------------------------------------------------------------- */
/* Zero `n_slots` float64 sum-of-squares slots and `n_acc` float64 accumulators (loss, counters)
 * and add `inc` to the device counter *counter (int64, nullable). First node of a captured train
 * step: the counter is Adam's `t` (optimizers.py:236), advanced on the device by the number of
 * layers the step will update so that a replayed graph needs no host-side argument. */
int b200nn_step_begin(double* ss, int n_slots, double* acc, int n_acc, long long* counter, long long inc,
                      void* stream);
/* dst[r, :] = src[idx[r], :] for r < rows (row_elems floats per row): the shuffled mini-batch gather of
 * Sequential.fit (models.py:386-402, utils.py:34-50) done on the device against a resident dataset. */
int b200nn_gather_rows(const float* src, const long long* idx, float* dst, long long rows, long long row_elems,
                       void* stream);
int b200nn_zero_f64(double* p, int n, void* stream);

/* ---- host <-> device plumbing of the public entry points (models.py:251-264: the batch goes in, the loss comes out) ---- */
int b200nn_upload(void* dst_dev, const void* src_host, size_t bytes, void* stream);          /* async H2D on `stream` */
int b200nn_read_sync(void* dst_host, const void* src_dev, size_t bytes, void* stream);      /* D2H + stream sync */

/* ---- CUDA graph helpers (capture every launch made on `stream` between begin and end) ------ */
int b200nn_graph_begin(void* stream);
int b200nn_graph_end(void* stream, void** graph_exec_out);
int b200nn_graph_launch(void* graph_exec, void* stream);
int b200nn_graph_destroy(void* graph_exec);

/* ---- elementwise -------------------------------------------------------------------------- */
/* Activation.forward_pass (layers.py:231-234): y = f(x). */
int b200nn_act_fwd(int act, float alpha, const float* x, float* y, long long n, void* stream);
/* Activation.backward_pass (layers.py:236-237) with the preceding clip (models.py:214) folded in:
 * dx = chain(err) * f'(.)  where f' is evaluated from the cached OUTPUT y (equivalent for every
 * activation above). ss_out (nullable) += sum(dx^2). */
int b200nn_act_bwd(int act, float alpha, const float* y, const float* err, const double* ss, uint64_t chain,
                   float* dx, double* ss_out, long long n, void* stream);
/* clip_gradients (preprocessing.py:220-224) materialised: out = chain(in). */
int b200nn_scale_apply(const float* in, const double* ss, uint64_t chain, float* out, long long n, void* stream);
/* ss_out += sum(x^2) */
int b200nn_sumsq(const float* x, long long n, double* ss_out, void* stream);
/* Softmax over axis 1 (activations.py:69-71). */
int b200nn_softmax_fwd(const float* x, float* p, int rows, int cols, void* stream);
/* Loss value + error for the last layer (models.py:255-256, 198-210):
 *   loss_acc[0] += the loss numerator (host divides: CCE by rows, BCE/MSE by rows*cols);
 *   fused_pair != 0 (Softmax+CCE or Sigmoid+BCE): err = p - y (models.py:200-205, not /B);
 *   fused_pair == 0: err = LossFunction.derivative(y, p) (losses.py:83-84, 96-99, 111-115);
 *   loss_acc[1] += number of rows whose argmax(p) == argmax(y) (cols == 1: (p >= 0.5) == y)
 *   (metrics.py:84-89, first-index tie break);  ss_out (nullable) += sum(err^2). */
int b200nn_loss_fwd_bwd(int loss, int fused_pair, const float* p, const float* y, float* err, double* loss_acc,
                        double* ss_out, int rows, int cols, void* stream);
/* The same with (a) `global_rows`: the rows of the WHOLE batch when this call sees one data-parallel shard of it —
 * MeanSquaredError.derivative divides by y_true.size of the whole batch (losses.py:84); (b) `param`: Huber's delta
 * (losses.py:158-176; MAE :146-155). SparseCategoricalCrossentropy (losses.py:123-143) is CCE on host-built one-hot rows. */
int b200nn_loss_fwd_bwd_ex(int loss, int fused_pair, const float* p, const float* y, float* err, double* loss_acc,
                           double* ss_out, int rows, int cols, long long global_rows, float param, void* stream);

/* ---- Dense (layers.py:152-198) ------------------------------------------------------------- */
/* y[M,N] = act(x[M,K] . w[K,N] + b[N])   layers.py:173 (+ the Activation Sequential.add appends).
 * `ws` is a float workspace of at least b200nn_dense_ws_bytes(M,N,K) bytes (split-K partials; the same
 * size serves the forward and the backward of that layer). */
size_t b200nn_dense_ws_bytes(int M, int N, int K);
/* the same for a given math mode: B200NN_MATH_TF32X3 cuts its accumulation chains through split-K and needs more */
size_t b200nn_dense_ws_bytes_math(int math, int M, int N, int K);
int b200nn_dense_fwd(int math, const float* x, const float* w, const float* b, float* y, int M, int N, int K,
                     int act, float alpha, float* ws, void* stream);
/* Classifier head as one kernel (N <= 32): p = softmax(x . w + b) (layers.py:173, activations.py:69-71), loss_acc[0] += CCE
 * numerator (losses.py:106-109), err = p - y_true (models.py:200-205), loss_acc[1] += argmax hits, ss_out += sum(err^2). */
int b200nn_dense_softmax_cce(const float* x, const float* w, const float* b, const float* y_true, float* p, float* err,
                             double* loss_acc, double* ss_out, int M, int N, int K, float* dw_partial, void* stream);
/* dw_partial (nullable): [b200nn_dense_head_partials(M, K)][(K + 1) * N] per-CTA partials of x^T . err (+ "ones" row = db),
 * formed while err is still in registers; b200nn_dense_bwd_head then replaces b200nn_dense_bwd for that layer: it adds the
 * partials in fixed order (times the pending chain) instead of re-reading x and err. 0 partials = shape not supported. */
int b200nn_dense_head_partials(int M, int K);
int b200nn_dense_bwd_head(const float* x, const float* w, const float* err, const double* ss, uint64_t chain, float* dx, float* dw,
                          float* db, int M, int N, int K, int prev_act, float prev_alpha, double* ss_out0, double* ss_out1,
                          const float* dw_partial, void* stream);
/* g = chain(err)[M,N];  dx[M,K] = g . w^T (layers.py:189);  dw[K,N] = x^T . g (:196);  db[N] = sum_rows g (:197).
 * dx == NULL skips dx (first parametrised layer inside fit). If prev_act != NONE the backward of the
 * PRECEDING Activation layer (whose output is this layer's input x) is fused into the dx epilogue:
 *   ss_out[0] += sum(dx^2) (the tensor the reference clips before Activation.backward_pass),
 *   dx <- dx * f'(x),  ss_out[1] += sum(dx^2).   With prev_act == NONE only ss_out[0] is written.
 * The caller then chains {slot0, slot1}. */
int b200nn_dense_bwd(int math, const float* x, const float* w, const float* err, const double* ss, uint64_t chain,
                     float* dx, float* dw, float* db, int M, int N, int K, int prev_act, float prev_alpha,
                     double* ss_out0, double* ss_out1, float* ws, void* stream);

/* ---- Conv2D (layers.py:442-530, preprocessing.py:34-126) ----------------------------------- */
typedef struct {
    int N, H, W, C;      /* input NHWC */
    int F, kh, kw;       /* filters, kernel */
    int sh, sw;          /* strides */
    int ph, pw;          /* symmetric zero pad (layers.py:450-465) */
    int OH, OW;          /* output spatial size */
} b200nn_conv_geom;
/* y = act(conv(x, w) + b). w is the reference's (kh,kw,C,F) buffer READ with K order (c,ky,kx)
 * (preprocessing.py:77 vs layers.py:474 — SURVEY Q1). */
size_t b200nn_conv2d_ws_bytes(const b200nn_conv_geom* g);
/* the same for a given math mode. With math != FP32, stride-1 'same' convolutions on power-of-two images with C and F
 * multiples of 32 run as implicit GEMMs on tcgen05 (operands fetched by 4-D TMA boxes, padding = TMA zero fill); every
 * other geometry keeps the FFMA path in every math mode. */
size_t b200nn_conv2d_ws_bytes_math(int math, const b200nn_conv_geom* g);
int b200nn_conv2d_fwd(int math, const b200nn_conv_geom* g, const float* x, const float* w, const float* b, float* y,
                      int act, float alpha, float* ws, void* stream);
/* layers.py:513-528: db[F] = sum g, dw (same K-order view), dx = col2im(g . w^T). Same dx epilogue
 * contract as b200nn_dense_bwd. */
int b200nn_conv2d_bwd(int math, const b200nn_conv_geom* g, const float* x, const float* w, const float* err,
                      const double* ss, uint64_t chain, float* dx, float* dw, float* db, int prev_act,
                      float prev_alpha, double* ss_out0, double* ss_out1, float* ws, void* stream);

/* ---- Conv2DTranspose (layers.py:2588-2659) ------------------------------------------------- */
typedef struct {
    int N, H, W, C;      /* input NHWC */
    int F, kh, kw, sh, sw;
    int pt, pl;          /* crop offsets of the forward (layers.py:2597-2609) */
    int OH, OW;          /* output spatial size after the crop */
} b200nn_convt_geom;
/* w is (kh,kw,F,C). y = act(scatter(x,w)[crop] + b). */
size_t b200nn_convt_ws_bytes(const b200nn_convt_geom* g);
/* math != FP32: the tensor-core path (column buffer + tcgen05 contractions, see gemm_simt.cu) needs the larger workspace */
size_t b200nn_convt_ws_bytes_math(int math, const b200nn_convt_geom* g);
int b200nn_convt_fwd(int math, const b200nn_convt_geom* g, const float* x, const float* w, const float* b,
                     float* y, int act, float alpha, float* ws, void* stream);
/* Reference backward: windows read from the cropped error with NO pad offset; an input pixel whose
 * kh x kw window is truncated by the border gets zero d_input and contributes nothing to dw
 * (layers.py:2651). db = sum over (n,h,w). */
int b200nn_convt_bwd(int math, const b200nn_convt_geom* g, const float* x, const float* w, const float* err,
                     const double* ss, uint64_t chain, float* dx, float* dw, float* db, int prev_act,
                     float prev_alpha, double* ss_out0, double* ss_out1, float* ws, void* stream);

/* ---- BatchNormalization (layers.py:1230-1276): statistics per POSITION over axis 0 ---------- */
/* x viewed as [B, P]. Training: xc = clip(x,+-10); mean, var over B; batch_var = var + eps;
 * running = momentum*running + (1-momentum)*{mean, batch_var}; std = sqrt(batch_var + eps);
 * y = gamma*(xc-mean)/std + beta. Saves mean[P] and invstd[P] for the backward. */
int b200nn_bn_fwd_train(const float* x, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, float* save_mean, float* save_invstd, float* y, int B, long long P,
                        float momentum, float eps, void* stream);
/* Inference: mean/var = running stats, std = sqrt(running_var + eps). */
int b200nn_bn_fwd_infer(const float* x, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float* y, int B, long long P, float eps, void* stream);
/* e = chain(err); dgamma = sum e*xhat; dbeta = sum e;
 * dx = gamma*invstd*(e - xhat*dgamma/B - dbeta/B)  (algebraically layers.py:1264-1274; the
 * mean(-2*centered) term is identically 0). xhat is recomputed from x, save_mean, save_invstd.
 * Fused dx epilogue as in b200nn_dense_bwd (prev_act evaluated on x, the BN input = activation output). */
int b200nn_bn_bwd(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                  const float* save_mean, const float* save_invstd, float* dx, float* dgamma, float* dbeta, int B,
                  long long P, int prev_act, float prev_alpha, double* ss_out0, double* ss_out1, void* stream);

/* ---- MaxPooling2D (layers.py:570-643) ------------------------------------------------------ */
typedef struct {
    int N, H, W, C;
    int ph, pw, sh, sw;  /* pool size, strides */
    int pad_h, pad_w;    /* zero padding ('same'; value 0 not -inf, layers.py:582-585) */
    int OH, OW;
} b200nn_pool_geom;
int b200nn_maxpool_fwd(const b200nn_pool_geom* g, const float* x, float* y, void* stream);
/* dx[n,h,w,c] = sum over windows containing (h,w) of chain(err)[window] * (x == window max):
 * EVERY tied maximum receives the gradient (layers.py:631-637). y is the forward output. */
int b200nn_maxpool_bwd(const b200nn_pool_geom* g, const float* x, const float* y, const float* err,
                       const double* ss, uint64_t chain, float* dx, double* ss_out, void* stream);

/* ---- fused [BatchNorm -> MaxPool 2x2/s2] forward and [MaxPool -> BatchNorm -> Activation] backward
 *      (the conv block of BASELINE configs 2 and 3; one pass over the conv output each way) ----- */
/* Largest batch the register-resident fused kernels accept (callers fall back to the unfused ops beyond it). */
int b200nn_fused_block_max_batch(void);
/* x [B,H,W,C] (H, W even) -> y_pool [B,H/2,W/2,C]; also updates the running stats and saves mean / invstd. */
int b200nn_bn_pool_fwd_train(const float* x, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, float* save_mean, float* save_invstd, float* y_pool, int B, int H,
                             int W, int C, float momentum, float eps, void* stream);
/* err is the error w.r.t. the pooled output. Emits r0 = BNbwd(poolbwd(chain(err))) * f'(x) and the three
 * sums of squares of the reference's three clip points (pool-bwd output, BN-bwd output, act-bwd output)
 * in ss_out0/1/2, plus dgamma/dbeta. The caller chains {slot0, slot1, slot2}. */
int b200nn_pool_bn_act_bwd(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                           const float* beta, const float* save_mean, const float* save_invstd, float* r0,
                           float* dgamma, float* dbeta, int B, int H, int W, int C, int prev_act, float prev_alpha,
                           double* ss_out0, double* ss_out1, double* ss_out2, void* stream);

/* ---- fused FIRST BLOCK: Conv2D(3x3, stride 1, 'same', C <= 3, F % 8 == 0) + Activation -> BatchNormalization(train) ->
 *      MaxPooling2D(2x2/2), the conv output never touching HBM (recomputed from the input image wherever it is needed;
 *      csrc/conv_block.cu). Replaces layers.py:442-489, :231-237, :1230-1276, :570-643 and the weight-gradient half of
 *      :491-530 for this pattern. `mode`: 0 single GPU; 1 / 2 the two halves around the data-parallel sum of
 *      `stats` (forward, fp64 [2][H*W*F]) resp. dgamma / dbeta (backward); B_total = global batch. */
struct b200nn_comm;   /* data-parallel communicator, declared below */
int b200nn_convblock_supported(const b200nn_conv_geom* g, int act, float alpha);
size_t b200nn_convblock_ws_bytes(const b200nn_conv_geom* g);
int b200nn_convblock_fwd(const b200nn_conv_geom* g, const float* x, const float* w, const float* b, int act, float alpha,
                         const float* gamma, const float* beta, float* running_mean, float* running_var, float* save_mean,
                         float* save_invstd, double* stats, int mode, long long B_total, float* y_pool, float momentum,
                         float eps, struct b200nn_comm* comm, int site, size_t region_off, void* stream);
/* err: error w.r.t. the pooled output, pending clips (ss, chain). Emits d_gamma / d_beta, the three sums of squares of the
 * reference's clip points (pool-bwd, BN-bwd, activation-bwd outputs) and dw / db scaled by chain_out (= those three
 * slots, complete once the main kernel has run). r_out (nullable): the error w.r.t. the conv output, materialised only
 * when a data gradient is wanted (it still owes chain_out). ws: b200nn_convblock_ws_bytes.
 * mode 3 ("peer-fused", both directions): ONE kernel per direction under data parallelism — every CTA merges its BatchNorm
 * batch reductions with the other GPUs' through CUDA-IPC peer memory over NVLink inside the kernel (persistent launch; other
 * CTAs of the SM compute while one waits). comm / site / region_off: a region of b200nn_convblock_region_bytes() bytes from
 * b200nn_comm_region_create; NULL / -1 / 0 for modes 0-2. */
size_t b200nn_convblock_region_bytes(const b200nn_conv_geom* g, int world);
int b200nn_convblock_bwd(const b200nn_conv_geom* g, const float* x, const float* w, const float* b, int act, float alpha,
                         const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                         const float* err, const double* ss, uint64_t chain, float* dgamma, float* dbeta, float* dw, float* db,
                         float* r_out, int mode, long long B_total, double* ss_out0, double* ss_out1, double* ss_out2,
                         uint64_t chain_out, float* ws, struct b200nn_comm* comm, int site, size_t region_off, void* stream);

/* the reduction of the weight-gradient partials as its own launch: under data parallelism it has to follow the exchange
 * of the clip slots in chain_out (call b200nn_convblock_bwd with mode 2 and dw = NULL first) */
int b200nn_convblock_wgrad_reduce(const b200nn_conv_geom* g, const float* ws, const double* ss, uint64_t chain_out, float* dw,
                                  float* db, void* stream);

/* ---- optimizers (optimizers.py:47-50, 167-245) and gradient clip (models.py:240-242) -------- */
typedef struct {
    float* p;          /* parameter tensor (updated in place, optimizers.py:223) */
    const float* g;    /* gradient */
    float* m;          /* Adam first moment  (unused by SGD) */
    float* v;          /* Adam second moment (unused by SGD) */
    long long n;       /* elements */
    int t_index;       /* Adam: 1-based position j of this tensor's LAYER among the L layers updated per step,
                          last layer first; t = *t_counter - L + j (optimizers.py:236) */
    int clip;          /* 1: scale g by 1/||g|| when ||g|| > 1 (models.py:240-242) */
    uint64_t chain;    /* error clips still pending on g (deferred clips; 0 = none): g is multiplied by chain(ss) first */
} b200nn_param;
/* `params` is a DEVICE array of n_params descriptors; `gss` a device array of n_params + 1 float64
 * scratch slots, ZERO on first use (slot i receives ||g_i||^2; the extra slot holds the two counters of the
 * single-launch variant used for small parameter sets, which re-zeroes the array itself); hyper is a DEVICE float64 array
 * {lr, beta1, beta2, eps} (device-resident so a LearningRateScheduler can change lr without
 * re-capturing the step graph); *t_counter is Adam's device-resident `t` AFTER this step's updates (the
 * caller advanced it by n_layers, e.g. through b200nn_step_begin). block_map is a DEVICE
 * array of total_blocks (tensor index, chunk index) int pairs, total_blocks = sum over tensors of
 * ceil(n / B200NN_OPT_CHUNK). */
#define B200NN_OPT_CHUNK 4096
int b200nn_adam_multi(const b200nn_param* params, int n_params, int n_layers, double* gss, const double* hyper,
                      const long long* t_counter, const int* block_map, int total_blocks, const double* ss, void* stream);
int b200nn_sgd_multi(const b200nn_param* params, int n_params, double* gss, const double* hyper,
                     const int* block_map, int total_blocks, const double* ss, void* stream);
/* The other update rules of optimizers.py on the same descriptors: Adam with its own clip_norm / clip_value (:190-202),
 * Momentum (:63-111; velocity in `m`), RMSprop (:114-164; squared-gradient average in `v`), AdaBelief (:286-402; `v` = s),
 * RAdam (:405-523). hyper is a DEVICE float64[6] {lr, beta1 | momentum | rho, beta2, eps, clip_norm, clip_value} (clip_* <= 0:
 * off); t_counter as for b200nn_adam_multi (NULL for Momentum / RMSprop, which keep no step count). */
#define B200NN_OPT_ADAM 0
#define B200NN_OPT_MOMENTUM 1
#define B200NN_OPT_RMSPROP 2
#define B200NN_OPT_ADABELIEF 3
#define B200NN_OPT_RADAM 4
int b200nn_opt_multi(int kind, const b200nn_param* params, int n_params, int n_layers, double* gss, const double* hyper,
                     const long long* t_counter, const int* block_map, int total_blocks, const double* ss, void* stream);

/* ---- data parallelism: one process per GPU on one NVLink/NVSwitch box -----------------------
 * The reference is single-process; under a batch shard per GPU its batch-global quantities become sums over ranks:
 * gradients (models.py:200-205 — never divided by the batch, so SUM not mean), the L2 norms behind every clip
 * (models.py:214, 240-242), BatchNormalization's per-position statistics (layers.py:1237-1238) and the reductions of its
 * backward (layers.py:1261-1270). One primitive serves them all: an in-place sum-allreduce of a device buffer, either
 * through NCCL (bandwidth-bound gradient buckets) or through a one-shot exchange over CUDA-IPC-mapped peer memory
 * (latency-bound scalars and statistic vectors; every rank sums the contributions in rank order, so all ranks obtain
 * bit-identical results). Both are asynchronous on `stream` and capturable into the step graph. */
typedef struct b200nn_comm b200nn_comm;
/* dlopen libnccl (NULL: the copy already loaded into the process, e.g. PyTorch's) */
int b200nn_comm_load_nccl(const char* path);
int b200nn_comm_nccl_version(void);
/* rank 0 creates the 128-byte NCCL id; the host side broadcasts it (torch.distributed) */
int b200nn_comm_unique_id(void* out128);
int b200nn_comm_create(const void* unique_id128, int rank, int world, b200nn_comm** out);
int b200nn_comm_destroy(b200nn_comm* comm);
/* symmetric peer buffer: allocate `bytes` (zeroed) and return its CUDA IPC handle; after an all-gather of the
 * handles on the host side, map every peer. Optional: without it every allreduce goes through NCCL. */
size_t b200nn_comm_ipc_handle_bytes(void);
int b200nn_comm_peer_alloc(b200nn_comm* comm, size_t bytes, void* handle_out);
int b200nn_comm_peer_open(b200nn_comm* comm, const void* all_handles);
int b200nn_comm_has_peer(const b200nn_comm* comm);
/* peer waits that gave up after 10 s because a rank never arrived (synchronises the device); non-zero = results invalid */
int b200nn_comm_timeouts(const b200nn_comm* comm);
/* reserve an exchange site (double-buffered inbox of world x payload_bytes) in the symmetric buffer; every rank
 * must create the same sites in the same order */
int b200nn_comm_site_create(b200nn_comm* comm, size_t payload_bytes, int* site_out);
/* a raw zero-initialised region of the symmetric buffer (same offset on every rank) with its own round counter, for kernels
 * that exchange over peer memory themselves (b200nn_convblock_* mode 3) */
int b200nn_comm_region_create(b200nn_comm* comm, size_t bytes, int* site_out, size_t* offset_out);
/* buf[0..n) <- sum over ranks, in place. dtype 0 = float32, 1 = float64. site >= 0: one-shot peer-memory exchange
 * through that site; site < 0: ncclAllReduce. */
int b200nn_comm_allreduce_sum(b200nn_comm* comm, int site, void* buf, long long n, int dtype, void* stream);
/* two payloads in one latency exchange (site >= 0 only; (n_a rounded up to even) + 2 n_b <= 1024 words, site payload >= 8 bytes per
 * word): n_a float32 values and n_b float64 values, each summed over the ranks in rank order like b200nn_comm_allreduce_sum */
int b200nn_comm_allreduce_sum2(b200nn_comm* comm, int site, float* a, int n_a, double* b, int n_b, void* stream);

/* ---- BatchNormalization under data parallelism: each batch-coupled kernel splits at its reduction ---------------
 * forward:  b200nn_bn_stats_partial  ->  sum `stats` over ranks (fp64)  ->  b200nn_bn_fwd_apply | b200nn_bn_pool_fwd_apply
 * backward: *_bwd_partial            ->  sum dgamma / dbeta over ranks  ->  *_bwd_apply
 * `stats` is [2][P] float64: sum clip(x) and sum clip(x)^2 over the LOCAL batch rows (the second moment is formed from
 * an exact two-pass local M2). B_total is the global batch (layers.py:1237-1238 take mean / var over all of it). */
int b200nn_bn_stats_partial(const float* x, double* stats, int B, long long P, void* stream);
int b200nn_bn_fwd_apply(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                        float* save_mean, float* save_invstd, const double* stats, float* y, int B, long long P,
                        long long B_total, float momentum, float eps, void* stream);
int b200nn_bn_pool_fwd_apply(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                             float* save_mean, float* save_invstd, const double* stats, float* y_pool, int B, int H, int W,
                             int C, long long B_total, float momentum, float eps, void* stream);
/* dgamma / dbeta receive the LOCAL sums (layers.py:1261-1262); after the sum over ranks the apply half reads them. */
int b200nn_bn_bwd_partial(const float* x, const float* err, const double* ss, uint64_t chain, const float* save_mean,
                          const float* save_invstd, float* dgamma, float* dbeta, int B, long long P, void* stream);
int b200nn_bn_bwd_apply(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                        const float* save_mean, const float* save_invstd, float* dx, const float* dgamma, const float* dbeta,
                        int B, long long P, long long B_total, int prev_act, float prev_alpha, double* ss_out0,
                        double* ss_out1, void* stream);
int b200nn_pool_bn_bwd_partial(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                               const float* beta, const float* save_mean, const float* save_invstd, float* dgamma,
                               float* dbeta, int B, int H, int W, int C, double* ss_out0, void* stream);
int b200nn_pool_bn_act_bwd_apply(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                                 const float* beta, const float* save_mean, const float* save_invstd, float* r0,
                                 const float* dgamma, const float* dbeta, int B, int H, int W, int C, long long B_total,
                                 int prev_act, float prev_alpha, double* ss_out1, double* ss_out2, void* stream);

/* ---- sibling layers of the same kernel families (csrc/layers_extra.cu) ---------------------------------
 * Every backward follows the lazy-clip protocol: dx = chain(ss) * (...), ss_out (nullable) += sum(dx^2).
 * b200nn_mul: out = chain * a * b — Dropout (layers.py:264-342) with the host-generated mask / (1 - rate) as `b`.
 * AveragePooling2D (layers.py:908-1011): mean over each window of the zero-padded input. GlobalAveragePooling2D
 * (:1419-1442) over HW positions. UpSampling2D nearest (:2695-2815): np.repeat forward, block sums backward.
 * b200nn_im2col / b200nn_col2im: columns ordered (ky, kx, c) for any channel count; with kh == 1 this is im2col_1d / col2im_1d
 * (preprocessing.py:128-186), which puts Conv1D (layers.py:667-816) on the Dense kernels. */
int b200nn_mul(const float* a, const float* b, const double* ss, uint64_t chain, float* out, double* ss_out, long long n, void* stream);
int b200nn_avgpool2d_fwd(const b200nn_pool_geom* g, const float* x, float* y, void* stream);
int b200nn_avgpool2d_bwd(const b200nn_pool_geom* g, const float* err, const double* ss, uint64_t chain, float* dx, double* ss_out,
                         void* stream);
/* Autoencoder.backward_pass's clip rule (models.py:1033-1048) on one tensor: norm clip to `threshold`, clamp to [-10, 10], divide by
 * (np.std + 1e-8); threshold <= 0 copies. scratch3: three device float64 words (zeroed by the call). in-place allowed. */
int b200nn_ae_clip(const float* g, float* out, long long n, float threshold, double* scratch3, void* stream);
/* out = a + s * b (the Autoencoder's scaled skip connection, models.py:1014-1019); in-place allowed */
int b200nn_axpy(const float* a, const float* b, float s, float* out, long long n, void* stream);
int b200nn_gap2d_fwd(const float* x, float* y, int N, int HW, int C, void* stream);
int b200nn_gap2d_bwd(const float* err, const double* ss, uint64_t chain, float* dx, double* ss_out, int N, int HW, int C, void* stream);
int b200nn_upsample2d_fwd(const float* x, float* y, int N, int H, int W, int C, int fh, int fw, void* stream);
int b200nn_upsample2d_bwd(const float* err, const double* ss, uint64_t chain, float* dx, double* ss_out, int N, int H, int W, int C,
                          int fh, int fw, void* stream);
int b200nn_im2col(const b200nn_conv_geom* g, const float* x, float* col, void* stream);
int b200nn_col2im(const b200nn_conv_geom* g, const float* dcol, const double* ss, uint64_t chain, float* dx, double* ss_out, void* stream);

/* STREAMING forms of the batch-coupled kernels for large batches (any batch size; thread = (pool window, channel quad) or 4
 * adjacent positions, blockIdx.y = a chunk of batch rows, every access a whole 128-byte line). `ws`: float workspace of
 * *_ws_bytes. The statistics are merged over the chunks in a fixed order (deterministic, no float atomics).
 *   b200nn_bn_stats_stream            = b200nn_bn_stats_partial in ONE sweep over x (shifted sums, Chan merge in fp64)
 *   b200nn_pool_bn_bwd_stream_partial = b200nn_pool_bn_bwd_partial   (dgamma / dbeta receive the LOCAL sums)
 *   b200nn_pool_bn_act_bwd_stream_apply = b200nn_pool_bn_act_bwd_apply */
size_t b200nn_bn_stats_stream_ws_bytes(int B, long long P);
int b200nn_bn_stats_stream(const float* x, double* stats, float* ws, int B, long long P, void* stream);
size_t b200nn_pool_bn_bwd_stream_ws_bytes(int B, int H, int W, int C);
int b200nn_pool_bn_bwd_stream_partial(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                                      const float* beta, const float* save_mean, const float* save_invstd, float* dgamma,
                                      float* dbeta, float* ws, int B, int H, int W, int C, double* ss_out0, void* stream);
int b200nn_pool_bn_act_bwd_stream_apply(const float* x, const float* err, const double* ss, uint64_t chain, const float* gamma,
                                        const float* beta, const float* save_mean, const float* save_invstd, float* r0,
                                        const float* dgamma, const float* dbeta, int B, int H, int W, int C, long long B_total,
                                        int prev_act, float prev_alpha, double* ss_out1, double* ss_out2, void* stream);

/* ---- single-launch train step of a small MLP (models.py:159-264 for Dense [+ Activation] ... Dense + Softmax, CCE, Adam) -------
 * One persistent cooperative kernel runs step_begin, every Dense forward (layers.py:173), the classifier head with
 * softmax / CCE / p - y (models.py:200-205), every Dense backward (layers.py:189-197) with the clips of models.py:214 carried as
 * the usual sum-of-squares slots, the per-tensor gradient clip (models.py:240-242) and Adam (optimizers.py:204-245, per-layer t),
 * separated by grid-wide barriers instead of kernel boundaries (config 1: 16 graph nodes -> 1). Results are those of the
 * layer-by-layer kernels (same slots, same stored gradients). Limits: batch <= 64, 2..8 Dense layers, last layer N <= 32.
 * `layer[l]`: w[K,N], b[N] (nullable), dw, db, out[B,N] (activation output; the last layer's is the prediction), err[B,N] (error
 * entering the layer's backward: written by the layer above; the last layer's is p - y), act / alpha of the Activation fused
 * behind the layer (ignored for the last layer: softmax), chain = the clip slots pending on err, s_out0 / s_out1 = slots that
 * receive ||dx||^2 before / after the activation backward of the layer below (-1: none), gi_w / gi_b = index of dw / db in the
 * optimizer descriptor table `params` (b200nn_param, as for b200nn_adam_multi; block_map / gss likewise). `sync`: 64 zeroed
 * device words (barrier counters in [0], [1], re-zeroed by the kernel; [2..31] = 15 x 8 bytes of per-phase %globaltimer stamps
 * written by CTA 0). */
#define B200NN_MLP_MAX_LAYERS 8
#define B200NN_MLP_MAX_BATCH 64
typedef struct {
    const float* w; const float* b; float* dw; float* db; float* out; float* err;
    int K, N, act; float alpha;
    uint64_t chain;
    int s_out0, s_out1, gi_w, gi_b;
} b200nn_mlp_layer;
typedef struct {
    int L, B, n_slots, n_params, n_layers_opt, total_blocks;
    const float* x; const float* y;
    double* ss; double* acc; double* gss; const double* hyper; long long* t_counter;
    const b200nn_param* params; const int* block_map; unsigned* sync;
    int s0;                                   /* slot of ||p - y||^2 */
    b200nn_mlp_layer layer[B200NN_MLP_MAX_LAYERS];
} b200nn_mlp_step;
int b200nn_mlp_train_step(const b200nn_mlp_step* step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200NN_H_ */
