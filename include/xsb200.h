/* libxsb200 -- C ABI of the B200-native backend for the XShinnosuke ("XS") data-parallel hot path.
 *
 * The reference (E1eveNn/xshinnosuke) has no FFI: its only backend seam is the module pointer
 * GLOBAL.np (xs/core/__global.py:8). The drop-in boundary is therefore the Python function level
 * (xs/nn/functional.py, xs/nn/grad_fn.py, xs/optim/*.py); each entry point below is what the
 * replacement of one of those functions binds through ctypes. The "replaces" note on every entry
 * names the reference code whose arithmetic it performs (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure (xsb_last_error() has the text);
 *   - all pointers except xsb_memcpy_* host sides are DEVICE pointers, fp32 unless stated;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the device;
 *   - activations are channels-last [N,H,W,C], filters [OC,KH,KW,C]; xsb_nchw_to_nhwc /
 *     xsb_nhwc_to_nchw convert from/to the reference's logical NCHW at the tensor boundary;
 *   - `accumulate` / `acc_*` = 1 means "+=" (the reference's closures always accumulate into
 *     `in_bound.grad`), 0 means "=" (first write into a freshly zeroed gradient, no memset needed);
 *   - `ws` is a caller-owned scratch buffer: zero-initialised once, reusable by every call on one stream.
 */
#ifndef XSB200_H
#define XSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XSB_OK 0
#define XSB_ERR_CUDA 1
#define XSB_ERR_ARG 2
#define XSB_ERR_UNSUPPORTED 3

/* math modes of the GEMM-shaped entry points */
#define XSB_MATH_FP32 0 /* CUDA-core FFMA, fp32 accumulate: <= 1e-5 vs the float64 reference      */
#define XSB_MATH_TF32 1 /* tcgen05 kind::tf32, fp32 accumulate in TMEM: <= 2e-3 vs the reference  */
/* fp32-accurate products ON the tensor cores ("fp32x3", <= 1e-5 vs the reference) are a host-level composition of
 * XSB_MATH_TF32 calls: xsb_split_tf32 writes x ~ hi + lo with both parts exact in tf32, and a product x.w becomes
 * lo.hi + hi.lo + hi.hi -- three accumulating calls (xsb_conv2d_fwd_acc, xsb_conv2d_dgrad / _wgrad / xsb_gemm with
 * accumulate = 1). The products are exact in fp32; dropped: lo.lo and the rounding of lo, ~2^-22 relative, unbiased
 * (xshinnosuke_b200/nn/x3.py). */

/* ---- runtime (no reference counterpart; plumbing) ---- */
int xsb_version(void);
const char* xsb_last_error(void);
int xsb_init(int device);
int xsb_device_sm_count(void);
/* n > 0: the grid-sized (persistent, one CTA per SM) kernels plan for n SMs, leaving the rest to a concurrent collective
 * (data parallelism: NCCL's all-reduce CTAs); 0: all SMs. */
int xsb_set_sm_budget(int n);
int xsb_stream_sync(void* stream);
int xsb_memcpy_h2d(void* dst_dev, const void* src_host, int64_t bytes, void* stream);
int xsb_memcpy_d2h(void* dst_host, const void* src_dev, int64_t bytes, void* stream);

/* ---- layout: logical NCHW (xs/core/base.py:69-75 `.eval` contract) <-> device channels-last ---- */
int xsb_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int accumulate, void* stream);
int xsb_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, int accumulate, void* stream);

/* ---- ReLU. replaces xs/nn/functional.py:297-313 + xs/nn/td_functional.py:59-60 (forward),
 *      xs/nn/grad_fn.py:146-157 + xs/layers/activators.py:37-50 (backward; mask is x < 0). ---- */
int xsb_relu_fwd(const float* x, float* y, uint8_t* mask /* nullable, ceil(n/8) bytes */, int64_t n, void* stream);
int xsb_relu_bwd(const float* dy, const float* x, float* dx, int64_t n, int accumulate, void* stream);
int xsb_relu_bwd_mask(const float* dy, const uint8_t* mask, float* dx, int64_t n, int accumulate, void* stream);

/* ---- add / accumulate / fill. replaces xs/nn/functional.py:6-21, xs/nn/grad_fn.py:6-16,
 *      xs/core/base.py:312-316 (zero_grad). ---- */
int xsb_add(const float* a, const float* b, float* y, int64_t n, void* stream);
int xsb_axpy(float* dst, const float* src, float alpha, int64_t n, void* stream);
/* residual add fused with the in-place ReLU that follows it: y = max(0, a+b), mask bit e = [(a+b)[e] < 0] */
int xsb_add_relu(const float* a, const float* b, float* y, uint8_t* mask, int64_t n, void* stream);
int xsb_fill(float* dst, float value, int64_t n, void* stream);
/* hi = x rounded to tf32, lo = (x - hi) rounded to tf32 (both exact in tf32, so the tensor core's operand truncation is a
 * no-op on them): operands of the 3xTF32 split */
int xsb_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);

/* ---- SURVEY 8f-4 ops around the hot path (csrc/extra_ops.cu) ----
 * Sigmoid / Tanh (kind 0 / 1). replaces xs/nn/td_functional.py:63-70 and xs/nn/grad_fn.py:161-170 (derivative written in y). */
int xsb_unary_fwd(int kind, const float* x, float* y, int64_t n, void* stream);
int xsb_unary_bwd(int kind, const float* dy, const float* y, float* dx, int64_t n, int accumulate, void* stream);
/* Dropout2D. replaces xs/nn/functional.py:650-674 (y = x*mask/keep_prob) and xs/nn/grad_fn.py:319-325 (dx += dy*mask*scale,
 * the reference passes scale = keep_prob). mask: 1 bit per element, set = kept. The draw is a stateless hash of
 * (seed, *step_dev, element index) -- NOT NumPy's binomial stream (parity is defined given the mask). */
int xsb_dropout_fwd(const float* x, float* y, uint8_t* mask, int64_t n, float keep_prob, int64_t seed,
                    const int* step_dev /* nullable */, void* stream);
int xsb_dropout_bwd(const float* dy, const uint8_t* mask, float* dx, int64_t n, float scale, int accumulate, void* stream);
/* ZeroPadding2D. replaces xs/nn/functional.py:629-647 and xs/nn/grad_fn.py:312-316 (channels-last tensors). */
int xsb_pad2d_fwd(const float* x, float* y, int N, int H, int W, int C, int ph, int pw, void* stream);
int xsb_pad2d_bwd(const float* dy, float* dx, int N, int H, int W, int C, int ph, int pw, int accumulate, void* stream);
/* LayerNorm over rows of x[N, F] in LOGICAL (NCHW row-major) order, per-element gamma / beta [F]. replaces
 * xs/nn/functional.py:743-775, xs/nn/grad_fn.py:414-447. mean / rstd [N] stand in for the cached xmu / sqrtvar. */
int xsb_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int N,
                      int64_t F, float eps, void* stream);
int xsb_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      float* dx /* nullable */, float* dgamma /* nullable */, float* dbeta /* nullable */, int N, int64_t F,
                      int acc_dx, int acc_dgamma, int acc_dbeta, void* stream);
/* GroupNorm on channels-last x[N, HW, C], G groups, per-channel gamma / beta. replaces xs/nn/functional.py:778-812,
 * xs/nn/grad_fn.py:450-490. mean / rstd [N*G]. */
int xsb_groupnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int N,
                      int HW, int C, int G, float eps, void* stream);
int xsb_groupnorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      float* dx /* nullable */, float* dgamma /* nullable */, float* dbeta /* nullable */, int N, int HW, int C,
                      int G, int acc_dx, int acc_dgamma, int acc_dbeta, void* stream);

/* ---- mean/sum over one axis and its backward. replaces xs/nn/functional.py mean/sum + grad_fn.py:122-137 ---- */
int xsb_reduce_axis(const float* x, float* y, int64_t outer, int64_t red, int64_t inner, float scale, void* stream);
int xsb_broadcast_axis(const float* dy, float* dx, int64_t outer, int64_t red, int64_t inner, float scale,
                       int accumulate, void* stream);

/* ---- max_pool2d with argmax. replaces xs/nn/functional.py:445-491 (zero padding, first max wins,
 *      argmax = ky*k+kx in (n,oh,ow,c) order) and xs/nn/grad_fn.py:206-235. ---- */
int xsb_maxpool2d_fwd(const float* x, float* y, uint8_t* argmax /* nullable */, int N, int H, int W, int C, int k,
                      int sh, int sw, int pad, void* stream);
/* BatchNormalization -> ReLU(inplace) -> MaxPooling2D fused forward: y = maxpool(max(0, bn(x))) with argmax and the
 * ReLU bit mask, without writing the intermediate tensor (xs/nn/functional.py:726-737 + activators.py:37-50 + 445-491).
 * Returns XSB_ERR_UNSUPPORTED, launching nothing, when the geometry / channel count is outside the kernel's envelope. */
int xsb_bn_relu_maxpool_fwd(const float* x, const float* mean, const float* invstd, const float* gamma,
                            const float* beta, float* y, uint8_t* argmax /* nullable */,
                            uint8_t* relu_mask /* nullable */, int N, int H, int W, int C, int k, int sh, int sw, int pad,
                            void* stream);
int xsb_maxpool2d_bwd(const float* dy, const uint8_t* argmax, float* dx, int N, int H, int W, int C, int k, int sh,
                      int sw, int pad, int accumulate, void* stream);

/* ---- avg_pool2d (padding 0). replaces xs/nn/functional.py:539-583, xs/nn/grad_fn.py:284-309 ---- */
int xsb_avgpool2d_fwd(const float* x, float* y, int N, int H, int W, int C, int k, int sh, int sw, void* stream);
int xsb_avgpool2d_bwd(const float* dy, float* dx, int N, int H, int W, int C, int k, int sh, int sw, int accumulate,
                      void* stream);

/* ---- BatchNorm, training branch, on [R, C] (R = N*H*W or N). replaces xs/nn/functional.py:676-706,730-740
 *      and xs/nn/grad_fn.py:379-411. save_mean/save_invstd stand in for the cached xhat/sqrtvar. ---- */
int64_t xsb_bn_workspace_floats(int64_t R, int C);
int xsb_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float* y, float* save_mean, float* save_invstd, int64_t R, int C, float eps, float momentum,
                     float* ws, void* stream);
/* inference branch, xs/nn/functional.py:708-725 */
int xsb_bn_eval_fwd(const float* x, const float* gamma, const float* beta, const float* moving_mean,
                    const float* moving_var, float* y, int64_t R, int C, float eps, float* ws, void* stream);
/* the two halves of xsb_bn_train_fwd; `relu` fuses the BatchNormalization -> ReLU(inplace=True) pair of
 * xs/layers/activators.py:37-50 (y = max(0, bn(x)), bit mask of bn(x) < 0 written in the same pass). */
int xsb_bn_stats(const float* x, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                 int64_t R, int C, float eps, float momentum, float* ws, void* stream);
int xsb_bn_apply(const float* x, float* y, const float* mean, const float* invstd, const float* gamma,
                 const float* beta, int64_t R, int C, int relu, uint8_t* relu_mask /* nullable */, void* stream);
/* relu_mask (nullable): dy is masked with the in-place ReLU's bit mask on the fly (xs/nn/grad_fn.py:147-151 fused). */
int xsb_bn_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
               float* dx /* nullable */, float* dgamma /* nullable */, float* dbeta /* nullable */, int64_t R, int C,
               int acc_dx, int acc_dgamma, int acc_dbeta, const uint8_t* relu_mask /* nullable */, float* ws,
               void* stream);

/* xsb_bn_bwd that also gathers dxsum[c] (+)= sum_r dx_contribution[r,c] while dx is written: the bias gradient of the
 * convolution feeding this BatchNorm (xs/nn/grad_fn.py:196-197 without a second pass over dx). *dxsum_done (HOST int)
 * = 1 when produced, 0 when the shape ran the generic kernels (then call xsb_colsum). */
int xsb_bn_bwd_dxsum(const float* dy, const float* x, const float* gamma, const float* save_mean,
                     const float* save_invstd, float* dx, float* dgamma /* nullable */, float* dbeta /* nullable */,
                     int64_t R, int C, int acc_dx, int acc_dgamma, int acc_dbeta,
                     const uint8_t* relu_mask /* nullable */, float* ws, float* dxsum, int acc_dxsum, int* dxsum_done,
                     void* stream);
/* The tail of a residual block -- BatchNormalization -> add -> ReLU(inplace) (xs/nn/functional.py:727-738 + 6-21 + 297-313)
 * -- in one pass: z = max(0, bn_a(xa) + rhs) and the (sum < 0) bit mask; rhs = b (mean_b == NULL: identity shortcut) or a
 * second BatchNorm applied on the fly, bn_b(b) (down-sampling shortcut). Bit-identical to xsb_bn_apply (+ xsb_bn_apply)
 * + xsb_add_relu; the BatchNorm outputs are not written. XSB_ERR_UNSUPPORTED (3): nothing launched. */
int xsb_bn_add_relu(const float* xa, const float* mean_a, const float* invstd_a, const float* gamma_a, const float* beta_a,
                    const float* b, const float* mean_b /* nullable */, const float* invstd_b, const float* gamma_b,
                    const float* beta_b, float* z, uint8_t* relu_mask, int64_t R, int C, void* stream);
/* xsb_bn_bwd_dxsum for a BatchNorm whose (in-place rectified) output feeds a 3x3 / stride 2 / pad 1 max-pool: dy is the
 * pool's backward gather of dpool[N,OH,OW,C] through argmax (xs/nn/grad_fn.py:206-235), masked by relu_mask
 * (xs/nn/grad_fn.py:146-157) -- formed on the fly in both passes, the full-size dy[N,H,W,C] is never written.
 * x, dx: [N,H,W,C]. dxsum nullable. XSB_ERR_UNSUPPORTED (3): nothing launched, run xsb_maxpool2d_bwd + xsb_bn_bwd. */
int xsb_bn_bwd_pool_dxsum(const float* dpool, const uint8_t* argmax, const float* x, const float* gamma,
                          const float* save_mean, const float* save_invstd, float* dx, float* dgamma /* nullable */,
                          float* dbeta /* nullable */, int N, int H, int W, int C, int acc_dx, int acc_dgamma, int acc_dbeta,
                          const uint8_t* relu_mask /* nullable */, float* ws, float* dxsum /* nullable */, int acc_dxsum,
                          void* stream);
/* xsb_bn_bwd_dxsum (dxsum nullable) that also writes dy_masked_out[R,C] = dy with the relu_mask elements zeroed, from the
 * apply pass that reads the masked dy anyway: the gradient of the identity shortcut of a residual block (AddBackward behind an
 * in-place ReLU, xs/nn/grad_fn.py:6-16 + 146-157) without a pass of its own. *side_done (HOST int) = 1 when written. */
int xsb_bn_bwd_dxsum_side(const float* dy, const float* x, const float* gamma, const float* save_mean,
                          const float* save_invstd, float* dx, float* dgamma /* nullable */, float* dbeta /* nullable */,
                          int64_t R, int C, int acc_dx, int acc_dgamma, int acc_dbeta, const uint8_t* relu_mask /* nullable */,
                          float* ws, float* dxsum /* nullable */, int acc_dxsum, int* dxsum_done, float* dy_masked_out,
                          int* side_done, void* stream);
/* batch statistics from the per-CTA (count, mean, M2) partials a convolution's epilogue left (xsb_conv2d_fwd_stats):
 * same outputs and running-statistics update as xsb_bn_stats (xs/nn/functional.py:689-706), no pass over x. */
int xsb_bn_finalize(const float* partials, int P, int64_t R, int C, float* running_mean, float* running_var,
                    float* save_mean, float* save_invstd, float eps, float momentum, void* stream);

/* ---- out[c] (+)= sum_r x[r,c]: conv / Dense bias gradient. replaces xs/nn/grad_fn.py:196-197, 89-95 ---- */
int xsb_colsum(const float* x, float* out, int64_t R, int C, int accumulate, float* ws, void* stream);

/* ---- Dense. replaces xs/nn/functional.py:81-120 (mm / addmm) and xs/nn/grad_fn.py:73-99.
 *      C[M,N] (+)= op(A).op(B) (+ bias[N]); transA: A stored [K,M]; transB: B stored [N,K].
 *      math_mode 1 (TF32): tcgen05 kernel when the row pitches of A, B and C are multiples of 16 bytes (the operands are
 *      TMA sources as stored, W[in,out] is never transposed), the FP32 FFMA kernel otherwise (counted by xsb_tc_stats). ---- */
int xsb_gemm(const float* A, const float* B, float* C, const float* bias /* nullable */, int M, int N, int K,
             int transA, int transB, int accumulate, int math_mode, float* ws, int64_t ws_floats, void* stream);

/* Dense(out, activation='relu') (xs/layers/linear.py:43-60: the activation layer chained behind addmm) as ONE kernel:
 * C = op(A).op(B) + bias and act = max(0, C) both leave from the GEMM epilogue (C keeps the pre-activation the ReLU backward
 * takes its x < 0 mask from, xs/nn/grad_fn.py:152-157). */
int xsb_gemm_bias_relu(const float* A, const float* B, float* C, const float* bias /* nullable */, float* act, int M, int N,
                       int K, int transA, int transB, int math_mode, float* ws, int64_t ws_floats, void* stream);

/* ---- conv2d as implicit GEMM (no materialised im2col). replaces xs/nn/functional.py:405-442 +
 *      xs/utils/common.py:35-46 (forward), xs/nn/grad_fn.py:186-203 + xs/utils/common.py:49-65 (backward).
 *      x [N,H,W,C], w [OC,KH,KW,C], y/dy [N,OH,OW,OC], OH = (H-KH+2*pad)/sh+1. ---- */
int xsb_conv2d_fwd(const float* x, const float* w, const float* bias /* nullable */, float* y, int N, int H, int W,
                   int C, int OC, int KH, int KW, int sh, int sw, int pad, int math_mode, float* ws, int64_t ws_floats,
                   void* stream);
/* Conv2D(..., activation='relu') (xs/layers/convolutional.py:61-83): y = conv(x, w) + bias and act = max(0, y); fused into the
 * epilogue of the fp32 kernel, one extra elementwise pass behind the tcgen05 kernels. */
int xsb_conv2d_fwd_relu(const float* x, const float* w, const float* bias /* nullable */, float* y, float* act, int N, int H,
                        int W, int C, int OC, int KH, int KW, int sh, int sw, int pad, int math_mode, float* ws,
                        int64_t ws_floats, void* stream);
/* y (+)= conv(x, w) (+ bias): xsb_conv2d_fwd with the accumulate flag of the gradient entry points */
int xsb_conv2d_fwd_acc(const float* x, const float* w, const float* bias /* nullable */, float* y, int N, int H, int W,
                       int C, int OC, int KH, int KW, int sh, int sw, int pad, int accumulate, int math_mode, float* ws,
                       int64_t ws_floats, void* stream);
/* forward conv that also leaves BatchNorm partial statistics of y in stats[p][3][OC] (tensor-core kernels only);
 * *n_partials (HOST int) = p, or 0 when the kernel that served the call gathers none -> use xsb_bn_stats then. */
int xsb_conv2d_fwd_stats(const float* x, const float* w, const float* bias /* nullable */, float* y, int N, int H, int W,
                         int C, int OC, int KH, int KW, int sh, int sw, int pad, int math_mode, float* ws,
                         int64_t ws_floats, float* stats, int64_t stats_floats, int* n_partials, void* stream);
int xsb_conv2d_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int C, int OC, int KH, int KW,
                     int sh, int sw, int pad, int accumulate, int math_mode, float* ws, int64_t ws_floats,
                     void* stream);
int xsb_conv2d_wgrad(const float* dy, const float* x, float* dw, int N, int H, int W, int C, int OC, int KH, int KW,
                     int sh, int sw, int pad, int accumulate, int math_mode, float* ws, int64_t ws_floats,
                     void* stream);

/* counters (no reference counterpart): out2[0] = tcgen05 kernel launches so far, out2[1] = GEMM-shaped calls
 * that ran the fp32 FFMA kernel while in XSB_MATH_TF32 mode (shape outside the tensor-core envelope). */
int xsb_tc_stats(uint64_t* out2);
/* kernel-selection switches for tests / A-B measurements: key 0 = CTA-pair (tcgen05 cta_group::2) forward kernels (0/1),
 * key 1 = dgrad reads the forward filter as stored (MN-major operand) instead of a flipped / transposed copy (0/1),
 * key 2 = persistent multi-job convolution kernel: 0 never, 1 strided dgrads (default), 2 every im2col forward / dgrad */
int xsb_tc_config(int key, int value);

/* ---- losses. replaces xs/nn/functional.py:984-1010 + xs/nn/td_functional.py:73-78,115-128 +
 *      xs/nn/grad_fn.py:604-620 (cross entropy; labels int64), xs/nn/functional.py:843-873 +
 *      xs/nn/grad_fn.py:503-520 (XS mse). reduction: 0 = mean, 1 = sum. ---- */
/* bad_label (nullable, device int): a label outside [0, C) is never dereferenced -- the loss becomes NaN and *bad_label = 1
 * (the reference's NumPy indexing raises IndexError, xs/nn/td_functional.py:119-127; the host raises it on the first read). */
int xsb_softmax_xent_fwd(const float* logits, const int64_t* labels, float* probs, float* loss, int N, int C,
                         int reduction, int* bad_label /* nullable */, void* stream);
int xsb_softmax_xent_bwd(const float* probs, const int64_t* labels, const float* dloss, float* dlogits, int N, int C,
                         int reduction, int accumulate, void* stream);
int xsb_mse_fwd(const float* a, const float* b, float* loss, int64_t n, int reduction, void* stream);
int xsb_mse_bwd(const float* a, const float* b, const float* dloss, float* da /* nullable */, float* db /* nullable */,
                int64_t n, int reduction, int acc_a, int acc_b, void* stream);

/* ---- fused multi-tensor optimizer over the flat parameter buffer. replaces xs/optim/sgd.py:5-9 and
 *      xs/optim/adam.py:13-30. grad_scale = 1/world for all-reduced (summed) gradients. weight_decay: every rule sees
 *      g' = grad_scale*g + weight_decay*p (L2 term, applied in the same pass). The reference accepts the argument
 *      (xs/optim/optimizer.py:9) and never applies it; 0 reproduces the reference's arithmetic exactly. ---- */
int xsb_sgd_step(float* p, const float* g, int64_t n, float lr, float grad_scale, float weight_decay, void* stream);
int xsb_adam_step(float* p, const float* g, float* v /* 1st moment */, float* m /* 2nd moment */, int64_t n, float lr,
                  float beta1, float beta2, float eps, float grad_scale, float weight_decay, const int* step_dev,
                  void* stream);
int xsb_counter_inc(int* counter_dev, void* stream);
/* Momentum (xs/optim/sgd.py:12-29), RMSprop (rmsprop.py:4-21), AdaGrad (adagrad.py:4-18), AdaDelta
 * (adadelta.py:4-26) -- kind 0..3; s1/s2 = per-parameter state buffers (s2: AdaDelta's delta_x only). */
int xsb_rule_step(int kind, float* p, const float* g, float* s1, float* s2, int64_t n, float lr, float rho, float eps,
                  float grad_scale, float weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XSB200_H */
