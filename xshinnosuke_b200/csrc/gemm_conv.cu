// Dense GEMM and conv2d forward / dgrad / wgrad entry points.
// math_mode XSB_MATH_FP32 (0): CUDA-core FFMA implicit GEMM (igemm_ffma.cuh) -- fp32 in, fp32 FMA accumulate.
// math_mode XSB_MATH_TF32 (1): tcgen05 tensor-core implicit GEMM (conv_tc.cu / gemm_tc.cu) where the shape is
//   supported by those kernels; the entry point reports XSB_ERR_UNSUPPORTED otherwise (no silent fallback).
//
// Layouts (device): activations channels-last [N,H,W,C]; filters [OC,KH,KW,C]; Dense weight [in,out] row-major
// (XS stores Dense kernels un-transposed, xs/layers/linear.py:32). API-level NCHW <-> channels-last conversion
// is done by xsb_nchw_to_nhwc / xsb_nhwc_to_nchw at the Tensor boundary, not here.
//
// Reference arithmetic: xs/nn/functional.py:405-442 (conv2d = im2col . W^T + b), xs/nn/grad_fn.py:186-203
// (dW = dY^T . col, db = sum dY, dX = col2im(dY . W)), xs/nn/functional.py:81-120 + grad_fn.py:73-99 (Dense).
// FLOPs: 2*N*OH*OW*C*KH*KW*OC for each of fwd / dgrad / wgrad.
#include "igemm_ffma.cuh"

using namespace igemm;

// implemented in conv_tc.cu (tcgen05). Return XSB_ERR_UNSUPPORTED when the shape is outside their envelope.
int xsb_tc_gemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int transA,
                int transB, int accumulate, float* ws, int64_t ws_floats, cudaStream_t s, float* act);
int xsb_tc_conv_fwd(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g, float* ws,
                    int64_t ws_floats, cudaStream_t s, float* stats, int64_t stats_floats, int* n_partials, int accumulate);
int xsb_tc_conv_dgrad(const float* dy, const float* w, float* dx, const ConvGeom& g, int accumulate, float* ws,
                      int64_t ws_floats, cudaStream_t s);
int xsb_tc_conv_wgrad(const float* dy, const float* x, float* dw, const ConvGeom& g, int accumulate, float* ws,
                      int64_t ws_floats, cudaStream_t s);
void xsb_tc_note_ffma();

// TF32 mode: run the tensor-core kernel when the shape is inside its envelope; XSB_ERR_UNSUPPORTED means
// "not my shape" and the fp32 FFMA kernel (a strictly more accurate CUDA kernel of this library) takes the
// call, counted in xsb_tc_stats. Any other math_mode value is an error.
#define XSB_TRY_TC(call)                                                     \
    do {                                                                     \
        if (math_mode == 1) {                                                \
            const int rc__ = (call);                                         \
            if (rc__ != XSB_ERR_UNSUPPORTED) return rc__;                    \
            xsb_tc_note_ffma();                                              \
        } else if (math_mode != 0) {                                         \
            xsb_set_error("unknown math_mode %d", math_mode);                \
            return XSB_ERR_ARG;                                              \
        }                                                                    \
    } while (0)

namespace {

bool fill_geom(ConvGeom& g, int N, int H, int W, int C, int OC, int KH, int KW, int sh, int sw, int pad) {
    g.N = N; g.H = H; g.W = W; g.C = C; g.OC = OC; g.KH = KH; g.KW = KW; g.sh = sh; g.sw = sw; g.pad = pad;
    g.OH = (H - KH + 2 * pad) / sh + 1;  // xs/nn/functional.py:415-416
    g.OW = (W - KW + 2 * pad) / sw + 1;
    return N > 0 && C > 0 && OC > 0 && KH > 0 && KW > 0 && sh > 0 && sw > 0 && pad >= 0 && g.OH > 0 && g.OW > 0 &&
           (int64_t)N * H * W < (int64_t)1 << 31 && (int64_t)N * g.OH * g.OW < (int64_t)1 << 31 &&
           (int64_t)KH * KW * C < (int64_t)1 << 31;
}

}  // namespace

extern "C" {

static int gemm_impl(const float* A, const float* B, float* C, const float* bias, float* act, int M, int N, int K, int transA,
                     int transB, int accumulate, int math_mode, float* ws, int64_t ws_floats, cudaStream_t s) {
    XSB_TRY_TC(xsb_tc_gemm(A, B, C, bias, M, N, K, transA, transB, accumulate, ws, ws_floats, s, act));
    const bool split_ok = true;
    if (!transA && !transB) {
        MatK la{A, K, M, K, K % 4 == 0 && aligned16(A)};
        MatR lb{B, N, N, K, N % 4 == 0 && aligned16(B)};
        return launch(la, lb, C, N, bias, M, N, K, accumulate, split_ok, ws, ws_floats, s, act);
    } else if (!transA && transB) {
        MatK la{A, K, M, K, K % 4 == 0 && aligned16(A)};
        MatK lb{B, K, N, K, K % 4 == 0 && aligned16(B)};
        return launch(la, lb, C, N, bias, M, N, K, accumulate, split_ok, ws, ws_floats, s, act);
    } else if (transA && !transB) {
        MatR la{A, M, M, K, M % 4 == 0 && aligned16(A)};
        MatR lb{B, N, N, K, N % 4 == 0 && aligned16(B)};
        return launch(la, lb, C, N, bias, M, N, K, accumulate, split_ok, ws, ws_floats, s, act);
    } else {
        MatR la{A, M, M, K, M % 4 == 0 && aligned16(A)};
        MatK lb{B, K, N, K, K % 4 == 0 && aligned16(B)};
        return launch(la, lb, C, N, bias, M, N, K, accumulate, split_ok, ws, ws_floats, s, act);
    }
}

// C[M,N] (+)= op(A) . op(B) (+ bias[N]); row-major. transA: A stored [K,M]; transB: B stored [N,K].
int xsb_gemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int transA,
             int transB, int accumulate, int math_mode, float* ws, int64_t ws_floats, void* stream) {
    return gemm_impl(A, B, C, bias, nullptr, M, N, K, transA, transB, accumulate, math_mode, ws, ws_floats, xsb_stream(stream));
}

// Dense(..., activation='relu') in one kernel: C = op(A).op(B) + bias (the pre-activation the ReLU backward reads its x < 0
// mask from) and act = max(0, C), both written by the GEMM epilogue -- no ReLU pass over C.
int xsb_gemm_bias_relu(const float* A, const float* B, float* C, const float* bias, float* act, int M, int N, int K,
                       int transA, int transB, int math_mode, float* ws, int64_t ws_floats, void* stream) {
    XSB_CHECK_ARG(act != nullptr, "gemm_bias_relu: act must not be null");
    return gemm_impl(A, B, C, bias, act, M, N, K, transA, transB, 0, math_mode, ws, ws_floats, xsb_stream(stream));
}

// y[N,OH,OW,OC] = conv(x[N,H,W,C], w[OC,KH,KW,C]) (+ bias[OC]). With `stats`: the tensor-core kernels also leave
// per-CTA (count, mean, M2) partials of every output channel in stats[p][3][OC] and *n_partials = p > 0, which
// xsb_bn_finalize turns into the batch statistics of a following BatchNorm without reading y again; *n_partials = 0
// when the kernel that served the call gathers none (FFMA path, buffer too small) -- then use xsb_bn_stats.
int xsb_conv2d_fwd_stats(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int C, int OC,
                         int KH, int KW, int sh, int sw, int pad, int math_mode, float* ws, int64_t ws_floats,
                         float* stats, int64_t stats_floats, int* n_partials, void* stream) {
    ConvGeom g;
    XSB_CHECK_ARG(fill_geom(g, N, H, W, C, OC, KH, KW, sh, sw, pad), "conv2d_fwd: bad geometry");
    cudaStream_t s = xsb_stream(stream);
    if (n_partials) *n_partials = 0;
    XSB_TRY_TC(xsb_tc_conv_fwd(x, w, bias, y, g, ws, ws_floats, s, stats, stats_floats, n_partials, 0));
    if (n_partials) *n_partials = 0;
    const int P = N * g.OH * g.OW, J = KH * KW * C;
    PatchA la{Patch{x, g, P, J, C % 4 == 0 && aligned16(x)}};
    MatK lb{w, J, OC, J, J % 4 == 0 && aligned16(w)};
    return launch(la, lb, y, OC, bias, P, OC, J, 0, true, ws, ws_floats, s);
}

// y[N,OH,OW,OC] = conv(x[N,H,W,C], w[OC,KH,KW,C]) (+ bias[OC])
int xsb_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int C, int OC,
                   int KH, int KW, int sh, int sw, int pad, int math_mode, float* ws, int64_t ws_floats,
                   void* stream) {
    ConvGeom g;
    XSB_CHECK_ARG(fill_geom(g, N, H, W, C, OC, KH, KW, sh, sw, pad), "conv2d_fwd: bad geometry");
    cudaStream_t s = xsb_stream(stream);
    XSB_TRY_TC(xsb_tc_conv_fwd(x, w, bias, y, g, ws, ws_floats, s, nullptr, 0, nullptr, 0));
    const int P = N * g.OH * g.OW, J = KH * KW * C;
    PatchA la{Patch{x, g, P, J, C % 4 == 0 && aligned16(x)}};
    MatK lb{w, J, OC, J, J % 4 == 0 && aligned16(w)};
    return launch(la, lb, y, OC, bias, P, OC, J, 0, true, ws, ws_floats, s);
}

// Conv2D(..., activation='relu'): y = conv(x, w) + bias AND act = max(0, y). The FFMA kernel writes both from its epilogue;
// the tcgen05 kernels write y and the rectified copy is made by one elementwise pass over it (xsb_relu_fwd).
int xsb_relu_fwd(const float* x, float* y, uint8_t* mask, int64_t n, void* stream);
int xsb_conv2d_fwd_relu(const float* x, const float* w, const float* bias, float* y, float* act, int N, int H, int W, int C,
                        int OC, int KH, int KW, int sh, int sw, int pad, int math_mode, float* ws, int64_t ws_floats,
                        void* stream) {
    ConvGeom g;
    XSB_CHECK_ARG(fill_geom(g, N, H, W, C, OC, KH, KW, sh, sw, pad) && act, "conv2d_fwd_relu: bad geometry");
    cudaStream_t s = xsb_stream(stream);
    if (math_mode == 1) {
        const int rc = xsb_tc_conv_fwd(x, w, bias, y, g, ws, ws_floats, s, nullptr, 0, nullptr, 0);
        if (rc == XSB_OK) return xsb_relu_fwd(y, act, nullptr, (int64_t)N * g.OH * g.OW * OC, stream);
        if (rc != XSB_ERR_UNSUPPORTED) return rc;
        xsb_tc_note_ffma();
    } else if (math_mode != 0) {
        xsb_set_error("unknown math_mode %d", math_mode);
        return XSB_ERR_ARG;
    }
    const int P = N * g.OH * g.OW, J = KH * KW * C;
    PatchA la{Patch{x, g, P, J, C % 4 == 0 && aligned16(x)}};
    MatK lb{w, J, OC, J, J % 4 == 0 && aligned16(w)};
    return launch(la, lb, y, OC, bias, P, OC, J, 0, true, ws, ws_floats, s, act);
}

// y (+)= conv(x, w) (+ bias): the forward convolution with the gradient kernels' accumulate flag. One product of the
// 3xTF32 split of the fp32x3 math mode (x = x_hi + x_lo, w = w_hi + w_lo; y = x_lo*w_hi + x_hi*w_lo + x_hi*w_hi) is one call.
int xsb_conv2d_fwd_acc(const float* x, const float* w, const float* bias, float* y, int N, int H, int W, int C, int OC,
                       int KH, int KW, int sh, int sw, int pad, int accumulate, int math_mode, float* ws, int64_t ws_floats,
                       void* stream) {
    ConvGeom g;
    XSB_CHECK_ARG(fill_geom(g, N, H, W, C, OC, KH, KW, sh, sw, pad), "conv2d_fwd_acc: bad geometry");
    cudaStream_t s = xsb_stream(stream);
    XSB_TRY_TC(xsb_tc_conv_fwd(x, w, bias, y, g, ws, ws_floats, s, nullptr, 0, nullptr, accumulate));
    const int P = N * g.OH * g.OW, J = KH * KW * C;
    PatchA la{Patch{x, g, P, J, C % 4 == 0 && aligned16(x)}};
    MatK lb{w, J, OC, J, J % 4 == 0 && aligned16(w)};
    return launch(la, lb, y, OC, bias, P, OC, J, accumulate, true, ws, ws_floats, s);
}

// dx[N,H,W,C] (+)= sum_{kh,kw,oc} dy[n,(h+p-kh)/sh,(w+p-kw)/sw,oc] * w[oc,kh,kw,c]
int xsb_conv2d_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int C, int OC, int KH, int KW,
                     int sh, int sw, int pad, int accumulate, int math_mode, float* ws, int64_t ws_floats,
                     void* stream) {
    ConvGeom g;
    XSB_CHECK_ARG(fill_geom(g, N, H, W, C, OC, KH, KW, sh, sw, pad), "conv2d_dgrad: bad geometry");
    cudaStream_t s = xsb_stream(stream);
    XSB_TRY_TC(xsb_tc_conv_dgrad(dy, w, dx, g, accumulate, ws, ws_floats, s));
    const int P = N * H * W, J = KH * KW * OC;
    DgradA la{dy, g, P, J, OC % 4 == 0 && aligned16(dy)};
    DgradB lb{w, g, J, C % 4 == 0 && aligned16(w)};
    return launch(la, lb, dx, C, nullptr, P, C, J, accumulate, true, ws, ws_floats, s);
}

// dw[OC,KH,KW,C] (+)= sum_{n,oh,ow} dy[n,oh,ow,oc] * x[n,oh*sh-p+kh,ow*sw-p+kw,c]
int xsb_conv2d_wgrad(const float* dy, const float* x, float* dw, int N, int H, int W, int C, int OC, int KH, int KW,
                     int sh, int sw, int pad, int accumulate, int math_mode, float* ws, int64_t ws_floats,
                     void* stream) {
    ConvGeom g;
    XSB_CHECK_ARG(fill_geom(g, N, H, W, C, OC, KH, KW, sh, sw, pad), "conv2d_wgrad: bad geometry");
    cudaStream_t s = xsb_stream(stream);
    XSB_TRY_TC(xsb_tc_conv_wgrad(dy, x, dw, g, accumulate, ws, ws_floats, s));
    const int P = N * g.OH * g.OW, J = KH * KW * C;
    MatR la{dy, OC, OC, P, OC % 4 == 0 && aligned16(dy)};  // A[m=oc][k=pixel] = dy[pixel*OC + oc]
    PatchB lb{Patch{x, g, P, J, C % 4 == 0 && aligned16(x)}};
    return launch(la, lb, dw, J, nullptr, OC, J, P, accumulate, true, ws, ws_floats, s);
}

}  // extern "C"
