// tcgen05 tensor-core path (placeholder until the kernels land): every entry reports UNSUPPORTED so that a
// caller asking for XSB_MATH_TF32 fails loudly instead of silently running the FFMA path.
#include "common.cuh"
#include "conv_geom.cuh"

int xsb_tc_gemm(const float*, const float*, float*, const float*, int, int, int, int, int, int, float*, int64_t,
                cudaStream_t) {
    xsb_set_error("tcgen05 GEMM not built yet");
    return XSB_ERR_UNSUPPORTED;
}
int xsb_tc_conv_fwd(const float*, const float*, const float*, float*, const ConvGeom&, float*, int64_t, cudaStream_t) {
    xsb_set_error("tcgen05 conv fwd not built yet");
    return XSB_ERR_UNSUPPORTED;
}
int xsb_tc_conv_dgrad(const float*, const float*, float*, const ConvGeom&, int, float*, int64_t, cudaStream_t) {
    xsb_set_error("tcgen05 conv dgrad not built yet");
    return XSB_ERR_UNSUPPORTED;
}
int xsb_tc_conv_wgrad(const float*, const float*, float*, const ConvGeom&, int, float*, int64_t, cudaStream_t) {
    xsb_set_error("tcgen05 conv wgrad not built yet");
    return XSB_ERR_UNSUPPORTED;
}
