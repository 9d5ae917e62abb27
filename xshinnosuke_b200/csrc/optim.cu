// Fused multi-tensor optimizer step over ONE flat parameter buffer (all parameters of the model live
// contiguously, gradients in a parallel flat buffer in the same order).
// Roofline: HBM. Algorithmic bytes: SGD 3e*P (p, g read; p write), Adam 7e*P (p, g, m, v read; p, m, v write).
// Reference: xs/optim/sgd.py:5-9 (p -= lr*g, weight_decay never applied, Q8);
// xs/optim/adam.py:13-30 (v = b1 v + (1-b1) g; m = b2 m + (1-b2) g^2;
//   p -= lr * (v/(1-b1^t)) / (sqrt(m/(1-b2^t)) + eps); v is the FIRST moment, m the second).
// grad_scale folds the data-parallel 1/world averaging of the all-reduced gradient sum into the update.
// weight_decay: the reference accepts the argument and never applies it (Q8); here wd != 0 adds the
// classic L2 term to the gradient every rule sees, g' = grad_scale*g + wd*p, in the same pass (no extra traffic);
// wd == 0 takes the exact arithmetic of the reference.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

__global__ void __launch_bounds__(kThreads)
sgd_vec(float4* p, const float4* __restrict__ g, int64_t n4, float step, float lr_wd) {
    const int64_t base = (int64_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
    float4 pv[kUnroll], gv[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
        const int64_t i = base + (int64_t)j * kThreads;
        if (i < n4) { pv[j] = p[i]; gv[j] = ldg_stream(g + i); }
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
        const int64_t i = base + (int64_t)j * kThreads;
        if (i < n4) {
            if (lr_wd != 0.f) {     // p -= lr * (gs*g + wd*p)
                pv[j].x -= step * gv[j].x + lr_wd * pv[j].x; pv[j].y -= step * gv[j].y + lr_wd * pv[j].y;
                pv[j].z -= step * gv[j].z + lr_wd * pv[j].z; pv[j].w -= step * gv[j].w + lr_wd * pv[j].w;
            } else {
                pv[j].x -= step * gv[j].x; pv[j].y -= step * gv[j].y; pv[j].z -= step * gv[j].z; pv[j].w -= step * gv[j].w;
            }
            p[i] = pv[j];
        }
    }
}
__global__ void sgd_scalar(float* p, const float* __restrict__ g, int64_t n, float step, float lr_wd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] -= lr_wd != 0.f ? step * g[i] + lr_wd * p[i] : step * g[i];
}

__device__ __forceinline__ float adam1(float& p, float g, float& v, float& m, float b1, float b2, float lr,
                                       float inv_c1, float inv_c2, float eps) {
    v = b1 * v + (1.f - b1) * g;
    m = b2 * m + (1.f - b2) * g * g;
    p -= lr * ((v * inv_c1) / (sqrtf(m * inv_c2) + eps));
    return p;
}

// `step` is a device-resident counter (already incremented for this update) so the kernel can be
// replayed from a CUDA graph with a changing bias correction.
__global__ void __launch_bounds__(kThreads)
adam_vec(float4* p, const float4* __restrict__ g, float4* v, float4* m, int64_t n4, float lr, float b1, float b2,
         float eps, float gscale, float wd, const int* __restrict__ step) {
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        const double t = (double)step[0];
        s_c[0] = (float)(1.0 / (1.0 - pow((double)b1, t)));
        s_c[1] = (float)(1.0 / (1.0 - pow((double)b2, t)));
    }
    __syncthreads();
    const float ic1 = s_c[0], ic2 = s_c[1];
    const int64_t base = (int64_t)blockIdx.x * (kThreads * 2) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int64_t i = base + (int64_t)j * kThreads;
        if (i < n4) {
            float4 pv = p[i], vv = v[i], mv = m[i];
            float4 gv = ldg_stream(g + i);
            gv.x *= gscale; gv.y *= gscale; gv.z *= gscale; gv.w *= gscale;
            if (wd != 0.f) { gv.x += wd * pv.x; gv.y += wd * pv.y; gv.z += wd * pv.z; gv.w += wd * pv.w; }
            adam1(pv.x, gv.x, vv.x, mv.x, b1, b2, lr, ic1, ic2, eps);
            adam1(pv.y, gv.y, vv.y, mv.y, b1, b2, lr, ic1, ic2, eps);
            adam1(pv.z, gv.z, vv.z, mv.z, b1, b2, lr, ic1, ic2, eps);
            adam1(pv.w, gv.w, vv.w, mv.w, b1, b2, lr, ic1, ic2, eps);
            p[i] = pv; v[i] = vv; m[i] = mv;
        }
    }
}
__global__ void adam_scalar(float* p, const float* __restrict__ g, float* v, float* m, int64_t n, float lr, float b1,
                            float b2, float eps, float gscale, float wd, const int* __restrict__ step) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double t = (double)step[0];
    const float ic1 = (float)(1.0 / (1.0 - pow((double)b1, t)));
    const float ic2 = (float)(1.0 / (1.0 - pow((double)b2, t)));
    adam1(p[i], wd != 0.f ? g[i] * gscale + wd * p[i] : g[i] * gscale, v[i], m[i], b1, b2, lr, ic1, ic2, eps);
}

__global__ void counter_inc_k(int* c) { c[0] += 1; }

// Momentum / RMSprop / AdaGrad / AdaDelta over the flat buffer (xs/optim/sgd.py:12-29, rmsprop.py:4-21,
// adagrad.py:4-18, adadelta.py:4-26). s1, s2: per-parameter state (s2 only for AdaDelta).
template <int KIND>
__device__ __forceinline__ void rule1(float& p, float g, float& s1, float& s2, float lr, float rho, float eps) {
    if (KIND == 0) {          // momentum: v = rho v + (1-rho) g ; p -= lr v
        s1 = rho * s1 + (1.f - rho) * g;
        p -= lr * s1;
    } else if (KIND == 1) {   // rmsprop: s = rho s + (1-rho) g^2 ; p -= lr g / sqrt(s + eps)
        s1 = rho * s1 + (1.f - rho) * g * g;
        p -= lr * g / sqrtf(s1 + eps);
    } else if (KIND == 2) {   // adagrad: s += g^2 ; p -= lr g / sqrt(s + eps)
        s1 += g * g;
        p -= lr * g / sqrtf(s1 + eps);
    } else {                  // adadelta: s = rho s + (1-rho) g^2 ; d = sqrt((x+eps)/(s+eps)) g ; p -= d ; x = rho x + (1-rho) d^2
        s1 = rho * s1 + (1.f - rho) * g * g;
        const float d = sqrtf((s2 + eps) / (s1 + eps)) * g;
        p -= d;
        s2 = rho * s2 + (1.f - rho) * d * d;
    }
}

template <int KIND>
__global__ void __launch_bounds__(kThreads)
rule_vec(float4* p, const float4* __restrict__ g, float4* s1, float4* s2, int64_t n4, float lr, float rho, float eps,
         float gscale, float wd) {
    const int64_t base = (int64_t)blockIdx.x * (kThreads * 2) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int64_t i = base + (int64_t)j * kThreads;
        if (i < n4) {
            float4 pv = p[i], a = s1[i], b = KIND == 3 ? s2[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 gv = ldg_stream(g + i);
            gv.x *= gscale; gv.y *= gscale; gv.z *= gscale; gv.w *= gscale;
            if (wd != 0.f) { gv.x += wd * pv.x; gv.y += wd * pv.y; gv.z += wd * pv.z; gv.w += wd * pv.w; }
            rule1<KIND>(pv.x, gv.x, a.x, b.x, lr, rho, eps);
            rule1<KIND>(pv.y, gv.y, a.y, b.y, lr, rho, eps);
            rule1<KIND>(pv.z, gv.z, a.z, b.z, lr, rho, eps);
            rule1<KIND>(pv.w, gv.w, a.w, b.w, lr, rho, eps);
            p[i] = pv;
            s1[i] = a;
            if (KIND == 3) s2[i] = b;
        }
    }
}

}  // namespace

extern "C" {

// p[i] -= lr * (grad_scale * g[i] + weight_decay * p[i]) over the whole flat buffer (n floats)
int xsb_sgd_step(float* p, const float* g, int64_t n, float lr, float grad_scale, float weight_decay, void* stream) {
    if (n <= 0) return XSB_OK;
    cudaStream_t s = xsb_stream(stream);
    const float step = lr * grad_scale, lr_wd = lr * weight_decay;
    if (n % 4 == 0 && aligned16(p) && aligned16(g))
        sgd_vec<<<(unsigned)ceil_div64(n / 4, (int64_t)kThreads * kUnroll), kThreads, 0, s>>>((float4*)p,
                                                                                             (const float4*)g, n / 4, step, lr_wd);
    else
        sgd_scalar<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(p, g, n, step, lr_wd);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// `step_dev`: device int holding t (1-based) for THIS update; see xsb_counter_inc.
int xsb_adam_step(float* p, const float* g, float* v, float* m, int64_t n, float lr, float beta1, float beta2,
                  float eps, float grad_scale, float weight_decay, const int* step_dev, void* stream) {
    if (n <= 0) return XSB_OK;
    cudaStream_t s = xsb_stream(stream);
    if (n % 4 == 0 && aligned16(p) && aligned16(g) && aligned16(v) && aligned16(m))
        adam_vec<<<(unsigned)ceil_div64(n / 4, (int64_t)kThreads * 2), kThreads, 0, s>>>(
            (float4*)p, (const float4*)g, (float4*)v, (float4*)m, n / 4, lr, beta1, beta2, eps, grad_scale, weight_decay,
            step_dev);
    else
        adam_scalar<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(p, g, v, m, n, lr, beta1, beta2, eps, grad_scale,
                                                                 weight_decay, step_dev);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// kind: 0 momentum, 1 rmsprop, 2 adagrad, 3 adadelta. n must be a multiple of 4 (the flat buffers are padded).
int xsb_rule_step(int kind, float* p, const float* g, float* s1, float* s2, int64_t n, float lr, float rho, float eps,
                  float grad_scale, float weight_decay, void* stream) {
    if (n <= 0) return XSB_OK;
    XSB_CHECK_ARG(n % 4 == 0 && aligned16(p) && aligned16(g) && aligned16(s1) && (kind != 3 || aligned16(s2)),
                  "rule_step: flat buffers must be 16-byte aligned with n %% 4 == 0");
    XSB_CHECK_ARG(kind >= 0 && kind <= 3, "rule_step: unknown kind %d", kind);
    cudaStream_t s = xsb_stream(stream);
    const unsigned blocks = (unsigned)ceil_div64(n / 4, (int64_t)kThreads * 2);
    float4 *P = (float4*)p, *S1 = (float4*)s1, *S2 = (float4*)s2;
    const float4* G = (const float4*)g;
    if (kind == 0) rule_vec<0><<<blocks, kThreads, 0, s>>>(P, G, S1, S2, n / 4, lr, rho, eps, grad_scale, weight_decay);
    else if (kind == 1) rule_vec<1><<<blocks, kThreads, 0, s>>>(P, G, S1, S2, n / 4, lr, rho, eps, grad_scale, weight_decay);
    else if (kind == 2) rule_vec<2><<<blocks, kThreads, 0, s>>>(P, G, S1, S2, n / 4, lr, rho, eps, grad_scale, weight_decay);
    else rule_vec<3><<<blocks, kThreads, 0, s>>>(P, G, S1, S2, n / 4, lr, rho, eps, grad_scale, weight_decay);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

int xsb_counter_inc(int* counter_dev, void* stream) {
    counter_inc_k<<<1, 1, 0, xsb_stream(stream)>>>(counter_dev);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

}  // extern "C"
