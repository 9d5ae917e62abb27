// The ops SURVEY.md 8f-4 lists around the hot path: Sigmoid / Tanh, Dropout2D, ZeroPadding2D, LayerNorm, GroupNorm.
// All are HBM-bound elementwise / row-reduction kernels: 128-bit accesses where the layout allows, warp-shuffle
// reductions, one pass per algorithmic read. Algorithmic bytes (e = 4): unary fwd 2e*n, bwd 3e*n; dropout fwd
// 2e*n + n/8, bwd 2e*n + n/8; pad 2e*n; LayerNorm / GroupNorm fwd 2e*n (rows stay in L2 between the statistics and the
// apply sweep), bwd 3e*n.
//
// Reference arithmetic: xs/nn/td_functional.py:63-70 + xs/nn/grad_fn.py:161-170 (sigmoid, tanh);
// xs/nn/functional.py:650-674 + grad_fn.py:319-325 (dropout2d: forward x*mask/keep_prob, backward dy*mask*keep_prob --
// the reference MULTIPLIES by keep_prob on the way back, kept); functional.py:629-647 + grad_fn.py:312-316 (pad2d);
// functional.py:743-775 + grad_fn.py:414-447 (layer_norm: statistics over every axis but 0, per-element gamma/beta);
// functional.py:778-812 + grad_fn.py:450-490 (group_norm: statistics per (sample, group), per-channel gamma/beta).
// Both norms: biased variance (np.var), two passes (mean, then squared deviations), eps inside the sqrt; the input
// gradient is the closed form of the reference's dvar / dmean chain:
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* smem) {   // result valid in every thread
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) r += smem[i];
    return r;
}

// ---------------------------------------------------------------- sigmoid / tanh
template <int KIND>
__device__ __forceinline__ float unary(float x) {
    return KIND == 0 ? 1.f / (1.f + expf(-x)) : tanhf(x);
}
template <int KIND>
__device__ __forceinline__ float unary_grad(float y) {      // derivative expressed in the OUTPUT, as the reference does
    return KIND == 0 ? y * (1.f - y) : 1.f - y * y;
}
template <int KIND>
__global__ void __launch_bounds__(kThreads) unary_fwd_k(const float* __restrict__ x, float* __restrict__ y, int64_t n, int vec) {
    const int64_t n4 = vec ? (n >> 2) : 0;      // unaligned buffers: the scalar loop takes everything
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = i; v < n4; v += stride) {
        const float4 a = ldg_stream(reinterpret_cast<const float4*>(x) + v);
        stg_stream(reinterpret_cast<float4*>(y) + v, make_float4(unary<KIND>(a.x), unary<KIND>(a.y), unary<KIND>(a.z), unary<KIND>(a.w)));
    }
    for (int64_t t = (n4 << 2) + i; t < n; t += stride) y[t] = unary<KIND>(x[t]);
}
template <int KIND>
__global__ void __launch_bounds__(kThreads) unary_bwd_k(const float* __restrict__ dy, const float* __restrict__ y, float* dx,
                                                        int64_t n, int acc, int vec) {
    const int64_t n4 = vec ? (n >> 2) : 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = i; v < n4; v += stride) {
        const float4 g = ldg_stream(reinterpret_cast<const float4*>(dy) + v);
        const float4 o = ldg_stream(reinterpret_cast<const float4*>(y) + v);
        float4 r = make_float4(g.x * unary_grad<KIND>(o.x), g.y * unary_grad<KIND>(o.y), g.z * unary_grad<KIND>(o.z),
                               g.w * unary_grad<KIND>(o.w));
        float4* p = reinterpret_cast<float4*>(dx) + v;
        if (acc) {
            const float4 d = *p;
            r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;
        }
        *p = r;
    }
    for (int64_t t = (n4 << 2) + i; t < n; t += stride) {
        const float r = dy[t] * unary_grad<KIND>(y[t]);
        dx[t] = acc ? dx[t] + r : r;
    }
}

// ---------------------------------------------------------------- dropout (counter-based RNG: stateless, replayable)
__device__ __forceinline__ uint32_t mix64(uint64_t z) {      // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}
// one thread per 8 elements -> one mask byte (bit e = element kept)
__global__ void __launch_bounds__(kThreads) dropout_fwd_k(const float* __restrict__ x, float* __restrict__ y,
                                                          uint8_t* __restrict__ mask, int64_t n, uint32_t threshold,
                                                          float inv_keep, uint64_t seed, const int* __restrict__ step) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 8 >= n) return;
    const uint64_t stream = seed + (step ? (uint64_t)(uint32_t)step[0] * 0x100000001B3ull : 0ull);
    unsigned bits = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int64_t i = b * 8 + e;
        if (i < n) {
            const bool keep = mix64(stream ^ ((uint64_t)i * 0xD6E8FEB86659FD93ull)) < threshold;
            bits |= keep ? (1u << e) : 0u;
            y[i] = keep ? x[i] * inv_keep : 0.f;
        }
    }
    mask[b] = (uint8_t)bits;
}
__global__ void __launch_bounds__(kThreads) dropout_bwd_k(const float* __restrict__ dy, const uint8_t* __restrict__ mask, float* dx,
                                                          int64_t n, float scale, int acc) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 8 >= n) return;
    const unsigned bits = mask[b];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int64_t i = b * 8 + e;
        if (i < n) {
            const float g = ((bits >> e) & 1u) ? dy[i] * scale : 0.f;
            dx[i] = acc ? dx[i] + g : g;
        }
    }
}

// ---------------------------------------------------------------- zero padding (channels-last [N,H,W,C])
__global__ void __launch_bounds__(kThreads) pad2d_fwd_k(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                                        int ph, int pw) {
    const int OH = H + 2 * ph, OW = W + 2 * pw;
    const int64_t total = (int64_t)N * OH * OW * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        int64_t r = t / C;
        const int ow = (int)(r % OW);
        r /= OW;
        const int oh = (int)(r % OH);
        const int n = (int)(r / OH);
        const int h = oh - ph, w = ow - pw;
        y[t] = (h >= 0 && h < H && w >= 0 && w < W) ? x[(((int64_t)n * H + h) * W + w) * C + c] : 0.f;
    }
}
__global__ void __launch_bounds__(kThreads) pad2d_bwd_k(const float* __restrict__ dy, float* dx, int N, int H, int W, int C, int ph,
                                                        int pw, int acc) {
    const int OH = H + 2 * ph, OW = W + 2 * pw;
    const int64_t total = (int64_t)N * H * W * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        int64_t r = t / C;
        const int w = (int)(r % W);
        r /= W;
        const int h = (int)(r % H);
        const int n = (int)(r / H);
        const float g = dy[(((int64_t)n * OH + h + ph) * OW + w + pw) * C + c];
        dx[t] = acc ? dx[t] + g : g;
    }
}

// ---------------------------------------------------------------- LayerNorm on rows [N, F] (logical row-major)
// one block per row: mean, then sum of squared deviations (np.var's two passes; the row is L2-resident), then apply
__global__ void __launch_bounds__(kThreads) layernorm_fwd_k(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int64_t F, float eps) {
    __shared__ float red[kThreads / 32];
    const float* xr = x + (int64_t)blockIdx.x * F;
    float* yr = y + (int64_t)blockIdx.x * F;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < F; i += kThreads) s += xr[i];
    const float mu = block_sum(s, red) / (float)F;
    float q = 0.f;
    for (int64_t i = threadIdx.x; i < F; i += kThreads) {
        const float d = xr[i] - mu;
        q = fmaf(d, d, q);
    }
    const float rs = rsqrtf(block_sum(q, red) / (float)F + eps);
    if (threadIdx.x == 0) {
        mean[blockIdx.x] = mu;
        rstd[blockIdx.x] = rs;
    }
    for (int64_t i = threadIdx.x; i < F; i += kThreads) yr[i] = fmaf(gamma[i], (xr[i] - mu) * rs, beta[i]);
}
__global__ void __launch_bounds__(kThreads) layernorm_bwd_dx_k(const float* __restrict__ dy, const float* __restrict__ x,
                                                               const float* __restrict__ gamma, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, float* dx, int64_t F, int acc) {
    __shared__ float red[kThreads / 32];
    const int64_t off = (int64_t)blockIdx.x * F;
    const float mu = mean[blockIdx.x], rs = rstd[blockIdx.x];
    float s1 = 0.f, s2 = 0.f;
    for (int64_t i = threadIdx.x; i < F; i += kThreads) {
        const float g = dy[off + i] * gamma[i];
        s1 += g;
        s2 = fmaf(g, (x[off + i] - mu) * rs, s2);
    }
    const float m1 = block_sum(s1, red) / (float)F;
    const float m2 = block_sum(s2, red) / (float)F;
    for (int64_t i = threadIdx.x; i < F; i += kThreads) {
        const float xh = (x[off + i] - mu) * rs;
        const float r = rs * (dy[off + i] * gamma[i] - m1 - xh * m2);
        dx[off + i] = acc ? dx[off + i] + r : r;
    }
}
// dgamma[f] (+)= sum_n dy*xhat, dbeta[f] (+)= sum_n dy: one thread per feature, rows walked with coalesced loads
__global__ void __launch_bounds__(kThreads) layernorm_bwd_params_k(const float* __restrict__ dy, const float* __restrict__ x,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   float* dgamma, float* dbeta, int N, int64_t F, int acc_g, int acc_b) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float sg = 0.f, sb = 0.f;
    for (int n = 0; n < N; ++n) {
        const float g = dy[(int64_t)n * F + f];
        sb += g;
        sg = fmaf(g, (x[(int64_t)n * F + f] - mean[n]) * rstd[n], sg);
    }
    if (dgamma) dgamma[f] = acc_g ? dgamma[f] + sg : sg;
    if (dbeta) dbeta[f] = acc_b ? dbeta[f] + sb : sb;
}

// ---------------------------------------------------------------- GroupNorm on channels-last [N, HW, C], G groups
// one block per (sample, group); the group's cg channels are contiguous inside every pixel
__global__ void __launch_bounds__(kThreads) groupnorm_fwd_k(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int HW, int C, int G,
                                                            float eps) {
    __shared__ float red[kThreads / 32];
    const int n = blockIdx.x / G, g = blockIdx.x % G, cg = C / G;
    const int64_t base = (int64_t)n * HW * C + (int64_t)g * cg;
    const int64_t cnt = (int64_t)HW * cg;
    float s = 0.f;
    for (int64_t t = threadIdx.x; t < cnt; t += kThreads) s += x[base + (t / cg) * C + (t % cg)];
    const float mu = block_sum(s, red) / (float)cnt;
    float q = 0.f;
    for (int64_t t = threadIdx.x; t < cnt; t += kThreads) {
        const float d = x[base + (t / cg) * C + (t % cg)] - mu;
        q = fmaf(d, d, q);
    }
    const float rs = rsqrtf(block_sum(q, red) / (float)cnt + eps);
    if (threadIdx.x == 0) {
        mean[blockIdx.x] = mu;
        rstd[blockIdx.x] = rs;
    }
    for (int64_t t = threadIdx.x; t < cnt; t += kThreads) {
        const int j = (int)(t % cg);
        const int64_t i = base + (t / cg) * C + j;
        y[i] = fmaf(gamma[g * cg + j], (x[i] - mu) * rs, beta[g * cg + j]);
    }
}
__global__ void __launch_bounds__(kThreads) groupnorm_bwd_dx_k(const float* __restrict__ dy, const float* __restrict__ x,
                                                               const float* __restrict__ gamma, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, float* dx, int HW, int C, int G, int acc) {
    __shared__ float red[kThreads / 32];
    const int n = blockIdx.x / G, g = blockIdx.x % G, cg = C / G;
    const int64_t base = (int64_t)n * HW * C + (int64_t)g * cg;
    const int64_t cnt = (int64_t)HW * cg;
    const float mu = mean[blockIdx.x], rs = rstd[blockIdx.x];
    float s1 = 0.f, s2 = 0.f;
    for (int64_t t = threadIdx.x; t < cnt; t += kThreads) {
        const int j = (int)(t % cg);
        const int64_t i = base + (t / cg) * C + j;
        const float v = dy[i] * gamma[g * cg + j];
        s1 += v;
        s2 = fmaf(v, (x[i] - mu) * rs, s2);
    }
    const float m1 = block_sum(s1, red) / (float)cnt;
    const float m2 = block_sum(s2, red) / (float)cnt;
    for (int64_t t = threadIdx.x; t < cnt; t += kThreads) {
        const int j = (int)(t % cg);
        const int64_t i = base + (t / cg) * C + j;
        const float r = rs * (dy[i] * gamma[g * cg + j] - m1 - (x[i] - mu) * rs * m2);
        dx[i] = acc ? dx[i] + r : r;
    }
}
// dgamma[c] (+)= sum_{n,p} dy*xhat, dbeta[c] (+)= sum dy. grid (row chunks): a thread owns one channel of its chunk's rows
// (coalesced across the block), partial sums leave through atomics into zeroed / live accumulators.
__global__ void __launch_bounds__(kThreads) groupnorm_bwd_params_k(const float* __restrict__ dy, const float* __restrict__ x,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   float* dgamma, float* dbeta, int64_t R, int HW, int C, int G,
                                                                   int rows_per_block) {
    const int cg = C / G;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > R) r1 = R;
    for (int c = threadIdx.x; c < C; c += kThreads) {
        float sg = 0.f, sb = 0.f;
        for (int64_t r = r0; r < r1; ++r) {
            const int ng = (int)(r / HW) * G + c / cg;
            const float v = dy[r * C + c];
            sb += v;
            sg = fmaf(v, (x[r * C + c] - mean[ng]) * rstd[ng], sg);
        }
        if (dgamma) atomicAdd(dgamma + c, sg);
        if (dbeta) atomicAdd(dbeta + c, sb);
    }
}

unsigned blocks_for(int64_t work, int per_thread = 4) {
    int64_t b = ceil_div64(work, (int64_t)kThreads * per_thread);
    const int64_t cap = (int64_t)xsb_sm_count_cached() * 16;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

// kind 0 = sigmoid, 1 = tanh
int xsb_unary_fwd(int kind, const float* x, float* y, int64_t n, void* stream) {
    if (n <= 0) return XSB_OK;
    XSB_CHECK_ARG(kind == 0 || kind == 1, "unary_fwd: unknown kind %d", kind);
    cudaStream_t s = xsb_stream(stream);
    const int vec = aligned16(x) && aligned16(y);
    if (kind == 0) unary_fwd_k<0><<<blocks_for(n, 16), kThreads, 0, s>>>(x, y, n, vec);
    else unary_fwd_k<1><<<blocks_for(n, 16), kThreads, 0, s>>>(x, y, n, vec);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// dx (+)= dy * f'(.) with the derivative written in the OUTPUT y (sigmoid: y(1-y), tanh: 1-y^2), as grad_fn.py:161-170
int xsb_unary_bwd(int kind, const float* dy, const float* y, float* dx, int64_t n, int accumulate, void* stream) {
    if (n <= 0) return XSB_OK;
    XSB_CHECK_ARG(kind == 0 || kind == 1, "unary_bwd: unknown kind %d", kind);
    cudaStream_t s = xsb_stream(stream);
    const int vec = aligned16(dy) && aligned16(y) && aligned16(dx);
    if (kind == 0) unary_bwd_k<0><<<blocks_for(n, 16), kThreads, 0, s>>>(dy, y, dx, n, accumulate, vec);
    else unary_bwd_k<1><<<blocks_for(n, 16), kThreads, 0, s>>>(dy, y, dx, n, accumulate, vec);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// y = x * keep / keep_prob, mask bit e = element e kept (ceil(n/8) bytes). The keep decision of element i is a hash of
// (seed, *step_dev, i): stateless, so a CUDA-graph replay that bumps the device step counter draws a fresh mask.
int xsb_dropout_fwd(const float* x, float* y, uint8_t* mask, int64_t n, float keep_prob, int64_t seed, const int* step_dev,
                    void* stream) {
    if (n <= 0) return XSB_OK;
    XSB_CHECK_ARG(keep_prob > 0.f && keep_prob <= 1.f && mask, "dropout: keep_prob must be in (0, 1] and mask non-null");
    const double th = (double)keep_prob * 4294967296.0;
    const uint32_t threshold = th >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)th;
    dropout_fwd_k<<<(unsigned)ceil_div64(ceil_div64(n, 8), kThreads), kThreads, 0, xsb_stream(stream)>>>(
        x, y, mask, n, threshold, 1.0f / keep_prob, (uint64_t)seed, step_dev);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// dx (+)= dy * mask * scale (the reference passes scale = keep_prob, grad_fn.py:319-325)
int xsb_dropout_bwd(const float* dy, const uint8_t* mask, float* dx, int64_t n, float scale, int accumulate, void* stream) {
    if (n <= 0) return XSB_OK;
    dropout_bwd_k<<<(unsigned)ceil_div64(ceil_div64(n, 8), kThreads), kThreads, 0, xsb_stream(stream)>>>(dy, mask, dx, n, scale,
                                                                                                       accumulate);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// y[N, H+2ph, W+2pw, C] = zero-padded x[N,H,W,C]; dx (+)= the interior of dy
int xsb_pad2d_fwd(const float* x, float* y, int N, int H, int W, int C, int ph, int pw, void* stream) {
    XSB_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && ph >= 0 && pw >= 0, "pad2d: bad geometry");
    pad2d_fwd_k<<<blocks_for((int64_t)N * (H + 2 * ph) * (W + 2 * pw) * C), kThreads, 0, xsb_stream(stream)>>>(x, y, N, H, W, C, ph, pw);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}
int xsb_pad2d_bwd(const float* dy, float* dx, int N, int H, int W, int C, int ph, int pw, int accumulate, void* stream) {
    XSB_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && ph >= 0 && pw >= 0, "pad2d: bad geometry");
    pad2d_bwd_k<<<blocks_for((int64_t)N * H * W * C), kThreads, 0, xsb_stream(stream)>>>(dy, dx, N, H, W, C, ph, pw, accumulate);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// LayerNorm over rows of x[N, F] (F = every axis but 0, logical row-major); gamma / beta [F]; mean / rstd [N] saved
int xsb_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int N,
                      int64_t F, float eps, void* stream) {
    XSB_CHECK_ARG(N > 0 && F > 0, "layernorm: empty input");
    layernorm_fwd_k<<<N, kThreads, 0, xsb_stream(stream)>>>(x, gamma, beta, y, mean, rstd, F, eps);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}
int xsb_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      float* dx /* nullable */, float* dgamma /* nullable */, float* dbeta /* nullable */, int N, int64_t F,
                      int acc_dx, int acc_dgamma, int acc_dbeta, void* stream) {
    XSB_CHECK_ARG(N > 0 && F > 0, "layernorm: empty input");
    cudaStream_t s = xsb_stream(stream);
    if (dgamma || dbeta) {
        layernorm_bwd_params_k<<<(unsigned)ceil_div64(F, kThreads), kThreads, 0, s>>>(dy, x, mean, rstd, dgamma, dbeta, N, F,
                                                                                     acc_dgamma, acc_dbeta);
        XSB_LAUNCH_CHECK();
    }
    if (dx) {
        layernorm_bwd_dx_k<<<N, kThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, F, acc_dx);
        XSB_LAUNCH_CHECK();
    }
    return XSB_OK;
}

// GroupNorm on channels-last x[N, HW, C]: statistics per (sample, group of C/G channels); gamma / beta [C]; mean / rstd [N*G]
int xsb_groupnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int N,
                      int HW, int C, int G, float eps, void* stream) {
    XSB_CHECK_ARG(N > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, "groupnorm: C must be a multiple of groups");
    groupnorm_fwd_k<<<N * G, kThreads, 0, xsb_stream(stream)>>>(x, gamma, beta, y, mean, rstd, HW, C, G, eps);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}
int xsb_groupnorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      float* dx /* nullable */, float* dgamma /* nullable */, float* dbeta /* nullable */, int N, int HW, int C,
                      int G, int acc_dx, int acc_dgamma, int acc_dbeta, void* stream) {
    XSB_CHECK_ARG(N > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, "groupnorm: C must be a multiple of groups");
    cudaStream_t s = xsb_stream(stream);
    if (dgamma || dbeta) {
        if (dgamma && !acc_dgamma) XSB_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * sizeof(float), s));
        if (dbeta && !acc_dbeta) XSB_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * sizeof(float), s));
        const int64_t R = (int64_t)N * HW;
        int64_t nb = ceil_div64(R, 64);
        const int64_t cap = (int64_t)xsb_sm_count_cached() * 4;
        if (nb > cap) nb = cap;
        const int rows_per_block = (int)ceil_div64(R, nb);
        nb = ceil_div64(R, rows_per_block);
        groupnorm_bwd_params_k<<<(unsigned)nb, kThreads, 0, s>>>(dy, x, mean, rstd, dgamma, dbeta, R, HW, C, G, rows_per_block);
        XSB_LAUNCH_CHECK();
    }
    if (dx) {
        groupnorm_bwd_dx_k<<<N * G, kThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, HW, C, G, acc_dx);
        XSB_LAUNCH_CHECK();
    }
    return XSB_OK;
}

}  // extern "C"
