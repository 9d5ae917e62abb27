// Softmax cross-entropy and MSE loss, forward + backward. Tiny tensors ([N, classes]); latency-bound,
// deterministic single-CTA reductions.
// Reference: xs/nn/functional.py:984-1010 (cross_entropy -> softmax -> log -> nll),
// xs/nn/td_functional.py:73-78,115-128, xs/nn/grad_fn.py:604-620 ((p - onehot)/N * dLoss),
// xs/nn/functional.py:843-873 + grad_fn.py:503-520 (XS mse 'mean' = sum((a-b)^2)/(2*numel), Q9).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ double block_sum_d(double v, double* smem) {
    v = warp_sum_d(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < (blockDim.x >> 5) ? smem[lane] : 0.0;
        r = warp_sum_d(r);
    }
    return r;  // valid in warp 0
}

// one warp per row (looping); probs[N,C] written; loss[0] = | -sum_i log p[i, y_i] / denom |
__global__ void __launch_bounds__(1024)
softmax_xent_fwd_k(const float* __restrict__ z, const int64_t* __restrict__ labels, float* __restrict__ probs,
                   float* __restrict__ loss, int N, int C, float inv_denom, int* bad_label) {
    __shared__ double s_part[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double acc = 0.0;
    for (int row = wid; row < N; row += nw) {
        const float* zr = z + (int64_t)row * C;
        float m = -INFINITY;
        for (int j = lane; j < C; j += 32) m = fmaxf(m, zr[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < C; j += 32) s += expf(zr[j] - m);
        s = warp_sum(s);
        const float inv = 1.0f / s;
        float* pr = probs + (int64_t)row * C;
        for (int j = lane; j < C; j += 32) pr[j] = expf(zr[j] - m) * inv;
        if (lane == 0) {
            const int64_t y = labels[row];
            if (y < 0 || y >= C) {      // the reference's NumPy indexing raises IndexError here: never read out of bounds,
                acc = NAN;              // poison the loss and leave a flag the host turns into that IndexError
                if (bad_label) *bad_label = 1;
            } else {
                acc += (double)(zr[y] - m) - log((double)s);
            }
        }
    }
    const double tot = block_sum_d(acc, s_part);
    if (threadIdx.x == 0) loss[0] = (float)fabs(-tot * (double)inv_denom);
}

// dz (+)= (p - onehot) * scale * dloss[0]
__global__ void softmax_xent_bwd_k(const float* __restrict__ probs, const int64_t* __restrict__ labels,
                                   const float* __restrict__ dloss, float* dz, int N, int C, float scale, int acc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * C) return;
    const int row = (int)(t / C), j = (int)(t % C);
    const float g = (probs[t] - (labels[row] == j ? 1.f : 0.f)) * scale * dloss[0];
    dz[t] = acc ? dz[t] + g : g;
}

__global__ void __launch_bounds__(kThreads)
mse_fwd_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ loss, int64_t n, double scale) {
    __shared__ double s_part[kThreads / 32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = (double)a[i] - (double)b[i];
        acc += d * d;
    }
    const double tot = block_sum_d(acc, s_part);
    if (threadIdx.x == 0) loss[0] = (float)(tot * scale);
}

// da (+)= (a-b)*scale*dloss ; db (+)= -(a-b)*scale*dloss   (either output nullable)
__global__ void mse_bwd_k(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ dloss,
                          float* da, float* db, int64_t n, float scale, int acc_a, int acc_b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = (a[i] - b[i]) * scale * dloss[0];
    if (da) da[i] = acc_a ? da[i] + g : g;
    if (db) db[i] = acc_b ? db[i] - g : -g;
}

}  // namespace

extern "C" {

// reduction: 0 = mean (divide by N = logits.shape[0]), 1 = sum
int xsb_softmax_xent_fwd(const float* logits, const int64_t* labels, float* probs, float* loss, int N, int C,
                         int reduction, int* bad_label, void* stream) {
    XSB_CHECK_ARG(N > 0 && C > 0, "xent: empty input");
    const float inv = reduction == 0 ? 1.0f / (float)N : 1.0f;
    // one CTA (deterministic row order in the final sum), one warp per row: 32 warps for the usual N >= 32
    const int threads = N >= 32 ? 1024 : N >= 16 ? 512 : kThreads;
    softmax_xent_fwd_k<<<1, threads, 0, xsb_stream(stream)>>>(logits, labels, probs, loss, N, C, inv, bad_label);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

int xsb_softmax_xent_bwd(const float* probs, const int64_t* labels, const float* dloss, float* dlogits, int N, int C,
                         int reduction, int accumulate, void* stream) {
    XSB_CHECK_ARG(N > 0 && C > 0, "xent: empty input");
    const float scale = reduction == 0 ? 1.0f / (float)N : 1.0f;
    const int64_t n = (int64_t)N * C;
    softmax_xent_bwd_k<<<(unsigned)ceil_div64(n, 256), 256, 0, xsb_stream(stream)>>>(probs, labels, dloss, dlogits, N,
                                                                                     C, scale, accumulate);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// reduction: 0 = XS 'mean' (sum/2/numel), 1 = 'sum'
int xsb_mse_fwd(const float* a, const float* b, float* loss, int64_t n, int reduction, void* stream) {
    XSB_CHECK_ARG(n > 0, "mse: empty input");
    const double scale = reduction == 0 ? 0.5 / (double)n : 1.0;
    mse_fwd_k<<<1, kThreads, 0, xsb_stream(stream)>>>(a, b, loss, n, scale);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

int xsb_mse_bwd(const float* a, const float* b, const float* dloss, float* da, float* db, int64_t n, int reduction,
                int acc_a, int acc_b, void* stream) {
    XSB_CHECK_ARG(n > 0, "mse: empty input");
    const float scale = reduction == 0 ? 1.0f / (float)n : 1.0f;
    mse_bwd_k<<<(unsigned)ceil_div64(n, 256), 256, 0, xsb_stream(stream)>>>(a, b, dloss, da, db, n, scale, acc_a,
                                                                            acc_b);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

}  // extern "C"
