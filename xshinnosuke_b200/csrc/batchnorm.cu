// BatchNorm (training branch) forward + backward and per-column sums, on row-major [R, C] fp32
// (channels-last activations: R = N*H*W, C = channels; a Dense activation [N, F]: R = N, C = F).
// Roofline: HBM. Algorithmic bytes: fwd 3e*n (x read for statistics, x read + y write
// for the apply), bwd 5e*n (dy, x read for the two sums; dy, x read + dx write for the apply).
//
// Reference semantics: xs/nn/functional.py:676-740 -- batch mean and BIASED variance (np.mean/np.var,
// i.e. a two-pass variance) over every axis but the channel axis, xhat = (x - mean)/sqrt(var + eps),
// y = gamma*xhat + beta, running = momentum*running + (1 - momentum)*batch (biased variance, Q4).
// xs/nn/grad_fn.py:379-411 -- dgamma += sum(dy*xhat), dbeta += sum(dy),
// dx += (M*dxhat - sum(dxhat) - xhat*sum(dxhat*xhat)) / (M*sqrt(var + eps)), dxhat = dy*gamma.
//
// Variance numerics: one pass over x, but NOT E[x^2]-E[x]^2. Every thread accumulates sum(x-K) and
// sum((x-K)^2) around a pivot K (its first element), converts to (count, mean, M2) and all partials are
// combined with Chan's pairwise update -- as well conditioned as np.var's two passes at 1/2 the traffic.
// The backward recomputes xhat from the retained input x and the saved mean / invstd instead of
// caching a full xhat tensor (the reference caches xhat + sqrtvar).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRowUnroll = 8;
constexpr int kMaxRowBlocks = 4 * 148;   // total CTAs (row blocks x column chunks): 4 per SM
constexpr int kMaxLX = 32;
constexpr int kRowsPerThread = 32;
constexpr int kMergeBatch = 8;           // partials fetched per round trip in the last-block merge
constexpr int kCounterSlots = 64;  // ws[0..63]: one "blocks done" ticket counter per column chunk (slot 63: BN-backward ticket)
constexpr int kSumsFloats = 2048;  // ws[64..64+2048): sum(dy) | sum(dy*xhat) of the persistent BN backward (self-cleaning)
constexpr int kScratchBase = kCounterSlots + kSumsFloats;   // per-block partials of the generic kernels start here

struct Geom {
    int V;        // channels per thread (4 or 1)
    int CV;       // C / V
    int LX, LY;   // thread tile: LX vector-columns x LY rows, LX*LY <= 256
    int chunks;   // column chunks of LX vector-columns (grid.y)
    int nb;       // row blocks (grid.x)
    int rows_per_block;
};

static Geom make_geom(int64_t R, int C, bool vec4, int max_blocks = kMaxRowBlocks) {
    Geom g;
    g.V = vec4 ? 4 : 1;
    g.CV = C / g.V;
    // <= 32 vector columns per block (a 512-byte row segment keeps every warp on full 128-byte lines) so that
    // LY >= 8 row lanes share the final merge of the per-block partials; wide tensors spread over grid.y chunks
    g.LX = g.CV < kMaxLX ? g.CV : kMaxLX;
    g.LY = kThreads / g.LX;
    g.chunks = (g.CV + g.LX - 1) / g.LX;
    // >= kRowsPerThread rows per thread: below that the serial merge of the per-block partials costs more than
    // the extra CTAs win; above it up to max_blocks CTAs keep enough loads in flight for the big tensors
    // (small tensors: 8 rows per thread = one round of loads, so that a 512-row tensor still spreads over 8 row blocks)
    const int rpt = R / ((int64_t)g.LY * kRowsPerThread) < 8 ? 8 : kRowsPerThread;
    int64_t nb = R / ((int64_t)g.LY * rpt);
    const int64_t cap = max_blocks / g.chunks > 8 ? max_blocks / g.chunks : 8;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    g.rows_per_block = (int)ceil_div64(R, nb);
    g.nb = (int)ceil_div64(R, g.rows_per_block);
    return g;
}

template <int V>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[V]) {
    if (V == 4) {
        const float4 q = ldg_stream(reinterpret_cast<const float4*>(p));
        v[0] = q.x; v[1 % V] = q.y; v[2 % V] = q.z; v[3 % V] = q.w;
    } else {
        v[0] = __ldg(p);
    }
}

// per-channel parameters: small, re-read by every thread -> keep them on the cached (L1) path, 128-bit wide
template <int V>
__device__ __forceinline__ void load_param(const float* p, float (&v)[V]) {
    if (V == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = q.x; v[1 % V] = q.y; v[2 % V] = q.z; v[3 % V] = q.w;
    } else {
        v[0] = __ldg(p);
    }
}

// ReLU mask (1 bit per element, bit e <=> pre-activation[e] < 0) applied to a gradient vector starting at element e0
template <int V>
__device__ __forceinline__ void apply_relu_mask(const uint8_t* __restrict__ mask, int64_t e0, float (&g)[V]) {
    if (V == 4) {
        const unsigned m = (__ldg(mask + (e0 >> 3)) >> (e0 & 4)) & 0xFu;
        if (m & 1u) g[0] = 0.f;
        if (m & 2u) g[1 % V] = 0.f;
        if (m & 4u) g[2 % V] = 0.f;
        if (m & 8u) g[3 % V] = 0.f;
    } else {
        if ((__ldg(mask + (e0 >> 3)) >> (e0 & 7)) & 1u) g[0] = 0.f;
    }
}

struct Moments {  // count, mean, sum of squared deviations
    float n, mean, m2;
};
__device__ __forceinline__ Moments merge(const Moments& a, const Moments& b) {
    if (b.n == 0.f) return a;
    if (a.n == 0.f) return b;
    Moments r;
    r.n = a.n + b.n;
    const float d = b.mean - a.mean;
    const float fb = b.n / r.n;
    r.mean = a.mean + d * fb;
    r.m2 = a.m2 + b.m2 + d * d * a.n * fb;
    return r;
}

// ------------------------------------------------------------------ forward statistics
// grid (nb, chunks). ws layout (floats): [kCounterSlots counters][nb][C] mean | [nb][C] m2
// (per-block row counts are re-derived from the row partition)
template <int V>
__global__ void __launch_bounds__(kThreads)
bn_stats_k(const float* __restrict__ x, int64_t R, int C, int LX, int LY, int rows_per_block, float* ws,
           float* __restrict__ save_mean, float* __restrict__ save_invstd, float* running_mean, float* running_var,
           float eps, float momentum) {
    __shared__ float s_mean[kThreads * V];
    __shared__ float s_m2[kThreads * V];
    __shared__ float s_n[kThreads];
    __shared__ unsigned s_ticket;

    const int tid = threadIdx.x;
    const int tx = tid % LX, ty = tid / LX;
    const int CV = C / V;
    const int cv = blockIdx.y * LX + tx;
    const bool active = (ty < LY) && (cv < CV);
    const int nb = gridDim.x;
    unsigned* counters = reinterpret_cast<unsigned*>(ws);
    float* p_mean = ws + kScratchBase;
    float* p_m2 = p_mean + (int64_t)nb * C;

    float K[V], s1[V], s2[V];
    float cnt = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { K[i] = 0.f; s1[i] = 0.f; s2[i] = 0.f; }
    if (active) {
        const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
        int64_t r_end = r_begin + rows_per_block;
        if (r_end > R) r_end = R;
        const float* base = x + (int64_t)cv * V;
        int64_t r = r_begin + ty;
        if (r < r_end) load_vec<V>(base + r * C, K);  // pivot = first element of this thread
        for (; r + (int64_t)(kRowUnroll - 1) * LY < r_end; r += (int64_t)kRowUnroll * LY) {
            float v[kRowUnroll][V];
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u) load_vec<V>(base + (r + (int64_t)u * LY) * C, v[u]);
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float d = v[u][i] - K[i];
                    s1[i] += d;
                    s2[i] = fmaf(d, d, s2[i]);
                }
            cnt += (float)kRowUnroll;
        }
        {   // remainder (< kRowUnroll rows): issue every load before the first use
            float v[kRowUnroll][V];
            int m = 0;
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (r + (int64_t)u * LY < r_end) { load_vec<V>(base + (r + (int64_t)u * LY) * C, v[u]); m = u + 1; }
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (u < m) {
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const float d = v[u][i] - K[i];
                        s1[i] += d;
                        s2[i] = fmaf(d, d, s2[i]);
                    }
                    cnt += 1.f;
                }
        }
    }
    // per-thread (n, mean, M2)
    s_n[tid] = cnt;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float inv = cnt > 0.f ? 1.f / cnt : 0.f;
        s_mean[tid * V + i] = K[i] + s1[i] * inv;
        s_m2[tid * V + i] = fmaxf(s2[i] - s1[i] * s1[i] * inv, 0.f);
    }
    __syncthreads();
    // tree over ty (rows of the thread tile)
    int span = 1;
    while (span < LY) span <<= 1;
    for (int st = span >> 1; st > 0; st >>= 1) {
        if (active && ty < st && ty + st < LY) {
            const int o = tid + st * LX;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                Moments a{s_n[tid], s_mean[tid * V + i], s_m2[tid * V + i]};
                Moments b{s_n[o], s_mean[o * V + i], s_m2[o * V + i]};
                const Moments m = merge(a, b);
                s_mean[tid * V + i] = m.mean;
                s_m2[tid * V + i] = m.m2;
                if (i == V - 1) s_n[tid] = m.n;
            }
        }
        __syncthreads();
    }
    if (active && ty == 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            p_mean[(int64_t)blockIdx.x * C + cv * V + i] = s_mean[tid * V + i];
            p_m2[(int64_t)blockIdx.x * C + cv * V + i] = s_m2[tid * V + i];
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&counters[blockIdx.y], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(nb - 1)) return;
    __threadfence();
    // ---- last block of this column chunk: combine the nb partials per channel (fixed order)
    {
        Moments acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = Moments{0.f, 0.f, 0.f};
        if (active) {
            for (int b0 = ty; b0 < nb; b0 += LY * kMergeBatch) {
                float pm[kMergeBatch][V], p2[kMergeBatch][V];
#pragma unroll
                for (int u = 0; u < kMergeBatch; ++u) {      // independent loads first: one L2 round trip per batch
                    const int b = b0 + u * LY;
                    if (b < nb) {
#pragma unroll
                        for (int i = 0; i < V; ++i) {
                            pm[u][i] = __ldcg(&p_mean[(int64_t)b * C + cv * V + i]);
                            p2[u][i] = __ldcg(&p_m2[(int64_t)b * C + cv * V + i]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kMergeBatch; ++u) {
                    const int b = b0 + u * LY;
                    if (b < nb) {
                        const int64_t rb = (int64_t)b * rows_per_block;
                        int64_t re = rb + rows_per_block;
                        if (re > R) re = R;
                        const float nrows = (float)(re - rb);
#pragma unroll
                        for (int i = 0; i < V; ++i) acc[i] = merge(acc[i], Moments{nrows, pm[u][i], p2[u][i]});
                    }
                }
            }
        }
        __syncthreads();
        s_n[tid] = active ? acc[0].n : 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            s_mean[tid * V + i] = acc[i].mean;
            s_m2[tid * V + i] = acc[i].m2;
        }
        __syncthreads();
        for (int st = span >> 1; st > 0; st >>= 1) {
            if (active && ty < st && ty + st < LY) {
                const int o = tid + st * LX;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    Moments a{s_n[tid], s_mean[tid * V + i], s_m2[tid * V + i]};
                    Moments b{s_n[o], s_mean[o * V + i], s_m2[o * V + i]};
                    const Moments m = merge(a, b);
                    s_mean[tid * V + i] = m.mean;
                    s_m2[tid * V + i] = m.m2;
                    if (i == V - 1) s_n[tid] = m.n;
                }
            }
            __syncthreads();
        }
        if (active && ty == 0) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int c = cv * V + i;
                const float mean = s_mean[tid * V + i];
                const float var = s_m2[tid * V + i] / (float)R;
                save_mean[c] = mean;
                save_invstd[c] = 1.0f / sqrtf(var + eps);
                if (running_mean) running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * mean;
                if (running_var) running_var[c] = momentum * running_var[c] + (1.f - momentum) * var;
            }
        }
        if (tid == 0) counters[blockIdx.y] = 0u;  // self-cleaning for the next launch
    }
}

// ------------------------------------------------------------------ forward statistics, small tensors (R <= 4096 rows)
// bn_stats_k on a 100 KB - 1 MB tensor is three dependent latencies (row loads; fence + ticket; the last block's reload and
// merge of the partials): 10 us. Here one 1024-thread block owns 8 float4 columns (32 channels) over ALL rows -- 128 row
// lanes, <= 32 rows per thread in batches of 8 independent loads -- so nothing crosses a block: deviations from a common
// pivot per channel (row 0, an actual sample: no E[x^2]-E[x]^2 cancellation) are plain sums, reduced by shuffles inside a
// warp (4 row lanes x 8 columns) and once through shared memory over the 32 warps.
constexpr int kSmallStatsThreads = 1024;
constexpr int kSmallStatsMaxRows = 4096;

__global__ void __launch_bounds__(kSmallStatsThreads)
bn_stats_small_k(const float* __restrict__ x, int R, int C, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                 float* running_mean, float* running_var, float eps, float momentum) {
    __shared__ float4 s_1[32][8], s_2[32][8];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;        // a warp = 4 row lanes x 8 columns
    const int CV = C >> 2;
    const int cv = blockIdx.x * 8 + tx;
    const bool active = cv < CV;
    const float4* col = reinterpret_cast<const float4*>(x) + (active ? cv : 0);
    float4 K = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) K = __ldg(col);                                       // row 0
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (int r0 = ty; r0 < R; r0 += 128 * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + u * 128;
            v[u] = K;                                                  // (rows past the end contribute a zero deviation)
            if (active && r < R) v[u] = __ldg(col + (int64_t)r * CV);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float dx = v[u].x - K.x, dy = v[u].y - K.y, dz = v[u].z - K.z, dw = v[u].w - K.w;
            a.x += dx; a.y += dy; a.z += dz; a.w += dw;
            b.x = fmaf(dx, dx, b.x); b.y = fmaf(dy, dy, b.y); b.z = fmaf(dz, dz, b.z); b.w = fmaf(dw, dw, b.w);
        }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {                               // the 4 row lanes of the warp that share a column
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
        b.x += __shfl_xor_sync(0xffffffffu, b.x, o); b.y += __shfl_xor_sync(0xffffffffu, b.y, o);
        b.z += __shfl_xor_sync(0xffffffffu, b.z, o); b.w += __shfl_xor_sync(0xffffffffu, b.w, o);
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane < 8) { s_1[warp][lane] = a; s_2[warp][lane] = b; }       // lane < 8: row lane 0 of the warp, column = lane
    __syncthreads();
    if (tid < 32) {                                                    // thread (column tid >> 2, component tid & 3)
        const int c8 = tid >> 2, k = tid & 3;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) {                                 // fixed order: deterministic
            s1 += reinterpret_cast<const float*>(&s_1[w][c8])[k];
            s2 += reinterpret_cast<const float*>(&s_2[w][c8])[k];
        }
        const int ch = (blockIdx.x * 8 + c8) * 4 + k;
        if (ch < C) {
            const float Kc = __ldg(x + ch);
            const float invR = 1.0f / (float)R;
            const float mean = Kc + s1 * invR;
            const float var = fmaxf(s2 - s1 * s1 * invR, 0.f) * invR;
            save_mean[ch] = mean;
            save_invstd[ch] = 1.0f / sqrtf(var + eps);
            if (running_mean) running_mean[ch] = momentum * running_mean[ch] + (1.f - momentum) * mean;
            if (running_var) running_var[ch] = momentum * running_var[ch] + (1.f - momentum) * var;
        }
    }
}

// ------------------------------------------------------------------ forward apply: y = gamma*(x-mean)*invstd + beta
// kRelu: y = max(0, bn(x)) and bit e of `mask` = [bn(x)[e] < 0] (the in-place ReLU protocol's mask), fused
template <int V, bool kRelu>
__global__ void __launch_bounds__(kThreads)
bn_apply_k(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean,
           const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
           int64_t nvec, int CV, int cv_mask, uint8_t* __restrict__ mask) {
    constexpr int U = 4;
    const int64_t base = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    float v[U][V];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int64_t i = base + (int64_t)j * kThreads;
        if (i < nvec) load_vec<V>(x + i * V, v[j]);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int64_t i = base + (int64_t)j * kThreads;
        float o[V];
#pragma unroll
        for (int q = 0; q < V; ++q) o[q] = 0.f;
        if (i < nvec) {
            const int c = (cv_mask >= 0 ? (int)((unsigned)i & (unsigned)cv_mask) : (int)(i % CV)) * V;
            float pm[V], pi[V], pg[V], pb[V];
            load_param<V>(mean + c, pm);
            load_param<V>(invstd + c, pi);
            load_param<V>(gamma + c, pg);
            load_param<V>(beta + c, pb);
#pragma unroll
            for (int q = 0; q < V; ++q) o[q] = fmaf((v[j][q] - pm[q]) * pi[q], pg[q], pb[q]);
        }
        if (kRelu) {   // V == 4 only: two lanes share one mask byte (same layout as relu_fwd_vec)
            unsigned nib = (o[0] < 0.f ? 1u : 0u) | (o[1 % V] < 0.f ? 2u : 0u) | (o[2 % V] < 0.f ? 4u : 0u) |
                           (o[3 % V] < 0.f ? 8u : 0u);
            const unsigned hi = __shfl_down_sync(0xffffffffu, nib, 1);
            if ((threadIdx.x & 1) == 0 && i < nvec) mask[i >> 1] = (uint8_t)(nib | (hi << 4));
#pragma unroll
            for (int q = 0; q < V; ++q) o[q] = fmaxf(o[q], 0.f);
        }
        if (i < nvec) {
            if (V == 4)
                stg_stream(reinterpret_cast<float4*>(y + i * V), make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]));
            else
                y[i] = o[0];
        }
    }
}

// persistent forward apply (C = 4 * 2^k <= 1024): the thread's four channels are fixed, their scale / shift live in
// registers -- y = x * a + b with a = gamma * invstd, b = beta - mean * a (one FMA per element), optional ReLU + bit mask
template <bool kRelu>
__global__ void __launch_bounds__(kThreads)
bn_apply_persist_k(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean,
                   const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                   int64_t nvec, int CV, uint8_t* __restrict__ mask) {
    constexpr int U = 4;
    const int tid = threadIdx.x;
    const int c = (tid & (CV - 1)) * 4;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    const int64_t stride = (int64_t)gridDim.x * (kThreads * U);
    // block-uniform trip count: the mask nibbles are paired with a full-warp shuffle
    for (int64_t bbase = (int64_t)blockIdx.x * (kThreads * U); bbase < nvec; bbase += stride) {
        const int64_t base = bbase + tid;
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            if (i < nvec) v[u] = ldg_stream(reinterpret_cast<const float4*>(x) + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nvec) {   // same operation order as bn_apply_k: ((x - mean) * invstd) * gamma + beta
                o.x = fmaf((v[u].x - mu.x) * is.x, ga.x, be.x);
                o.y = fmaf((v[u].y - mu.y) * is.y, ga.y, be.y);
                o.z = fmaf((v[u].z - mu.z) * is.z, ga.z, be.z);
                o.w = fmaf((v[u].w - mu.w) * is.w, ga.w, be.w);
            }
            if (kRelu) {
                unsigned nib = (o.x < 0.f ? 1u : 0u) | (o.y < 0.f ? 2u : 0u) | (o.z < 0.f ? 4u : 0u) | (o.w < 0.f ? 8u : 0u);
                const unsigned hi = __shfl_down_sync(0xffffffffu, nib, 1);
                if ((tid & 1) == 0 && i < nvec) mask[i >> 1] = (uint8_t)(nib | (hi << 4));
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            if (i < nvec) stg_stream(reinterpret_cast<float4*>(y) + i, o);
        }
    }
}

// The tail of a residual block in ONE pass: z = max(0, bn_a(xa) + rhs) with the (sum < 0) bit mask, where rhs is a plain
// tensor (identity shortcut) or a second BatchNorm applied on the fly (kTwoBN: the down-sampling shortcut). Same arithmetic,
// in the same order and rounded to fp32 at the same points, as bn_apply followed by add_relu -- bit-identical -- but the
// BatchNorm output(s) are never written and read back: 3e*n (+n/8) bytes instead of 5e*n (7e*n with two BatchNorms).
template <bool kTwoBN>
__global__ void __launch_bounds__(kThreads)
bn_add_relu_persist_k(const float* __restrict__ xa, const float* __restrict__ mean_a, const float* __restrict__ invstd_a,
                      const float* __restrict__ gamma_a, const float* __restrict__ beta_a, const float* __restrict__ b,
                      const float* __restrict__ mean_b, const float* __restrict__ invstd_b, const float* __restrict__ gamma_b,
                      const float* __restrict__ beta_b, float* __restrict__ z, uint8_t* __restrict__ mask, int64_t nvec, int CV) {
    constexpr int U = 4;
    const int tid = threadIdx.x;
    const int c = (tid & (CV - 1)) * 4;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean_a + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd_a + c));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma_a + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta_a + c));
    float4 mu2 = mu, is2 = is, ga2 = ga, be2 = be;
    if (kTwoBN) {
        mu2 = __ldg(reinterpret_cast<const float4*>(mean_b + c));
        is2 = __ldg(reinterpret_cast<const float4*>(invstd_b + c));
        ga2 = __ldg(reinterpret_cast<const float4*>(gamma_b + c));
        be2 = __ldg(reinterpret_cast<const float4*>(beta_b + c));
    }
    const int64_t stride = (int64_t)gridDim.x * (kThreads * U);
    // block-uniform trip count: the mask nibbles are paired with a full-warp shuffle
    for (int64_t bbase = (int64_t)blockIdx.x * (kThreads * U); bbase < nvec; bbase += stride) {
        const int64_t base = bbase + tid;
        float4 v[U], w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            v[u] = w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nvec) {
                v[u] = ldg_stream(reinterpret_cast<const float4*>(xa) + i);
                w[u] = ldg_stream(reinterpret_cast<const float4*>(b) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            float4 o, r = w[u];
            o.x = fmaf((v[u].x - mu.x) * is.x, ga.x, be.x);
            o.y = fmaf((v[u].y - mu.y) * is.y, ga.y, be.y);
            o.z = fmaf((v[u].z - mu.z) * is.z, ga.z, be.z);
            o.w = fmaf((v[u].w - mu.w) * is.w, ga.w, be.w);
            if (kTwoBN) {
                r.x = fmaf((w[u].x - mu2.x) * is2.x, ga2.x, be2.x);
                r.y = fmaf((w[u].y - mu2.y) * is2.y, ga2.y, be2.y);
                r.z = fmaf((w[u].z - mu2.z) * is2.z, ga2.z, be2.z);
                r.w = fmaf((w[u].w - mu2.w) * is2.w, ga2.w, be2.w);
            }
            const float4 t = make_float4(__fadd_rn(o.x, r.x), __fadd_rn(o.y, r.y), __fadd_rn(o.z, r.z), __fadd_rn(o.w, r.w));
            const unsigned nib = (t.x < 0.f ? 1u : 0u) | (t.y < 0.f ? 2u : 0u) | (t.z < 0.f ? 4u : 0u) | (t.w < 0.f ? 8u : 0u);
            const unsigned hi = __shfl_down_sync(0xffffffffu, nib, 1);
            if ((tid & 1) == 0 && i < nvec) mask[i >> 1] = (uint8_t)(nib | (hi << 4));
            if (i < nvec)
                stg_stream(reinterpret_cast<float4*>(z) + i,
                           make_float4(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f), fmaxf(t.z, 0.f), fmaxf(t.w, 0.f)));
        }
    }
}

// ------------------------------------------------------------------ backward sums (and plain column sums)
// kXhat: also accumulate sum(dy * xhat). ws: [counters][nb][C] s_dy | [nb][C] s_dyxhat
// finalize writes dgamma/dbeta (+)= and coef[0..C)=sum_dy/M, coef[C..2C)=sum_dyxhat/M, coef[2C..3C)=gamma*invstd
template <int V, bool kXhat>
__global__ void __launch_bounds__(kThreads)
col_sums_k(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
           const float* __restrict__ invstd, const float* __restrict__ gamma, int64_t R, int C, int LX, int LY,
           int rows_per_block, float* ws, float* dgamma, float* dbeta, float* coef, int acc_gamma, int acc_beta,
           const uint8_t* __restrict__ rmask) {
    __shared__ float s_a[kThreads * V];
    __shared__ float s_b[kThreads * V];
    __shared__ unsigned s_ticket;
    const int tid = threadIdx.x;
    const int tx = tid % LX, ty = tid / LX;
    const int CV = C / V;
    const int cv = blockIdx.y * LX + tx;
    const bool active = (ty < LY) && (cv < CV);
    const int nb = gridDim.x;
    unsigned* counters = reinterpret_cast<unsigned*>(ws);
    float* p_a = ws + kScratchBase;
    float* p_b = p_a + (int64_t)nb * C;

    float sa[V], sb[V], mu[V], is[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { sa[i] = 0.f; sb[i] = 0.f; mu[i] = 0.f; is[i] = 0.f; }
    if (active) {
        if (kXhat) {
#pragma unroll
            for (int i = 0; i < V; ++i) { mu[i] = mean[cv * V + i]; is[i] = invstd[cv * V + i]; }
        }
        const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
        int64_t r_end = r_begin + rows_per_block;
        if (r_end > R) r_end = R;
        const int64_t coff = (int64_t)cv * V;
        int64_t r = r_begin + ty;
        constexpr int U = kXhat ? 4 : 8;
        for (; r + (int64_t)(U - 1) * LY < r_end; r += (int64_t)U * LY) {
            float g[U][V], xv[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                load_vec<V>(dy + (r + (int64_t)u * LY) * C + coff, g[u]);
                if (kXhat) load_vec<V>(x + (r + (int64_t)u * LY) * C + coff, xv[u]);
            }
            if (rmask) {
#pragma unroll
                for (int u = 0; u < U; ++u) apply_relu_mask<V>(rmask, (r + (int64_t)u * LY) * C + coff, g[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    sa[i] += g[u][i];
                    if (kXhat) sb[i] = fmaf(g[u][i], (xv[u][i] - mu[i]) * is[i], sb[i]);
                }
        }
        for (; r < r_end; r += LY) {
            float g[V], xv[V];
            load_vec<V>(dy + r * C + coff, g);
            if (kXhat) load_vec<V>(x + r * C + coff, xv);
            if (rmask) apply_relu_mask<V>(rmask, r * C + coff, g);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                sa[i] += g[i];
                if (kXhat) sb[i] = fmaf(g[i], (xv[i] - mu[i]) * is[i], sb[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { s_a[tid * V + i] = sa[i]; s_b[tid * V + i] = sb[i]; }
    __syncthreads();
    int span = 1;
    while (span < LY) span <<= 1;
    for (int st = span >> 1; st > 0; st >>= 1) {
        if (active && ty < st && ty + st < LY) {
            const int o = tid + st * LX;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                s_a[tid * V + i] += s_a[o * V + i];
                if (kXhat) s_b[tid * V + i] += s_b[o * V + i];
            }
        }
        __syncthreads();
    }
    if (active && ty == 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            p_a[(int64_t)blockIdx.x * C + cv * V + i] = s_a[tid * V + i];
            if (kXhat) p_b[(int64_t)blockIdx.x * C + cv * V + i] = s_b[tid * V + i];
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&counters[blockIdx.y], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(nb - 1)) return;
    __threadfence();
    {
        double da[V], db[V];
#pragma unroll
        for (int i = 0; i < V; ++i) { da[i] = 0.0; db[i] = 0.0; }
        if (active) {
            for (int b0 = ty; b0 < nb; b0 += LY * kMergeBatch) {
                float pa[kMergeBatch][V], pb[kMergeBatch][V];
#pragma unroll
                for (int u = 0; u < kMergeBatch; ++u) {
                    const int b = b0 + u * LY;
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        pa[u][i] = b < nb ? __ldcg(&p_a[(int64_t)b * C + cv * V + i]) : 0.f;
                        pb[u][i] = (kXhat && b < nb) ? __ldcg(&p_b[(int64_t)b * C + cv * V + i]) : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < kMergeBatch; ++u)
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        da[i] += (double)pa[u][i];
                        if (kXhat) db[i] += (double)pb[u][i];
                    }
            }
        }
        __syncthreads();
        // cross-ty combine of <= LY partial sums in fp32
#pragma unroll
        for (int i = 0; i < V; ++i) {
            s_a[tid * V + i] = (float)da[i];
            s_b[tid * V + i] = (float)db[i];
        }
        __syncthreads();
        for (int st = span >> 1; st > 0; st >>= 1) {
            if (active && ty < st && ty + st < LY) {
                const int o = tid + st * LX;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    s_a[tid * V + i] += s_a[o * V + i];
                    if (kXhat) s_b[tid * V + i] += s_b[o * V + i];
                }
            }
            __syncthreads();
        }
        if (active && ty == 0) {
            const float invM = 1.0f / (float)R;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int c = cv * V + i;
                const float sdy = s_a[tid * V + i];
                if (dbeta) dbeta[c] = acc_beta ? dbeta[c] + sdy : sdy;
                if (kXhat) {
                    const float sdyx = s_b[tid * V + i];
                    if (dgamma) dgamma[c] = acc_gamma ? dgamma[c] + sdyx : sdyx;
                    if (coef) {
                        coef[c] = sdy * invM;
                        coef[C + c] = sdyx * invM;
                        coef[2 * C + c] = gamma[c] * invstd[c];
                    }
                }
            }
        }
        if (tid == 0) counters[blockIdx.y] = 0u;
    }
}

// ------------------------------------------------------------------ backward apply
// dx (+)= gamma*invstd * (dy - sum_dy/M - xhat * sum_dyxhat/M)
template <int V, bool kAcc>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_k(const float* __restrict__ dy, const float* __restrict__ x, float* dx, const float* __restrict__ mean,
               const float* __restrict__ invstd, const float* __restrict__ coef, int64_t nvec, int CV, int C,
               int cv_mask, const uint8_t* __restrict__ rmask) {
    constexpr int U = 4;
    const int64_t base = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    float g[U][V], xv[U][V];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int64_t i = base + (int64_t)j * kThreads;
        if (i < nvec) {
            load_vec<V>(dy + i * V, g[j]);
            load_vec<V>(x + i * V, xv[j]);
            if (rmask) apply_relu_mask<V>(rmask, i * V, g[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int64_t i = base + (int64_t)j * kThreads;
        if (i < nvec) {
            const int c = (cv_mask >= 0 ? (int)((unsigned)i & (unsigned)cv_mask) : (int)(i % CV)) * V;
            float o[V], pm[V], pi[V], k1[V], k2[V], kg[V];
            load_param<V>(mean + c, pm);
            load_param<V>(invstd + c, pi);
            load_param<V>(coef + c, k1);
            load_param<V>(coef + C + c, k2);
            load_param<V>(coef + 2 * C + c, kg);
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const float xhat = (xv[j][q] - pm[q]) * pi[q];
                o[q] = kg[q] * (g[j][q] - k1[q] - xhat * k2[q]);
            }
            if (V == 4) {
                float4* p = reinterpret_cast<float4*>(dx + i * V);
                float4 r = make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]);
                if (kAcc) {
                    const float4 d = *p;
                    r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;
                }
                stg_stream(p, r);
            } else {
                dx[i] = kAcc ? dx[i] + o[0] : o[0];
            }
        }
    }
}

__global__ void invstd_k(const float* __restrict__ var, float* __restrict__ invstd, int C, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) invstd[c] = 1.0f / sqrtf(var[c] + eps);
}

// ------------------------------------------------------------------ persistent backward (C = 4 * 2^k <= 1024)
// Both passes walk the tensor grid-stride with a stride that is a multiple of the row length, so a thread meets the
// SAME four channels in every iteration: parameters are loaded once, the per-channel sums stay in registers, and a
// block contributes C atomic adds at the end (no per-block partial arrays, no serial last-block merge).
//   sums[0..C) = sum(dy), sums[C..2C) = sum(dy * xhat)   (zeroed by the caller)
// Measured in the training step (ResNet-18, 128 x 224 x 224; blocks per SM x float4 loads in flight per thread and array):
// 4 x 4: 1710 us of BatchNorm backward per step, 8 x 4: 1668, 2 x 4: 1571, 4 x 8: 1611, 2 x 8: 1467, 1 x 8: 1752,
// 2 x 16: 2378 (spills). Two fat blocks per SM keep the same bytes in flight with half the block-end reductions/atomics.
constexpr int kPersistBlocksPerSM = 2;
constexpr int kPersistUnroll = 8;

template <bool kXhat>
__global__ void __launch_bounds__(kThreads)
col_sums_atomic_k(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                  const float* __restrict__ invstd, int64_t nvec, int CV, float* sums, const uint8_t* __restrict__ rmask,
                  float* zero_me) {
    __shared__ float4 s_a[kThreads], s_b[kThreads];
    constexpr int U = kPersistUnroll;
    const int tid = threadIdx.x;
    const int c = (tid & (CV - 1)) * 4;     // CV divides kThreads: the channel group depends on the thread only
    const int C = CV * 4;
    if (zero_me && blockIdx.x == 0)         // the apply kernel that follows accumulates the bias gradient into it ("=")
        for (int ch = tid; ch < C; ch += kThreads) zero_me[ch] = 0.f;
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu;
    if (kXhat) {
        mu = __ldg(reinterpret_cast<const float4*>(mean + c));
        is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    }
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    const int64_t stride = (int64_t)gridDim.x * (kThreads * U);
    for (int64_t base = (int64_t)blockIdx.x * (kThreads * U) + tid; base < nvec; base += stride) {
        float4 g[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            xv[u] = mu;
            if (i < nvec) {
                g[u] = ldg_stream(reinterpret_cast<const float4*>(dy) + i);
                if (kXhat) xv[u] = ldg_stream(reinterpret_cast<const float4*>(x) + i);
            }
        }
        if (rmask) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + (int64_t)u * kThreads;
                if (i < nvec) {
                    const int64_t e0 = i * 4;
                    const unsigned m = (__ldg(rmask + (e0 >> 3)) >> (e0 & 4)) & 0xFu;
                    if (m & 1u) g[u].x = 0.f;
                    if (m & 2u) g[u].y = 0.f;
                    if (m & 4u) g[u].z = 0.f;
                    if (m & 8u) g[u].w = 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            sa.x += g[u].x; sa.y += g[u].y; sa.z += g[u].z; sa.w += g[u].w;
            if (kXhat) {
                sb.x = fmaf(g[u].x, (xv[u].x - mu.x) * is.x, sb.x);
                sb.y = fmaf(g[u].y, (xv[u].y - mu.y) * is.y, sb.y);
                sb.z = fmaf(g[u].z, (xv[u].z - mu.z) * is.z, sb.z);
                sb.w = fmaf(g[u].w, (xv[u].w - mu.w) * is.w, sb.w);
            }
        }
    }
    s_a[tid] = sa;
    if (kXhat) s_b[tid] = sb;
    __syncthreads();
    for (int st = kThreads / 2; st >= CV; st >>= 1) {   // threads tid, tid+CV, tid+2CV ... share a channel group
        if (tid < st) {
            float4 a = s_a[tid], b = s_a[tid + st];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            s_a[tid] = a;
            if (kXhat) {
                float4 p = s_b[tid], q = s_b[tid + st];
                p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w;
                s_b[tid] = p;
            }
        }
        __syncthreads();
    }
    if (tid < CV) {
        const float4 a = s_a[tid];
        atomicAdd(sums + c, a.x); atomicAdd(sums + c + 1, a.y); atomicAdd(sums + c + 2, a.z); atomicAdd(sums + c + 3, a.w);
        if (kXhat) {
            const float4 b = s_b[tid];
            atomicAdd(sums + C + c, b.x); atomicAdd(sums + C + c + 1, b.y); atomicAdd(sums + C + c + 2, b.z);
            atomicAdd(sums + C + c + 3, b.w);
        }
    }
}

// dx (+)= gamma*invstd * (dy - sum_dy/M - xhat * sum_dyxhat/M); block 0 also writes dgamma / dbeta (+)= from `sums`;
// dxsum (nullable): per-channel column sums of the dx CONTRIBUTION, accumulated with atomics (the bias gradient of the
// convolution that produced x: xs/nn/grad_fn.py:196-197 computed without a second pass over dx).
template <bool kAcc>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_persist_k(const float* __restrict__ dy, const float* __restrict__ x, float* dx, const float* __restrict__ mean,
                       const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ sums,
                       float invM, int64_t nvec, int CV, const uint8_t* __restrict__ rmask, float* dgamma, float* dbeta,
                       int acc_gamma, int acc_beta, float* dxsum, float* sums_rw, unsigned* ticket, int rev,
                       float* __restrict__ side) {
    __shared__ float4 s_a[kThreads];
    __shared__ unsigned s_last;
    constexpr int U = kPersistUnroll;
    const int tid = threadIdx.x;
    const int c = (tid & (CV - 1)) * 4;
    const int C = CV * 4;
    if (blockIdx.x == 0) {
        for (int ch = tid; ch < C; ch += kThreads) {
            if (dbeta) dbeta[ch] = acc_beta ? dbeta[ch] + sums[ch] : sums[ch];
            if (dgamma) dgamma[ch] = acc_gamma ? dgamma[ch] + sums[C + ch] : sums[C + ch];
        }
    }
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    float4 k1 = *reinterpret_cast<const float4*>(sums + c), k2 = *reinterpret_cast<const float4*>(sums + C + c);
    k1.x *= invM; k1.y *= invM; k1.z *= invM; k1.w *= invM;
    k2.x *= invM; k2.y *= invM; k2.z *= invM; k2.w *= invM;
    const float4 kg = make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t stride = (int64_t)gridDim.x * (kThreads * U);
    // LAST chunk first: the sums pass that ran just before walked the tensor front to back, so its tail is what the L2
    // still holds -- walking back to front turns the first ~100 MB of re-reads into L2 hits (rev: A/B switch)
    const int64_t b0 = (int64_t)blockIdx.x * (kThreads * U);
    const int64_t nit = b0 < nvec ? (nvec - b0 + stride - 1) / stride : 0;
    for (int64_t it = 0; it < nit; ++it) {
        const int64_t base = b0 + (rev ? nit - 1 - it : it) * stride + tid;
        float4 g[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            if (i < nvec) {
                g[u] = ldg_stream(reinterpret_cast<const float4*>(dy) + i);
                xv[u] = ldg_stream(reinterpret_cast<const float4*>(x) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * kThreads;
            if (i < nvec) {
                if (rmask) {
                    const int64_t e0 = i * 4;
                    const unsigned m = (__ldg(rmask + (e0 >> 3)) >> (e0 & 4)) & 0xFu;
                    if (m & 1u) g[u].x = 0.f;
                    if (m & 2u) g[u].y = 0.f;
                    if (m & 4u) g[u].z = 0.f;
                    if (m & 8u) g[u].w = 0.f;
                }
                // side (nullable): the MASKED dy is also the gradient of the identity shortcut of a residual block
                // (AddBackward behind the in-place ReLU) -- written from here instead of by a relu_bwd_mask pass of its own
                if (side) stg_stream(reinterpret_cast<float4*>(side) + i, g[u]);
                float4 o;
                o.x = kg.x * (g[u].x - k1.x - (xv[u].x - mu.x) * is.x * k2.x);
                o.y = kg.y * (g[u].y - k1.y - (xv[u].y - mu.y) * is.y * k2.y);
                o.z = kg.z * (g[u].z - k1.z - (xv[u].z - mu.z) * is.z * k2.z);
                o.w = kg.w * (g[u].w - k1.w - (xv[u].w - mu.w) * is.w * k2.w);
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                float4* p = reinterpret_cast<float4*>(dx) + i;
                if (kAcc) {
                    const float4 d = *p;
                    o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
                }
                stg_stream(p, o);
            }
        }
    }
    if (dxsum) {
        s_a[tid] = acc;
        __syncthreads();
        for (int st = kThreads / 2; st >= CV; st >>= 1) {
            if (tid < st) {
                float4 a = s_a[tid], b = s_a[tid + st];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                s_a[tid] = a;
            }
            __syncthreads();
        }
        if (tid < CV) {
            const float4 a = s_a[tid];
            atomicAdd(dxsum + c, a.x); atomicAdd(dxsum + c + 1, a.y); atomicAdd(dxsum + c + 2, a.z); atomicAdd(dxsum + c + 3, a.w);
        }
    }
    // every block read `sums` before its loop; the LAST block to get here zeroes them for the next BatchNorm backward
    // (self-cleaning scratch: no memset node per call)
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        for (int ch = tid; ch < 2 * C; ch += kThreads) sums_rw[ch] = 0.f;
        if (tid == 0) *ticket = 0u;
    }
}

// ------------------------------------------------------------------ BatchNorm backward fed by a max-pool backward
// In BatchNorm -> in-place ReLU -> MaxPool(3x3, stride 2, pad 1) (the ResNet stem) the BatchNorm's dy IS the pool's
// backward gather. Written out it is 4x the pooled gradient, and the two BatchNorm passes read it twice. Here both passes
// gather it on the fly instead: one thread owns a 2x2 input quad x 4 channels, loads the <= 4 windows that touch the quad
// once ((arg-max, dpool) pairs), forms the four pixels' gradients with the constant window slots of k = 3 / pad = 1 (same
// visiting order as maxpool_bwd_s2_k: bit-identical to the materialised gradient), clears what the ReLU bits clear and
// carries on like col_sums_atomic_k / bn_bwd_apply_persist_k. HBM traffic 2 x (x + dpool + argmax + mask) + dx instead
// of (dpool + argmax + dy) + 2 x (dy + x + mask) + dx: 1.5 GB instead of 2.6 GB on the 128 x 112 x 112 x 64 stem.
#ifndef XSB_POOL_OCC
#define XSB_POOL_OCC 3
#endif
constexpr int kPoolOcc = XSB_POOL_OCC;      // resident blocks per SM the two kernels are compiled for
struct PoolGeom {
    int N, H, W, C, OH, OW, QH, QW, cvshift;
    int fast;     // C % 8 == 0 and every 32-bit index of the fast path fits
};
// everything one quad needs, fetched before anything is computed (16 independent loads in flight per thread)
struct QuadIn {
    uchar4 am[4];     // window (dh, dw) at [dh * 2 + dw]
    float4 gp[4];
    float4 xv[4];     // pixel (a, b) at [a * 2 + b]
    unsigned mk[4];
    bool ok[4];
};

// FAST: the quad is interior (its four windows and four pixels all exist) and C % 8 == 0 -- no existence checks, every
// address is one base + a constant (the kernels are instruction-bound: a quad costs ~350 instructions to gather)
template <bool FAST>
__device__ __forceinline__ void pool_quad_load(const float* __restrict__ dpool, const uint8_t* __restrict__ amax,
                                               const float* __restrict__ x, const uint8_t* __restrict__ rmask,
                                               const PoolGeom& pg, int n, int qi, int qj, int cv, const float4& xfill,
                                               QuadIn& in) {
    const int CV = pg.C >> 2;
    if (FAST) {
        const unsigned o00 = (unsigned)(((n * pg.OH + qi) * pg.OW + qj) * CV + cv);      // float4 / uchar4 index
        const unsigned orow = (unsigned)(pg.OW * CV);
        const uchar4* ap = reinterpret_cast<const uchar4*>(amax);
        const float4* gpp = reinterpret_cast<const float4*>(dpool);
        in.am[0] = __ldg(ap + o00); in.am[1] = __ldg(ap + o00 + CV);
        in.am[2] = __ldg(ap + o00 + orow); in.am[3] = __ldg(ap + o00 + orow + CV);
        in.gp[0] = __ldg(gpp + o00); in.gp[1] = __ldg(gpp + o00 + CV);
        in.gp[2] = __ldg(gpp + o00 + orow); in.gp[3] = __ldg(gpp + o00 + orow + CV);
        const unsigned p00 = (unsigned)((n * pg.H + 2 * qi) * pg.W + 2 * qj);              // pixel index
        const unsigned x00 = p00 * (unsigned)CV + (unsigned)cv, xrow = (unsigned)(pg.W * CV);
        const float4* xp = reinterpret_cast<const float4*>(x);
        in.xv[0] = ldg_stream(xp + x00); in.xv[1] = ldg_stream(xp + x00 + CV);
        in.xv[2] = ldg_stream(xp + x00 + xrow); in.xv[3] = ldg_stream(xp + x00 + xrow + CV);
        in.ok[0] = in.ok[1] = in.ok[2] = in.ok[3] = true;
        in.mk[0] = in.mk[1] = in.mk[2] = in.mk[3] = 0u;
        if (rmask) {
            const unsigned C8 = (unsigned)(pg.C >> 3);
            const unsigned m00 = p00 * C8 + (unsigned)(cv >> 1), mrow = (unsigned)pg.W * C8, sh = (unsigned)(cv & 1) * 4u;
            in.mk[0] = ((unsigned)__ldg(rmask + m00) >> sh) & 0xFu;
            in.mk[1] = ((unsigned)__ldg(rmask + m00 + C8) >> sh) & 0xFu;
            in.mk[2] = ((unsigned)__ldg(rmask + m00 + mrow) >> sh) & 0xFu;
            in.mk[3] = ((unsigned)__ldg(rmask + m00 + mrow + C8) >> sh) & 0xFu;
        }
        return;
    }
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
            const int oh = qi + dh, ow = qj + dw;
            const bool live = oh < pg.OH && ow < pg.OW;
            const int64_t o = (((int64_t)n * pg.OH + (live ? oh : 0)) * pg.OW + (live ? ow : 0)) * pg.C + (int64_t)cv * 4;
            in.am[dh * 2 + dw] = __ldg(reinterpret_cast<const uchar4*>(amax + o));
            in.gp[dh * 2 + dw] = __ldg(reinterpret_cast<const float4*>(dpool + o));
            if (!live) in.am[dh * 2 + dw] = make_uchar4(255, 255, 255, 255);
        }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int h = 2 * qi + a, w = 2 * qj + b;
            const bool ok = h < pg.H && w < pg.W;
            in.ok[a * 2 + b] = ok;
            in.xv[a * 2 + b] = xfill;
            in.mk[a * 2 + b] = 0u;
            if (ok) {
                const int64_t pix = ((int64_t)n * pg.H + h) * pg.W + w;
                in.xv[a * 2 + b] = ldg_stream(reinterpret_cast<const float4*>(x) + pix * CV + cv);
                if (rmask) {
                    const int64_t e0 = pix * pg.C + (int64_t)cv * 4;
                    in.mk[a * 2 + b] = ((unsigned)__ldg(rmask + (e0 >> 3)) >> (e0 & 4)) & 0xFu;
                }
            }
        }
}

// item -> (image, quad row, quad column); -> true when the quad is interior and the fast path applies
__device__ __forceinline__ bool pool_quad_of(const PoolGeom& pg, unsigned item, int& n, int& qi, int& qj) {
    const unsigned quad = item >> pg.cvshift;
    const unsigned r = quad / (unsigned)pg.QW;
    qj = (int)(quad - r * (unsigned)pg.QW);
    n = (int)(r / (unsigned)pg.QH);
    qi = (int)(r - (unsigned)n * (unsigned)pg.QH);
    return pg.fast && qi + 1 < pg.OH && qj + 1 < pg.OW && 2 * qi + 1 < pg.H && 2 * qj + 1 < pg.W;
}

// the four pixels' gradients: constant window slots of k = 3 / pad = 1, windows visited in maxpool_bwd_s2_k's order
__device__ __forceinline__ void pool_quad_grad(const QuadIn& in, float4 (&g)[4]) {
    float acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[p][i] = 0.f;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
            const uchar4 am = in.am[dh * 2 + dw];
            const float4 gp = in.gp[dh * 2 + dw];
            const int a4[4] = {am.x, am.y, am.z, am.w};
            const float g4[4] = {gp.x, gp.y, gp.z, gp.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const int yy = 1 - 2 * dh + a, xx = 1 - 2 * dw + b;      // window-relative position of the pixel
                        if (yy >= 0 && yy < 3 && xx >= 0 && xx < 3 && a4[i] == yy * 3 + xx) acc[a * 2 + b][i] += g4[i];
                    }
        }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float4 v = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        const unsigned m = in.ok[p] ? in.mk[p] : 0xFu;
        if (m & 1u) v.x = 0.f;
        if (m & 2u) v.y = 0.f;
        if (m & 4u) v.z = 0.f;
        if (m & 8u) v.w = 0.f;
        g[p] = v;
    }
}

__global__ void __launch_bounds__(kThreads, kPoolOcc)
pool_bn_sums_k(const float* __restrict__ dpool, const uint8_t* __restrict__ amax, const float* __restrict__ x,
               const float* __restrict__ mean, const float* __restrict__ invstd, const PoolGeom pg, unsigned nitems, float* sums,
               const uint8_t* __restrict__ rmask, float* zero_me) {
    __shared__ float4 s_a[kThreads], s_b[kThreads];
    const int tid = threadIdx.x;
    const int CV = pg.C / 4;
    const int cv = tid & (CV - 1);
    const int c = cv * 4, C = pg.C;
    if (zero_me && blockIdx.x == 0)
        for (int ch = tid; ch < C; ch += kThreads) zero_me[ch] = 0.f;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    const unsigned stride = gridDim.x * kThreads;
    for (unsigned item = blockIdx.x * kThreads + tid; item < nitems; item += stride) {
        QuadIn in;
        int n, qi, qj;
        if (pool_quad_of(pg, item, n, qi, qj))      // (a warp holds two neighbouring quads: uniform away from the borders)
            pool_quad_load<true>(dpool, amax, x, rmask, pg, n, qi, qj, cv, mu, in);
        else
            pool_quad_load<false>(dpool, amax, x, rmask, pg, n, qi, qj, cv, mu, in);
        float4 g[4];
        pool_quad_grad(in, g);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            sa.x += g[p].x; sa.y += g[p].y; sa.z += g[p].z; sa.w += g[p].w;
            sb.x = fmaf(g[p].x, (in.xv[p].x - mu.x) * is.x, sb.x);
            sb.y = fmaf(g[p].y, (in.xv[p].y - mu.y) * is.y, sb.y);
            sb.z = fmaf(g[p].z, (in.xv[p].z - mu.z) * is.z, sb.z);
            sb.w = fmaf(g[p].w, (in.xv[p].w - mu.w) * is.w, sb.w);
        }
    }
    s_a[tid] = sa;
    s_b[tid] = sb;
    __syncthreads();
    for (int st = kThreads / 2; st >= CV; st >>= 1) {
        if (tid < st) {
            float4 a = s_a[tid], b = s_a[tid + st];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            s_a[tid] = a;
            float4 p = s_b[tid], q = s_b[tid + st];
            p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w;
            s_b[tid] = p;
        }
        __syncthreads();
    }
    if (tid < CV) {
        const float4 a = s_a[tid], b = s_b[tid];
        atomicAdd(sums + c, a.x); atomicAdd(sums + c + 1, a.y); atomicAdd(sums + c + 2, a.z); atomicAdd(sums + c + 3, a.w);
        atomicAdd(sums + C + c, b.x); atomicAdd(sums + C + c + 1, b.y); atomicAdd(sums + C + c + 2, b.z);
        atomicAdd(sums + C + c + 3, b.w);
    }
}

template <bool kAcc>
__global__ void __launch_bounds__(kThreads, kPoolOcc)
pool_bn_apply_k(const float* __restrict__ dpool, const uint8_t* __restrict__ amax, const float* __restrict__ x, float* dx,
                const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                const float* __restrict__ sums, float invM, const PoolGeom pg, unsigned nitems, const uint8_t* __restrict__ rmask,
                float* dgamma, float* dbeta, int acc_gamma, int acc_beta, float* dxsum, float* sums_rw, unsigned* ticket) {
    __shared__ float4 s_a[kThreads];
    __shared__ unsigned s_last;
    const int tid = threadIdx.x;
    const int CV = pg.C / 4;
    const int cv = tid & (CV - 1);
    const int c = cv * 4, C = pg.C;
    if (blockIdx.x == 0) {
        for (int ch = tid; ch < C; ch += kThreads) {
            if (dbeta) dbeta[ch] = acc_beta ? dbeta[ch] + sums[ch] : sums[ch];
            if (dgamma) dgamma[ch] = acc_gamma ? dgamma[ch] + sums[C + ch] : sums[C + ch];
        }
    }
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    float4 k1 = *reinterpret_cast<const float4*>(sums + c), k2 = *reinterpret_cast<const float4*>(sums + C + c);
    k1.x *= invM; k1.y *= invM; k1.z *= invM; k1.w *= invM;
    k2.x *= invM; k2.y *= invM; k2.z *= invM; k2.w *= invM;
    const float4 kg = make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned stride = gridDim.x * kThreads;
    for (unsigned item = blockIdx.x * kThreads + tid; item < nitems; item += stride) {
        QuadIn in;
        int n, qi, qj;
        if (pool_quad_of(pg, item, n, qi, qj))      // (a warp holds two neighbouring quads: uniform away from the borders)
            pool_quad_load<true>(dpool, amax, x, rmask, pg, n, qi, qj, cv, mu, in);
        else
            pool_quad_load<false>(dpool, amax, x, rmask, pg, n, qi, qj, cv, mu, in);
        float4 g[4];
        pool_quad_grad(in, g);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (!in.ok[p]) continue;
            const int h = 2 * qi + (p >> 1), w = 2 * qj + (p & 1);
            float4 o;
            o.x = kg.x * (g[p].x - k1.x - (in.xv[p].x - mu.x) * is.x * k2.x);
            o.y = kg.y * (g[p].y - k1.y - (in.xv[p].y - mu.y) * is.y * k2.y);
            o.z = kg.z * (g[p].z - k1.z - (in.xv[p].z - mu.z) * is.z * k2.z);
            o.w = kg.w * (g[p].w - k1.w - (in.xv[p].w - mu.w) * is.w * k2.w);
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            float4* q = reinterpret_cast<float4*>(dx) + (((int64_t)n * pg.H + h) * pg.W + w) * CV + cv;
            if (kAcc) {
                const float4 d = *q;
                o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
            }
            stg_stream(q, o);
        }
    }
    if (dxsum) {
        s_a[tid] = acc;
        __syncthreads();
        for (int st = kThreads / 2; st >= CV; st >>= 1) {
            if (tid < st) {
                float4 a = s_a[tid], b = s_a[tid + st];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                s_a[tid] = a;
            }
            __syncthreads();
        }
        if (tid < CV) {
            const float4 a = s_a[tid];
            atomicAdd(dxsum + c, a.x); atomicAdd(dxsum + c + 1, a.y); atomicAdd(dxsum + c + 2, a.z); atomicAdd(dxsum + c + 3, a.w);
        }
    }
    // every block read `sums` before its loop; the LAST block to get here zeroes them for the next BatchNorm backward
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        for (int ch = tid; ch < 2 * C; ch += kThreads) sums_rw[ch] = 0.f;
        if (tid == 0) *ticket = 0u;
    }
}

// ------------------------------------------------------------------ fused backward for tensors that fit the SMs' shared memory
// One COOPERATIVE launch (grid = #SMs, 512 threads): phase 1 streams dy (masked) and x once, accumulates the two sums and
// parks what it read in shared memory -- dy only (CACHE = 1) or dy and x (CACHE = 2), each thread in its own slots;
// grid-wide barrier; phase 2 computes dx from the parked copies. HBM traffic 4e*n / 3e*n instead of 5e*n and one launch
// instead of two: the layer3 / layer4 BatchNorms of ResNet-18 (12-26 MB tensors), where the two-launch form runs at
// 3.1-4.1 TB/s of algorithmic traffic, and every BatchNorm of the 56 x 56 configuration. Same arithmetic per element.
constexpr int kFusedThreads = 512;
constexpr int kFusedU = 4;                       // float4 per thread, array and iteration: chunk = 2048 float4 = 32 KB
constexpr int kFusedSmemMax = 200 * 1024;        // parked copies (the two reduction scratch arrays come on top)

template <int CACHE, bool kAcc>
__global__ void __launch_bounds__(kFusedThreads, 1)
bn_bwd_fused_k(const float* __restrict__ dy, const float* __restrict__ x, float* dx, const float* __restrict__ mean,
               const float* __restrict__ invstd, const float* __restrict__ gamma, float* sums, float invM, int64_t nvec, int CV,
               const uint8_t* __restrict__ rmask, float* dgamma, float* dbeta, int acc_gamma, int acc_beta, float* dxsum,
               int zero_dxsum, unsigned* ticket, int iters, float* __restrict__ side) {
    extern __shared__ float4 park[];                       // [CACHE][iters][U][threads]
    __shared__ float4 s_a[kFusedThreads], s_b[kFusedThreads];
    __shared__ unsigned s_last;
    constexpr int U = kFusedU, T = kFusedThreads;
    const int tid = threadIdx.x;
    const int c = (tid & (CV - 1)) * 4;
    const int C = CV * 4;
    const int64_t nchunks = (nvec + T * U - 1) / (T * U);
    float4* park_x = park + (int64_t)iters * U * T;
    if (dxsum && zero_dxsum && blockIdx.x == 0)
        for (int ch = tid; ch < C; ch += T) dxsum[ch] = 0.f;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    for (int it = 0; it < iters; ++it) {
        const int64_t chunk = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
        if (chunk >= nchunks) break;
        const int64_t base = chunk * (T * U) + tid;
        float4 g[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * T;
            g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            xv[u] = mu;
            if (i < nvec) {
                g[u] = ldg_stream(reinterpret_cast<const float4*>(dy) + i);
                xv[u] = ldg_stream(reinterpret_cast<const float4*>(x) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * T;
            if (rmask && i < nvec) {
                const int64_t e0 = i * 4;
                const unsigned m = (__ldg(rmask + (e0 >> 3)) >> (e0 & 4)) & 0xFu;
                if (m & 1u) g[u].x = 0.f;
                if (m & 2u) g[u].y = 0.f;
                if (m & 4u) g[u].z = 0.f;
                if (m & 8u) g[u].w = 0.f;
            }
            park[(it * U + u) * T + tid] = g[u];            // the MASKED gradient: phase 2 does not look at the mask again
            if (CACHE == 2) park_x[(it * U + u) * T + tid] = xv[u];
            sa.x += g[u].x; sa.y += g[u].y; sa.z += g[u].z; sa.w += g[u].w;
            sb.x = fmaf(g[u].x, (xv[u].x - mu.x) * is.x, sb.x);
            sb.y = fmaf(g[u].y, (xv[u].y - mu.y) * is.y, sb.y);
            sb.z = fmaf(g[u].z, (xv[u].z - mu.z) * is.z, sb.z);
            sb.w = fmaf(g[u].w, (xv[u].w - mu.w) * is.w, sb.w);
        }
    }
    s_a[tid] = sa;
    s_b[tid] = sb;
    __syncthreads();
    for (int st = T / 2; st >= CV; st >>= 1) {      // threads tid, tid+CV, tid+2CV ... share a channel group
        if (tid < st) {
            float4 a = s_a[tid], b = s_a[tid + st];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            s_a[tid] = a;
            float4 p = s_b[tid], q = s_b[tid + st];
            p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w;
            s_b[tid] = p;
        }
        __syncthreads();
    }
    if (tid < CV) {
        const float4 a = s_a[tid], b = s_b[tid];
        atomicAdd(sums + c, a.x); atomicAdd(sums + c + 1, a.y); atomicAdd(sums + c + 2, a.z); atomicAdd(sums + c + 3, a.w);
        atomicAdd(sums + C + c, b.x); atomicAdd(sums + C + c + 1, b.y); atomicAdd(sums + C + c + 2, b.z);
        atomicAdd(sums + C + c + 3, b.w);
    }
    __threadfence();
    cooperative_groups::this_grid().sync();
    // ---- phase 2
    if (blockIdx.x == 0) {
        for (int ch = tid; ch < C; ch += T) {
            const float sd = __ldcg(sums + ch), sx = __ldcg(sums + C + ch);
            if (dbeta) dbeta[ch] = acc_beta ? dbeta[ch] + sd : sd;
            if (dgamma) dgamma[ch] = acc_gamma ? dgamma[ch] + sx : sx;
        }
    }
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    float4 k1 = __ldcg(reinterpret_cast<const float4*>(sums + c)), k2 = __ldcg(reinterpret_cast<const float4*>(sums + C + c));
    k1.x *= invM; k1.y *= invM; k1.z *= invM; k1.w *= invM;
    k2.x *= invM; k2.y *= invM; k2.z *= invM; k2.w *= invM;
    const float4 kg = make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        const int64_t chunk = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
        if (chunk >= nchunks) break;
        const int64_t base = chunk * (T * U) + tid;
        float4 xv[U];
        if (CACHE != 2) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + (int64_t)u * T;
                xv[u] = mu;
                if (i < nvec) xv[u] = ldg_stream(reinterpret_cast<const float4*>(x) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * T;
            if (i < nvec) {
                const float4 g = park[(it * U + u) * T + tid];
                if (side) stg_stream(reinterpret_cast<float4*>(side) + i, g);      // (see bn_bwd_apply_persist_k)
                const float4 xx = CACHE == 2 ? park_x[(it * U + u) * T + tid] : xv[u];
                float4 o;
                o.x = kg.x * (g.x - k1.x - (xx.x - mu.x) * is.x * k2.x);
                o.y = kg.y * (g.y - k1.y - (xx.y - mu.y) * is.y * k2.y);
                o.z = kg.z * (g.z - k1.z - (xx.z - mu.z) * is.z * k2.z);
                o.w = kg.w * (g.w - k1.w - (xx.w - mu.w) * is.w * k2.w);
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                float4* p = reinterpret_cast<float4*>(dx) + i;
                if (kAcc) {
                    const float4 d = *p;
                    o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
                }
                stg_stream(p, o);
            }
        }
    }
    if (dxsum) {
        __syncthreads();
        s_a[tid] = acc;
        __syncthreads();
        for (int st = T / 2; st >= CV; st >>= 1) {
            if (tid < st) {
                float4 a = s_a[tid], b = s_a[tid + st];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                s_a[tid] = a;
            }
            __syncthreads();
        }
        if (tid < CV) {
            const float4 a = s_a[tid];
            atomicAdd(dxsum + c, a.x); atomicAdd(dxsum + c + 1, a.y); atomicAdd(dxsum + c + 2, a.z); atomicAdd(dxsum + c + 3, a.w);
        }
    }
    // every block has read `sums`; the LAST block to get here zeroes them for the next BatchNorm backward
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        for (int ch = tid; ch < 2 * C; ch += T) sums[ch] = 0.f;
        if (tid == 0) *ticket = 0u;
    }
}

// ------------------------------------------------------------------ statistics from conv-epilogue partials
// part[p][{n, mean, M2}][C], p < P  ->  save_mean / save_invstd and the running-statistics update. block (32, TY).
__global__ void __launch_bounds__(1024) bn_finalize_k(const float* __restrict__ part, int P, int C, float R, float* __restrict__ save_mean,
                              float* __restrict__ save_invstd, float* running_mean, float* running_var, float eps,
                              float momentum) {
    __shared__ float s_n[1152], s_mean[1152], s_m2[1152];     // [TY][TX + 1], (TX, TY) = (32, <= 32) or (8, 128)
    const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
    const int c = blockIdx.x * TX + tx;
    const int me = ty * (TX + 1) + tx;
    float an = 0.f, amean = 0.f, am2 = 0.f;
    if (c < C) {
        for (int p0 = ty; p0 < P; p0 += TY * 8) {
            float pn[8], pm[8], p2[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = p0 + u * TY;
                pn[u] = 0.f; pm[u] = 0.f; p2[u] = 0.f;
                if (p < P) {
                    const float* q = part + (int64_t)p * 3 * C + c;
                    pn[u] = __ldg(q); pm[u] = __ldg(q + C); p2[u] = __ldg(q + 2 * C);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const Moments m = merge(Moments{an, amean, am2}, Moments{pn[u], pm[u], p2[u]});
                an = m.n; amean = m.mean; am2 = m.m2;
            }
        }
    }
    s_n[me] = an; s_mean[me] = amean; s_m2[me] = am2;
    __syncthreads();
    for (int st = TY / 2; st > 0; st >>= 1) {
        if (ty < st) {
            const int other = me + st * (TX + 1);
            const Moments m = merge(Moments{s_n[me], s_mean[me], s_m2[me]}, Moments{s_n[other], s_mean[other], s_m2[other]});
            s_n[me] = m.n; s_mean[me] = m.mean; s_m2[me] = m.m2;
        }
        __syncthreads();
    }
    if (ty == 0 && c < C) {
        const float mean = s_mean[tx];
        const float var = s_m2[tx] / R;
        save_mean[c] = mean;
        save_invstd[c] = 1.0f / sqrtf(var + eps);
        if (running_mean) running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * mean;
        if (running_var) running_var[c] = momentum * running_var[c] + (1.f - momentum) * var;
    }
}

inline bool use_vec4(int C, const void* a, const void* b = nullptr, const void* c = nullptr) {
    return (C % 4 == 0) && aligned16(a) && (!b || aligned16(b)) && (!c || aligned16(c));
}

}  // namespace

// -> XSB_ERR_UNSUPPORTED when the tensor does not fit the shared memory of one wave (the two-launch kernels take it)
static int launch_bn_bwd_fused(const float* dy, const float* x, float* dx, const float* mean, const float* invstd,
                               const float* gamma, float* sums, int64_t R, int64_t nvec, int CV, const uint8_t* rmask,
                               float* dgamma, float* dbeta, int acc_dx, int acc_gamma, int acc_beta, float* dxsum, int acc_dxsum,
                               unsigned* ticket, cudaStream_t s, float* side = nullptr) {
    static const int enabled = [] { const char* e = getenv("XSB_BN_FUSED"); return (e && e[0] == '0') ? 0 : 1; }();
    if (!enabled || CV > kFusedThreads) return XSB_ERR_UNSUPPORTED;
    const int64_t chunk = (int64_t)kFusedThreads * kFusedU;
    const int64_t nchunks = ceil_div64(nvec, chunk);
    int grid = xsb_sm_count_cached();
    if (grid > nchunks) grid = (int)nchunks;
    const int iters = (int)ceil_div64(nchunks, grid);
    const int64_t per_array = (int64_t)iters * chunk * 16;        // bytes one parked array takes per block
    int cache = 0;
    if (2 * per_array <= kFusedSmemMax) cache = 2;
    else if (per_array <= kFusedSmemMax) cache = 1;
    if (cache == 0) return XSB_ERR_UNSUPPORTED;
    const size_t smem = (size_t)cache * per_array;
    float invM = 1.0f / (float)R;
    int zero_dxsum = (dxsum && !acc_dxsum) ? 1 : 0;
    int iters_arg = iters;
    void* args[] = {(void*)&dy, (void*)&x, (void*)&dx, (void*)&mean, (void*)&invstd, (void*)&gamma, (void*)&sums, (void*)&invM,
                    (void*)&nvec, (void*)&CV, (void*)&rmask, (void*)&dgamma, (void*)&dbeta, (void*)&acc_gamma, (void*)&acc_beta,
                    (void*)&dxsum, (void*)&zero_dxsum, (void*)&ticket, (void*)&iters_arg, (void*)&side};
    const void* fn;
    if (cache == 2) fn = acc_dx ? (const void*)bn_bwd_fused_k<2, true> : (const void*)bn_bwd_fused_k<2, false>;
    else fn = acc_dx ? (const void*)bn_bwd_fused_k<1, true> : (const void*)bn_bwd_fused_k<1, false>;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(bn_bwd_fused_k<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
        XSB_CUDA(cudaFuncSetAttribute(bn_bwd_fused_k<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
        XSB_CUDA(cudaFuncSetAttribute(bn_bwd_fused_k<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
        XSB_CUDA(cudaFuncSetAttribute(bn_bwd_fused_k<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
        configured = true;
    }
    XSB_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(kFusedThreads), args, smem, s));
    return XSB_OK;
}

extern "C" {

// Number of floats the `ws` scratch buffer must hold for a [R, C] problem. The buffer must be
// zero-initialised once (the ticket counters at its head are self-cleaning afterwards).
int64_t xsb_bn_workspace_floats(int64_t R, int C) {
    (void)R;
    return kScratchBase + (int64_t)kMaxRowBlocks * C * 3 + kMaxRowBlocks + 3 * (int64_t)C;
}

static int launch_bn_apply(const float* x, float* y, const float* mean, const float* invstd, const float* gamma,
                           const float* beta, int64_t R, int C, uint8_t* relu_mask, bool relu, cudaStream_t s) {
    const bool v4 = use_vec4(C, x, y);
    const int V = v4 ? 4 : 1, CV = C / V;
    const int64_t nvec = R * CV;
    const int cv_mask = (CV & (CV - 1)) == 0 ? CV - 1 : -1;
    const unsigned blocks = (unsigned)ceil_div64(nvec, (int64_t)kThreads * 4);
    if (relu) XSB_CHECK_ARG(v4 && nvec % 2 == 0 && relu_mask, "bn_apply: fused ReLU needs C %% 4 == 0, an even vector count and a mask");
    if (v4 && CV <= kThreads && (CV & (CV - 1)) == 0 && aligned16(mean) && aligned16(invstd) && aligned16(gamma) &&
        aligned16(beta)) {
        int64_t pb = ceil_div64(nvec, (int64_t)kThreads * 4 * 2);
        const int64_t cap = (int64_t)xsb_sm_count_cached() * 8;
        if (pb > cap) pb = cap;
        if (pb < 1) pb = 1;
        if (relu)
            bn_apply_persist_k<true><<<(unsigned)pb, kThreads, 0, s>>>(x, y, mean, invstd, gamma, beta, nvec, CV, relu_mask);
        else
            bn_apply_persist_k<false><<<(unsigned)pb, kThreads, 0, s>>>(x, y, mean, invstd, gamma, beta, nvec, CV, nullptr);
        XSB_LAUNCH_CHECK();
        return XSB_OK;
    }
    if (relu) {
        bn_apply_k<4, true><<<blocks, kThreads, 0, s>>>(x, y, mean, invstd, gamma, beta, nvec, CV, cv_mask, relu_mask);
    } else if (v4) {
        bn_apply_k<4, false><<<blocks, kThreads, 0, s>>>(x, y, mean, invstd, gamma, beta, nvec, CV, cv_mask, nullptr);
    } else {
        bn_apply_k<1, false><<<blocks, kThreads, 0, s>>>(x, y, mean, invstd, gamma, beta, nvec, CV, cv_mask, nullptr);
    }
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// z = max(0, bn_a(xa) + rhs) + ReLU bit mask (see bn_add_relu_persist_k); rhs = b, or bn_b(b) when mean_b != NULL.
// XSB_ERR_UNSUPPORTED (nothing launched) outside the envelope: the caller runs xsb_bn_apply (x2) + xsb_add_relu.
int xsb_bn_add_relu(const float* xa, const float* mean_a, const float* invstd_a, const float* gamma_a, const float* beta_a,
                    const float* b, const float* mean_b, const float* invstd_b, const float* gamma_b, const float* beta_b,
                    float* z, uint8_t* relu_mask, int64_t R, int C, void* stream) {
    const int CV = C / 4;
    const int64_t nvec = R * CV;
    const bool two = mean_b != nullptr;
    if (R <= 0 || C <= 0 || C % 4 != 0 || CV > kThreads || (CV & (CV - 1)) != 0 || nvec % 2 != 0 || !relu_mask ||
        !aligned16(xa) || !aligned16(b) || !aligned16(z) || !aligned16(mean_a) || !aligned16(invstd_a) || !aligned16(gamma_a) ||
        !aligned16(beta_a) ||
        (two && (!invstd_b || !gamma_b || !beta_b || !aligned16(mean_b) || !aligned16(invstd_b) || !aligned16(gamma_b) ||
                 !aligned16(beta_b))))
        return XSB_ERR_UNSUPPORTED;
    int64_t pb = ceil_div64(nvec, (int64_t)kThreads * 4 * 2);
    const int64_t cap = (int64_t)xsb_sm_count_cached() * 8;
    if (pb > cap) pb = cap;
    if (pb < 1) pb = 1;
    cudaStream_t s = xsb_stream(stream);
    if (two)
        bn_add_relu_persist_k<true><<<(unsigned)pb, kThreads, 0, s>>>(xa, mean_a, invstd_a, gamma_a, beta_a, b, mean_b, invstd_b,
                                                                     gamma_b, beta_b, z, relu_mask, nvec, CV);
    else
        bn_add_relu_persist_k<false><<<(unsigned)pb, kThreads, 0, s>>>(xa, mean_a, invstd_a, gamma_a, beta_a, b, nullptr, nullptr,
                                                                      nullptr, nullptr, z, relu_mask, nvec, CV);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// Batch statistics of x [R,C]: save_mean / save_invstd and the running-statistics update (functional.py:689-706).
int xsb_bn_stats(const float* x, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                 int64_t R, int C, float eps, float momentum, float* ws, void* stream) {
    XSB_CHECK_ARG(R > 0 && C > 0, "bn: empty input");
    cudaStream_t s = xsb_stream(stream);
    const bool v4 = use_vec4(C, x);
    static const int small_on = [] { const char* e = getenv("XSB_BN_STATS_SMALL"); return (e && e[0] == '0') ? 0 : 1; }();
    if (small_on && v4 && R <= kSmallStatsMaxRows && R >= 64) {
        bn_stats_small_k<<<(C / 4 + 7) / 8, kSmallStatsThreads, 0, s>>>(x, (int)R, C, save_mean, save_invstd, running_mean,
                                                                         running_var, eps, momentum);
        XSB_LAUNCH_CHECK();
        return XSB_OK;
    }
    const Geom g = make_geom(R, C, v4);
    XSB_CHECK_ARG(g.chunks < kCounterSlots, "bn: C=%d too large", C);
    dim3 grid(g.nb, g.chunks);
    if (v4)
        bn_stats_k<4><<<grid, kThreads, 0, s>>>(x, R, C, g.LX, g.LY, g.rows_per_block, ws, save_mean, save_invstd,
                                                running_mean, running_var, eps, momentum);
    else
        bn_stats_k<1><<<grid, kThreads, 0, s>>>(x, R, C, g.LX, g.LY, g.rows_per_block, ws, save_mean, save_invstd,
                                                running_mean, running_var, eps, momentum);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// y = gamma * (x - mean) * invstd + beta, optionally followed by ReLU with the x<0 bit mask written in the same
// pass (relu != 0: the fusion of BatchNormalization -> ReLU(inplace=True); needs C % 4 == 0).
int xsb_bn_apply(const float* x, float* y, const float* mean, const float* invstd, const float* gamma,
                 const float* beta, int64_t R, int C, int relu, uint8_t* relu_mask, void* stream) {
    XSB_CHECK_ARG(R > 0 && C > 0, "bn: empty input");
    return launch_bn_apply(x, y, mean, invstd, gamma, beta, R, C, relu_mask, relu != 0, xsb_stream(stream));
}

// Training-mode forward = xsb_bn_stats + xsb_bn_apply. x,y [R,C]; gamma,beta,running_*,save_* [C]. running_* nullable.
int xsb_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float* y, float* save_mean, float* save_invstd, int64_t R, int C, float eps, float momentum,
                     float* ws, void* stream) {
    int rc = xsb_bn_stats(x, running_mean, running_var, save_mean, save_invstd, R, C, eps, momentum, ws, stream);
    if (rc != XSB_OK) return rc;
    return launch_bn_apply(x, y, save_mean, save_invstd, gamma, beta, R, C, nullptr, false, xsb_stream(stream));
}

// Inference branch (xs/nn/functional.py:708-725): y = gamma * (x - moving_mean) / sqrt(moving_var + eps) + beta.
int xsb_bn_eval_fwd(const float* x, const float* gamma, const float* beta, const float* moving_mean,
                    const float* moving_var, float* y, int64_t R, int C, float eps, float* ws, void* stream) {
    XSB_CHECK_ARG(R > 0 && C > 0, "bn: empty input");
    cudaStream_t s = xsb_stream(stream);
    float* invstd = ws + kScratchBase;   // partial-sum area of the scratch buffer; free between launches
    invstd_k<<<(C + 255) / 256, 256, 0, s>>>(moving_var, invstd, C, eps);
    XSB_LAUNCH_CHECK();
    return launch_bn_apply(x, y, moving_mean, invstd, gamma, beta, R, C, nullptr, false, s);
}

// statistics of a conv output from the partials its epilogue left (xsb_conv2d_fwd_stats): replaces xsb_bn_stats
int xsb_bn_finalize(const float* partials, int P, int64_t R, int C, float* running_mean, float* running_var,
                    float* save_mean, float* save_invstd, float eps, float momentum, void* stream) {
    XSB_CHECK_ARG(P > 0 && R > 0 && C > 0, "bn_finalize: empty input");
    // many partials (one per 128-row tile of a large layer): 8 channels x 128 partial rows per block, so that C / 8 blocks
    // share the serial Chan merges; few partials (persistent producers): 32 channels per block
    const bool tall = P > 64;
    const int TX = tall ? 8 : 32, TY = tall ? 128 : (P > 32 ? 32 : 8);
    bn_finalize_k<<<(C + TX - 1) / TX, dim3(TX, TY), 0, xsb_stream(stream)>>>(partials, P, C, (float)R, save_mean, save_invstd,
                                                                              running_mean, running_var, eps, momentum);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

static int bn_bwd_impl(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                       float* dx, float* dgamma, float* dbeta, int64_t R, int C, int acc_dx, int acc_dgamma, int acc_dbeta,
                       const uint8_t* relu_mask, float* ws, float* dxsum, int acc_dxsum, int* dxsum_done, void* stream,
                       float* side = nullptr, int* side_done = nullptr);

// Backward. dx nullable (input needs no grad); dgamma/dbeta nullable. acc_* : 1 = "+=", 0 = "=".
// relu_mask (nullable): dy is first masked with the in-place ReLU's bit mask (fused ReLUBackward).
int xsb_bn_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
               float* dx, float* dgamma, float* dbeta, int64_t R, int C, int acc_dx, int acc_dgamma, int acc_dbeta,
               const uint8_t* relu_mask, float* ws, void* stream) {
    return bn_bwd_impl(dy, x, gamma, save_mean, save_invstd, dx, dgamma, dbeta, R, C, acc_dx, acc_dgamma, acc_dbeta, relu_mask,
                       ws, nullptr, 0, nullptr, stream);
}

// Same, and dxsum[c] (+)= sum_r (dx contribution)[r, c]: the bias gradient of the convolution that feeds this BatchNorm
// (xs/nn/grad_fn.py:196-197), gathered while dx is written. *dxsum_done (host) = 1 when it was produced, 0 when this
// shape runs the generic kernels (then call xsb_colsum on dx).
int xsb_bn_bwd_dxsum(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                     float* dx, float* dgamma, float* dbeta, int64_t R, int C, int acc_dx, int acc_dgamma, int acc_dbeta,
                     const uint8_t* relu_mask, float* ws, float* dxsum, int acc_dxsum, int* dxsum_done, void* stream) {
    return bn_bwd_impl(dy, x, gamma, save_mean, save_invstd, dx, dgamma, dbeta, R, C, acc_dx, acc_dgamma, acc_dbeta, relu_mask,
                       ws, dxsum, acc_dxsum, dxsum_done, stream);
}

// Same (dxsum nullable), and dy_masked_out[R, C] = dy with the relu_mask elements zeroed, written by the apply pass that
// reads the masked dy anyway: the gradient of the identity shortcut of a residual block (xs/nn/grad_fn.py:6-16 behind
// 146-157) without a relu_bwd_mask pass of its own. *side_done (host) = 1 when written, 0 when this shape runs the generic
// kernels (then call xsb_relu_bwd_mask).
int xsb_bn_bwd_dxsum_side(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                          float* dx, float* dgamma, float* dbeta, int64_t R, int C, int acc_dx, int acc_dgamma, int acc_dbeta,
                          const uint8_t* relu_mask, float* ws, float* dxsum, int acc_dxsum, int* dxsum_done,
                          float* dy_masked_out, int* side_done, void* stream) {
    return bn_bwd_impl(dy, x, gamma, save_mean, save_invstd, dx, dgamma, dbeta, R, C, acc_dx, acc_dgamma, acc_dbeta, relu_mask,
                       ws, dxsum, acc_dxsum, dxsum_done, stream, dy_masked_out, side_done);
}

static int bn_bwd_impl(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                       float* dx, float* dgamma, float* dbeta, int64_t R, int C, int acc_dx, int acc_dgamma, int acc_dbeta,
                       const uint8_t* relu_mask, float* ws, float* dxsum, int acc_dxsum, int* dxsum_done, void* stream,
                       float* side, int* side_done) {
    XSB_CHECK_ARG(R > 0 && C > 0, "bn: empty input");
    cudaStream_t s = xsb_stream(stream);
    if (dxsum_done) *dxsum_done = 0;
    if (side_done) *side_done = 0;
    if (side && !aligned16(side)) side = nullptr;
    const bool v4 = use_vec4(C, dy, x, dx);
    const int CV4 = C / 4;
    if (v4 && dx && C % 4 == 0 && CV4 <= kThreads && (CV4 & (CV4 - 1)) == 0 && aligned16(save_mean) &&
        aligned16(save_invstd) && aligned16(gamma)) {
        // scratch: sums[2C] at a FIXED offset (independent of C) so that the zeros one call leaves are the zeros the
        // next call finds; the buffer is zero-initialised once by the caller, afterwards the apply kernel cleans up
        float* sums = ws + kCounterSlots;
        unsigned* ticket = reinterpret_cast<unsigned*>(ws) + (kCounterSlots - 1);
        const int64_t nvec = R * CV4;
        {
            const int rc = launch_bn_bwd_fused(dy, x, dx, save_mean, save_invstd, gamma, sums, R, nvec, CV4, relu_mask, dgamma,
                                               dbeta, acc_dx, acc_dgamma, acc_dbeta, dxsum, acc_dxsum, ticket, s, side);
            if (rc != XSB_ERR_UNSUPPORTED) {
                if (rc == XSB_OK && dxsum && dxsum_done) *dxsum_done = 1;
                if (rc == XSB_OK && side && side_done) *side_done = 1;
                return rc;
            }
        }
        // >= 4 grid-stride iterations per thread: a block ends with 2C (+C) atomics, keep them few on small tensors
        int64_t blocks = ceil_div64(nvec, (int64_t)kThreads * 4 * 4);
        const int64_t cap = (int64_t)xsb_sm_count_cached() * kPersistBlocksPerSM;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        col_sums_atomic_k<true><<<(unsigned)blocks, kThreads, 0, s>>>(dy, x, save_mean, save_invstd, nvec, CV4, sums, relu_mask,
                                                                      (dxsum && !acc_dxsum) ? dxsum : nullptr);
        XSB_LAUNCH_CHECK();
        const float invM = 1.0f / (float)R;
        static const int rev = [] { const char* e = getenv("XSB_BN_BWD_REV"); return (e && e[0] == '0') ? 0 : 1; }();
        if (acc_dx)
            bn_bwd_apply_persist_k<true><<<(unsigned)blocks, kThreads, 0, s>>>(dy, x, dx, save_mean, save_invstd, gamma, sums, invM,
                                                                              nvec, CV4, relu_mask, dgamma, dbeta, acc_dgamma,
                                                                              acc_dbeta, dxsum, sums, ticket, rev, side);
        else
            bn_bwd_apply_persist_k<false><<<(unsigned)blocks, kThreads, 0, s>>>(dy, x, dx, save_mean, save_invstd, gamma, sums, invM,
                                                                               nvec, CV4, relu_mask, dgamma, dbeta, acc_dgamma,
                                                                               acc_dbeta, dxsum, sums, ticket, rev, side);
        XSB_LAUNCH_CHECK();
        if (dxsum && dxsum_done) *dxsum_done = 1;
        if (side && side_done) *side_done = 1;
        return XSB_OK;
    }
    const Geom g = make_geom(R, C, v4, 2 * 148);   // two streamed arrays per thread: fewer, fatter CTAs measured faster
    XSB_CHECK_ARG(g.chunks < kCounterSlots, "bn: C=%d too large", C);
    float* coef = ws + kScratchBase + (int64_t)kMaxRowBlocks * C * 3 + kMaxRowBlocks;
    dim3 grid(g.nb, g.chunks);
    if (v4)
        col_sums_k<4, true><<<grid, kThreads, 0, s>>>(dy, x, save_mean, save_invstd, gamma, R, C, g.LX, g.LY,
                                                      g.rows_per_block, ws, dgamma, dbeta, coef, acc_dgamma, acc_dbeta,
                                                      relu_mask);
    else
        col_sums_k<1, true><<<grid, kThreads, 0, s>>>(dy, x, save_mean, save_invstd, gamma, R, C, g.LX, g.LY,
                                                      g.rows_per_block, ws, dgamma, dbeta, coef, acc_dgamma, acc_dbeta,
                                                      relu_mask);
    XSB_LAUNCH_CHECK();
    if (dx) {
        const int64_t nvec = R * g.CV;
        const int cv_mask = (g.CV & (g.CV - 1)) == 0 ? g.CV - 1 : -1;
        const unsigned blocks = (unsigned)ceil_div64(nvec, (int64_t)kThreads * 4);
        if (v4) {
            if (acc_dx)
                bn_bwd_apply_k<4, true><<<blocks, kThreads, 0, s>>>(dy, x, dx, save_mean, save_invstd, coef, nvec, g.CV, C, cv_mask, relu_mask);
            else
                bn_bwd_apply_k<4, false><<<blocks, kThreads, 0, s>>>(dy, x, dx, save_mean, save_invstd, coef, nvec, g.CV, C, cv_mask, relu_mask);
        } else {
            if (acc_dx)
                bn_bwd_apply_k<1, true><<<blocks, kThreads, 0, s>>>(dy, x, dx, save_mean, save_invstd, coef, nvec, g.CV, C, cv_mask, relu_mask);
            else
                bn_bwd_apply_k<1, false><<<blocks, kThreads, 0, s>>>(dy, x, dx, save_mean, save_invstd, coef, nvec, g.CV, C, cv_mask, relu_mask);
        }
        XSB_LAUNCH_CHECK();
    }
    return XSB_OK;
}

// out[c] (+)= sum_r x[r, c]   (bias gradients: conv db over N*OH*OW rows, Dense db over N rows)
// BatchNorm backward whose dy is the backward gather of the 3x3 / stride 2 / pad 1 max-pool that follows the BatchNorm's
// in-place ReLU (see pool_quad_grad): xsb_maxpool2d_bwd + xsb_bn_bwd_dxsum without the full-size dy in between.
// XSB_ERR_UNSUPPORTED (nothing launched) outside the envelope: the caller materialises dy and runs the plain kernels.
int xsb_bn_bwd_pool_dxsum(const float* dpool, const uint8_t* argmax, const float* x, const float* gamma,
                          const float* save_mean, const float* save_invstd, float* dx, float* dgamma, float* dbeta, int N,
                          int H, int W, int C, int acc_dx, int acc_dgamma, int acc_dbeta, const uint8_t* relu_mask, float* ws,
                          float* dxsum, int acc_dxsum, void* stream) {
    const int CV4 = C / 4;
    if (C % 4 != 0 || CV4 > kThreads || (CV4 & (CV4 - 1)) != 0 || !dx || !aligned16(dpool) || !aligned16(x) || !aligned16(dx) ||
        !aligned16(save_mean) || !aligned16(save_invstd) || !aligned16(gamma) || ((uintptr_t)argmax & 3u) != 0 ||
        2 * C > kSumsFloats || N < 1 || H < 1 || W < 1)
        return XSB_ERR_UNSUPPORTED;
    cudaStream_t s = xsb_stream(stream);
    PoolGeom pg;
    pg.N = N; pg.H = H; pg.W = W; pg.C = C;
    pg.OH = (H + 2 - 3) / 2 + 1; pg.OW = (W + 2 - 3) / 2 + 1;
    pg.QH = (H + 1) / 2; pg.QW = (W + 1) / 2;
    pg.cvshift = 0;
    while ((1 << pg.cvshift) < CV4) ++pg.cvshift;
    pg.fast = (C % 8 == 0 && (int64_t)N * H * W * CV4 < ((int64_t)1 << 31)) ? 1 : 0;
    const int64_t nitems64 = (int64_t)N * pg.QH * pg.QW * CV4;
    if (nitems64 >= ((int64_t)1 << 31)) return XSB_ERR_UNSUPPORTED;
    const unsigned nitems = (unsigned)nitems64;
    float* sums = ws + kCounterSlots;
    unsigned* ticket = reinterpret_cast<unsigned*>(ws) + (kCounterSlots - 1);
    // >= 8 quads per thread (a block ends with 2C (+C) atomics), many more blocks than SMs: the block scheduler keeps
    // every SM at 3 blocks and there is no persistent-loop tail
    // (small tensors: one quad per thread until every SM holds 6 blocks -- 49 blocks of 8 quads per thread took 25 us per
    // pass on the 32 x 28 x 28 x 64 stem of the 56 x 56 configuration)
    const int64_t sms = xsb_sm_count_cached();
    int64_t blocks = ceil_div64(nitems64, (int64_t)kThreads);
    if (blocks > sms * 6) {
        blocks = ceil_div64(nitems64, (int64_t)kThreads * 8);
        if (blocks < sms * 6) blocks = sms * 6;
    }
    const int64_t cap = sms * 24;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pool_bn_sums_k<<<(unsigned)blocks, kThreads, 0, s>>>(dpool, argmax, x, save_mean, save_invstd, pg, nitems, sums, relu_mask,
                                                        (dxsum && !acc_dxsum) ? dxsum : nullptr);
    XSB_LAUNCH_CHECK();
    const float invM = 1.0f / ((float)N * (float)H * (float)W);
    if (acc_dx)
        pool_bn_apply_k<true><<<(unsigned)blocks, kThreads, 0, s>>>(dpool, argmax, x, dx, save_mean, save_invstd, gamma, sums, invM, pg,
                                                                   nitems, relu_mask, dgamma, dbeta, acc_dgamma, acc_dbeta, dxsum,
                                                                   sums, ticket);
    else
        pool_bn_apply_k<false><<<(unsigned)blocks, kThreads, 0, s>>>(dpool, argmax, x, dx, save_mean, save_invstd, gamma, sums, invM, pg,
                                                                    nitems, relu_mask, dgamma, dbeta, acc_dgamma, acc_dbeta, dxsum,
                                                                    sums, ticket);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

int xsb_colsum(const float* x, float* out, int64_t R, int C, int accumulate, float* ws, void* stream) {
    XSB_CHECK_ARG(R > 0 && C > 0, "colsum: empty input");
    cudaStream_t s = xsb_stream(stream);
    const bool v4 = use_vec4(C, x);
    const Geom g = make_geom(R, C, v4);
    XSB_CHECK_ARG(g.chunks < kCounterSlots, "colsum: C=%d too large", C);
    dim3 grid(g.nb, g.chunks);
    if (v4)
        col_sums_k<4, false><<<grid, kThreads, 0, s>>>(x, nullptr, nullptr, nullptr, nullptr, R, C, g.LX, g.LY,
                                                       g.rows_per_block, ws, nullptr, out, nullptr, 0, accumulate, nullptr);
    else
        col_sums_k<1, false><<<grid, kThreads, 0, s>>>(x, nullptr, nullptr, nullptr, nullptr, R, C, g.LX, g.LY,
                                                       g.rows_per_block, ws, nullptr, out, nullptr, 0, accumulate, nullptr);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

}  // extern "C"
