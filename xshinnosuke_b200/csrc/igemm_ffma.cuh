// FP32 implicit-GEMM on CUDA cores (FFMA): the exact-fp32 math mode (XSB_MATH_FP32).
// C[M,N] (+)= A[M,K] * B[K,N] (+ bias[N]) where A and B are *loaders* -- plain matrices or on-the-fly
// im2col / transposed-filter views of channels-last tensors, so no patch matrix is ever materialised
// (the reference materialises a float64 col buffer: xs/utils/common.py:35-46).
// Block tile BM x BN x 16, 256 threads, (BM/16)x(BN/16) register micro-tile, register-staged double buffering.
#pragma once
#include "common.cuh"
#include "conv_geom.cuh"

namespace igemm {

using ::ConvGeom;

constexpr int BK = 16;
constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// Loaders. load4(r, k, v): if kContigK -> v = {X[r][k..k+3]} else v = {X[r..r+3][k]}; out-of-range = 0.
// r is a row of A (m) or a column of B (n); R / K are the logical extents.
// ------------------------------------------------------------------------------------------------
struct MatK {  // X[r][k] = p[r*ld + k]   (contiguous along k)
    static constexpr bool kContigK = true;
    const float* p; int64_t ld; int R, K; bool vec;
    __device__ __forceinline__ void load4(int r, int k, float (&v)[4]) const {
        if (r < R && vec && k + 3 < K) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p + (int64_t)r * ld + k));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (r < R && k + i < K) ? __ldg(p + (int64_t)r * ld + k + i) : 0.f;
        }
    }
};

struct MatR {  // X[r][k] = p[k*ld + r]   (contiguous along r)
    static constexpr bool kContigK = false;
    const float* p; int64_t ld; int R, K; bool vec;
    __device__ __forceinline__ void load4(int r, int k, float (&v)[4]) const {
        if (k < K && vec && r + 3 < R) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p + (int64_t)k * ld + r));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (k < K && r + i < R) ? __ldg(p + (int64_t)k * ld + r + i) : 0.f;
        }
    }
};


// patch(pixel=(n,oh,ow), j=(kh,kw,c)) = x[n, oh*sh-pad+kh, ow*sw-pad+kw, c]   (zero outside)
struct Patch {
    const float* x; ConvGeom g; int P, J; bool vec;  // P = N*OH*OW, J = KH*KW*C
    __device__ __forceinline__ const float* addr(int pix, int j, bool& ok) const {
        const int c = j % g.C;
        const int t = j / g.C;
        const int kw = t % g.KW, kh = t / g.KW;
        const int ow = pix % g.OW;
        const int t2 = pix / g.OW;
        const int oh = t2 % g.OH, n = t2 / g.OH;
        const int h = oh * g.sh - g.pad + kh, w = ow * g.sw - g.pad + kw;
        ok = (pix < P) && (j < J) && h >= 0 && h < g.H && w >= 0 && w < g.W;
        return x + (((int64_t)n * g.H + h) * g.W + w) * g.C + c;
    }
    // 4 consecutive j (same pixel)
    __device__ __forceinline__ void load4_j(int pix, int j, float (&v)[4]) const {
        bool ok;
        if (vec) {  // C % 4 == 0 and j % 4 == 0: one (kh,kw), 4 contiguous channels
            const float* a = addr(pix, j, ok);
            if (ok) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(a));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
                v[0] = v[1] = v[2] = v[3] = 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float* a = addr(pix, j + i, ok);
                v[i] = ok ? __ldg(a) : 0.f;
            }
        }
    }
};

struct PatchA {  // conv forward A: rows = pixels, k = (kh,kw,c)
    static constexpr bool kContigK = true;
    Patch pt;
    __device__ __forceinline__ void load4(int r, int k, float (&v)[4]) const { pt.load4_j(r, k, v); }
};
struct PatchB {  // conv wgrad B: columns n = (kh,kw,c), k = pixels
    static constexpr bool kContigK = false;
    Patch pt;
    __device__ __forceinline__ void load4(int r, int k, float (&v)[4]) const { pt.load4_j(k, r, v); }
};

// conv dgrad A: rows = input pixels (n,h,w), k = (kh,kw,oc):
//   dy[n, (h+pad-kh)/sh, (w+pad-kw)/sw, oc] when both divisions are exact and in range, else 0
struct DgradA {
    static constexpr bool kContigK = true;
    const float* dy; ConvGeom g; int P, J; bool vec;  // P = N*H*W, J = KH*KW*OC
    __device__ __forceinline__ const float* addr(int pix, int j, bool& ok) const {
        const int oc = j % g.OC;
        const int t = j / g.OC;
        const int kw = t % g.KW, kh = t / g.KW;
        const int w = pix % g.W;
        const int t2 = pix / g.W;
        const int h = t2 % g.H, n = t2 / g.H;
        const int hh = h + g.pad - kh, ww = w + g.pad - kw;
        const int oh = hh / g.sh, ow = ww / g.sw;
        ok = (pix < P) && (j < J) && hh >= 0 && ww >= 0 && (hh - oh * g.sh == 0) && (ww - ow * g.sw == 0) &&
             oh < g.OH && ow < g.OW;
        return dy + (((int64_t)n * g.OH + oh) * g.OW + ow) * g.OC + oc;
    }
    __device__ __forceinline__ void load4(int r, int k, float (&v)[4]) const {
        bool ok;
        if (vec) {
            const float* a = addr(r, k, ok);
            if (ok) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(a));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
                v[0] = v[1] = v[2] = v[3] = 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float* a = addr(r, k + i, ok);
                v[i] = ok ? __ldg(a) : 0.f;
            }
        }
    }
};

// conv dgrad B: columns n = c, k = (kh,kw,oc): w[oc, kh, kw, c]  (filter stored [OC,KH,KW,C])
struct DgradB {
    static constexpr bool kContigK = false;
    const float* w; ConvGeom g; int J; bool vec;  // J = KH*KW*OC
    __device__ __forceinline__ void load4(int r, int k, float (&v)[4]) const {
        const int oc = k % g.OC;
        const int t = k / g.OC;  // kh*KW + kw
        const float* a = w + ((int64_t)oc * g.KH * g.KW + t) * g.C + r;
        if (k < J && vec && r + 3 < g.C) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(a));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (k < J && r + i < g.C) ? __ldg(a + i) : 0.f;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Kernel. grid = (m tiles, n tiles, splits). splits > 1 -> raw partial sums go to `part[z][M][N]`.
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, class LA, class LB>
__global__ void __launch_bounds__(kThreads)
igemm_ffma_k(const LA la, const LB lb, float* C, int64_t ldc, const float* __restrict__ bias, int M, int N, int K,
             int k_per_split, int accumulate, float* part, float* act) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int APAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN + APAD];
    constexpr int A_LOADS = BM * BK / 4 / kThreads;  // float4-groups per thread
    constexpr int B_LOADS = BN * BK / 4 / kThreads;
    static_assert(A_LOADS >= 1 && B_LOADS >= 1, "tile too small");

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int k_begin = blockIdx.z * k_per_split;
    int k_end = k_begin + k_per_split;
    if (k_end > K) k_end = K;

    float ra[A_LOADS][4], rb[B_LOADS][4];
    auto fetch = [&](int kb) {
#pragma unroll
        for (int l = 0; l < A_LOADS; ++l) {
            const int g = tid + l * kThreads;
            if (LA::kContigK) {
                const int r = g / (BK / 4), kq = (g % (BK / 4)) * 4;
                la.load4(m0 + r, kb + kq, ra[l]);
                if (kb + kq + 3 >= k_end) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (kb + kq + i >= k_end) ra[l][i] = 0.f;
                }
            } else {
                const int kk = g / (BM / 4), rq = (g % (BM / 4)) * 4;
                if (kb + kk < k_end) la.load4(m0 + rq, kb + kk, ra[l]);
                else ra[l][0] = ra[l][1] = ra[l][2] = ra[l][3] = 0.f;
            }
        }
#pragma unroll
        for (int l = 0; l < B_LOADS; ++l) {
            const int g = tid + l * kThreads;
            if (LB::kContigK) {
                const int r = g / (BK / 4), kq = (g % (BK / 4)) * 4;
                lb.load4(n0 + r, kb + kq, rb[l]);
                if (kb + kq + 3 >= k_end) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (kb + kq + i >= k_end) rb[l][i] = 0.f;
                }
            } else {
                const int kk = g / (BN / 4), rq = (g % (BN / 4)) * 4;
                if (kb + kk < k_end) lb.load4(n0 + rq, kb + kk, rb[l]);
                else rb[l][0] = rb[l][1] = rb[l][2] = rb[l][3] = 0.f;
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int l = 0; l < A_LOADS; ++l) {
            const int g = tid + l * kThreads;
            if (LA::kContigK) {
                const int r = g / (BK / 4), kq = (g % (BK / 4)) * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) As[buf][kq + i][r] = ra[l][i];
            } else {
                const int kk = g / (BM / 4), rq = (g % (BM / 4)) * 4;
                *reinterpret_cast<float4*>(&As[buf][kk][rq]) = make_float4(ra[l][0], ra[l][1], ra[l][2], ra[l][3]);
            }
        }
#pragma unroll
        for (int l = 0; l < B_LOADS; ++l) {
            const int g = tid + l * kThreads;
            if (LB::kContigK) {
                const int r = g / (BK / 4), kq = (g % (BK / 4)) * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) Bs[buf][kq + i][r] = rb[l][i];
            } else {
                const int kk = g / (BN / 4), rq = (g % (BN / 4)) * 4;
                *reinterpret_cast<float4*>(&Bs[buf][kk][rq]) = make_float4(rb[l][0], rb[l][1], rb[l][2], rb[l][3]);
            }
        }
    };

    // two-level accumulation: `acc` is flushed into `tot` every kFlush k-blocks so the fp32 rounding
    // error grows like sqrt(chunk) + sqrt(K/chunk) instead of sqrt(K) (wgrad reduces over N*OH*OW ~ 1e6)
    constexpr int kFlush = 32;
    float acc[TM][TN], tot[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }
    int since_flush = 0;

    const int ty = tid / 16, tx = tid % 16;
    int buf = 0;
    if (k_begin < k_end) {
        fetch(k_begin);
        stash(0);
    }
    __syncthreads();
    for (int kb = k_begin; kb < k_end; kb += BK) {
        const bool more = kb + BK < k_end;
        if (more) fetch(kb + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 q = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + i * 16]);
                a[i] = q.x; a[i + 1] = q.y; a[i + 2] = q.z; a[i + 3] = q.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 q = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + j * 16]);
                b[j] = q.x; b[j + 1] = q.y; b[j + 2] = q.z; b[j + 3] = q.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (++since_flush == kFlush) {
            since_flush = 0;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += tot[i][j];

    // micro-tile rows: ty*4 + (i/4)*64 + i%4 ; cols: tx*4 + (j/4)*64 + j%4
    const bool split = part != nullptr;
    float* out = split ? part + (int64_t)blockIdx.z * M * N : C;
    const int64_t ldo = split ? N : ldc;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * 4 + (i / 4) * 64 + (i % 4);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * 4 + (j / 4) * 64 + (j % 4);
            if (n >= N) continue;
            float v = acc[i][j];
            float* o = out + (int64_t)m * ldo + n;
            if (!split) {
                if (bias) v += __ldg(bias + n);
                if (accumulate) v += *o;
                if (act) act[(int64_t)m * ldc + n] = fmaxf(v, 0.f);     // fused activation: second output, same pitch
            }
            *o = v;
        }
    }
}

// C (+)= sum_z part[z] (+ bias)
static __global__ void splitk_reduce_k(const float* __restrict__ part, float* C, int64_t ldc, const float* __restrict__ bias,
                                int M, int N, int splits, int accumulate, float* act) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)M * N) return;
    const int m = (int)(t / N), n = (int)(t % N);
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += part[(int64_t)z * M * N + t];
    if (bias) v += bias[n];
    float* o = C + (int64_t)m * ldc + n;
    v = accumulate ? *o + v : v;
    *o = v;
    if (act) act[(int64_t)m * ldc + n] = fmaxf(v, 0.f);
}

// Host-side launcher. ws/ws_floats: scratch for split-K partials (may be null -> no split).
template <class LA, class LB>
int launch(const LA& la, const LB& lb, float* C, int64_t ldc, const float* bias, int M, int N, int K, int accumulate,
           bool allow_split, float* ws, int64_t ws_floats, cudaStream_t s, float* act = nullptr) {
    if (M <= 0 || N <= 0) return XSB_OK;
    if (K <= 0) {
        xsb_set_error("igemm: K must be positive");
        return XSB_ERR_ARG;
    }
    const int sms = xsb_sm_count_cached();
    const bool big = (int64_t)M * N >= (int64_t)128 * 128 * sms;  // enough 128x128 tiles for a full wave
    const int BM = big ? 128 : 64, BN = big ? 128 : 64;
    const int mt = (M + BM - 1) / BM, nt = (N + BN - 1) / BN;
    int splits = 1;
    if (allow_split && ws && mt * nt < 2 * sms && K >= 8 * BK) {
        splits = (4 * sms + mt * nt - 1) / (mt * nt);
        const int max_by_k = K / (4 * BK);
        if (splits > max_by_k) splits = max_by_k;
        const int64_t max_by_ws = ws_floats / ((int64_t)M * N);
        if (splits > max_by_ws) splits = (int)max_by_ws;
        if (splits > 256) splits = 256;
        if (splits < 1) splits = 1;
    }
    int k_per = (K + splits - 1) / splits;
    k_per = (k_per + BK - 1) / BK * BK;
    splits = (K + k_per - 1) / k_per;
    float* part = splits > 1 ? ws : nullptr;
    dim3 grid(mt, nt, splits);
    if (nt > 65535) {
        xsb_set_error("igemm: N=%d too large", N);
        return XSB_ERR_ARG;
    }
    if (big)
        igemm_ffma_k<128, 128, LA, LB><<<grid, kThreads, 0, s>>>(la, lb, C, ldc, bias, M, N, K, k_per, accumulate, part, act);
    else
        igemm_ffma_k<64, 64, LA, LB><<<grid, kThreads, 0, s>>>(la, lb, C, ldc, bias, M, N, K, k_per, accumulate, part, act);
    XSB_LAUNCH_CHECK();
    if (splits > 1) {
        const int64_t tot = (int64_t)M * N;
        splitk_reduce_k<<<(unsigned)ceil_div64(tot, 256), 256, 0, s>>>(part, C, ldc, bias, M, N, splits, accumulate, act);
        XSB_LAUNCH_CHECK();
    }
    return XSB_OK;
}

}  // namespace igemm
