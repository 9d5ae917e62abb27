// max_pool2d (with argmax) and avg_pool2d, forward + backward, on channels-last [N,H,W,C] fp32.
// Roofline: HBM. Algorithmic bytes: fwd 4*(NHWC + N*OH*OW*C) + 1*N*OH*OW*C,
// bwd 4*(N*OH*OW*C + NHWC) + 1*N*OH*OW*C.
//
// Reference semantics (xs/nn/functional.py:445-491, xs/nn/grad_fn.py:206-235):
//  * the window is taken on a ZERO-padded input (not -inf), OH = (H + 2p - k)//s + 1;
//  * argmax is the index ky*k+kx of the FIRST maximum of the k*k window (np.argmax), stored flat in
//    (n, oh, ow, c) order -- exactly the channels-last order used here, so the u8 array this kernel
//    writes is the reference's int64 `pool_argmax` element for element;
//  * backward scatters dY to that slot and overlap-adds (col2im); slots that fall in the padding
//    are cropped away. Here: each input pixel GATHERS from the <= ceil(k/s)^2 windows covering it
//    (no atomics, deterministic, supports accumulate).
#include "common.cuh"

namespace {

// V = channels per thread (4 -> float4, 1 -> scalar). Grids are (ceil(W*CV/256), rows, N): no 64-bit div.
// KT > 0: window size known at compile time (unrolled taps, branch-free interior path)
template <int V, int KT = 0>
__global__ void __launch_bounds__(256)
maxpool_fwd_k(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ amax, int N, int H, int W,
              int C, int OH, int OW, int k, int sh, int sw, int pad) {
    const unsigned CV = C / V;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)OW * CV) return;
    const int cv = (int)(t % CV);
    const int ow = (int)(t / CV);
    const int oh = blockIdx.y;
    const int n = blockIdx.z;
    const int h0 = oh * sh - pad, w0 = ow * sw - pad;
    float best[V];
    int arg[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { best[i] = 0.f; arg[i] = 0; }
    if (KT > 0 && V == 4) {
        constexpr int T = KT > 0 ? KT : 1;
        float4 q[T * T];
        const float4* img = reinterpret_cast<const float4*>(x) + (int64_t)n * H * W * CV;
        if (h0 >= 0 && w0 >= 0 && h0 + T <= H && w0 + T <= W) {
            // interior window (all but the first/last output row and column): no clamps, no padding selects --
            // the kernel is issue-bound (368 instructions per output vector with them, 74 % SM busy in ncu)
            const float4* p0 = img + (unsigned)((h0 * W + w0) * (int)CV + cv);
            const unsigned row = (unsigned)W * CV;
#pragma unroll
            for (int ky = 0; ky < T; ++ky)
#pragma unroll
                for (int kx = 0; kx < T; ++kx) q[ky * T + kx] = __ldg(p0 + (ky * row + kx * CV));
        } else {
#pragma unroll
            for (int ky = 0; ky < T; ++ky)
#pragma unroll
                for (int kx = 0; kx < T; ++kx) {
                    const int h = h0 + ky, w = w0 + kx;
                    q[ky * T + kx] = make_float4(0.f, 0.f, 0.f, 0.f);   // zero padding takes part in the max
                    if (h >= 0 && h < H && w >= 0 && w < W) q[ky * T + kx] = __ldg(img + (unsigned)((h * W + w) * (int)CV + cv));
                }
        }
        best[0] = q[0].x; best[1 % V] = q[0].y; best[2 % V] = q[0].z; best[3 % V] = q[0].w;
#pragma unroll
        for (int t2 = 1; t2 < T * T; ++t2) {      // same tap order and tie rule (first maximum wins) as the generic loop
            const float v4[4] = {q[t2].x, q[t2].y, q[t2].z, q[t2].w};
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (v4[i] > best[i]) { best[i] = v4[i]; arg[i] = t2; }
        }
    }
    bool first = true;
    for (int ky = 0; ky < ((KT > 0 && V == 4) ? 0 : k); ++ky) {
        const int h = h0 + ky;
        for (int kx = 0; kx < k; ++kx) {
            const int w = w0 + kx;
            float v[V];
            if (h >= 0 && h < H && w >= 0 && w < W) {
                const float* p = x + (((int64_t)n * H + h) * W + w) * C + (int64_t)cv * V;
                if (V == 4) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                    v[0] = q.x; v[1 % V] = q.y; v[2 % V] = q.z; v[3 % V] = q.w;
                } else {
                    v[0] = __ldg(p);
                }
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = 0.f;  // zero padding takes part in the max
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                if (first || v[i] > best[i]) { best[i] = v[i]; arg[i] = ky * k + kx; }
            }
            first = false;
        }
    }
    const int64_t o = (((int64_t)n * OH + oh) * OW + ow) * C + (int64_t)cv * V;
    if (V == 4) {
        *reinterpret_cast<float4*>(y + o) = make_float4(best[0], best[1 % V], best[2 % V], best[3 % V]);
        if (amax)
            *reinterpret_cast<uchar4*>(amax + o) =
                make_uchar4((unsigned char)arg[0], (unsigned char)arg[1 % V], (unsigned char)arg[2 % V],
                            (unsigned char)arg[3 % V]);
    } else {
        y[o] = best[0];
        if (amax) amax[o] = (uint8_t)arg[0];
    }
}

template <int V, bool kAcc>
__global__ void __launch_bounds__(256)
maxpool_bwd_k(const float* __restrict__ dy, const uint8_t* __restrict__ amax, float* dx, int N, int H, int W, int C,
              int OH, int OW, int k, int sh, int sw, int pad) {
    const unsigned CV = C / V;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)W * CV) return;
    const int cv = (int)(t % CV);
    const int w = (int)(t / CV);
    const int h = blockIdx.y;
    const int n = blockIdx.z;
    // windows oh with oh*sh - pad <= h <= oh*sh - pad + k - 1
    const int hp = h + pad, wp = w + pad;
    int oh_lo = hp - k + 1;
    oh_lo = oh_lo <= 0 ? 0 : (oh_lo + sh - 1) / sh;
    int oh_hi = hp / sh;
    if (oh_hi > OH - 1) oh_hi = OH - 1;
    int ow_lo = wp - k + 1;
    ow_lo = ow_lo <= 0 ? 0 : (ow_lo + sw - 1) / sw;
    int ow_hi = wp / sw;
    if (ow_hi > OW - 1) ow_hi = OW - 1;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
            const int slot = (hp - oh * sh) * k + (wp - ow * sw);
            const int64_t o = (((int64_t)n * OH + oh) * OW + ow) * C + (int64_t)cv * V;
            if (V == 4) {
                const uchar4 a = __ldg(reinterpret_cast<const uchar4*>(amax + o));
                const float4 g = __ldg(reinterpret_cast<const float4*>(dy + o));
                if (a.x == slot) acc[0] += g.x;
                if (a.y == slot) acc[1 % V] += g.y;
                if (a.z == slot) acc[2 % V] += g.z;
                if (a.w == slot) acc[3 % V] += g.w;
            } else {
                if (__ldg(amax + o) == slot) acc[0] += __ldg(dy + o);
            }
        }
    }
    const int64_t xi = (((int64_t)n * H + h) * W + w) * C + (int64_t)cv * V;
    if (V == 4) {
        float4 out = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
        if (kAcc) {
            const float4 d = *reinterpret_cast<const float4*>(dx + xi);
            out.x += d.x; out.y += d.y; out.z += d.z; out.w += d.w;
        }
        *reinterpret_cast<float4*>(dx + xi) = out;
    } else {
        dx[xi] = kAcc ? dx[xi] + acc[0] : acc[0];
    }
}

// Stride-2 specialisation of the backward gather (the ResNet 3x3/2 pool): one thread owns a 2x2 input quad
// (rows 2i..2i+1, cols 2j..2j+1). Only the <= ceil((k+1)/2)^2 windows touching the quad are visited and each
// window's (argmax, dy) is loaded ONCE for the four pixels instead of once per pixel (the generic kernel is
// LSU-bound: ncu shows SM 77 % busy at 18 % DRAM).
template <bool kAcc, bool kSmall = false, bool kK3P1 = false>
__global__ void __launch_bounds__(256)
maxpool_bwd_s2_k(const float* __restrict__ dy, const uint8_t* __restrict__ amax, float* dx, int N, int H, int W, int C,
                 int OH, int OW, int k, int pad) {
    const unsigned CV = C / 4;
    const unsigned QW = (W + 1) / 2;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= QW * CV) return;
    const int cv = (int)(t % CV);
    const int qj = (int)(t / CV);
    const int qi = blockIdx.y;
    const int n = blockIdx.z;
    const int h0 = 2 * qi, w0 = 2 * qj;
    // windows oh with 2*oh - pad <= h <= 2*oh - pad + k - 1 for some h in {h0, h0+1}
    int oh_lo = h0 + pad - k + 1;
    oh_lo = oh_lo <= 0 ? 0 : (oh_lo + 1) / 2;
    int oh_hi = (h0 + 1 + pad) / 2;
    if (oh_hi > OH - 1) oh_hi = OH - 1;
    int ow_lo = w0 + pad - k + 1;
    ow_lo = ow_lo <= 0 ? 0 : (ow_lo + 1) / 2;
    int ow_hi = (w0 + 1 + pad) / 2;
    if (ow_hi > OW - 1) ow_hi = OW - 1;
    float acc[2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[a][b][i] = 0.f;
    // kSmall (k <= 3): at most 2 x 2 windows touch the quad -- fixed trip counts, so the (argmax, dy) loads of all four
    // windows are in flight together; windows past oh_hi / ow_hi contribute a slot that matches nothing
    const int noh = kSmall ? 2 : oh_hi - oh_lo + 1, now = kSmall ? 2 : ow_hi - ow_lo + 1;
#pragma unroll
    for (int dh = 0; dh < (kSmall ? 2 : noh); ++dh) {
        const int oh = oh_lo + dh;
        // window-relative row of pixel h0 (h0+1 -> ry+1); k = 3, pad = 1: oh_lo = qi, so ry is the constant 1 - 2*dh and
        // 9 of the 16 (window, quad pixel) pairs remain
        const int ry = kK3P1 ? 1 - 2 * dh : h0 + pad - 2 * oh;
#pragma unroll
        for (int dw = 0; dw < (kSmall ? 2 : now); ++dw) {
            const int ow = ow_lo + dw;
            const int rx = kK3P1 ? 1 - 2 * dw : w0 + pad - 2 * ow;
            const bool live = oh <= oh_hi && ow <= ow_hi;
            const int64_t o = (((int64_t)n * OH + (live ? oh : 0)) * OW + (live ? ow : 0)) * C + (int64_t)cv * 4;
            uchar4 am = __ldg(reinterpret_cast<const uchar4*>(amax + o));
            const float4 g = __ldg(reinterpret_cast<const float4*>(dy + o));
            if (!live) am = make_uchar4(255, 255, 255, 255);
            const int a4[4] = {am.x, am.y, am.z, am.w};
            const float g4[4] = {g.x, g.y, g.z, g.w};
            // window slot of each quad pixel (-1: the pixel is outside this window): compares, no divisions
            int slot[2][2];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int yy = ry + a, xx = rx + b, kk = kK3P1 ? 3 : k;
                    slot[a][b] = (yy >= 0 && yy < kk && xx >= 0 && xx < kk) ? yy * kk + xx : -1;
                }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b)
                        if (slot[a][b] >= 0 && a4[i] == slot[a][b]) acc[a][b][i] += g4[i];
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int h = h0 + a;
        if (h >= H) continue;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int w = w0 + b;
            if (w >= W) continue;
            float4* p = reinterpret_cast<float4*>(dx + (((int64_t)n * H + h) * W + w) * C + (int64_t)cv * 4);
            float4 out = make_float4(acc[a][b][0], acc[a][b][1], acc[a][b][2], acc[a][b][3]);
            if (kAcc) {
                const float4 d = *p;
                out.x += d.x; out.y += d.y; out.z += d.z; out.w += d.w;
            }
            *p = out;
        }
    }
}

// BatchNorm apply + ReLU (+ its bit mask) + max-pool with argmax in ONE pass over the conv output x: the normalised,
// rectified tensor z = max(0, gamma*(x-mean)*invstd + beta) -- which only the pool would read -- is never written.
// A block owns a TOH x TOW tile of outputs (all C <= 64 channels): phase 1 normalises the input tile + halo once into
// shared memory (zero padding stays zero) and writes the mask bits of the pixels the tile owns; phase 2 pools from
// shared memory. Same arithmetic, tap order and tie rule as bn_apply_k followed by maxpool_fwd_k.
constexpr int FP_TOH = 4, FP_TOW = 8;

__global__ void __launch_bounds__(256)
bn_relu_maxpool_fwd_k(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                      uint8_t* __restrict__ amax, uint8_t* __restrict__ rmask, int N, int H, int W, int C, int OH, int OW,
                      int k, int sh, int sw, int pad) {
    extern __shared__ float4 s_z[];                 // [rows][cols][CV]
    const int CV = C / 4;
    const int rows = (FP_TOH - 1) * sh + k, cols = (FP_TOW - 1) * sw + k;
    const int oh0 = blockIdx.y * FP_TOH, ow0 = blockIdx.x * FP_TOW;
    const int n = blockIdx.z;
    const int h_base = oh0 * sh - pad, w_base = ow0 * sw - pad;
    const int tid = threadIdx.x;
    const int total = rows * cols * CV;
    // 256 % CV == 0 (checked by the launcher): the thread's channel group is fixed, pixel index advances by 256/CV
    const int cv = tid % CV, ppi = 256 / CV;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + cv);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + cv);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + cv);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + cv);
    const int trips = (total + 255) / 256;          // block-uniform: the nibble pairing uses a full-warp shuffle
    constexpr int UB = 4;                           // loads in flight per thread
    for (int it0 = 0; it0 < trips; it0 += UB) {
        float4 q[UB];
        int hh[UB], ww[UB];
        bool ins[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int pix = (it0 + u) * ppi + tid / CV;           // pixel of the tile (row-major rows x cols)
            const int r = pix / cols, cidx = pix - r * cols;
            hh[u] = h_base + r;
            ww[u] = w_base + cidx;
            ins[u] = (it0 + u) < trips && pix < rows * cols && hh[u] >= 0 && hh[u] < H && ww[u] >= 0 && ww[u] < W;
            q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ins[u]) q[u] = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)n * H + hh[u]) * W + ww[u]) * C) + cv);
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            if (it0 + u >= trips) break;            // block-uniform
            const int i = (it0 + u) * 256 + tid;
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned nib = 0;
            if (ins[u]) {
                z.x = fmaf((q[u].x - mu.x) * is.x, ga.x, be.x);
                z.y = fmaf((q[u].y - mu.y) * is.y, ga.y, be.y);
                z.z = fmaf((q[u].z - mu.z) * is.z, ga.z, be.z);
                z.w = fmaf((q[u].w - mu.w) * is.w, ga.w, be.w);
                nib = (z.x < 0.f ? 1u : 0u) | (z.y < 0.f ? 2u : 0u) | (z.z < 0.f ? 4u : 0u) | (z.w < 0.f ? 8u : 0u);
                z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f);
            }
            if (i < total) s_z[i] = z;
            if (rmask) {
                const unsigned hi = __shfl_down_sync(0xffffffffu, nib, 1);
                // the tile owns pixel (h, w) when window (h / sh, w / sw) is one of its outputs
                const int oh = hh[u] / sh, ow = ww[u] / sw;
                const bool own = ins[u] && oh >= oh0 && oh < oh0 + FP_TOH && ow >= ow0 && ow < ow0 + FP_TOW;
                if (own && (cv & 1) == 0) {
                    const int64_t e0 = ((((int64_t)n * H + hh[u]) * W + ww[u]) * C) + (int64_t)cv * 4;
                    rmask[e0 >> 3] = (uint8_t)(nib | (hi << 4));
                }
            }
        }
    }
    __syncthreads();
    const int outs = FP_TOH * FP_TOW * CV;
    for (int o = tid; o < outs; o += 256) {
        const int cv = o % CV;
        const int t2 = o / CV;
        const int tow = t2 % FP_TOW, toh = t2 / FP_TOW;
        const int oh = oh0 + toh, ow = ow0 + tow;
        if (oh >= OH || ow >= OW) continue;
        float best[4] = {0.f, 0.f, 0.f, 0.f};
        int arg[4] = {0, 0, 0, 0};
        bool first = true;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) {
                const float4 q = s_z[((toh * sh + ky) * cols + tow * sw + kx) * CV + cv];
                const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (first || v[i] > best[i]) { best[i] = v[i]; arg[i] = ky * k + kx; }
                first = false;
            }
        const int64_t off = (((int64_t)n * OH + oh) * OW + ow) * C + (int64_t)cv * 4;
        *reinterpret_cast<float4*>(y + off) = make_float4(best[0], best[1], best[2], best[3]);
        if (amax)
            *reinterpret_cast<uchar4*>(amax + off) =
                make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
    }
}

// The same fusion for k = 3, stride 2, pad 1 (the ResNet stem pool) WITHOUT the shared-memory tile: one thread = one
// output pixel x 4 channels, the 3x3 window read straight from x (9 float4 loads in flight; the re-reads by the
// neighbouring windows hit L1 / L2 exactly as in maxpool_fwd_k), BatchNorm + ReLU applied in registers. The thread OWNS
// the input pixels (2oh, 2oh+1) x (2ow, 2ow+1) -- window slots {1,2} x {1,2} -- and their mask nibbles; lane pairs
// (even, odd channel group of one pixel: C % 8 == 0) join theirs into the mask BYTE with one shuffle. HBM traffic: x once +
// pooled y + argmax + mask instead of x, z (write), z (read), y, argmax, mask.
// The kernel is INSTRUCTION-bound (ncu: 69 % issue-active, top stall math-pipe throttle), so:
//  * interior windows (all but the first output row / column) run a separate, check-free instantiation (a real branch:
//    a warp holds two output pixels, it is uniform everywhere but at the image border);
//  * no per-element ReLU: max / first-argmax are taken over the raw BN values (padding = 0) -- if that maximum is > 0
//    it is also the rectified maximum at the same tap, otherwise every rectified value is 0 and the first tap wins.
template <bool INTERIOR>
__device__ __forceinline__ void bn_relu_pool_window(const float4* __restrict__ img, int H, int W, unsigned CV, int cv, int h0,
                                                    int w0, const float4& mu, const float4& is, const float4& ga,
                                                    const float4& be, float (&best)[4], int (&arg)[4], unsigned& nibs) {
    float4 q[9];
    if (INTERIOR) {
        const float4* p0 = img + (unsigned)((h0 * W + w0) * (int)CV + cv);
        const unsigned row = (unsigned)W * CV;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) q[ky * 3 + kx] = __ldg(p0 + (ky * row + kx * CV));
    } else {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int h = h0 + ky, w = w0 + kx;
                q[ky * 3 + kx] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (h >= 0 && h < H && w >= 0 && w < W) q[ky * 3 + kx] = __ldg(img + (unsigned)((h * W + w) * (int)CV + cv));
            }
    }
    nibs = 0;     // nibble of owned pixel (ky - 1) * 2 + (kx - 1) at bits 8 * that index
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int tp = ky * 3 + kx;
            const float4 v = q[tp];
            float z[4];
            // same arithmetic as bn_apply; zero PADDING stays zero and takes part in the max
            z[0] = fmaf((v.x - mu.x) * is.x, ga.x, be.x);
            z[1] = fmaf((v.y - mu.y) * is.y, ga.y, be.y);
            z[2] = fmaf((v.z - mu.z) * is.z, ga.z, be.z);
            z[3] = fmaf((v.w - mu.w) * is.w, ga.w, be.w);
            if (!INTERIOR) {
                const bool inside = h0 + ky >= 0 && h0 + ky < H && w0 + kx >= 0 && w0 + kx < W;
                if (!inside) z[0] = z[1] = z[2] = z[3] = 0.f;
            }
            if (ky >= 1 && kx >= 1) {
                const unsigned nib = (z[0] < 0.f ? 1u : 0u) | (z[1] < 0.f ? 2u : 0u) | (z[2] < 0.f ? 4u : 0u) |
                                     (z[3] < 0.f ? 8u : 0u);
                nibs |= nib << (8 * ((ky - 1) * 2 + (kx - 1)));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (tp == 0 || z[i] > best[i]) { best[i] = z[i]; arg[i] = tp; }    // first maximum wins
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (!(best[i] > 0.f)) { best[i] = 0.f; arg[i] = 0; }      // every rectified value of the window is 0
}

__global__ void __launch_bounds__(256, 4)
bn_relu_maxpool_k3s2_k(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                       uint8_t* __restrict__ amax, uint8_t* __restrict__ rmask, int N, int H, int W, int C, int OH, int OW) {
    const unsigned CV = C / 4;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < (unsigned)OW * CV;      // (dead lanes stay for the shuffle; blockDim and CV are even)
    const int cv = live ? (int)(t % CV) : 0;
    const int ow = live ? (int)(t / CV) : 0;
    const int oh = blockIdx.y;
    const int n = blockIdx.z;
    const int h0 = 2 * oh - 1, w0 = 2 * ow - 1;
    const float4* img = reinterpret_cast<const float4*>(x) + (int64_t)n * H * W * CV;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + cv);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + cv);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + cv);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + cv);
    unsigned nibs;
    float best[4] = {0.f, 0.f, 0.f, 0.f};
    int arg[4] = {0, 0, 0, 0};
    if (h0 >= 0 && w0 >= 0 && h0 + 3 <= H && w0 + 3 <= W)
        bn_relu_pool_window<true>(img, H, W, CV, cv, h0, w0, mu, is, ga, be, best, arg, nibs);
    else
        bn_relu_pool_window<false>(img, H, W, CV, cv, h0, w0, mu, is, ga, be, best, arg, nibs);
    if (live) {
        const int64_t o = (((int64_t)n * OH + oh) * OW + ow) * C + (int64_t)cv * 4;
        *reinterpret_cast<float4*>(y + o) = make_float4(best[0], best[1], best[2], best[3]);
        if (amax)
            *reinterpret_cast<uchar4*>(amax + o) =
                make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
    }
    if (rmask) {
        const unsigned hi = __shfl_down_sync(0xffffffffu, nibs, 1);     // the odd channel group of the same pixel
        if (live && (cv & 1) == 0) {
            const unsigned bytes4 = nibs | (hi << 4);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int h = 2 * oh + a, w = 2 * ow + b;
                    if (h < H && w < W) {
                        const int64_t e0 = (((int64_t)n * H + h) * W + w) * C + (int64_t)cv * 4;
                        rmask[e0 >> 3] = (uint8_t)((bytes4 >> (8 * (a * 2 + b))) & 0xFFu);
                    }
                }
        }
    }
}

// avg pool (padding == 0 only: the reference's padded variant is broken). V channels per thread.
template <int V>
__global__ void avgpool_fwd_k(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int OH,
                              int OW, int k, int sh, int sw) {
    const unsigned CV = C / V;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)OW * CV) return;
    const int c = (int)(t % CV) * V;
    const int ow = (int)(t / CV);
    const int oh = blockIdx.y;
    const int n = blockIdx.z;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int ky = 0; ky < k; ++ky)
        for (int kx = 0; kx < k; ++kx) {
            const float* p = x + (((int64_t)n * H + oh * sh + ky) * W + ow * sw + kx) * C + c;
            if (V == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                acc[0] += q.x; acc[1 % V] += q.y; acc[2 % V] += q.z; acc[3 % V] += q.w;
            } else {
                acc[0] += __ldg(p);
            }
        }
    const float inv = 1.0f / (float)(k * k);
    float* o = y + (((int64_t)n * OH + oh) * OW + ow) * C + c;
    if (V == 4)
        *reinterpret_cast<float4*>(o) = make_float4(acc[0] * inv, acc[1 % V] * inv, acc[2 % V] * inv, acc[3 % V] * inv);
    else
        o[0] = acc[0] * inv;
}

template <int V>
__global__ void avgpool_bwd_k(const float* __restrict__ dy, float* dx, int N, int H, int W, int C, int OH, int OW,
                              int k, int sh, int sw, int acc_flag) {
    const unsigned CV = C / V;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)W * CV) return;
    const int c = (int)(t % CV) * V;
    const int w = (int)(t / CV);
    const int h = blockIdx.y;
    const int n = blockIdx.z;
    int oh_lo = h - k + 1;
    oh_lo = oh_lo <= 0 ? 0 : (oh_lo + sh - 1) / sh;
    int oh_hi = h / sh;
    if (oh_hi > OH - 1) oh_hi = OH - 1;
    int ow_lo = w - k + 1;
    ow_lo = ow_lo <= 0 ? 0 : (ow_lo + sw - 1) / sw;
    int ow_hi = w / sw;
    if (ow_hi > OW - 1) ow_hi = OW - 1;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
            const float* p = dy + (((int64_t)n * OH + oh) * OW + ow) * C + c;
            if (V == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                acc[0] += q.x; acc[1 % V] += q.y; acc[2 % V] += q.z; acc[3 % V] += q.w;
            } else {
                acc[0] += __ldg(p);
            }
        }
    const float inv = 1.0f / (float)(k * k);
    float* p = dx + (((int64_t)n * H + h) * W + w) * C + c;
    if (V == 4) {
        float4 r = make_float4(acc[0] * inv, acc[1 % V] * inv, acc[2 % V] * inv, acc[3 % V] * inv);
        if (acc_flag) {
            const float4 d = *reinterpret_cast<const float4*>(p);
            r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;
        }
        *reinterpret_cast<float4*>(p) = r;
    } else {
        const float r = acc[0] * inv;
        p[0] = acc_flag ? p[0] + r : r;
    }
}

}  // namespace

extern "C" {

// x [N,H,W,C] -> y [N,OH,OW,C], argmax u8 [N,OH,OW,C] (nullable). OH = (H + 2*pad - k)/sh + 1.
int xsb_maxpool2d_fwd(const float* x, float* y, uint8_t* argmax, int N, int H, int W, int C, int k, int sh, int sw,
                      int pad, void* stream) {
    XSB_CHECK_ARG(k >= 1 && k <= 15 && sh >= 1 && sw >= 1 && pad >= 0, "maxpool: bad k/stride/pad %d %d %d %d", k, sh,
                  sw, pad);
    const int OH = (H + 2 * pad - k) / sh + 1, OW = (W + 2 * pad - k) / sw + 1;
    XSB_CHECK_ARG(OH > 0 && OW > 0, "maxpool: empty output");
    cudaStream_t s = xsb_stream(stream);
    const bool v4 = (C % 4 == 0) && aligned16(x) && aligned16(y) && (!argmax || ((uintptr_t)argmax & 3u) == 0);
    XSB_CHECK_ARG(N <= 65535 && OH <= 65535, "maxpool: N/OH exceed grid limits");
    if (v4) {
        dim3 grid((unsigned)ceil_div64((int64_t)OW * (C / 4), 256), OH, N);
        if (k == 3)
            maxpool_fwd_k<4, 3><<<grid, 256, 0, s>>>(x, y, argmax, N, H, W, C, OH, OW, k, sh, sw, pad);
        else if (k == 2)
            maxpool_fwd_k<4, 2><<<grid, 256, 0, s>>>(x, y, argmax, N, H, W, C, OH, OW, k, sh, sw, pad);
        else
            maxpool_fwd_k<4><<<grid, 256, 0, s>>>(x, y, argmax, N, H, W, C, OH, OW, k, sh, sw, pad);
    } else {
        dim3 grid((unsigned)ceil_div64((int64_t)OW * C, 256), OH, N);
        maxpool_fwd_k<1><<<grid, 256, 0, s>>>(x, y, argmax, N, H, W, C, OH, OW, k, sh, sw, pad);
    }
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// Fused BatchNorm apply + ReLU + max-pool forward (see bn_relu_maxpool_fwd_k). Returns XSB_ERR_UNSUPPORTED (nothing
// launched) when the geometry does not let every input pixel be owned by exactly one window or C % 8 != 0: the caller
// then runs xsb_bn_apply (relu) and xsb_maxpool2d_fwd.
int xsb_bn_relu_maxpool_fwd(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                            float* y, uint8_t* argmax, uint8_t* relu_mask, int N, int H, int W, int C, int k, int sh,
                            int sw, int pad, void* stream) {
    XSB_CHECK_ARG(k >= 1 && k <= 15 && sh >= 1 && sw >= 1 && pad >= 0, "maxpool: bad k/stride/pad %d %d %d %d", k, sh,
                  sw, pad);
    const int OH = (H + 2 * pad - k) / sh + 1, OW = (W + 2 * pad - k) / sw + 1;
    XSB_CHECK_ARG(OH > 0 && OW > 0, "maxpool: empty output");
    const bool covered = pad + sh - 1 < k && pad + sw - 1 < k && (H + sh - 1) / sh <= OH && (W + sw - 1) / sw <= OW;
    if (!covered || C % 8 != 0 || !aligned16(x) || !aligned16(y) || !aligned16(mean) || !aligned16(invstd) ||
        !aligned16(gamma) || !aligned16(beta) || (argmax && ((uintptr_t)argmax & 3u) != 0) || N > 65535 || OH > 65535)
        return XSB_ERR_UNSUPPORTED;
    if (k == 3 && sh == 2 && sw == 2 && pad == 1 && (int64_t)H * W * (C / 4) < ((int64_t)1 << 31)) {
        dim3 grid((unsigned)ceil_div64((int64_t)OW * (C / 4), 256), OH, N);
        bn_relu_maxpool_k3s2_k<<<grid, 256, 0, xsb_stream(stream)>>>(x, mean, invstd, gamma, beta, y, argmax, relu_mask, N, H, W,
                                                                    C, OH, OW);
        XSB_LAUNCH_CHECK();
        return XSB_OK;
    }
    const int rows = (FP_TOH - 1) * sh + k, cols = (FP_TOW - 1) * sw + k;
    const size_t smem = (size_t)rows * cols * C * sizeof(float);
    if (smem > 48 * 1024 || (OH + FP_TOH - 1) / FP_TOH > 65535 || 256 % (C / 4) != 0) return XSB_ERR_UNSUPPORTED;
    dim3 grid((OW + FP_TOW - 1) / FP_TOW, (OH + FP_TOH - 1) / FP_TOH, N);
    bn_relu_maxpool_fwd_k<<<grid, 256, smem, xsb_stream(stream)>>>(x, mean, invstd, gamma, beta, y, argmax, relu_mask, N, H, W,
                                                                   C, OH, OW, k, sh, sw, pad);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

int xsb_maxpool2d_bwd(const float* dy, const uint8_t* argmax, float* dx, int N, int H, int W, int C, int k, int sh,
                      int sw, int pad, int accumulate, void* stream) {
    const int OH = (H + 2 * pad - k) / sh + 1, OW = (W + 2 * pad - k) / sw + 1;
    XSB_CHECK_ARG(OH > 0 && OW > 0, "maxpool: empty output");
    cudaStream_t s = xsb_stream(stream);
    const bool v4 = (C % 4 == 0) && aligned16(dy) && aligned16(dx) && (((uintptr_t)argmax & 3u) == 0);
    XSB_CHECK_ARG(N <= 65535 && H <= 65535, "maxpool: N/H exceed grid limits");
    if (v4 && sh == 2 && sw == 2) {
        dim3 g((unsigned)ceil_div64((int64_t)((W + 1) / 2) * (C / 4), 256), (H + 1) / 2, N);
        if (k == 3 && pad == 1) {
            if (accumulate)
                maxpool_bwd_s2_k<true, true, true><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, pad);
            else
                maxpool_bwd_s2_k<false, true, true><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, pad);
        } else if (k <= 3) {
            if (accumulate)
                maxpool_bwd_s2_k<true, true><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, pad);
            else
                maxpool_bwd_s2_k<false, true><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, pad);
        } else if (accumulate)
            maxpool_bwd_s2_k<true><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, pad);
        else
            maxpool_bwd_s2_k<false><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, pad);
    } else if (v4) {
        dim3 g((unsigned)ceil_div64((int64_t)W * (C / 4), 256), H, N);
        if (accumulate)
            maxpool_bwd_k<4, true><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, sh, sw, pad);
        else
            maxpool_bwd_k<4, false><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, sh, sw, pad);
    } else {
        dim3 g((unsigned)ceil_div64((int64_t)W * C, 256), H, N);
        if (accumulate)
            maxpool_bwd_k<1, true><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, sh, sw, pad);
        else
            maxpool_bwd_k<1, false><<<g, 256, 0, s>>>(dy, argmax, dx, N, H, W, C, OH, OW, k, sh, sw, pad);
    }
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

int xsb_avgpool2d_fwd(const float* x, float* y, int N, int H, int W, int C, int k, int sh, int sw, void* stream) {
    const int OH = (H - k) / sh + 1, OW = (W - k) / sw + 1;
    XSB_CHECK_ARG(OH > 0 && OW > 0, "avgpool: empty output");
    XSB_CHECK_ARG(N <= 65535 && OH <= 65535, "avgpool: N/OH exceed grid limits");
    if (C % 4 == 0 && aligned16(x) && aligned16(y)) {
        dim3 grid((unsigned)ceil_div64((int64_t)OW * (C / 4), 128), OH, N);
        avgpool_fwd_k<4><<<grid, 128, 0, xsb_stream(stream)>>>(x, y, N, H, W, C, OH, OW, k, sh, sw);
    } else {
        dim3 grid((unsigned)ceil_div64((int64_t)OW * C, 256), OH, N);
        avgpool_fwd_k<1><<<grid, 256, 0, xsb_stream(stream)>>>(x, y, N, H, W, C, OH, OW, k, sh, sw);
    }
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

int xsb_avgpool2d_bwd(const float* dy, float* dx, int N, int H, int W, int C, int k, int sh, int sw, int accumulate,
                      void* stream) {
    const int OH = (H - k) / sh + 1, OW = (W - k) / sw + 1;
    XSB_CHECK_ARG(OH > 0 && OW > 0, "avgpool: empty output");
    XSB_CHECK_ARG(N <= 65535 && H <= 65535, "avgpool: N/H exceed grid limits");
    if (C % 4 == 0 && aligned16(dy) && aligned16(dx)) {
        dim3 grid((unsigned)ceil_div64((int64_t)W * (C / 4), 128), H, N);
        avgpool_bwd_k<4><<<grid, 128, 0, xsb_stream(stream)>>>(dy, dx, N, H, W, C, OH, OW, k, sh, sw, accumulate);
    } else {
        dim3 grid((unsigned)ceil_div64((int64_t)W * C, 256), H, N);
        avgpool_bwd_k<1><<<grid, 256, 0, xsb_stream(stream)>>>(dy, dx, N, H, W, C, OH, OW, k, sh, sw, accumulate);
    }
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

}  // extern "C"
