// tcgen05 tensor-core implicit-GEMM convolution for sm_100a (XSB_MATH_TF32).
//
//  * operands stay fp32 in HBM (channels-last activations, [OC,KH,KW,C] filters) and are fed to
//    `tcgen05.mma.cta_group::1.kind::tf32` directly: no conversion pass, no im2col buffer;
//  * A tiles are fetched by TMA in IM2COL mode (cp.async.bulk.tensor.4d...im2col): one instruction
//    gathers {32 channels x P pixels} for one filter tap, zero padding and the conv stride handled
//    by the tensor map; B tiles by tiled 2-D TMA; both land in SWIZZLE_128B shared memory;
//  * accumulators live in TMEM (fp32), read back with tcgen05.ld in the epilogue warps;
//  * warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2-5 =
//    epilogue; smem ring of STAGES stages guarded by full/empty mbarriers, tcgen05.commit frees stages.
//
//  fprop : D[pixel, oc]   = sum_{tap,c} patch[pixel,(tap,c)] * w[oc,(tap,c)]   A,B K-major
//  dgrad : the same kernel on dY (zero-inserted for strided convs) with the flipped/transposed filter
//          wt[c,kh',kw',oc], pad' = K-1-pad
//  wgrad : D[(tap,c), oc] = sum_pixel patch[pixel,(tap,c)] * dY[pixel,oc]      A,B MN-major, split-K
//
// Reference arithmetic: xs/nn/functional.py:405-442, xs/nn/grad_fn.py:186-203 (see gemm_conv.cu).
// Shapes outside the envelope (C or OC not multiples of 32/64, strided dgrad) return
// XSB_ERR_UNSUPPORTED from the xsb_tc_* probes and the caller runs the fp32 FFMA kernel instead.
#include "tc_common.cuh"

namespace tc {

// ------------------------------------------------------------------------------------------ fprop / dgrad
struct FpropParams {
    float* out;          // [M, ldo] row-major (channels-last output)
    const float* bias;   // [OCtot] or null
    int M, ldo;          // M = N*OH*OW output pixels, ldo = total output channels
    int C, KH, KW, OH, OW, sh, sw, pad;
    int accumulate;
};

template <int BN, int STAGES, int OCC>
__global__ void __launch_bounds__(kThreads, OCC)
conv_fprop_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int A_BYTES = BM * A_ROW_BYTES;
    constexpr int B_BYTES = BN * A_ROW_BYTES;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cblocks = p.C / 32;
    const int num_kb = p.KH * p.KW * cblocks;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // first pixel of the tile -> window origin in the input image
            const int hw = p.OH * p.OW;
            const int n_img = m0 / hw, rem = m0 - n_img * hw;
            const int oh0 = rem / p.OW, ow0 = rem - oh0 * p.OW;
            const int w0 = ow0 * p.sw - p.pad, h0 = oh0 * p.sh - p.pad;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                const int tap = kb / cblocks, cb = kb - tap * cblocks;
                const int kh = tap / p.KW, kw = tap - kh * p.KW;
                tma_load_im2col(sA + s * A_BYTES, &tmA, &full[s], cb * 32, w0, h0, n_img, (uint16_t)kw, (uint16_t)kh);
                tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], tap * p.C + cb * 32, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, 0, 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full[s], (kb / STAGES) & 1);
                tc_fence_after();
                const uint64_t da = make_desc(smem_u32(sA + s * A_BYTES), 16, 1024);
                const uint64_t db = make_desc(smem_u32(sB + s * B_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 4 x (K = 8 tf32 = 32 B) per 128-byte row
                    umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                umma_commit(&empty[s]);
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = q * 32 + lane;
        const int m = m0 + row;
        float* orow = p.out + (int64_t)m * p.ldo + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_wait_ld();
            if (m < p.M) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                           __uint_as_float(r[j + 3]));
                    if (p.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
                        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                    }
                    float4* dst = reinterpret_cast<float4*>(orow + c0 + j);
                    if (p.accumulate) {
                        const float4 o = *dst;
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    *dst = v;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
    }
}

// ------------------------------------------------------------------------------------------ wgrad
struct WgradParams {
    float* dw;           // [OC, KH*KW, C]
    int P;               // N*OH*OW pixels (the reduction dimension)
    int C, OC, KH, KW, OH, OW, sh, sw, pad;
    int kb_per_split;    // pixel blocks per grid.z slice
    int use_atomics;     // 1: red.add into dw (split-K or accumulate); 0: plain store
};

template <int BN, int BKP, int STAGES>
__global__ void __launch_bounds__(kThreads, (STAGES * (4 + BN / 32) * BKP * 128 <= 100 * 1024) ? 2 : 1)
conv_wgrad_tc_k(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int BOX_BYTES = BKP * A_ROW_BYTES;       // one {32 elements x BKP pixels} box
    constexpr int A_BYTES = 4 * BOX_BYTES;             // 128 rows of D = 4 boxes
    constexpr int B_BYTES = (BN / 32) * BOX_BYTES;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = p.KH * p.KW;
    const int r0 = blockIdx.x * BM;      // first row of D: rows are (tap, c)
    const int n0 = blockIdx.y * BN;      // first output channel
    const int total_kb = (p.P + BKP - 1) / BKP;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    int kb_end = kb_begin + p.kb_per_split;
    if (kb_end > total_kb) kb_end = total_kb;
    const int num_kb = kb_end - kb_begin;   // >= 1 by construction of the grid

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmDy);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int box_c[4], box_kh[4], box_kw[4], nvalid = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int mrow = r0 + i * 32;
                const int tap = mrow / p.C;
                box_c[i] = mrow - tap * p.C;
                box_kh[i] = tap / p.KW;
                box_kw[i] = tap - box_kh[i] * p.KW;
                if (tap < taps) nvalid = i + 1;   // valid boxes are a prefix
            }
            const int hw = p.OH * p.OW;
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], nvalid * BOX_BYTES + B_BYTES);
                const int p0 = (kb_begin + i) * BKP;
                const int n_img = p0 / hw, rem = p0 - n_img * hw;
                const int oh0 = rem / p.OW, ow0 = rem - oh0 * p.OW;
                const int w0 = ow0 * p.sw - p.pad, h0 = oh0 * p.sh - p.pad;
                for (int b = 0; b < nvalid; ++b)
                    tma_load_im2col(sA + s * A_BYTES + b * BOX_BYTES, &tmX, &full[s], box_c[b], w0, h0, n_img,
                                    (uint16_t)box_kw[b], (uint16_t)box_kh[b]);
#pragma unroll
                for (int b = 0; b < BN / 32; ++b)
                    tma_load_2d(sB + s * B_BYTES + b * BOX_BYTES, &tmDy, &full[s], n0 + b * 32, p0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, 1, 1);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&full[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint64_t da = make_desc(smem_u32(sA + s * A_BYTES), BOX_BYTES, 512, kLayoutSW128Base32B);
                const uint64_t db = make_desc(smem_u32(sB + s * B_BYTES), BOX_BYTES, 512, kLayoutSW128Base32B);
#pragma unroll
                for (int k = 0; k < BKP / 8; ++k)  // K = 8 pixels = two 4-row swizzle atoms = 1024 bytes
                    umma_tf32(tmem_base, da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), idesc, (i | k) != 0);
                umma_commit(&empty[s]);
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int mrow = r0 + q * 32 + lane;
        const int tap = mrow / p.C, c = mrow - tap * p.C;
        const bool valid = tap < taps;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_wait_ld();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int oc = n0 + c0 + j;
                    float* dst = p.dw + ((int64_t)oc * taps + tap) * p.C + c;   // lanes -> consecutive c: coalesced
                    if (p.use_atomics)
                        atomicAdd(dst, __uint_as_float(r[j]));
                    else
                        *dst = __uint_as_float(r[j]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
    }
}


// ------------------------------------------------------------------------------------------ stem (C <= 4)
// First-layer convolutions have 1..4 input channels (K = KH*KW*C = 147 for the 7x7/2 ResNet stem): no 32-channel
// k-block exists. The image is first copied into a zero-bordered 4-channel buffer xp[N,Hp,Wp,4]; then for a fixed
// filter row kh the taps (kw, c) of one output pixel are 32 CONTIGUOUS floats starting at pixel
// (oh*sh+kh, ow*sw) -- 8 pixels x 4 channels, columns kw >= KW / c >= C meet zero filter entries. A 5-D TILED tensor
// map with overlapping strides {ow: sw*16 B, oh: sh*Wp*16 B, kh: Wp*16 B, n} exposes exactly those rows, so the
// mainloop is the usual TMA -> tcgen05 pipeline with KH k-blocks of 32.
constexpr int STEM_TW = 16, STEM_TH = 8;    // fprop tile: 16 x 8 output pixels = 128 rows of D
constexpr int STEM_WTH = 4;                 // wgrad pixel block: 16 x 4 = 64 pixels (the K dimension)

struct StemParams {
    float* out;           // fprop: y [N,OH,OW,OC] ; wgrad: dw [OC,KH,KW,C]
    const float* bias;
    int N, OH, OW, OC, KH, KW, C;
    int tiles_w, tiles_h;
    int kb_per_split, use_atomics;
};

// xp[n, h+pad, w+pad, 0..3] = x[n,h,w,0..C-1] (zero elsewhere); every element of xp is written
__global__ void stem_pad_c4_k(const float* __restrict__ x, float4* __restrict__ xp, int N, int H, int W, int C, int Hp,
                              int Wp, int pad) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)N * Hp * Wp;
    if (t >= total) return;
    const int wp = (int)(t % Wp);
    const int64_t r = t / Wp;
    const int hp = (int)(r % Hp);
    const int n = (int)(r / Hp);
    const int h = hp - pad, w = wp - pad;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (h >= 0 && h < H && w >= 0 && w < W) {
        const float* src = x + (((int64_t)n * H + h) * W + w) * C;
        for (int c = 0; c < C; ++c) v[c] = __ldg(src + c);
    }
    xp[t] = make_float4(v[0], v[1], v[2], v[3]);
}

// wq[oc, kh, kw(8), c(4)] = w[oc, kh, kw, c] (zero for kw >= KW or c >= C)
__global__ void stem_pack_filter_k(const float* __restrict__ w, float* __restrict__ wq, int OC, int KH, int KW, int C) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= OC * KH * 32) return;
    const int j = t % 32, kh = (t / 32) % KH, oc = t / (32 * KH);
    const int kw = j / 4, c = j % 4;
    wq[t] = (kw < KW && c < C) ? w[((oc * KH + kh) * KW + kw) * C + c] : 0.f;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
stem_fprop_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int A_BYTES = BM * A_ROW_BYTES;
    constexpr int B_BYTES = BN * A_ROW_BYTES;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = p.KH;
    const int t = blockIdx.x;
    const int tw_i = t % p.tiles_w, th_i = (t / p.tiles_w) % p.tiles_h, n_img = t / (p.tiles_w * p.tiles_h);
    const int n0 = blockIdx.y * BN;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                tma_load_5d(sA + s * A_BYTES, &tmA, &full[s], 0, tw_i * STEM_TW, th_i * STEM_TH, kb, n_img);
                tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], kb * 32, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, 0, 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full[s], (kb / STAGES) & 1);
                tc_fence_after();
                const uint64_t da = make_desc(smem_u32(sA + s * A_BYTES), 16, 1024);
                const uint64_t db = make_desc(smem_u32(sB + s * B_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                umma_commit(&empty[s]);
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = q * 32 + lane;
        const int oh = th_i * STEM_TH + row / STEM_TW, ow = tw_i * STEM_TW + row % STEM_TW;
        const bool valid = oh < p.OH && ow < p.OW;
        float* orow = p.out + (((int64_t)n_img * p.OH + oh) * p.OW + ow) * p.OC + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_wait_ld();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                           __uint_as_float(r[j + 3]));
                    if (p.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
                        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                    }
                    *reinterpret_cast<float4*>(orow + c0 + j) = v;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
    }
}

// dwq[(kh, kw, c4), oc] = sum over pixel blocks of xp-rows^T . dY ; both operands MN-major (pixels = K)
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
stem_wgrad_tc_k(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy, const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int BKP = STEM_TW * STEM_WTH;            // 64 pixels per k-block
    constexpr int BOX_BYTES = BKP * A_ROW_BYTES;
    constexpr int A_BYTES = 4 * BOX_BYTES;
    constexpr int B_BYTES = (BN / 32) * BOX_BYTES;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kh0 = blockIdx.x * 4;                    // this CTA's 4 filter rows = 128 rows of D
    const int n0 = blockIdx.y * BN;
    const int blocks_per_img = p.tiles_w * p.tiles_h;  // tiles_h counts 4-row blocks here
    const int total_kb = p.N * blocks_per_img;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    int kb_end = kb_begin + p.kb_per_split;
    if (kb_end > total_kb) kb_end = total_kb;
    const int num_kb = kb_end - kb_begin;
    int nvalid = p.KH - kh0;
    if (nvalid > 4) nvalid = 4;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmDy);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], nvalid * BOX_BYTES + B_BYTES);
                const int kb = kb_begin + i;
                const int n_img = kb / blocks_per_img, rem = kb - n_img * blocks_per_img;
                const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
                for (int b = 0; b < nvalid; ++b)
                    tma_load_5d(sA + s * A_BYTES + b * BOX_BYTES, &tmX, &full[s], 0, tw_i * STEM_TW, th_i * STEM_WTH,
                                kh0 + b, n_img);
#pragma unroll
                for (int b = 0; b < BN / 32; ++b)
                    tma_load_4d(sB + s * B_BYTES + b * BOX_BYTES, &tmDy, &full[s], n0 + b * 32, tw_i * STEM_TW,
                                th_i * STEM_WTH, n_img);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, 1, 1);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&full[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint64_t da = make_desc(smem_u32(sA + s * A_BYTES), BOX_BYTES, 512, kLayoutSW128Base32B);
                const uint64_t db = make_desc(smem_u32(sB + s * B_BYTES), BOX_BYTES, 512, kLayoutSW128Base32B);
#pragma unroll
                for (int k = 0; k < BKP / 8; ++k)
                    umma_tf32(tmem_base, da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), idesc, (i | k) != 0);
                umma_commit(&empty[s]);
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = q * 32 + lane;
        const int kh = kh0 + row / 32, j = row % 32, kw = j / 4, c = j % 4;
        const bool valid = kh < p.KH && kw < p.KW && c < p.C;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_wait_ld();
            if (valid) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const int oc = n0 + c0 + jj;
                    float* dst = p.out + (((int64_t)oc * p.KH + kh) * p.KW + kw) * p.C + c;
                    if (p.use_atomics)
                        atomicAdd(dst, __uint_as_float(r[jj]));
                    else
                        *dst = __uint_as_float(r[jj]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
    }
}

// wt[c, KH-1-kh, KW-1-kw, oc] = w[oc, kh, kw, c]: the filter of the equivalent forward conv on dY
__global__ void flip_transpose_filter_k(const float* __restrict__ w, float* __restrict__ wt, int OC, int taps, int C) {
    __shared__ float tile[32][33];
    const int tap = blockIdx.z;
    const int c0 = blockIdx.x * 32, oc0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int oc = oc0 + j, c = c0 + threadIdx.x;
        if (oc < OC && c < C) tile[j][threadIdx.x] = w[((int64_t)oc * taps + tap) * C + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, oc = oc0 + threadIdx.x;
        if (oc < OC && c < C) wt[((int64_t)c * taps + (taps - 1 - tap)) * OC + oc] = tile[threadIdx.x][j];
    }
}

// up[n, oh*sh, ow*sw, :] = dy[n, oh, ow, :] (up is pre-zeroed): the zero-insertion that turns a strided
// dgrad into a stride-1 convolution of the up-sampled dY with the flipped filter.
__global__ void zero_insert_k(const float4* __restrict__ dy, float4* __restrict__ up, int OH, int OW, int C4, int Hu,
                              int Wu, int sh, int sw, int64_t total) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C4);
    int64_t r = t / C4;
    const int ow = (int)(r % OW);
    r /= OW;
    const int oh = (int)(r % OH);
    const int n = (int)(r / OH);
    up[(((int64_t)n * Hu + (int64_t)oh * sh) * Wu + (int64_t)ow * sw) * C4 + c] = dy[t];
}

// ------------------------------------------------------------------------------------------ host side
// im2col view of a channels-last activation [N,H,W,Cx]: box {32 channels, pixels_per_col output pixels};
// window origins span [-pad, W+pad-KW] (step = conv stride), taps are given per load as {kw, kh} offsets.
static bool make_map_im2col(CUtensorMap* m, const float* base, int N, int H, int W, int Cx, int KH, int KW, int sh,
                            int sw, int pad, uint32_t pixels_per_col,
                            CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int extra_w = 0, int extra_h = 0) {
    cuuint64_t dims[4] = {(cuuint64_t)Cx, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cx * 4, (cuuint64_t)W * Cx * 4, (cuuint64_t)H * W * Cx * 4};
    int lower[2] = {-pad, -pad};
    // extra_*: additional window origins past the symmetric-padding range (their taps read the zero OOB fill)
    int upper[2] = {pad - (KW - 1) + extra_w, pad - (KH - 1) + extra_h};
    cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
    CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, lower,
                                 upper, 32, pixels_per_col, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    // Same driver workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp): for tensors smaller
    // than 128 KiB drivers <= 13.1 set a descriptor bit that breaks im2col traversal.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (uint64_t)N * H * W * Cx * 4 < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
    return true;
}

static uint64_t g_tc_launches = 0, g_ffma_fallbacks = 0;

// XSB_TC_OCC=1|2 (env, read once): CTAs per SM for the fprop kernel. 2 (default) halves the smem ring of each CTA
// so that one CTA's epilogue overlaps the other's mainloop.
static int tc_occupancy() {
    static int occ = 0;
    if (occ == 0) {
        const char* e = getenv("XSB_TC_OCC");
        occ = (e && e[0] == '1') ? 1 : 2;
    }
    return occ;
}

template <int BN, int STAGES, int OCC>
static int launch_fprop(const CUtensorMap& a, const CUtensorMap& b, const FpropParams& p, int n_tiles, cudaStream_t s) {
    constexpr int smem = STAGES * (BM * A_ROW_BYTES + BN * A_ROW_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(conv_fprop_tc_k<BN, STAGES, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid((p.M + BM - 1) / BM, n_tiles);
    conv_fprop_tc_k<BN, STAGES, OCC><<<grid, kThreads, smem, s>>>(a, b, p);
    XSB_LAUNCH_CHECK();
    ++g_tc_launches;
    return XSB_OK;
}

// D[M, OCx] (+)= conv of channels-last `in` [N,H,W,Cx] with filter `flt` [OCx, KH*KW*Cx] (+ bias)
static int fprop_dispatch(const float* in, const float* flt, const float* bias, float* out, int N, int H, int W, int Cx,
                          int OCx, int KH, int KW, int sh, int sw, int pad, int OH, int OW, int accumulate,
                          cudaStream_t s, int extra_w = 0, int extra_h = 0) {
    if (!load_driver_entry_points()) {
        xsb_set_error("cuTensorMapEncode* entry points unavailable");
        return XSB_ERR_CUDA;
    }
    int BN = (OCx % 256 == 0) ? 256 : (OCx % 128 == 0) ? 128 : 64;
    {   // a 256-wide tile that leaves SMs idle loses to twice as many 128-wide CTAs (XSB_TC_NARROW=0 disables)
        static const int narrow = [] { const char* e = getenv("XSB_TC_NARROW"); return (e && e[0] == '1') ? 1 : 0; }();   // measured slower on layer4: off
        const int64_t ctas = (int64_t)((N * OH * OW + BM - 1) / BM) * (OCx / BN);
        if (narrow && BN == 256 && ctas < xsb_sm_count_cached()) BN = 128;
    }
    CUtensorMap ta, tb;
    if (!make_map_im2col(&ta, in, N, H, W, Cx, KH, KW, sh, sw, pad, BM, CU_TENSOR_MAP_SWIZZLE_128B, extra_w, extra_h) ||
        !make_map_2d(&tb, flt, (uint64_t)OCx, (uint64_t)KH * KW * Cx, (uint32_t)BN)) {
        xsb_set_error("cuTensorMapEncode failed (fprop N=%d H=%d W=%d C=%d OC=%d k=%dx%d s=%d p=%d)", N, H, W, Cx, OCx, KH,
                      KW, sh, pad);
        return XSB_ERR_CUDA;
    }
    FpropParams p{out, bias, N * OH * OW, OCx, Cx, KH, KW, OH, OW, sh, sw, pad, accumulate};
    if (tc_occupancy() == 1) {
        if (BN == 256) return launch_fprop<256, 4, 1>(ta, tb, p, OCx / 256, s);
        if (BN == 128) return launch_fprop<128, 6, 1>(ta, tb, p, OCx / 128, s);
        return launch_fprop<64, 8, 1>(ta, tb, p, OCx / 64, s);
    }
    // measured on B200 (tests/op_bench.py, R224): 2 CTAs/SM is +45 % for BN=64 and +19 % for BN=128, but -25 % for
    // BN=256 (only 2 stages fit) -> the widest tile keeps one CTA per SM with a 4-stage ring
    if (BN == 256) return launch_fprop<256, 4, 1>(ta, tb, p, OCx / 256, s);
    if (BN == 128) return launch_fprop<128, 3, 2>(ta, tb, p, OCx / 128, s);
    return launch_fprop<64, 4, 2>(ta, tb, p, OCx / 64, s);
}

template <int BN, int BKP, int STAGES>
static int launch_wgrad(const CUtensorMap& x, const CUtensorMap& dy, const WgradParams& p, dim3 grid, cudaStream_t s) {
    constexpr int smem = STAGES * ((4 + BN / 32) * BKP * A_ROW_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_k<BN, BKP, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    conv_wgrad_tc_k<BN, BKP, STAGES><<<grid, kThreads, smem, s>>>(x, dy, p);
    XSB_LAUNCH_CHECK();
    ++g_tc_launches;
    return XSB_OK;
}

struct StemPlan {
    int Hp, Wp;
    int64_t xp_floats, wq_floats;
};
static bool stem_ok(const ConvGeom& g) {
    return g.C <= 4 && g.KW <= 8 && g.KH <= 64 && g.OC % 64 == 0 && g.sw * 16 % 16 == 0;
}
static StemPlan stem_plan(const ConvGeom& g) {
    StemPlan p;
    p.Hp = g.H + 2 * g.pad;
    const int need_h = (g.OH - 1) * g.sh + g.KH;
    if (p.Hp < need_h) p.Hp = need_h;
    p.Wp = g.W + 2 * g.pad;
    const int need_w = (g.OW - 1) * g.sw + 8;
    if (p.Wp < need_w) p.Wp = need_w;
    p.xp_floats = ((int64_t)g.N * p.Hp * p.Wp * 4 + 255) / 256 * 256;
    p.wq_floats = ((int64_t)g.OC * g.KH * 32 + 255) / 256 * 256;
    return p;
}
// the overlapping-window view of xp: {32 floats, OW, OH, KH, N}
static bool stem_map_x(CUtensorMap* m, const float* xp, const ConvGeom& g, const StemPlan& pl, int box_w, int box_h,
                       CUtensorMapSwizzle swz) {
    cuuint64_t dims[5] = {32, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.KH, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)g.sw * 16, (cuuint64_t)g.sh * pl.Wp * 16, (cuuint64_t)pl.Wp * 16,
                             (cuuint64_t)pl.Hp * pl.Wp * 16};
    cuuint32_t box[5] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
    return make_map_tiled(m, xp, 5, dims, strides, box, swz);
}

static int stem_prepare(const float* x, const ConvGeom& g, const StemPlan& pl, float* xp, cudaStream_t s) {
    const int64_t total = (int64_t)g.N * pl.Hp * pl.Wp;
    stem_pad_c4_k<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(x, reinterpret_cast<float4*>(xp), g.N, g.H, g.W, g.C,
                                                                  pl.Hp, pl.Wp, g.pad);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

static int stem_fwd(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g, float* ws,
                    int64_t ws_floats, cudaStream_t s) {
    const StemPlan pl = stem_plan(g);
    if (!ws || ws_floats < pl.xp_floats + pl.wq_floats) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    float* xp = ws;
    float* wq = ws + pl.xp_floats;
    CUtensorMap ta, tb;
    if (!stem_map_x(&ta, xp, g, pl, STEM_TW, STEM_TH, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_map_2d(&tb, wq, (uint64_t)g.OC, (uint64_t)g.KH * 32, 64))
        return XSB_ERR_UNSUPPORTED;   // driver rejected the overlapping-stride map: FFMA kernel takes the call
    int rc = stem_prepare(x, g, pl, xp, s);
    if (rc != XSB_OK) return rc;
    stem_pack_filter_k<<<(g.OC * g.KH * 32 + 255) / 256, 256, 0, s>>>(w, wq, g.OC, g.KH, g.KW, g.C);
    XSB_LAUNCH_CHECK();
    StemParams p{y, bias, g.N, g.OH, g.OW, g.OC, g.KH, g.KW, g.C, (g.OW + STEM_TW - 1) / STEM_TW,
                 (g.OH + STEM_TH - 1) / STEM_TH, 0, 0};
    constexpr int STAGES = 4, BN = 64;
    constexpr int smem = STAGES * (BM * A_ROW_BYTES + BN * A_ROW_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(stem_fprop_tc_k<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(g.N * p.tiles_w * p.tiles_h, g.OC / BN);
    stem_fprop_tc_k<BN, STAGES><<<grid, kThreads, smem, s>>>(ta, tb, p);
    XSB_LAUNCH_CHECK();
    ++g_tc_launches;
    return XSB_OK;
}

static int stem_wgrad(const float* dy, const float* x, float* dw, const ConvGeom& g, int accumulate, float* ws,
                      int64_t ws_floats, cudaStream_t s) {
    const StemPlan pl = stem_plan(g);
    if (!ws || ws_floats < pl.xp_floats) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    float* xp = ws;
    CUtensorMap tx, tdy;
    cuuint64_t ddims[4] = {(cuuint64_t)g.OC, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
    cuuint64_t dstr[3] = {(cuuint64_t)g.OC * 4, (cuuint64_t)g.OW * g.OC * 4, (cuuint64_t)g.OH * g.OW * g.OC * 4};
    cuuint32_t dbox[4] = {32, STEM_TW, STEM_WTH, 1};
    if (!stem_map_x(&tx, xp, g, pl, STEM_TW, STEM_WTH, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map_tiled(&tdy, dy, 4, ddims, dstr, dbox, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return XSB_ERR_UNSUPPORTED;
    int rc = stem_prepare(x, g, pl, xp, s);
    if (rc != XSB_OK) return rc;
    constexpr int STAGES = 4, BN = 64;
    const int tiles_w = (g.OW + STEM_TW - 1) / STEM_TW, tiles_h = (g.OH + STEM_WTH - 1) / STEM_WTH;
    const int total_kb = g.N * tiles_w * tiles_h;
    const int m_tiles = (g.KH + 3) / 4, n_tiles = g.OC / BN;
    int splits = (2 * xsb_sm_count_cached() + m_tiles * n_tiles - 1) / (m_tiles * n_tiles);
    if (splits > total_kb / 4) splits = total_kb / 4;
    if (splits < 1) splits = 1;
    const int kb_per = (total_kb + splits - 1) / splits;
    splits = (total_kb + kb_per - 1) / kb_per;
    const int use_atomics = (splits > 1 || accumulate) ? 1 : 0;
    if (use_atomics && !accumulate)
        XSB_CUDA(cudaMemsetAsync(dw, 0, (size_t)g.OC * g.KH * g.KW * g.C * sizeof(float), s));
    StemParams p{dw, nullptr, g.N, g.OH, g.OW, g.OC, g.KH, g.KW, g.C, tiles_w, tiles_h, kb_per, use_atomics};
    constexpr int smem = STAGES * ((4 + BN / 32) * STEM_TW * STEM_WTH * A_ROW_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(stem_wgrad_tc_k<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(m_tiles, n_tiles, splits);
    stem_wgrad_tc_k<BN, STAGES><<<grid, kThreads, smem, s>>>(tx, tdy, p);
    XSB_LAUNCH_CHECK();
    ++g_tc_launches;
    return XSB_OK;
}

}  // namespace tc


using namespace tc;

// ------------------------------------------------------------------------------------------ entry points
// Envelope of the tensor-core kernels; everything else is answered with XSB_ERR_UNSUPPORTED and the
// caller (gemm_conv.cu) runs the fp32 FFMA kernel, counting the event in xsb_tc_stats.
static bool fprop_ok(int Cx, int OCx, const void* a, const void* b, const void* c) {
    return Cx % 32 == 0 && OCx % 64 == 0 && aligned16(a) && aligned16(b) && aligned16(c);
}

int xsb_tc_gemm(const float*, const float*, float*, const float*, int, int, int, int, int, int, float*, int64_t,
                cudaStream_t) {
    return XSB_ERR_UNSUPPORTED;   // Dense layers of the hot path are < 1 % of the FLOPs: FFMA kernel
}

int xsb_tc_conv_fwd(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g, float* ws,
                    int64_t ws_floats, cudaStream_t s) {
    if (stem_ok(g) && aligned16(x) && aligned16(w) && aligned16(y)) return stem_fwd(x, w, bias, y, g, ws, ws_floats, s);
    if (!fprop_ok(g.C, g.OC, x, w, y) || g.KW > 255 || g.KH > 255) return XSB_ERR_UNSUPPORTED;
    return fprop_dispatch(x, w, bias, y, g.N, g.H, g.W, g.C, g.OC, g.KH, g.KW, g.sh, g.sw, g.pad, g.OH, g.OW, 0, s);
}

int xsb_tc_conv_dgrad(const float* dy, const float* w, float* dx, const ConvGeom& g, int accumulate, float* ws,
                      int64_t ws_floats, cudaStream_t s) {
    // dX = conv_stride1(dY (zero-inserted when the forward conv was strided), flip/transpose(w)), pad' = K-1-pad
    if (g.pad > g.KH - 1 || g.pad > g.KW - 1 || g.KH != g.KW) return XSB_ERR_UNSUPPORTED;
    if (!fprop_ok(g.OC, g.C, dy, w, dx)) return XSB_ERR_UNSUPPORTED;
    const int64_t wn = (int64_t)g.OC * g.KH * g.KW * g.C;
    const bool strided = g.sh != 1 || g.sw != 1;
    const int Hu = (g.OH - 1) * g.sh + 1, Wu = (g.OW - 1) * g.sw + 1;
    const int64_t un = strided ? (int64_t)g.N * Hu * Wu * g.OC : 0;
    const int64_t wn_al = (wn + 255) / 256 * 256;
    if (!ws || ws_floats < wn_al + un) return XSB_ERR_UNSUPPORTED;
    const int taps = g.KH * g.KW;
    dim3 grid((g.C + 31) / 32, (g.OC + 31) / 32, taps), block(32, 8);
    flip_transpose_filter_k<<<grid, block, 0, s>>>(w, ws, g.OC, taps, g.C);
    XSB_LAUNCH_CHECK();
    const int padp = g.KH - 1 - g.pad;
    const float* src = dy;
    if (strided) {
        float* up = ws + wn_al;
        XSB_CUDA(cudaMemsetAsync(up, 0, (size_t)un * sizeof(float), s));
        const int64_t total = (int64_t)g.N * g.OH * g.OW * (g.OC / 4);
        zero_insert_k<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(dy),
                                                                      reinterpret_cast<float4*>(up), g.OH, g.OW,
                                                                      g.OC / 4, Hu, Wu, g.sh, g.sw, total);
        XSB_LAUNCH_CHECK();
        src = up;
    }
    // rows/cols of dX past the last forward window get no gradient: extend the im2col box so they are produced (as 0)
    const int extra_h = g.H - (Hu + 2 * padp - g.KH + 1), extra_w = g.W - (Wu + 2 * padp - g.KW + 1);
    if (extra_h < 0 || extra_w < 0 || extra_h > 64 || extra_w > 64) return XSB_ERR_UNSUPPORTED;
    return fprop_dispatch(src, ws, nullptr, dx, g.N, strided ? Hu : g.OH, strided ? Wu : g.OW, g.OC, g.C, g.KH, g.KW, 1, 1,
                          padp, g.H, g.W, accumulate, s, extra_w, extra_h);
}

int xsb_tc_conv_wgrad(const float* dy, const float* x, float* dw, const ConvGeom& g, int accumulate, float* ws,
                      int64_t ws_floats, cudaStream_t s) {
    if (stem_ok(g) && aligned16(dy) && aligned16(x)) return stem_wgrad(dy, x, dw, g, accumulate, ws, ws_floats, s);
    if (g.C % 32 != 0 || g.OC % 64 != 0 || !aligned16(dy) || !aligned16(x) || !aligned16(dw)) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) {
        xsb_set_error("cuTensorMapEncode* entry points unavailable");
        return XSB_ERR_CUDA;
    }
    const int P = g.N * g.OH * g.OW;
    static const int wg_bn_cap = [] { const char* e = getenv("XSB_TC_WG_BN"); return e ? atoi(e) : 128; }();   // measured: 128-wide tiles + 64-pixel k-blocks beat 256 x 32 (layer4 93 -> 73 us)
    int BN = (g.OC % 256 == 0) ? 256 : (g.OC % 128 == 0) ? 128 : 64;
    if (BN > wg_bn_cap) BN = wg_bn_cap;
    const int BKP = BN == 256 ? 32 : 64;
    CUtensorMap tx, tdy;
    if (!make_map_im2col(&tx, x, g.N, g.H, g.W, g.C, g.KH, g.KW, g.sh, g.sw, g.pad, (uint32_t)BKP,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map_2d(&tdy, dy, (uint64_t)P, (uint64_t)g.OC, (uint32_t)BKP, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) {
        xsb_set_error("cuTensorMapEncode failed (wgrad)");
        return XSB_ERR_CUDA;
    }
    const int rows = g.KH * g.KW * g.C;
    const int m_tiles = (rows + BM - 1) / BM, n_tiles = g.OC / BN;
    const int total_kb = (P + BKP - 1) / BKP;
    int splits = (2 * xsb_sm_count_cached() + m_tiles * n_tiles - 1) / (m_tiles * n_tiles);
    if (splits > total_kb / 4) splits = total_kb / 4;
    if (splits < 1) splits = 1;
    int kb_per = (total_kb + splits - 1) / splits;
    splits = (total_kb + kb_per - 1) / kb_per;
    const int use_atomics = (splits > 1 || accumulate) ? 1 : 0;
    if (use_atomics && !accumulate) XSB_CUDA(cudaMemsetAsync(dw, 0, (size_t)g.OC * rows * sizeof(float), s));
    WgradParams p{dw, P, g.C, g.OC, g.KH, g.KW, g.OH, g.OW, g.sh, g.sw, g.pad, kb_per, use_atomics};
    dim3 grid(m_tiles, n_tiles, splits);
    static const int wg_occ = [] { const char* e = getenv("XSB_TC_WG_OCC"); return (e && e[0] == '2') ? 2 : 1; }();   // measured: no gain from 2 CTAs/SM
    if (BN == 256) return launch_wgrad<256, 32, 4>(tx, tdy, p, grid, s);
    if (BN == 128) return launch_wgrad<128, 64, 3>(tx, tdy, p, grid, s);
    if (wg_occ == 2) return launch_wgrad<64, 64, 2>(tx, tdy, p, grid, s);   // 2 x 48 KB stages: two CTAs per SM
    return launch_wgrad<64, 64, 4>(tx, tdy, p, grid, s);
}

extern "C" {
// counters: [0] tensor-core kernel launches, [1] GEMM-shaped calls that ran the fp32 FFMA kernel in TF32 mode
int xsb_tc_stats(uint64_t* out2) {
    out2[0] = tc::g_tc_launches;
    out2[1] = tc::g_ffma_fallbacks;
    return XSB_OK;
}
}
void xsb_tc_note_ffma() { ++tc::g_ffma_fallbacks; }
