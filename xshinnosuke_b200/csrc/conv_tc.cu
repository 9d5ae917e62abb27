// tcgen05 tensor-core implicit-GEMM convolution for sm_100a (XSB_MATH_TF32).
//
//  * operands stay fp32 in HBM (channels-last activations, [OC,KH,KW,C] filters) and are fed to
//    `tcgen05.mma.kind::tf32` directly: no conversion pass, no im2col buffer;
//  * A tiles are fetched by TMA in IM2COL mode (cp.async.bulk.tensor.4d...im2col): one instruction
//    gathers {32 channels x 128 pixels} for one filter tap, zero padding and the conv stride handled
//    by the tensor map; B tiles by tiled 2-D TMA; both land in SWIZZLE_128B shared memory;
//  * accumulators live in TMEM (fp32); the epilogue warps read them with tcgen05.ld, add the bias, stage 32 x 32
//    blocks in swizzled shared memory (the idle operand ring) and leave through TMA stores / TMA reduce-adds,
//    gathering BatchNorm partial statistics on the way (epilogue_store_tma);
//  * warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2-5 = epilogue; the
//    role loops run warp-uniformly and issue under elect.sync (tc_common.cuh: why); smem ring of STAGES stages
//    guarded by full/empty mbarriers, tcgen05.commit frees stages.
//
//  fprop : D[pixel, oc]   = sum_{tap,c} patch[pixel,(tap,c)] * w[oc,(tap,c)]   A,B K-major
//          conv_fprop_tc_k (one CTA per 128 x N tile, N = a column part of <= 256), conv_fprop_tc2_k (CTA pairs,
//          cta_group::2, opt-in), conv_halo_tc_k (conv_halo_tc.cuh: filter-resident, persistent; stem and 64->64 3x3)
//  dgrad : stride-1 convolutions of dY, one per parity class of dX, with gathered sub-filters (tc_conv_dgrad)
//  wgrad : D[(tap,c), oc] = sum_pixel patch[pixel,(tap,c)] * dY[pixel,oc]      A,B MN-major, split-K
//          conv_wgrad_tc_k, stem_wgrad_tc_k, conv_wgrad_halo_k (conv_wgrad_halo_tc.cuh)
//
// Reference arithmetic: xs/nn/functional.py:405-442, xs/nn/grad_fn.py:186-203 (see gemm_conv.cu).
// Shapes outside the envelope (C not a multiple of 32 and > 4, OC not a multiple of 64) return
// XSB_ERR_UNSUPPORTED from the xsb_tc_* probes and the caller runs the fp32 FFMA kernel instead.
#include "conv_halo_tc.cuh"
#include "conv_wgrad_halo_tc.cuh"
#include "conv_pers_tc.cuh"
#include "gemm_tc.cuh"

namespace tc {

// ------------------------------------------------------------------------------------------ fprop / dgrad
struct FpropParams {
    float* out;          // [M, ldo] row-major (channels-last output)
    const float* bias;   // [OCtot] or null
    int M, ldo;          // M = N*OH*OW output pixels, ldo = total output channels
    int C, KH, KW, OH, OW, sh, sw, pad_h, pad_w;
    int accumulate;
    // strided destination (dgrad of a strided conv: one launch produces one parity class of dX):
    // GEMM row (n,i,j) lands on pixel (n, o_a + i*o_sh, o_b + j*o_sw) of an [N, o_H, o_W, ldo] tensor
    int o_strided, o_H, o_W, o_sh, o_sw, o_a, o_b;
    // BatchNorm partial statistics (null = off): stats[(m_tile*3 + {0:n, 1:mean, 2:M2}) * ldo + channel]
    float* stats;
    // this launch covers m-tiles m_tile0 .. m_tile0 + gridDim.x - 1; CTA column blockIdx.y computes output channels
    // part_n0[y] .. part_n0[y] + part_cols[y] - 1 (cols <= BN, multiple of 32): unequal column parts let a launch whose
    // tile count does not fill the SMs split its N range into as many parts as there are idle SMs
    int m_tile0;
    short part_n0[16], part_cols[16];
    // b_mn (dgrad): the B operand is read from the forward filter w[oc][kh][kw][c] AS STORED -- rows = oc (the K dimension
    // of a dgrad), 32 contiguous c per box = an MN-major tile ({32 c, 32 oc} boxes, SWIZZLE_128B_BASE32B). The k-block of
    // job tap (kh', kw') starts at float column b_col0 - kh'*b_step_h - kw'*b_step_w (the flipped / sub-sampled tap of the
    // forward filter this parity class meets): no transposed copy of the filter is made.
    int b_mn, b_col0, b_step_h, b_step_w;
    // split-K (ksplit > 1, gridDim.z = ksplit): CTA z multiplies k-blocks [num_kb*z/ksplit, num_kb*(z+1)/ksplit) and ADDS its
    // partial tile into out (TMA reduce-add; the launcher zero-fills out first when the call is not accumulating; z = 0
    // carries the bias; no BatchNorm statistics). For launches of a handful of tiles with a long K (the 2x2 / 4x4 feature
    // maps of the 56 x 56 configuration: 16 CTAs x 144 k-blocks = 18 us of pure latency on 11 % of the SMs).
    int ksplit;
};

// Non-strided epilogue shared by the 1-CTA and 2-CTA im2col kernels. Every MMA has completed, so the operand ring is
// idle: its first 32 KB become eight SWIZZLE_128B staging blocks (TWO per epilogue warp); a warp's 32 rows x 32 columns
// leave as ONE TMA store / reduce-add of a {32 ch, 32 pixel} box of out[M, ldo] (rows >= M are clipped by the tensor
// map); optional BatchNorm partial statistics are gathered from the staged block. This epilogue is EXPOSED (one tile per
// CTA), so it is software-pipelined: the TMEM load of chunk i+1 is in flight while chunk i is staged, and chunk i is
// staged while the TMA store of chunk i-1 still reads the other staging block (one store may stay pending).
constexpr int EPI_STAGE_BYTES = 8 * 4096;     // staging blocks; the statistics scratch ([4][3][BN] floats) follows
template <int BN>
__device__ __forceinline__ void epilogue_chunk(const FpropParams& p, const float* __restrict__ bias, const CUtensorMap& tmO,
                                               uint8_t* sA, uint8_t* stage, uint32_t (&r)[32], int q, int lane, int m0, int n0,
                                               int c0) {
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
    if (bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j));
            o[j] += b.x; o[j + 1] += b.y; o[j + 2] += b.z; o[j + 3] += b.w;
        }
    }
    if (lane == 0) bulk_wait_read<1>();      // the store that last used THIS block (two chunks ago) has read it
    __syncwarp();
    stage_rows_sw128(stage, lane, o);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && m0 + q * 32 < p.M) {
        if (p.accumulate)
            tma_reduce_add_2d(&tmO, stage, n0 + c0, m0 + q * 32);
        else
            tma_store_2d(&tmO, stage, n0 + c0, m0 + q * 32);
    }
    if (lane == 0) bulk_commit();            // (an empty group keeps the wait_group<1> bookkeeping uniform)
    if (p.stats) {   // this warp's 32 rows of columns c0..c0+31 -> (n, mean, M2) per column
        const int left = p.M - (m0 + q * 32);
        const uint32_t vm = left >= 32 ? 0xFFFFFFFFu : (left > 0 ? ((1u << left) - 1u) : 0u);
        const Moments mo = stage_col_moments(stage, lane, vm);
        float* sStat = reinterpret_cast<float*>(sA + EPI_STAGE_BYTES) + q * 3 * BN + c0 + lane;
        sStat[0] = mo.n;
        sStat[BN] = mo.mean;
        sStat[2 * BN] = mo.m2;
    }
}

template <int BN>
__device__ __forceinline__ void epilogue_store_tma(const FpropParams& p, const float* __restrict__ bias,
                                                   const CUtensorMap* tmOp, uint32_t tmem_base, uint8_t* sA, int q, int lane,
                                                   int m_tile, int m0, int n0, int ncols) {
    const CUtensorMap& tmO = *tmOp;
    const int row = q * 32 + lane;
    uint8_t* stage0 = sA + q * 8192;
    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ra[32], rb[32];
    tmem_ld32(t0, ra);
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 64) {
        tmem_wait_ld();
        tmem_regs_ready(ra);
        const bool has_b = c0 + 32 < ncols;
        if (has_b) tmem_ld32(t0 + (uint32_t)(c0 + 32), rb);
        epilogue_chunk<BN>(p, bias, tmO, sA, stage0, ra, q, lane, m0, n0, c0);
        if (has_b) {
            tmem_wait_ld();
            tmem_regs_ready(rb);
            if (c0 + 64 < ncols) tmem_ld32(t0 + (uint32_t)(c0 + 64), ra);
            epilogue_chunk<BN>(p, bias, tmO, sA, stage0 + 4096, rb, q, lane, m0, n0, c0 + 32);
        }
    }
    if (lane == 0) bulk_wait_all<0>();
    if (p.stats) {   // merge the four warps, one partial per CTA and column
        epilogue_bar();
        const float* sStat = reinterpret_cast<const float*>(sA + EPI_STAGE_BYTES);
        for (int col = row; col < ncols; col += 128) {
            Moments acc{0.f, 0.f, 0.f};
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4)
                acc = merge_moments(acc, Moments{sStat[w4 * 3 * BN + col], sStat[(w4 * 3 + 1) * BN + col],
                                                 sStat[(w4 * 3 + 2) * BN + col]});
            float* dst = p.stats + (int64_t)m_tile * 3 * p.ldo + n0 + col;
            dst[0] = acc.n;
            dst[p.ldo] = acc.mean;
            dst[2 * p.ldo] = acc.m2;
        }
    }
}

template <int BN, int STAGES, int OCC>
__global__ void __launch_bounds__(kThreads, OCC)
conv_fprop_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const FpropParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int A_BYTES = BM * A_ROW_BYTES;
    constexpr int B_BYTES = BN * A_ROW_BYTES;
    static_assert(STAGES * A_BYTES >= EPI_STAGE_BYTES + 12 * BN * 4, "epilogue staging + statistics scratch live in the A ring");
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
    const int m_tile = p.m_tile0 + blockIdx.x;
    const int m0 = m_tile * BM, n0 = p.part_n0[blockIdx.y], ncols = p.part_cols[blockIdx.y];
    // this CTA's k-blocks (split-K: gridDim.z parts; else all of them)
    const int kb_begin = p.ksplit > 1 ? (int)((int64_t)num_kb * blockIdx.z / p.ksplit) : 0;
    const int kb_end = p.ksplit > 1 ? (int)((int64_t)num_kb * (blockIdx.z + 1) / p.ksplit) : num_kb;

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
    constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // first pixel of the tile -> window origin in the input image
        const int hw = p.OH * p.OW;
        const int n_img = m0 / hw, rem = m0 - n_img * hw;
        const int oh0 = rem / p.OW, ow0 = rem - oh0 * p.OW;
        const int w0 = ow0 * p.sw - p.pad_w, h0 = oh0 * p.sh - p.pad_h;
        int s = 0;
        uint32_t ph = 0;
        const int tap0 = kb_begin / cblocks;
        int cb = kb_begin - tap0 * cblocks, kh = tap0 / p.KW, kw = tap0 - (tap0 / p.KW) * p.KW;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    if (elect_one()) {
                        if (p.b_mn) {
                            const int col = p.b_col0 - kh * p.b_step_h - kw * p.b_step_w + n0;
                            mbar_expect_tx(&full[s], A_BYTES + (ncols >> 5) * 4096);
                            tma_load_im2col(sA + s * A_BYTES, &tmA, &full[s], cb * 32, w0, h0, n_img, (uint16_t)kw, (uint16_t)kh);
                            for (int b = 0; b < (ncols >> 5); ++b)
                                tma_load_2d(sB + s * B_BYTES + b * 4096, &tmB, &full[s], col + b * 32, cb * 32);
                        } else {
                            mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                            tma_load_im2col(sA + s * A_BYTES, &tmA, &full[s], cb * 32, w0, h0, n_img, (uint16_t)kw, (uint16_t)kh);
                            tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], kb * 32, n0);   // column (tap*C + cb*32) = kb*32
                        }
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                    if (++cb == cblocks) {        // next k-block: (kh, kw, cb) in filter order
                        cb = 0;
                        if (++kw == p.KW) { kw = 0; ++kh; }
                    }
                }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc(ncols, 0, p.b_mn ? 1 : 0);   // N = this CTA's column count (the B tile holds BN >= N rows)
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        const uint32_t b_kstep = p.b_mn ? 64u : 2u;     // K = 8 advance: 8 rows of 128 B (MN-major) / 32 B (K-major)
        int s = 0;
        uint32_t ph = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da = make_desc(a_base + s * A_BYTES, 16, 1024);
                const uint64_t db = p.b_mn ? make_desc(b_base + s * B_BYTES, 4096, 512, kLayoutSW128Base32B)
                                           : make_desc(b_base + s * B_BYTES, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 4 x (K = 8 tf32 = 32 B) per 128-byte row
                    umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * b_kstep), idesc,
                              ((kb - kb_begin) | k) != 0);
                umma_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tmem_full);
        __syncwarp();
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = q * 32 + lane;
        const int m = m0 + row;
        if (!p.o_strided) {
            epilogue_store_tma<BN>(p, blockIdx.z == 0 ? p.bias : nullptr, &tmO, tmem_base, sA, q, lane, m_tile, m0, n0, ncols);
        } else {
            const int hw = p.OH * p.OW;
            const int n_img = m / hw, rem = m - n_img * hw;
            const int i = rem / p.OW, j2 = rem - i * p.OW;
            const int64_t pix = ((int64_t)n_img * p.o_H + p.o_a + i * p.o_sh) * p.o_W + p.o_b + j2 * p.o_sw;
            float* orow = p.out + pix * p.ldo + n0;
#pragma unroll 1
            for (int c0 = 0; c0 < ncols; c0 += 32) {
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
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ fprop, CTA pairs
// Same implicit GEMM on a cluster of two CTAs (adjacent m-tiles): each CTA loads its own A tile and HALF of the B
// tile (rows [rank*BN/2, +BN/2) of the column part), the leader issues M = 256 MMAs, accumulator rows 0..127 land in
// the leader's TMEM and 128..255 in the peer's; each CTA runs its own epilogue. Per SM and k-block: 16 + BN/2*128 B of
// fills and 32 + BN/8 operand wavefronts per MMA instead of 16 KB + BN*128 B and 32 + BN/4.
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_fprop_tc2_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const FpropParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int A_BYTES = BM * A_ROW_BYTES;
    constexpr int BH_BYTES = (BN / 2) * A_ROW_BYTES;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * BH_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cblocks = p.C / 32;
    const int num_kb = p.KH * p.KW * cblocks;
    const int m_tile = p.m_tile0 + blockIdx.x;
    const int m0 = m_tile * BM, n0 = p.part_n0[blockIdx.y], ncols = p.part_cols[blockIdx.y];   // ncols: multiple of 32

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmO);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    if (warp == 1) {
        tmem_alloc2(tmem_slot, TMEM_COLS);
        tmem_relinquish2();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // both CTAs' barriers are initialised before any remote complete_tx / commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        const int hw = p.OH * p.OW;
        const int m0l = m0 < p.M ? m0 : 0;                // odd tile count: the last pair's second CTA loads tile 0 again
        const int n_img = m0l / hw, rem = m0l - n_img * hw;
        const int oh0 = rem / p.OW, ow0 = rem - oh0 * p.OW;
        const int w0 = ow0 * p.sw - p.pad_w, h0 = oh0 * p.sh - p.pad_h;
        const int nb = n0 + (int)rank * (ncols / 2);      // this CTA's half of the B rows
        int s = 0, kb = 0;
        uint32_t ph = 0;
        for (int kh = 0; kh < p.KH; ++kh)
            for (int kw = 0; kw < p.KW; ++kw)
                for (int cb = 0; cb < cblocks; ++cb, ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(&full[s], 2 * (A_BYTES + BH_BYTES));
                        tma2_load_im2col(sA + s * A_BYTES, &tmA, &full[s], cb * 32, w0, h0, n_img, (uint16_t)kw, (uint16_t)kh);
                        tma2_load_2d(sB + s * BH_BYTES, &tmB, &full[s], kb * 32, nb);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = make_idesc2(ncols);
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_desc(a_base + s * A_BYTES, 16, 1024);
                    const uint64_t db = make_desc(b_base + s * BH_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma2_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma2_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma2_commit(tmem_full);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (m0 < p.M) epilogue_store_tma<BN>(p, p.bias, &tmO, tmem_base, sA, q, lane, m_tile, m0, n0, ncols);
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();        // the peer's shared memory / TMEM stay alive until both CTAs are done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
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
        // running pixel coordinates of the k-block (no divisions in the loop)
        int p0 = kb_begin * BKP;
        int n_img = p0 / hw, rem = p0 - n_img * hw;
        int oh0 = rem / p.OW, ow0 = rem - oh0 * p.OW;
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < num_kb; ++i) {
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full[s], nvalid * BOX_BYTES + B_BYTES);
                const int w0 = ow0 * p.sw - p.pad, h0 = oh0 * p.sh - p.pad;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (b < nvalid)
                        tma_load_im2col(sA + s * A_BYTES + b * BOX_BYTES, &tmX, &full[s], box_c[b], w0, h0, n_img,
                                        (uint16_t)box_kw[b], (uint16_t)box_kh[b]);
#pragma unroll
                for (int b = 0; b < BN / 32; ++b)
                    tma_load_2d(sB + s * B_BYTES + b * BOX_BYTES, &tmDy, &full[s], n0 + b * 32, p0);
            }
            p0 += BKP;
            ow0 += BKP;
            while (ow0 >= p.OW) {
                ow0 -= p.OW;
                if (++oh0 == p.OH) { oh0 = 0; ++n_img; }
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, 1, 1);
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < num_kb; ++i) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da = make_desc(a_base + s * A_BYTES, BOX_BYTES, 512, kLayoutSW128Base32B);
                const uint64_t db = make_desc(b_base + s * B_BYTES, BOX_BYTES, 512, kLayoutSW128Base32B);
#pragma unroll
                for (int k = 0; k < BKP / 8; ++k)  // K = 8 pixels = two 4-row swizzle atoms = 1024 bytes
                    umma_tf32(tmem_base, da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), idesc, (i | k) != 0);
                umma_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tmem_full);
        __syncwarp();
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
// map with overlapping strides {ow: sw*16 B, ...} exposes exactly those rows: fprop runs the filter-resident halo
// kernel (conv_halo_tc.cuh) on {32 floats, ow, input row, n}, wgrad a TMA -> tcgen05 pipeline over
// {32 floats, ow, oh, kh, n} with pixels as the K dimension.
constexpr int STEM_TW = 16;                 // wgrad pixel block is 16 wide
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
        int n_img = kb_begin / blocks_per_img, rem = kb_begin - n_img * blocks_per_img;
        int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < num_kb; ++i) {
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full[s], nvalid * BOX_BYTES + B_BYTES);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (b < nvalid)
                        tma_load_5d(sA + s * A_BYTES + b * BOX_BYTES, &tmX, &full[s], 0, tw_i * STEM_TW, th_i * STEM_WTH,
                                    kh0 + b, n_img);
#pragma unroll
                for (int b = 0; b < BN / 32; ++b)
                    tma_load_4d(sB + s * B_BYTES + b * BOX_BYTES, &tmDy, &full[s], n0 + b * 32, tw_i * STEM_TW,
                                th_i * STEM_WTH, n_img);
            }
            if (++tw_i == p.tiles_w) {
                tw_i = 0;
                if (++th_i == p.tiles_h) { th_i = 0; ++n_img; }
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, 1, 1);
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < num_kb; ++i) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da = make_desc(a_base + s * A_BYTES, BOX_BYTES, 512, kLayoutSW128Base32B);
                const uint64_t db = make_desc(b_base + s * B_BYTES, BOX_BYTES, 512, kLayoutSW128Base32B);
#pragma unroll
                for (int k = 0; k < BKP / 8; ++k)
                    umma_tf32(tmem_base, da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), idesc, (i | k) != 0);
                umma_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tmem_full);
        __syncwarp();
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

// dgrad as forward convolutions of dY. For a forward conv of stride (sh, sw) the input pixels split into sh*sw
// parity classes (a, b) = (h % sh, w % sw); class (a, b) only ever meets the filter taps kh = r_a + sh*t,
// kw = r_b + sw*t' (r_a = (a + pad) % sh), so
//   dX[n, a + sh*i, b + sw*j, c] = sum_{ua, ub, oc} dY[n, i - pad_a + ua, j - pad_b + ub, oc] * wt_ab[c, ua, ub, oc]
// is a STRIDE-1 convolution of dY with the T_a x T_b sub-filter wt_ab[c, ua, ub, oc] = w[oc, r_a + sh*(T_a-1-ua),
// r_b + sw*(T_b-1-ub), c] (flipped and transposed). Stride 1 is the single class with the whole flipped filter.
// No zero-inserted copy of dY, no multiplications by inserted zeros.
struct TapPlace {
    int base[64];     // float offset of the class' sub-filter in the workspace
    int pos[64];      // tap position (ua*T_b + ub) inside the sub-filter
    int ntaps[64];    // T_a*T_b of the class
};
__global__ void gather_subfilters_k(const float* __restrict__ w, float* __restrict__ wt, int OC, int taps, int C,
                                    const TapPlace tp) {
    __shared__ float tile[32][33];
    const int tap = blockIdx.z;
    const int c0 = blockIdx.x * 32, oc0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int oc = oc0 + j, c = c0 + threadIdx.x;
        if (oc < OC && c < C) tile[j][threadIdx.x] = w[((int64_t)oc * taps + tap) * C + c];
    }
    __syncthreads();
    float* dst = wt + tp.base[tap];
    const int pos = tp.pos[tap], nt = tp.ntaps[tap];
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, oc = oc0 + threadIdx.x;
        if (oc < OC && c < C) dst[((int64_t)c * nt + pos) * OC + oc] = tile[threadIdx.x][j];
    }
}

// dst[n, a + sh*i, b + sw*j, :] = 0 for one parity class (classes no filter tap reaches)
__global__ void zero_class_k(float4* __restrict__ dx, int Ha, int Wb, int C4, int H, int W, int sh, int sw, int a, int b,
                             int64_t total) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C4);
    int64_t r = t / C4;
    const int j = (int)(r % Wb);
    r /= Wb;
    const int i = (int)(r % Ha);
    const int n = (int)(r / Ha);
    dx[(((int64_t)n * H + a + (int64_t)i * sh) * W + b + (int64_t)j * sw) * C4 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------ host side
// im2col view of a channels-last activation [N,H,W,Cx]: box {32 channels, pixels_per_col window origins};
// origins run over [lo_w, W - 1 + up_w] x [lo_h, H - 1 + up_h] in steps of the conv stride (taps are given per
// load as {kw, kh} offsets from the origin; whatever falls outside the tensor reads as zero).
static bool make_map_im2col(CUtensorMap* m, const float* base, int N, int H, int W, int Cx, int sh, int sw, int lo_h,
                            int lo_w, int up_h, int up_w, uint32_t pixels_per_col,
                            CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    if (lo_h < -128 || lo_w < -128 || up_h < -128 || up_w < -128 || lo_h > 127 || lo_w > 127 || up_h > 127 || up_w > 127)
        return false;
    cuuint64_t dims[4] = {(cuuint64_t)Cx, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cx * 4, (cuuint64_t)W * Cx * 4, (cuuint64_t)H * W * Cx * 4};
    int lower[2] = {lo_w, lo_h};
    int upper[2] = {up_w, up_h};
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
static int g_pair_mode = -1;     // -1: read XSB_TC_PAIR on first use
static int g_dgrad_mn = -1;      // -1: read XSB_TC_DGRAD_MN on first use

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
static int launch_fprop(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const FpropParams& p, int m_tiles,
                        int n_parts, cudaStream_t s) {
    constexpr int smem = STAGES * (BM * A_ROW_BYTES + BN * A_ROW_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(conv_fprop_tc_k<BN, STAGES, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(m_tiles, n_parts, p.ksplit > 1 ? p.ksplit : 1);
    conv_fprop_tc_k<BN, STAGES, OCC><<<grid, kThreads, smem, s>>>(a, b, o, p);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

template <int BN, int STAGES>
static int launch_fprop2(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const FpropParams& p, int m_tiles,
                         int n_parts, cudaStream_t s) {
    constexpr int smem = STAGES * (BM * A_ROW_BYTES + (BN / 2) * A_ROW_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(conv_fprop_tc2_k<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(2 * ((m_tiles + 1) / 2), n_parts);     // clusters of 2 along x
    conv_fprop_tc2_k<BN, STAGES><<<grid, kThreads, smem, s>>>(a, b, o, p);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// One implicit-GEMM launch: D[(n,i,j), OCx] (+)= sum_{kh,kw,c} in[n, i*sh - pad_h + kh, j*sw - pad_w + kw, c] *
// flt[ocx, kh, kw, c] (+ bias) over the OH x OW window origins of every image of channels-last `in` [N,H,W,Cx].
struct FpropJob {
    const float* in; int N, H, W, Cx;
    const float* flt; int OCx, KH, KW, sh, sw, pad_h, pad_w, OH, OW;
    const float* bias; float* out; int accumulate;
    int o_strided, o_H, o_W, o_sh, o_sw, o_a, o_b;
    float* stats;          // BatchNorm partials (null = off)
    int* n_partials;       // host: receives the number of partials written
    // b_mn: `flt` is the FORWARD filter [b_rows = oc, b_cols = KH*KW*C floats] read as an MN-major B operand (see FpropParams)
    int b_mn = 0, b_rows = 0, b_cols = 0, b_col0 = 0, b_step_h = 0, b_step_w = 0;
};

static int fprop_dispatch(const FpropJob& j, cudaStream_t s) {
    if (!load_driver_entry_points()) {
        xsb_set_error("cuTensorMapEncode* entry points unavailable");
        return XSB_ERR_CUDA;
    }
    int BN = (j.OCx % 256 == 0) ? 256 : (j.OCx % 128 == 0) ? 128 : 64;
    // window origins: first at -pad, last such that exactly OH x OW of them exist
    const int lo_h = -j.pad_h, lo_w = -j.pad_w;
    const int up_h = lo_h + (j.OH - 1) * j.sh - (j.H - 1), up_w = lo_w + (j.OW - 1) * j.sw - (j.W - 1);
    CUtensorMap ta, tb;
    const bool ok_b = j.b_mn ? make_map_2d(&tb, j.flt, (uint64_t)j.b_rows, (uint64_t)j.b_cols, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                             : make_map_2d(&tb, j.flt, (uint64_t)j.OCx, (uint64_t)j.KH * j.KW * j.Cx, (uint32_t)BN);
    if (!make_map_im2col(&ta, j.in, j.N, j.H, j.W, j.Cx, j.sh, j.sw, lo_h, lo_w, up_h, up_w, BM) || !ok_b) {
        xsb_set_error("cuTensorMapEncode failed (fprop N=%d H=%d W=%d C=%d OC=%d k=%dx%d s=%d p=%d,%d)", j.N, j.H, j.W,
                      j.Cx, j.OCx, j.KH, j.KW, j.sh, j.pad_h, j.pad_w);
        return XSB_ERR_CUDA;
    }
    FpropParams p{j.out, j.bias, j.N * j.OH * j.OW, j.OCx, j.Cx, j.KH, j.KW, j.OH, j.OW, j.sh, j.sw, j.pad_h, j.pad_w,
                  j.accumulate, j.o_strided, j.o_H, j.o_W, j.o_sh, j.o_sw, j.o_a, j.o_b, j.o_strided ? nullptr : j.stats};
    p.b_mn = j.b_mn; p.b_col0 = j.b_col0; p.b_step_h = j.b_step_h; p.b_step_w = j.b_step_w;
    const int m_tiles = (p.M + BM - 1) / BM;
    if (p.stats && j.n_partials) *j.n_partials = m_tiles;
    CUtensorMap to = ta;   // unused by strided-destination launches
    if (!j.o_strided && !make_map_2d(&to, j.out, (uint64_t)p.M, (uint64_t)j.OCx, 32)) {
        xsb_set_error("cuTensorMapEncode failed (fprop output M=%d OC=%d)", p.M, j.OCx);
        return XSB_ERR_CUDA;
    }
    // one launch: m-tiles [m_tile0, m_tile0 + count) x `parts` column ranges of OCx (multiples of 32, widths differ by <= 32)
    auto launch_parts = [&](int m_tile0, int count, int parts) -> int {
        const int w32 = j.OCx / 32, base = w32 / parts, extra = w32 % parts;
        int n0 = 0, widest = 0;
        for (int i = 0; i < parts; ++i) {
            const int cols = (base + (i < extra ? 1 : 0)) * 32;
            p.part_n0[i] = (short)n0;
            p.part_cols[i] = (short)cols;
            n0 += cols;
            if (cols > widest) widest = cols;
        }
        p.m_tile0 = m_tile0;
        const int bn = widest <= 64 ? 64 : widest <= 96 ? 96 : widest <= 128 ? 128 : widest <= 192 ? 192 : 256;
        CUtensorMap tbp = tb;     // (the MN-major filter map does not depend on the tile width)
        if (!j.b_mn && !make_map_2d(&tbp, j.flt, (uint64_t)j.OCx, (uint64_t)j.KH * j.KW * j.Cx, (uint32_t)bn)) {
            xsb_set_error("cuTensorMapEncode failed (fprop filter box %d)", bn);
            return XSB_ERR_CUDA;
        }
        if (bn == 256) return launch_fprop<256, 4, 1>(ta, tbp, to, p, count, parts, s);
        if (bn == 192) return launch_fprop<192, 5, 1>(ta, tbp, to, p, count, parts, s);
        if (bn == 128) return launch_fprop<128, 6, 1>(ta, tbp, to, p, count, parts, s);
        if (bn == 96) return launch_fprop<96, 6, 1>(ta, tbp, to, p, count, parts, s);
        return launch_fprop<64, 8, 1>(ta, tbp, to, p, count, parts, s);
    };
    // CTA-pair variant of launch_parts (cta_group::2): `count` m-tiles as ceil(count/2) clusters
    auto launch_parts2 = [&](int m_tile0, int count, int parts) -> int {
        const int w32 = j.OCx / 32, base = w32 / parts, extra = w32 % parts;
        int n0 = 0, widest = 0;
        for (int i = 0; i < parts; ++i) {
            const int cols = (base + (i < extra ? 1 : 0)) * 32;
            p.part_n0[i] = (short)n0;
            p.part_cols[i] = (short)cols;
            n0 += cols;
            if (cols > widest) widest = cols;
        }
        p.m_tile0 = m_tile0;
        const int bn = widest <= 64 ? 64 : widest <= 128 ? 128 : widest <= 192 ? 192 : 256;
        CUtensorMap tbp;
        if (!make_map_2d(&tbp, j.flt, (uint64_t)j.OCx, (uint64_t)j.KH * j.KW * j.Cx, (uint32_t)(bn / 2))) {
            xsb_set_error("cuTensorMapEncode failed (fprop filter box %d)", bn / 2);
            return XSB_ERR_CUDA;
        }
        if (bn == 256) return launch_fprop2<256, 6>(ta, tbp, to, p, count, parts, s);
        if (bn == 192) return launch_fprop2<192, 7>(ta, tbp, to, p, count, parts, s);
        if (bn == 128) return launch_fprop2<128, 8>(ta, tbp, to, p, count, parts, s);
        return launch_fprop2<64, 8>(ta, tbp, to, p, count, parts, s);
    };
    const int S = xsb_sm_count_cached();
    // CTA pairs (cta_group::2) halve the B fills and operand reads per SM; measured on R224 they tie the 1-CTA kernels on
    // layer3 (53 vs 54 us), and lose where pairs leave SMs idle (layer4: 50 clusters on 74 pair slots) or where the 1-CTA
    // kernel runs two CTAs per SM (layer2): opt-in (XSB_TC_PAIR=1 / xsb_tc_config(0, 1)), exercised by the parity suite.
    if (g_pair_mode < 0) {
        const char* e = getenv("XSB_TC_PAIR");
        g_pair_mode = e ? atoi(e) : 0;
    }
    if (g_pair_mode > 0 && !j.o_strided && !j.b_mn && BN >= 128 && m_tiles >= 2) {
        const int SP = S / 2, pairs = (m_tiles + 1) / 2, n_tiles = j.OCx / BN;
        const int64_t units = (int64_t)pairs * n_tiles;
        int max_parts = j.OCx / 32;
        if (max_parts > 16) max_parts = 16;
        if (units < SP) {
            int parts = SP / pairs;
            if (parts > max_parts) parts = max_parts;
            if (parts > n_tiles) return launch_parts2(0, m_tiles, parts);
        } else if (n_tiles == 1 && units % SP != 0 && (units % SP) * 10 < (int64_t)SP * 7) {
            const int main_tiles = (int)(units / SP) * SP * 2, tail = m_tiles - main_tiles;
            int parts = SP / ((tail + 1) / 2);
            if (parts > max_parts) parts = max_parts;
            if (parts > 1 && tail > 0) {
                const int rc = launch_parts2(0, main_tiles, 1);
                if (rc != XSB_OK) return rc;
                return launch_parts2(main_tiles, tail, parts);
            }
        }
        return launch_parts2(0, m_tiles, n_tiles);
    }
    static const int wave_aware = [] { const char* e = getenv("XSB_TC_WAVES"); return (e && e[0] == '0') ? 0 : 1; }();
    if (wave_aware && (BN == 256 || (int64_t)m_tiles * (j.OCx / BN) * 2 <= S)) {
        // One CTA per SM: tile counts that are not a multiple of the SM count leave SMs idle (98 tiles on 148 SMs =
        // 66 %, 196 tiles = two waves at 66 %). Fewer tiles than SMs: split every tile's columns into floor(S/tiles)
        // parts (any tile width when the launch would fill less than half of the SMs: the 14x14 / 7x7 maps of the 56x56
        // configuration). A short tail wave after full waves (OC = 256): the tail m-tiles are split the same way.
        const int n_tiles = j.OCx / BN;
        const int64_t units = (int64_t)m_tiles * n_tiles;
        if (units < S) {
            int parts = S / m_tiles;
            if (parts > j.OCx / 32) parts = j.OCx / 32;
            if (parts > 16) parts = 16;
            if (parts > n_tiles) {
                // still only a fraction of the SMs, each CTA with a long serial K loop: split K as well
                static const int splitk_on = [] { const char* e = getenv("XSB_TC_SPLITK"); return (e && e[0] == '0') ? 0 : 1; }();
                const int ctas = m_tiles * parts, num_kb = j.KH * j.KW * (j.Cx / 32);
                int ks = S / ctas;
                if (ks > num_kb / 4) ks = num_kb / 4;
                if (ks > 32) ks = 32;
                if (splitk_on && !j.o_strided && ctas * 2 <= S && num_kb >= 16 && ks >= 2) {
                    if (!j.accumulate) XSB_CUDA(cudaMemsetAsync(j.out, 0, (size_t)p.M * j.OCx * sizeof(float), s));
                    p.accumulate = 1;
                    p.ksplit = ks;
                    p.stats = nullptr;
                    if (j.n_partials) *j.n_partials = 0;
                }
                return launch_parts(0, m_tiles, parts);
            }
        } else if (BN == 256 && n_tiles == 1 && units % S != 0 && (units % S) * 10 < (int64_t)S * 7) {
            const int main_tiles = (int)(units / S) * S, tail = m_tiles - main_tiles;
            int parts = S / tail;
            if (parts > j.OCx / 32) parts = j.OCx / 32;
            if (parts > 16) parts = 16;
            if (parts > 1) {
                const int rc = launch_parts(0, main_tiles, 1);
                if (rc != XSB_OK) return rc;
                return launch_parts(main_tiles, tail, parts);
            }
        }
    }
    const int parts = j.OCx / BN;
    if (parts > 16) return XSB_ERR_UNSUPPORTED;
    for (int i = 0; i < parts; ++i) {
        p.part_n0[i] = (short)(i * BN);
        p.part_cols[i] = (short)BN;
    }
    p.m_tile0 = 0;
    if (tc_occupancy() == 1) {
        if (BN == 256) return launch_fprop<256, 4, 1>(ta, tb, to, p, m_tiles, parts, s);
        if (BN == 128) return launch_fprop<128, 6, 1>(ta, tb, to, p, m_tiles, parts, s);
        return launch_fprop<64, 8, 1>(ta, tb, to, p, m_tiles, parts, s);
    }
    // measured on B200 (tests/op_bench.py, R224): 2 CTAs/SM is +45 % for BN=64 and +19 % for BN=128, but -25 % for
    // BN=256 (only 2 stages fit) -> the widest tile keeps one CTA per SM with a 4-stage ring
    if (BN == 256) return launch_fprop<256, 4, 1>(ta, tb, to, p, m_tiles, parts, s);
    if (BN == 128) return launch_fprop<128, 3, 2>(ta, tb, to, p, m_tiles, parts, s);
    return launch_fprop<64, 4, 2>(ta, tb, to, p, m_tiles, parts, s);
}

// ---- persistent multi-job launches (conv_pers_tc.cuh)
// g_pers_mode: 0 = never, 1 = strided dgrads (all parity classes in one launch), 2 = every im2col forward / dgrad too
static int g_pers_mode = -1;     // -1: read XSB_TC_PERS on first use (default 1)
static int pers_mode() {
    if (g_pers_mode < 0) {
        const char* e = getenv("XSB_TC_PERS");
        g_pers_mode = e ? atoi(e) : 1;
    }
    return g_pers_mode;
}

struct PersJobDesc {      // host description of one job: stride-(sh,sw) conv of `in` with flt[OCx, KH*KW*Cx]
    const float* flt;
    int KH, KW, sh, sw, pad_h, pad_w, OH, OW, o_a, o_b;
};

// in [N,H,W,Cx] channels-last, every job writes OCx channels of `out`. Jobs must be ordered heaviest first.
static int pers_dispatch(const float* in, int N, int H, int W, int Cx, int OCx, const PersJobDesc* jd, int n_jobs,
                         const float* bias, float* out, int accumulate, int o_strided, int o_H, int o_W, int o_sh, int o_sw,
                         float* stats, int* n_partials, cudaStream_t s) {
    if (n_jobs < 1 || n_jobs > PERS_MAX_JOBS || Cx % 32 != 0 || OCx % 32 != 0) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) {
        xsb_set_error("cuTensorMapEncode* entry points unavailable");
        return XSB_ERR_CUDA;
    }
    const int S = xsb_sm_count_cached();
    PersParams p{};
    PersMaps maps;
    p.out = out; p.bias = bias; p.stats = nullptr; p.ldo = OCx; p.C = Cx; p.accumulate = accumulate;
    p.o_strided = o_strided; p.o_H = o_H; p.o_W = o_W; p.o_sh = o_sh; p.o_sw = o_sw; p.n_jobs = n_jobs;
    int64_t m_tiles_total = 0;
    for (int j = 0; j < n_jobs; ++j) m_tiles_total += ((int64_t)N * jd[j].OH * jd[j].OW + BM - 1) / BM;
    // column parts: <= 256 columns each; fewer tiles than SMs -> more (narrower) parts so that every SM has a tile
    const int w32 = OCx / 32;
    int parts = (OCx + 255) / 256;
    if (m_tiles_total * parts < S) {
        int want = (int)(S / m_tiles_total);
        if (want > w32) want = w32;
        if (want > 16) want = 16;
        if (want > parts) parts = want;
    }
    if (parts > 16) return XSB_ERR_UNSUPPORTED;
    int widest = 0, n0 = 0;
    for (int i = 0; i < parts; ++i) {
        const int cols = (w32 / parts + (i < w32 % parts ? 1 : 0)) * 32;
        p.part_n0[i] = (short)n0;
        p.part_cols[i] = (short)cols;
        n0 += cols;
        if (cols > widest) widest = cols;
    }
    p.n_parts = parts;
    const int bn = widest <= 64 ? 64 : widest <= 96 ? 96 : widest <= 128 ? 128 : widest <= 192 ? 192 : 256;
    int tile = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const PersJobDesc& d = jd[j];
        PersJob& J = p.job[j];
        J.M = N * d.OH * d.OW; J.OH = d.OH; J.OW = d.OW; J.KH = d.KH; J.KW = d.KW; J.pad_h = d.pad_h; J.pad_w = d.pad_w;
        J.sh = d.sh; J.sw = d.sw; J.tile_begin = tile; J.o_a = d.o_a; J.o_b = d.o_b;
        tile += ((J.M + BM - 1) / BM) * parts;
        const int lo_h = -d.pad_h, lo_w = -d.pad_w;
        const int up_h = lo_h + (d.OH - 1) * d.sh - (H - 1), up_w = lo_w + (d.OW - 1) * d.sw - (W - 1);
        if (d.KH > 255 || d.KW > 255 ||
            !make_map_im2col(&maps.a[j], in, N, H, W, Cx, d.sh, d.sw, lo_h, lo_w, up_h, up_w, BM) ||
            !make_map_2d(&maps.b[j], d.flt, (uint64_t)OCx, (uint64_t)d.KH * d.KW * Cx, (uint32_t)bn)) {
            xsb_set_error("cuTensorMapEncode failed (persistent conv job %d: N=%d H=%d W=%d C=%d OC=%d k=%dx%d)", j, N, H, W, Cx,
                          OCx, d.KH, d.KW);
            return XSB_ERR_CUDA;
        }
    }
    p.total_tiles = tile;
    maps.o = maps.a[0];      // unused by strided-destination launches
    if (!o_strided) {
        if (n_jobs != 1) return XSB_ERR_UNSUPPORTED;
        if (!make_map_2d(&maps.o, out, (uint64_t)p.job[0].M, (uint64_t)OCx, 32)) {
            xsb_set_error("cuTensorMapEncode failed (persistent conv output M=%d OC=%d)", p.job[0].M, OCx);
            return XSB_ERR_CUDA;
        }
        if (stats) {
            p.stats = stats;
            if (n_partials) *n_partials = (p.job[0].M + BM - 1) / BM;
        }
    }
    const int grid = tile < S ? tile : S;
    if (bn == 256) return launch_pers<256, 1, 4>(maps, p, grid, s);
    if (bn == 192) return launch_pers<192, 1, 4>(maps, p, grid, s);
    if (bn == 128) return launch_pers<128, 2, 3>(maps, p, grid, s);
    if (bn == 96) return launch_pers<96, 2, 3>(maps, p, grid, s);
    return launch_pers<64, 2, 4>(maps, p, grid, s);
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

// The padded image lives at the TAIL of the workspace (every other user -- dgrad sub-filters, split-K partials -- grows
// from its head), so the copy made by the stem's forward is still there when its wgrad runs later in the same step:
// the tag remembers which input it holds. A forward always re-pads (the batch behind a pointer changes between steps).
struct StemXpTag {
    const float* x; const float* ws; int N, H, W, C, pad, Hp, Wp;
};
static StemXpTag g_xp_tag = {nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0};
static float* stem_xp_ptr(float* ws, int64_t ws_floats, const StemPlan& pl) {
    return ws + ((ws_floats - pl.xp_floats) & ~(int64_t)255);
}
static bool stem_xp_valid(const float* x, const float* ws, const ConvGeom& g, const StemPlan& pl) {
    const StemXpTag& t = g_xp_tag;
    return t.x == x && t.ws == ws && t.N == g.N && t.H == g.H && t.W == g.W && t.C == g.C && t.pad == g.pad &&
           t.Hp == pl.Hp && t.Wp == pl.Wp;
}

static int stem_prepare(const float* x, const ConvGeom& g, const StemPlan& pl, float* xp, cudaStream_t s) {
    const int64_t total = (int64_t)g.N * pl.Hp * pl.Wp;
    stem_pad_c4_k<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(x, reinterpret_cast<float4*>(xp), g.N, g.H, g.W, g.C,
                                                                  pl.Hp, pl.Wp, g.pad);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

static int stem_fwd(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g, float* ws,
                    int64_t ws_floats, cudaStream_t s, StatsReq* st, int accumulate) {
    const StemPlan pl = stem_plan(g);
    if (!ws || ws_floats < pl.xp_floats + pl.wq_floats) return XSB_ERR_UNSUPPORTED;
    if ((g.OC != 64 && g.OC != 128) || g.KH > HALO_MAX_KH) return XSB_ERR_UNSUPPORTED;
    const int rows = (HALO_TH - 1) * g.sh + g.KH;
    const int stages = halo_stages(g.KH, g.OC, rows * 1024);
    if (stages == 0 || rows > 256) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    if (ws_floats < 2 * pl.xp_floats + pl.wq_floats + 512) return XSB_ERR_UNSUPPORTED;
    float* xp = stem_xp_ptr(ws, ws_floats, pl);
    float* wq = ws;
    CUtensorMap ta, tb;
    // rows of 32 floats starting every sw pixels: {32, OW, Hp, N} with OVERLAPPING dim-1 stride
    cuuint64_t dims[4] = {32, (cuuint64_t)g.OW, (cuuint64_t)pl.Hp, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.sw * 16, (cuuint64_t)pl.Wp * 16, (cuuint64_t)pl.Hp * pl.Wp * 16};
    cuuint32_t box[4] = {32, HALO_TW, (cuuint32_t)rows, 1};
    if (!make_map_tiled(&ta, xp, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_map_2d(&tb, wq, (uint64_t)g.OC, (uint64_t)g.KH * 32, (uint32_t)g.OC))
        return XSB_ERR_UNSUPPORTED;   // driver rejected the overlapping-stride map: FFMA kernel takes the call
    int rc = stem_prepare(x, g, pl, xp, s);
    if (rc != XSB_OK) return rc;
    g_xp_tag = StemXpTag{x, ws, g.N, g.H, g.W, g.C, g.pad, pl.Hp, pl.Wp};
    stem_pack_filter_k<<<(g.OC * g.KH * 32 + 255) / 256, 256, 0, s>>>(w, wq, g.OC, g.KH, g.KW, g.C);
    XSB_LAUNCH_CHECK();
    HaloParams p{};
    p.out = y; p.bias = bias; p.N = g.N; p.OH = g.OH; p.OW = g.OW; p.ldo = g.OC;
    p.tiles_w = (g.OW + HALO_TW - 1) / HALO_TW; p.tiles_h = (g.OH + HALO_TH - 1) / HALO_TH;
    p.n_tiles = g.N * p.tiles_w * p.tiles_h;
    p.cblocks = 1; p.KWl = 1; p.KH = g.KH; p.sh = g.sh; p.pad_w = 0; p.bt_kh_stride = 1; p.h_off = 0;
    p.stage_bytes = rows * 1024; p.tx_bytes = rows * 1024; p.n_stages = stages; p.n_btiles = g.KH; p.accumulate = accumulate;
    halo_attach_stats(p, st);
    return g.OC == 64 ? launch_halo<64>(ta, tb, p, s) : launch_halo<128>(ta, tb, p, s);
}

static int stem_wgrad(const float* dy, const float* x, float* dw, const ConvGeom& g, int accumulate, float* ws,
                      int64_t ws_floats, cudaStream_t s) {
    const StemPlan pl = stem_plan(g);
    if (!ws || ws_floats < 2 * pl.xp_floats + 512) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    float* xp = stem_xp_ptr(ws, ws_floats, pl);
    CUtensorMap tx, tdy;
    cuuint64_t ddims[4] = {(cuuint64_t)g.OC, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
    cuuint64_t dstr[3] = {(cuuint64_t)g.OC * 4, (cuuint64_t)g.OW * g.OC * 4, (cuuint64_t)g.OH * g.OW * g.OC * 4};
    cuuint32_t dbox[4] = {32, STEM_TW, STEM_WTH, 1};
    if (!stem_map_x(&tx, xp, g, pl, STEM_TW, STEM_WTH, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map_tiled(&tdy, dy, 4, ddims, dstr, dbox, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return XSB_ERR_UNSUPPORTED;
    int rc = XSB_OK;
    if (stem_xp_valid(x, ws, g, pl)) {
        g_xp_tag.x = nullptr;            // one forward, one wgrad: the next wgrad on this pointer pads again
    } else {
        rc = stem_prepare(x, g, pl, xp, s);
        if (rc != XSB_OK) return rc;
    }
    static const int stem_halo = [] { const char* e = getenv("XSB_TC_WGRAD_HALO"); return (e && e[0] == '0') ? 0 : 1; }();
    if (stem_halo) {   // persistent halo kernel: every filter row from one load, dY read once
        if (!accumulate) XSB_CUDA(cudaMemsetAsync(dw, 0, (size_t)g.OC * g.KH * g.KW * g.C * sizeof(float), s));
        rc = wgrad_halo_stem(dy, xp, dw, g, pl.Hp, pl.Wp, s);
        if (rc != XSB_ERR_UNSUPPORTED) return rc;
    }
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

// Dense products (gemm_tc.cuh). XSB_TC_GEMM=0 (env, read once) keeps them on the FFMA kernel for A/B measurements.
int xsb_tc_gemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int transA, int transB,
                int accumulate, float*, int64_t, cudaStream_t s, float* act) {
    static const int on = [] { const char* e = getenv("XSB_TC_GEMM"); return (e && e[0] == '0') ? 0 : 1; }();
    if (!on) return XSB_ERR_UNSUPPORTED;
    const int rc = gemm_tf32(A, B, C, bias, M, N, K, transA, transB, accumulate, s, act);
    if (rc == XSB_OK) ++g_tc_launches;
    return rc;
}

// XSB_TC_HALO=0 (env, read once) routes everything through the im2col kernel (A/B measurements)
static bool halo_enabled() {
    static const int on = [] { const char* e = getenv("XSB_TC_HALO"); return (e && e[0] == '0') ? 0 : 1; }();
    return on != 0;
}

// g_tc_launches counts ENTRY-POINT calls served by tensor-core kernels (a strided dgrad is several launches)
static int counted(int rc) {
    if (rc == XSB_OK) ++g_tc_launches;
    return rc;
}

static int tc_conv_fwd(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g, float* ws,
                       int64_t ws_floats, cudaStream_t s, StatsReq* st, int accumulate) {
    if (stem_ok(g) && aligned16(x) && aligned16(w) && aligned16(y))
        return stem_fwd(x, w, bias, y, g, ws, ws_floats, s, st, accumulate);
    if (!fprop_ok(g.C, g.OC, x, w, y) || g.KW > 255 || g.KH > 255) return XSB_ERR_UNSUPPORTED;
    if (g.sh == 1 && g.sw == 1 && halo_enabled()) {
        const int rc = halo_conv_s1(x, g.N, g.H, g.W, g.C, w, g.OC, g.KH, g.KW, g.pad, g.pad, g.OH, g.OW, bias, y, accumulate, s, st);
        if (rc != XSB_ERR_UNSUPPORTED) return rc;
    }
    const int64_t m_tiles = ((int64_t)g.N * g.OH * g.OW + BM - 1) / BM;
    if (pers_mode() >= 2) {
        const PersJobDesc jd{w, g.KH, g.KW, g.sh, g.sw, g.pad, g.pad, g.OH, g.OW, 0, 0};
        const bool want_stats = st && st->buf && st->floats >= m_tiles * 3 * g.OC;
        const int rc = pers_dispatch(x, g.N, g.H, g.W, g.C, g.OC, &jd, 1, bias, y, accumulate, 0, 0, 0, 1, 1,
                                     want_stats ? st->buf : nullptr, want_stats ? st->n_partials : nullptr, s);
        if (rc != XSB_ERR_UNSUPPORTED) return rc;
    }
    FpropJob j{x, g.N, g.H, g.W, g.C, w, g.OC, g.KH, g.KW, g.sh, g.sw, g.pad, g.pad, g.OH, g.OW, bias, y, accumulate,
               0, 0, 0, 1, 1, 0, 0, nullptr, nullptr};
    if (st && st->buf && st->floats >= m_tiles * 3 * g.OC) {
        j.stats = st->buf;
        j.n_partials = st->n_partials;
    }
    return fprop_dispatch(j, s);
}

// accumulate = 1: y += conv(x, w) (+ bias) -- the later passes of the 3xTF32 split (XSB_MATH_FP32X3); statistics describe
// the pass's own contribution, so callers ask for them only with accumulate = 0
int xsb_tc_conv_fwd(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g, float* ws,
                    int64_t ws_floats, cudaStream_t s, float* stats, int64_t stats_floats, int* n_partials, int accumulate) {
    StatsReq st{stats, stats_floats, n_partials};
    return counted(tc_conv_fwd(x, w, bias, y, g, ws, ws_floats, s, (stats && !accumulate) ? &st : nullptr, accumulate));
}

static int tc_conv_dgrad(const float* dy, const float* w, float* dx, const ConvGeom& g, int accumulate, float* ws,
                         int64_t ws_floats, cudaStream_t s) {
    // see gather_subfilters_k: one stride-1 convolution of dY per parity class of dX
    if (!fprop_ok(g.OC, g.C, dy, w, dx)) return XSB_ERR_UNSUPPORTED;
    const int taps = g.KH * g.KW;
    const int64_t wn = (int64_t)g.OC * taps * g.C;
    if (taps > 64 || g.sh > 8 || g.sw > 8 || wn >= ((int64_t)1 << 31) || !ws || ws_floats < wn) return XSB_ERR_UNSUPPORTED;
    struct Cls { int r, T, padp, n; };   // first tap, #taps, padding of the equivalent conv, #rows (cols) of dX
    Cls ch[8], cw[8];
    for (int a = 0; a < g.sh; ++a) {
        const int r = (a + g.pad) % g.sh, T = g.KH > r ? (g.KH - r + g.sh - 1) / g.sh : 0;
        ch[a] = {r, T, T - 1 - (a + g.pad - r) / g.sh, g.H > a ? (g.H - a + g.sh - 1) / g.sh : 0};
    }
    for (int b = 0; b < g.sw; ++b) {
        const int r = (b + g.pad) % g.sw, T = g.KW > r ? (g.KW - r + g.sw - 1) / g.sw : 0;
        cw[b] = {r, T, T - 1 - (b + g.pad - r) / g.sw, g.W > b ? (g.W - b + g.sw - 1) / g.sw : 0};
    }
    // sub-filters packed back to back: class (a,b) at cls_base[a][b], [C, T_a*T_b, OC] each (offsets stay 16-byte
    // aligned because C*OC is a multiple of 32*64)
    int64_t cls_base[8][8];
    int64_t off = 0;
    for (int a = 0; a < g.sh; ++a)
        for (int b = 0; b < g.sw; ++b) {
            cls_base[a][b] = off;
            off += (int64_t)g.C * ch[a].T * cw[b].T * g.OC;
        }
    TapPlace tp;
    for (int kh = 0; kh < g.KH; ++kh)
        for (int kw = 0; kw < g.KW; ++kw) {
            const int ra = kh % g.sh, rb = kw % g.sw;
            const int a = ((ra - g.pad) % g.sh + g.sh) % g.sh, b = ((rb - g.pad) % g.sw + g.sw) % g.sw;
            const int ua = ch[a].T - 1 - kh / g.sh, ub = cw[b].T - 1 - kw / g.sw;
            const int t = kh * g.KW + kw;
            tp.base[t] = (int)cls_base[a][b];
            tp.pos[t] = ua * cw[b].T + ub;
            tp.ntaps[t] = ch[a].T * cw[b].T;
        }
    const int strided = (g.sh != 1 || g.sw != 1) ? 1 : 0;
    // The im2col kernel can read the forward filter as stored (MN-major B operand, FpropParams::b_mn) instead of the
    // flipped / transposed copy gather_subfilters_k makes. Measured on R224 (same box, alternating runs): 5.912 ms/step
    // against 5.882 with the copy -- N/32 boxes of 4 KB per k-block cost more TMA issue slots than 10 gathers of ~5 us
    // save -- so it is opt-in (XSB_TC_DGRAD_MN=1 / xsb_tc_config(1, 1)), parity-tested, and the copy stays the default.
    if (g_dgrad_mn < 0) {
        const char* e = getenv("XSB_TC_DGRAD_MN");
        g_dgrad_mn = (e && e[0] == '1') ? 1 : 0;
    }
    const bool halo_first = !strided && halo_enabled() && halo_s1_accepts(g.OC, g.C, g.KH, g.KW, g.H, g.W);
    const bool gathered = halo_first || !g_dgrad_mn;
    if (gathered) {
        dim3 grid((g.C + 31) / 32, (g.OC + 31) / 32, taps), block(32, 8);
        gather_subfilters_k<<<grid, block, 0, s>>>(w, ws, g.OC, taps, g.C, tp);
        XSB_LAUNCH_CHECK();
    }
    // all parity classes (or the single stride-1 class the halo kernel does not take) in ONE persistent launch
    if (gathered && !halo_first && (strided ? pers_mode() >= 1 : pers_mode() >= 2)) {
        PersJobDesc jd[PERS_MAX_JOBS];
        int nj = 0;
        bool fits = true;
        for (int a = 0; a < g.sh && fits; ++a)
            for (int b = 0; b < g.sw; ++b) {
                if (ch[a].n == 0 || cw[b].n == 0 || ch[a].T == 0 || cw[b].T == 0) continue;
                if (nj == PERS_MAX_JOBS) { fits = false; break; }
                jd[nj++] = PersJobDesc{ws + cls_base[a][b], ch[a].T, cw[b].T, 1, 1, ch[a].padp, cw[b].padp, ch[a].n, cw[b].n, a, b};
            }
        if (fits && nj > 0) {
            for (int i = 1; i < nj; ++i)           // heaviest job first (insertion sort by tap count)
                for (int k = i; k > 0 && jd[k].KH * jd[k].KW > jd[k - 1].KH * jd[k - 1].KW; --k) {
                    const PersJobDesc t = jd[k]; jd[k] = jd[k - 1]; jd[k - 1] = t;
                }
            const int rc = pers_dispatch(dy, g.N, g.OH, g.OW, g.OC, g.C, jd, nj, nullptr, dx, accumulate, strided, g.H, g.W,
                                         g.sh, g.sw, nullptr, nullptr, s);
            if (rc != XSB_ERR_UNSUPPORTED) {
                if (rc != XSB_OK) return rc;
                if (!accumulate)                   // classes no filter tap reaches: their gradient is zero
                    for (int a = 0; a < g.sh; ++a)
                        for (int b = 0; b < g.sw; ++b) {
                            const int Ha = ch[a].n, Wb = cw[b].n;
                            if (Ha == 0 || Wb == 0 || (ch[a].T != 0 && cw[b].T != 0)) continue;
                            const int64_t total = (int64_t)g.N * Ha * Wb * (g.C / 4);
                            zero_class_k<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(reinterpret_cast<float4*>(dx), Ha, Wb,
                                                                                         g.C / 4, g.H, g.W, g.sh, g.sw, a, b, total);
                            XSB_LAUNCH_CHECK();
                        }
                return XSB_OK;
            }
        }
    }
    for (int a = 0; a < g.sh; ++a)
        for (int b = 0; b < g.sw; ++b) {
            const int Ha = ch[a].n, Wb = cw[b].n;
            if (Ha == 0 || Wb == 0) continue;
            if (ch[a].T == 0 || cw[b].T == 0) {      // no filter tap reaches this class: its gradient is zero
                if (!accumulate) {
                    const int64_t total = (int64_t)g.N * Ha * Wb * (g.C / 4);
                    zero_class_k<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(reinterpret_cast<float4*>(dx), Ha, Wb,
                                                                                 g.C / 4, g.H, g.W, g.sh, g.sw, a, b, total);
                    XSB_LAUNCH_CHECK();
                }
                continue;
            }
            if (halo_first) {
                const int rc = halo_conv_s1(dy, g.N, g.OH, g.OW, g.OC, ws, g.C, g.KH, g.KW, ch[0].padp, cw[0].padp, g.H, g.W,
                                            nullptr, dx, accumulate, s);
                if (rc != XSB_ERR_UNSUPPORTED) return rc;
            }
            FpropJob j{dy, g.N, g.OH, g.OW, g.OC, ws + cls_base[a][b], g.C, ch[a].T, cw[b].T, 1, 1, ch[a].padp,
                       cw[b].padp, Ha, Wb, nullptr, dx, accumulate, strided, g.H, g.W, g.sh, g.sw, a, b, nullptr, nullptr};
            if (!gathered) {
                // job tap (ua, ub) of this class = forward tap (r_a + sh*(T_a-1-ua), r_b + sw*(T_b-1-ub))
                j.flt = w;
                j.b_mn = 1;
                j.b_rows = g.OC;
                j.b_cols = taps * g.C;
                j.b_col0 = ((ch[a].r + g.sh * (ch[a].T - 1)) * g.KW + (cw[b].r + g.sw * (cw[b].T - 1))) * g.C;
                j.b_step_h = g.sh * g.KW * g.C;
                j.b_step_w = g.sw * g.C;
            }
            const int rc = fprop_dispatch(j, s);
            if (rc != XSB_OK) return rc;
        }
    return XSB_OK;
}

int xsb_tc_conv_dgrad(const float* dy, const float* w, float* dx, const ConvGeom& g, int accumulate, float* ws,
                      int64_t ws_floats, cudaStream_t s) {
    return counted(tc_conv_dgrad(dy, w, dx, g, accumulate, ws, ws_floats, s));
}
static int tc_conv_wgrad(const float* dy, const float* x, float* dw, const ConvGeom& g, int accumulate, float* ws,
                         int64_t ws_floats, cudaStream_t s) {
    if (stem_ok(g) && aligned16(dy) && aligned16(x)) return stem_wgrad(dy, x, dw, g, accumulate, ws, ws_floats, s);
    if (g.C % 32 != 0 || g.OC % 64 != 0 || !aligned16(dy) || !aligned16(x) || !aligned16(dw)) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) {
        xsb_set_error("cuTensorMapEncode* entry points unavailable");
        return XSB_ERR_CUDA;
    }
    static const int wg_halo = [] { const char* e = getenv("XSB_TC_WGRAD_HALO"); return (e && e[0] == '0') ? 0 : 1; }();
    if (wg_halo && halo_enabled()) {
        if (!accumulate) XSB_CUDA(cudaMemsetAsync(dw, 0, (size_t)g.OC * g.KH * g.KW * g.C * sizeof(float), s));
        int rc = wgrad_halo(dy, x, dw, g, s);
        if (rc != XSB_ERR_UNSUPPORTED) return rc;
        static const int wg_strip = [] { const char* e = getenv("XSB_TC_WGRAD_STRIP"); return (e && e[0] == '0') ? 0 : 1; }();
        if (wg_strip) {
            rc = wgrad_strip(dy, x, dw, g, s);
            if (rc != XSB_ERR_UNSUPPORTED) return rc;
        }
    }
    const int P = g.N * g.OH * g.OW;
    static const int wg_bn_cap = [] { const char* e = getenv("XSB_TC_WG_BN"); return e ? atoi(e) : 128; }();   // measured: 128-wide tiles + 64-pixel k-blocks beat 256 x 32 (layer4 93 -> 73 us)
    int BN = (g.OC % 256 == 0) ? 256 : (g.OC % 128 == 0) ? 128 : 64;
    if (BN > wg_bn_cap) BN = wg_bn_cap;
    const int BKP = BN == 256 ? 32 : 64;
    CUtensorMap tx, tdy;
    if (!make_map_im2col(&tx, x, g.N, g.H, g.W, g.C, g.sh, g.sw, -g.pad, -g.pad, g.pad - (g.KH - 1), g.pad - (g.KW - 1),
                         (uint32_t)BKP, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
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

int xsb_tc_conv_wgrad(const float* dy, const float* x, float* dw, const ConvGeom& g, int accumulate, float* ws,
                      int64_t ws_floats, cudaStream_t s) {
    return counted(tc_conv_wgrad(dy, x, dw, g, accumulate, ws, ws_floats, s));
}

extern "C" {
// kernel-selection switches (tests / A-B measurements): key 0 = CTA-pair (cta_group::2) forward kernels on/off,
// key 1 = dgrad reads the forward filter as an MN-major operand (no transposed copy) on/off
int xsb_tc_config(int key, int value) {
    if (key == 0) {
        tc::g_pair_mode = value ? 1 : 0;
        return XSB_OK;
    }
    if (key == 1) {
        tc::g_dgrad_mn = value ? 1 : 0;
        return XSB_OK;
    }
    if (key == 2) {
        tc::g_pers_mode = value;
        return XSB_OK;
    }
    xsb_set_error("xsb_tc_config: unknown key %d", key);
    return XSB_ERR_ARG;
}
// counters: [0] GEMM-shaped calls served by tensor-core kernels, [1] GEMM-shaped calls that ran the fp32 FFMA kernel in TF32 mode
int xsb_tc_stats(uint64_t* out2) {
    out2[0] = tc::g_tc_launches;
    out2[1] = tc::g_ffma_fallbacks;
    return XSB_OK;
}
}
void xsb_tc_note_ffma() { ++tc::g_ffma_fallbacks; }
