// Persistent, multi-job tcgen05 implicit-GEMM convolution (XSB_MATH_TF32).
//
// conv_fprop_tc_k (conv_tc.cu) is one CTA per 128 x N tile: every tile pays TMEM allocation, barrier set-up, the
// pipeline ramp and an epilogue nothing overlaps with, and a strided dgrad needs one such launch per parity class of dX
// (tiles with as few as 4 k-blocks: 141 TFLOP/s on the 3x3/2 layers of ResNet-18, 27 on the 1x1/2 shortcuts). Here
//  * ONE launch covers up to PERS_MAX_JOBS implicit GEMMs ("jobs": the parity classes of a strided dgrad, or a single
//    forward convolution) that share the input tensor, the output tensor and the channel counts; each job has its own
//    im2col tensor map (window origins / padding), filter matrix and tap count; tiles are numbered job-major with the
//    job of most taps first, so the tail of the launch is made of the short tiles;
//  * CTAs are persistent (grid = #SMs): tile t = blockIdx.x + i*gridDim.x. The TMA producer runs ahead across tile
//    boundaries (the operand ring never drains), accumulators are double-buffered in TMEM so the 4 epilogue warps
//    drain tile i while tile i+1 is multiplied;
//  * a pipeline stage holds KBS k-blocks (KBS*4 MMAs per mbarrier wait of the issuing thread, see tc_common.cuh);
//  * epilogues as in conv_tc.cu: TMEM -> registers (+bias) -> swizzled staging -> TMA store / reduce-add (+ BatchNorm
//    partial statistics per m-tile), or float4 stores to a strided destination (parity class of dX).
// Reference arithmetic: xs/nn/functional.py:405-442 (forward), xs/utils/common.py:49-65 + xs/nn/grad_fn.py:199-202 (dgrad).
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int PERS_MAX_JOBS = 9;

struct PersJob {
    int M;                 // rows of this job's GEMM = N * OH * OW window origins
    int OH, OW;            // window-origin grid per image
    int KH, KW;            // taps
    int pad_h, pad_w;      // origin (oh, ow) reads input rows oh*sh - pad_h + kh
    int sh, sw;
    int tile_begin;        // first global tile of this job
    int o_a, o_b;          // strided destination: row (n,i,j) -> pixel (n, o_a + i*o_sh, o_b + j*o_sw)
};

struct PersParams {
    float* out;            // [*, ldo] channels-last
    const float* bias;     // [ldo] or null
    float* stats;          // BatchNorm partials per m-tile (single-job, non-strided launches), or null
    int ldo, C;            // output channels (row pitch of out), input channels (K = taps * C)
    int accumulate;
    int o_strided, o_H, o_W, o_sh, o_sw;
    int n_jobs, total_tiles, n_parts;
    short part_n0[16], part_cols[16];
    PersJob job[PERS_MAX_JOBS];
};

struct PersMaps {
    CUtensorMap a[PERS_MAX_JOBS];   // im2col views of the input, one per job
    CUtensorMap b[PERS_MAX_JOBS];   // filter matrices [ldo rows, taps*C columns], box {32, BN}
    CUtensorMap o;                  // out[M, ldo], box {32, 32} (non-strided launches)
};

struct PersTile {
    int job, m_tile, n0, ncols;
};
__device__ __forceinline__ PersTile pers_decode(const PersParams& p, int t) {
    int j = 0;
    while (j + 1 < p.n_jobs && t >= p.job[j + 1].tile_begin) ++j;
    const int r = t - p.job[j].tile_begin;
    const int m_tile = r / p.n_parts, part = r - m_tile * p.n_parts;   // the parts of one m-tile run side by side: A hits L2
    return PersTile{j, m_tile, p.part_n0[part], p.part_cols[part]};
}

template <int BN, int KBS, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
conv_pers_tc_k(const __grid_constant__ PersMaps maps, const __grid_constant__ PersParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int A_BYTES = BM * A_ROW_BYTES;
    constexpr int B_BYTES = BN * A_ROW_BYTES;
    constexpr int BUF = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;       // TMEM columns per accumulator buffer
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sB = sA + STAGES * KBS * A_BYTES;
    uint8_t* sOut = sB + STAGES * KBS * B_BYTES;                      // 4 x 4 KB SWIZZLE_128B staging blocks
    float* sStat = reinterpret_cast<float*>(sOut + 16384);            // [4 warps][3][BN]
    uint64_t* full = reinterpret_cast<uint64_t*>(sStat + 12 * BN);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cblocks = p.C / 32;

    if (warp == 0 && lane == 0) {
        for (int j = 0; j < p.n_jobs; ++j) {
            prefetch_tmap(&maps.a[j]);
            prefetch_tmap(&maps.b[j]);
        }
        prefetch_tmap(&maps.o);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);      // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 2 * BUF);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const PersTile T = pers_decode(p, t);
            const PersJob& J = p.job[T.job];
            const CUtensorMap* ta = &maps.a[T.job];
            const CUtensorMap* tb = &maps.b[T.job];
            const int m0 = T.m_tile * BM, hw = J.OH * J.OW;
            const int n_img = m0 / hw, rem = m0 - n_img * hw;
            const int oh0 = rem / J.OW, ow0 = rem - oh0 * J.OW;
            const int w0 = ow0 * J.sw - J.pad_w, h0 = oh0 * J.sh - J.pad_h;
            const int num_kb = J.KH * J.KW * cblocks;
            for (int kb = 0; kb < num_kb; kb += KBS) {
                const int nk = num_kb - kb < KBS ? num_kb - kb : KBS;
                mbar_wait(&empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[s], (uint32_t)nk * (A_BYTES + B_BYTES));
#pragma unroll
                    for (int i = 0; i < KBS; ++i)
                        if (i < nk) {
                            const int k = kb + i;
                            const int tap = k / cblocks, cb = k - tap * cblocks;
                            const int kh = tap / J.KW, kw = tap - kh * J.KW;
                            tma_load_im2col(sA + (s * KBS + i) * A_BYTES, ta, &full[s], cb * 32, w0, h0, n_img, (uint16_t)kw,
                                            (uint16_t)kh);
                            tma_load_2d(sB + (s * KBS + i) * B_BYTES, tb, &full[s], k * 32, T.n0);
                        }
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        int s = 0, buf = 0;
        uint32_t ph = 0, tph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const PersTile T = pers_decode(p, t);
            const PersJob& J = p.job[T.job];
            const int num_kb = J.KH * J.KW * cblocks;
            const uint32_t idesc = make_idesc(T.ncols, 0, 0);
            const uint32_t tacc = tmem_base + (uint32_t)(buf * BUF);
            mbar_wait(&tmem_empty[buf], tph ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < num_kb; kb += KBS) {
                const int nk = num_kb - kb < KBS ? num_kb - kb : KBS;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int i = 0; i < KBS; ++i)
                        if (i < nk) {
                            const uint64_t da = make_desc(a_base + (s * KBS + i) * A_BYTES, 16, 1024);
                            const uint64_t db = make_desc(b_base + (s * KBS + i) * B_BYTES, 16, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // 4 x (K = 8 tf32 = 32 B) per 128-byte row
                                umma_tf32(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | i | k) != 0);
                        }
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&tmem_full[buf]);
            __syncwarp();
            buf ^= 1;
            if (buf == 0) tph ^= 1;
        }
    } else {
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        uint8_t* stage = sOut + q * 4096;
        int buf = 0;
        uint32_t tph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const PersTile T = pers_decode(p, t);
            const PersJob& J = p.job[T.job];
            const int m0 = T.m_tile * BM, n0 = T.n0, ncols = T.ncols;
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BUF);
            mbar_wait(&tmem_full[buf], tph);
            tc_fence_after();
            if (!p.o_strided) {
#pragma unroll 1
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tacc + (uint32_t)c0, r);
                    tmem_wait_ld();
                    float o[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
                            o[j] += b.x; o[j + 1] += b.y; o[j + 2] += b.z; o[j + 3] += b.w;
                        }
                    }
                    if (lane == 0) bulk_wait_read<0>();     // the previous store has finished reading the staging block
                    __syncwarp();
                    stage_rows_sw128(stage, lane, o);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && m0 + q * 32 < J.M) {
                        if (p.accumulate)
                            tma_reduce_add_2d(&maps.o, stage, n0 + c0, m0 + q * 32);
                        else
                            tma_store_2d(&maps.o, stage, n0 + c0, m0 + q * 32);
                        bulk_commit();
                    }
                    if (p.stats) {   // this warp's 32 rows of columns c0..c0+31 -> (n, mean, M2) per column
                        const int left = J.M - (m0 + q * 32);
                        const uint32_t vm = left >= 32 ? 0xFFFFFFFFu : (left > 0 ? ((1u << left) - 1u) : 0u);
                        const Moments mo = stage_col_moments(stage, lane, vm);
                        float* w = sStat + q * 3 * BN + c0 + lane;
                        w[0] = mo.n;
                        w[BN] = mo.mean;
                        w[2 * BN] = mo.m2;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);      // the accumulator buffer may be overwritten
                if (p.stats) {   // merge the four warps: one partial per m-tile and column
                    epilogue_bar();
                    for (int col = row; col < ncols; col += 128) {
                        Moments acc{0.f, 0.f, 0.f};
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4)
                            acc = merge_moments(acc, Moments{sStat[w4 * 3 * BN + col], sStat[(w4 * 3 + 1) * BN + col],
                                                             sStat[(w4 * 3 + 2) * BN + col]});
                        float* dst = p.stats + (int64_t)T.m_tile * 3 * p.ldo + n0 + col;
                        dst[0] = acc.n;
                        dst[p.ldo] = acc.mean;
                        dst[2 * p.ldo] = acc.m2;
                    }
                    epilogue_bar();   // sStat is rewritten by the next tile
                }
            } else {
                const int m = m0 + row;
                const int hw = J.OH * J.OW;
                const int n_img = m / hw, rem = m - n_img * hw;
                const int i = rem / J.OW, j2 = rem - i * J.OW;
                const int64_t pix = ((int64_t)n_img * p.o_H + J.o_a + i * p.o_sh) * p.o_W + J.o_b + j2 * p.o_sw;
                float* orow = p.out + pix * p.ldo + n0;
#pragma unroll 1
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tacc + (uint32_t)c0, r);
                    tmem_wait_ld();
                    if (m < J.M) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                   __uint_as_float(r[j + 3]));
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
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            }
            buf ^= 1;
            if (buf == 0) tph ^= 1;
        }
        if (lane == 0) bulk_wait_all<0>();     // global writes complete before the CTA exits
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BUF);
    }
}

template <int BN, int KBS, int STAGES>
static int launch_pers(const PersMaps& maps, const PersParams& p, int grid, cudaStream_t s) {
    constexpr int BUFC = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    (void)BUFC;
    constexpr int smem = STAGES * KBS * (BM * A_ROW_BYTES + BN * A_ROW_BYTES) + 16384 + 12 * BN * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
    static_assert(smem <= 227 * 1024, "persistent conv kernel: shared memory budget");
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(conv_pers_tc_k<BN, KBS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    conv_pers_tc_k<BN, KBS, STAGES><<<grid, kThreads, smem, s>>>(maps, p);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

}  // namespace tc
