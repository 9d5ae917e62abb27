// Persistent, filter-resident tcgen05 convolution for layers whose WHOLE filter fits in shared memory
// (the 3-channel stem and the 64->64 3x3 layers of ResNet-18: the layers the im2col kernel runs L2-bound).
//
// Why: conv_fprop_tc_k re-fetches, for every 128-pixel tile, the im2col patch (KH*KW x the unique input)
// and the filter: 21 FLOP per L2->SM byte at C = OC = 64, and the L2->SM fabric tops out at ~12 TB/s. Here
//  * the filter is loaded ONCE per CTA and stays in shared memory (B tiles [OC rows x 32 k] K-major, SWIZZLE_128B);
//  * CTAs are persistent (grid = #SMs) and walk 8-wide x 16-high output tiles of one image;
//  * the A operand of filter row kh is a SHIFTED VIEW of one shared-memory halo block: a TMA box
//    {32 k-elements, 8 output columns, (16-1)*sh + KH input rows} lands as rows*1024 bytes, output row r / filter
//    row kh is the 8-row swizzle group at byte (r*sh + kh)*1024, so the UMMA descriptor of k-block kh is
//    {start = stage + kh*1024, SBO = sh*1024}: every start stays 1024-byte aligned, the swizzle phase is intact;
//    one load feeds KH k-blocks (KH x less A traffic than im2col);
//  * accumulators are double-buffered in TMEM: the 4 epilogue warps drain tile i while tile i+1 is multiplied.
//
// 64->64 3x3: 2 loads per tile (one {32 ch, 10 w, 18 rows} halo block per channel block, zero padding by out-of-bounds
// fill) x 9 k-blocks: filter column kw is the view shifted by kw*128 B (see HaloParams::kwd). Stem (C <= 4): the image is pre-padded to 4 channels (stem_pad_c4_k) and the
// tensor map walks OVERLAPPING rows {32 floats = 8 pixels x 4 ch, stride sw pixels}: 1 load x KH k-blocks.
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_MAX_KH = 8;
constexpr int HALO_OUT_BYTES = 4 * 4096;

struct HaloParams {
    float* out;              // [N, OH, OW, ldo]
    const float* bias;       // [ldo] or null
    int N, OH, OW, ldo;
    int tiles_w, tiles_h, n_tiles;
    // per tile: cblocks x KWl loads (channel block cb, filter column kw), each followed by KH k-blocks (filter rows);
    // load (cb, kw) has TMA coordinates {cb*32, w0 + kw - pad_w, h0, n}; k-block kh multiplies the view at stage +
    // kh*1024 with resident filter tile kh*bt_kh_stride + kw*cblocks + cb
    int cblocks, KWl, KH, sh, pad_w, bt_kh_stride;
    int h_off;               // TMA row coordinate of output row 0, filter row 0 (= -pad, or 0 for a pre-padded image)
    int stage_bytes, n_stages, n_btiles;
    int accumulate;
    // kwd > 0: ONE halo load {32 ch, 8 + KW - 1 = kwd columns, rows} per channel block feeds all KW filter columns: the
    // view of (kh, kw) starts at stage + (kh*kwd + kw)*128 B with SBO = kwd*128 B -- neither is a multiple of the 1024-byte
    // swizzle period. Measured (parity suite): the MMA unit derives the SWIZZLE_128B phase of every operand row from its
    // absolute shared-memory address, exactly as TMA wrote it, so any 128-byte-aligned start / SBO is a valid view.
    // KW x less A traffic than one load per filter column (XSB_TC_HALO_KWD=0 restores that form for A/B runs).
    int kwd, tx_bytes;
    // BatchNorm partial statistics (null = off): stats[(blockIdx.x*3 + {0:n, 1:mean, 2:M2}) * ldo + channel]
    float* stats;
};

// KH_T > 0: filter height known at compile time (k-block loop fully unrolled); 0: runtime p.KH
template <int BN, int KH_T>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int B_TILE = BN * A_ROW_BYTES;
    constexpr int MAXS = 8;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sB = smem;
    uint8_t* sA = smem + p.n_btiles * B_TILE;
    uint8_t* sOut = sA + p.n_stages * p.stage_bytes;      // 4 x 4 KB: one SWIZZLE_128B staging block per epilogue warp
    uint64_t* full = reinterpret_cast<uint64_t*>(sOut + HALO_OUT_BYTES);
    uint64_t* empty = full + MAXS;
    uint64_t* bfull = empty + MAXS;
    uint64_t* tmem_full = bfull + 1;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.n_stages;
    const int KH = KH_T > 0 ? KH_T : p.KH;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmO);
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(bfull, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);   // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 2 * BN < 32 ? 32 : 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(bfull, p.n_btiles * B_TILE);
            for (int t = 0; t < p.n_btiles; ++t) tma_load_2d(sB + t * B_TILE, &tmB, bfull, t * 32, 0);
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 0;
        int n_img = blockIdx.x / tiles_per_img, rem = blockIdx.x - n_img * tiles_per_img;   // gridDim.x <= n_tiles
        int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int step_img = gridDim.x / tiles_per_img, step_rem = gridDim.x - step_img * tiles_per_img;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int w0 = tw * HALO_TW - p.pad_w, h0 = th * HALO_TH * p.sh + p.h_off;
            const int n_kw_loads = p.kwd ? 1 : p.KWl;
            for (int cb = 0; cb < p.cblocks; ++cb)
                for (int kw = 0; kw < n_kw_loads; ++kw) {
                    mbar_wait(&empty[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full[s], p.tx_bytes);
                        tma_load_4d(sA + s * p.stage_bytes, &tmA, &full[s], cb * 32, w0 + kw, h0, n_img);
                    }
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            // next tile of this CTA: (n_img, th, tw) += gridDim.x tiles
            rem += step_rem;
            n_img += step_img;
            if (rem >= tiles_per_img) { rem -= tiles_per_img; ++n_img; }
            th = rem / p.tiles_w;
            tw = rem - th * p.tiles_w;
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, 0, 0);
        constexpr uint32_t B_ENC = B_TILE >> 4;
        // descriptors as {lo = start address | LBO, hi = SBO | version | layout}: only lo moves
        const uint32_t a_hi = desc_hi((uint32_t)p.sh * 1024, kLayoutSW128), b_hi = desc_hi(1024, kLayoutSW128);
        const uint32_t a_lo0 = desc_lo(smem_u32(sA), 16), b_lo0 = desc_lo(smem_u32(sB), 16);
        const uint32_t stage_enc = (uint32_t)p.stage_bytes >> 4;
        const uint32_t bt_kh_enc = (uint32_t)p.bt_kh_stride * B_ENC;
        mbar_wait(bfull, 0);
        int s = 0, buf = 0;
        uint32_t ph = 0, tph = 0, a_lo_s = a_lo0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[buf], tph ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
            uint32_t fresh = 0;   // 0 for the first MMA of the tile (overwrite the accumulator), 1 afterwards
            if (p.kwd) {
                const uint32_t a_hi_d = desc_hi((uint32_t)p.kwd * 128, kLayoutSW128);
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        uint32_t b_lo_l = b_lo0 + (uint32_t)cb * B_ENC;
                        for (int kw = 0; kw < p.KWl; ++kw, b_lo_l += (uint32_t)p.cblocks * B_ENC)
                            for (int kh = 0; kh < KH; ++kh)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_tf32_split(tacc, a_lo_s + (uint32_t)((kh * p.kwd + kw) * 8 + k * 2), a_hi_d,
                                                    b_lo_l + (uint32_t)kh * bt_kh_enc + (uint32_t)(k * 2), b_hi, idesc,
                                                    (kw | kh | k) ? 1u : fresh);
                        umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    fresh = 1;
                    a_lo_s += stage_enc;
                    if (++s == S) { s = 0; ph ^= 1; a_lo_s = a_lo0; }
                }
            } else
            for (int cb = 0; cb < p.cblocks; ++cb) {
                uint32_t b_lo_l = b_lo0 + (uint32_t)cb * B_ENC;
                for (int kw = 0; kw < p.KWl; ++kw, b_lo_l += (uint32_t)p.cblocks * B_ENC) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        if (KH_T > 0) {
#pragma unroll
                            for (int kh = 0; kh < (KH_T > 0 ? KH_T : 1); ++kh)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_tf32_split(tacc, a_lo_s + (uint32_t)(kh * 64 + k * 2), a_hi,
                                                    b_lo_l + (uint32_t)kh * bt_kh_enc + (uint32_t)(k * 2), b_hi, idesc,
                                                    (kh | k) ? 1u : fresh);
                        } else {
                            for (int kh = 0; kh < KH; ++kh)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_tf32_split(tacc, a_lo_s + (uint32_t)(kh * 64 + k * 2), a_hi,
                                                    b_lo_l + (uint32_t)kh * bt_kh_enc + (uint32_t)(k * 2), b_hi, idesc,
                                                    (kh | k) ? 1u : fresh);
                        }
                        umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    fresh = 1;
                    a_lo_s += stage_enc;
                    if (++s == S) { s = 0; ph ^= 1; a_lo_s = a_lo0; }
                }
            }
            if (elect_one()) umma_commit(&tmem_full[buf]);
            __syncwarp();
            buf ^= 1;
            if (buf == 0) tph ^= 1;
        }
    } else {
        // epilogue warp q owns accumulator rows 32q..32q+31 = output rows 4q..4q+3 of the tile (8 pixels each): TMEM ->
        // registers (+bias) -> swizzled staging block -> ONE TMA store (or reduce-add) of a {32 ch, 8 w, 4 h} box; the
        // tensor map clips rows / columns past the image, so partial tiles need no predication. The TMEM load of chunk
        // i+1 is in flight while chunk i is staged. (Measured: a second staging block per warp -- store i-1 still reading
        // while chunk i is staged -- changes nothing on the stem and costs the 64-channel 3x3 layers 3 us: their MMAs are
        // bound by the same shared-memory port the faster epilogue then competes for.)
        const int q = warp & 3;
        uint8_t* stage = sOut + q * 4096;
        int buf = 0;
        uint32_t tph = 0;
        // BatchNorm statistics. REG_STATS (BN = 64): every thread keeps sum / sum of squares of ITS accumulator row for all
        // 64 columns over all tiles of the CTA in registers (bias-free values: the bias is the pivot) -- no shared-memory
        // reads, no dependent chains; lanes and warps are merged ONCE per CTA. Otherwise: per-chunk column moments from
        // the staged block, Chan-merged per tile.
        constexpr bool REG_STATS = (BN == 64);
        constexpr int NRS = REG_STATS ? BN : 1, NRUN = REG_STATS ? 1 : BN / 32;
        float rs1[NRS], rs2[NRS], rcnt = 0.f;
#pragma unroll
        for (int i = 0; i < NRS; ++i) { rs1[i] = 0.f; rs2[i] = 0.f; }
        Moments run[NRUN];   // lane j: running statistics of columns 32*i + j over this warp's rows of every tile
#pragma unroll
        for (int i = 0; i < NRUN; ++i) run[i] = Moments{0.f, 0.f, 0.f};
        int n_img = blockIdx.x / tiles_per_img, rem = blockIdx.x - n_img * tiles_per_img;
        const int step_img = gridDim.x / tiles_per_img, step_rem = gridDim.x - step_img * tiles_per_img;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
            const int oh0 = th * HALO_TH + q * 4, ow0 = tw * HALO_TW;
            uint32_t vm = 0;   // valid rows of this warp's {8 w, 4 h} box
            if (p.stats) {
                const uint32_t wmask = ow0 + HALO_TW <= p.OW ? 0xFFu : ((1u << (p.OW - ow0)) - 1u);
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    if (oh0 + h < p.OH) vm |= wmask << (8 * h);
            }
            const bool mine = (vm >> lane) & 1u;
            if (REG_STATS && mine) rcnt += 1.f;
            mbar_wait(&tmem_full[buf], tph);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
            uint32_t v[2][32];
            tmem_ld32(t0, v[0]);
#pragma unroll
            for (int ci = 0; ci < BN / 32; ++ci) {
                const int c0 = ci * 32;
                uint32_t (&r)[32] = v[ci & 1];
                tmem_wait_ld();
                tmem_regs_ready(r);
                if (ci + 1 < BN / 32) tmem_ld32(t0 + (uint32_t)(c0 + 32), v[(ci + 1) & 1]);
                float o[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
                if (REG_STATS) {
                    if (mine) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            rs1[(REG_STATS ? c0 : 0) + j] += o[j];
                            rs2[(REG_STATS ? c0 : 0) + j] = fmaf(o[j], o[j], rs2[(REG_STATS ? c0 : 0) + j]);
                        }
                    }
                }
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
                        o[j] += b.x; o[j + 1] += b.y; o[j + 2] += b.z; o[j + 3] += b.w;
                    }
                }
                if (lane == 0) bulk_wait_read<0>();     // the previous store has finished reading the staging block
                __syncwarp();
                stage_rows_sw128(stage, lane, o);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (oh0 < p.OH) {
                        if (p.accumulate)
                            tma_reduce_add_4d(&tmO, stage, c0, ow0, oh0, n_img);
                        else
                            tma_store_4d(&tmO, stage, c0, ow0, oh0, n_img);
                    }
                    bulk_commit();
                }
                if (!REG_STATS && p.stats)
                    run[REG_STATS ? 0 : ci] = merge_moments(run[REG_STATS ? 0 : ci], stage_col_moments(stage, lane, vm));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            buf ^= 1;
            if (buf == 0) tph ^= 1;
            rem += step_rem;
            n_img += step_img;
            if (rem >= tiles_per_img) { rem -= tiles_per_img; ++n_img; }
        }
        if (lane == 0) bulk_wait_all<0>();   // global writes complete before the CTA exits
        if (p.stats) {   // one partial per CTA and column: merge the four warps through the staging area
            __syncwarp();
            epilogue_bar();   // ... once EVERY warp's last store has finished reading its staging block
            float* sStat = reinterpret_cast<float*>(sOut);
            if (REG_STATS) {
                // thread -> (n, mean, M2) per column (n is the same for all columns of a thread), then a transposing
                // butterfly over the warp: level l exchanges half of the remaining columns with lane ^ (16 >> l), so
                // after 5 levels lane L holds columns 2L, 2L+1 merged over all 32 lanes (62 merges, not 64 x 5)
                float n = rcnt;
                const float inv = n > 0.f ? 1.f / n : 0.f;
#pragma unroll
                for (int i = 0; i < NRS; ++i) {
                    const float mean = rs1[i] * inv;
                    rs2[i] = fmaxf(rs2[i] - rs1[i] * mean, 0.f);
                    rs1[i] = mean;
                }
#pragma unroll
                for (int lvl = 0; lvl < 5; ++lvl) {
                    const int mask = 16 >> lvl, half = (NRS / 2) >> lvl;
                    const bool up = (lane & mask) != 0;
                    const float n_o = __shfl_xor_sync(0xFFFFFFFFu, n, mask);
                    const float n_t = n + n_o;
                    const float f = n_t > 0.f ? n_o / n_t : 0.f;     // weight of the partner's half
                    const float nf = n * f;
#pragma unroll
                    for (int i = 0; i < half; ++i) {
                        const float keep_m = up ? rs1[i + half] : rs1[i], send_m = up ? rs1[i] : rs1[i + half];
                        const float keep_q = up ? rs2[i + half] : rs2[i], send_q = up ? rs2[i] : rs2[i + half];
                        const float om = __shfl_xor_sync(0xFFFFFFFFu, send_m, mask);
                        const float oq = __shfl_xor_sync(0xFFFFFFFFu, send_q, mask);
                        const float d = om - keep_m;
                        rs1[i] = keep_m + d * f;
                        rs2[i] = keep_q + oq + d * d * nf;
                    }
                    n = n_t;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int col = 2 * lane + i;
                    sStat[q * 3 * BN + col] = n;
                    sStat[(q * 3 + 1) * BN + col] = rs1[REG_STATS ? i : 0] + (p.bias ? __ldg(p.bias + col) : 0.f);
                    sStat[(q * 3 + 2) * BN + col] = rs2[REG_STATS ? i : 0];
                }
            } else {
#pragma unroll
                for (int i = 0; i < NRUN; ++i) {
                    sStat[q * 3 * BN + i * 32 + lane] = run[i].n;
                    sStat[(q * 3 + 1) * BN + i * 32 + lane] = run[i].mean;
                    sStat[(q * 3 + 2) * BN + i * 32 + lane] = run[i].m2;
                }
            }
            epilogue_bar();
            for (int col = q * 32 + lane; col < BN; col += 128) {
                Moments acc{0.f, 0.f, 0.f};
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4)
                    acc = merge_moments(acc, Moments{sStat[w4 * 3 * BN + col], sStat[(w4 * 3 + 1) * BN + col],
                                                     sStat[(w4 * 3 + 2) * BN + col]});
                float* dst = p.stats + (int64_t)blockIdx.x * 3 * p.ldo + col;
                dst[0] = acc.n;
                dst[p.ldo] = acc.mean;
                dst[2 * p.ldo] = acc.m2;
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN < 32 ? 32 : 2 * BN);
    }
}

constexpr int HALO_SMEM_MAX = 227 * 1024;


// -> number of A stages that fit next to the resident filter (0: does not fit)
static int halo_stages(int n_btiles, int BN, int stage_bytes) {
    const int fixed = 1024 + n_btiles * BN * A_ROW_BYTES + HALO_OUT_BYTES + 256;
    int s = (HALO_SMEM_MAX - fixed) / stage_bytes;
    if (s > 6) s = 6;
    return s >= 2 ? s : 0;
}

template <int BN, int KH_T>
static int launch_halo_t(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const HaloParams& p,
                         cudaStream_t s) {
    const int smem = 1024 + p.n_btiles * BN * A_ROW_BYTES + p.n_stages * p.stage_bytes + HALO_OUT_BYTES + 256;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(conv_halo_tc_k<BN, KH_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX));
        configured = true;
    }
    int grid = xsb_sm_count_cached();
    if (grid > p.n_tiles) grid = p.n_tiles;
    conv_halo_tc_k<BN, KH_T><<<grid, kThreads, smem, s>>>(a, b, o, p);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// output tensor [N, OH, OW, OC] as a TMA-store target: box {32 ch, 8 w, 4 h} (one epilogue warp's rows)
static bool halo_out_map(CUtensorMap* m, float* out, int N, int OH, int OW, int OC) {
    cuuint64_t dims[4] = {(cuuint64_t)OC, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)OC * 4, (cuuint64_t)OW * OC * 4, (cuuint64_t)OH * OW * OC * 4};
    cuuint32_t box[4] = {32, HALO_TW, 4, 1};
    return make_map_tiled(m, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN>
static int launch_halo(const CUtensorMap& a, const CUtensorMap& b, const HaloParams& p, cudaStream_t s) {
    CUtensorMap o;
    if (!halo_out_map(&o, p.out, p.N, p.OH, p.OW, p.ldo)) return XSB_ERR_UNSUPPORTED;
    if (p.KH == 3) return launch_halo_t<BN, 3>(a, b, o, p, s);
    if (p.KH == 7) return launch_halo_t<BN, 7>(a, b, o, p, s);
    return launch_halo_t<BN, 0>(a, b, o, p, s);
}

// a caller's request for BatchNorm partial statistics of the conv output (see xsb_conv2d_fwd_stats)
struct StatsReq {
    float* buf;          // device, `floats` long
    int64_t floats;
    int* n_partials;     // host, receives the number of partials written (0: the kernel that ran produces none)
};
static void halo_attach_stats(HaloParams& p, StatsReq* st) {
    int grid = xsb_sm_count_cached();
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (st && st->buf && st->floats >= (int64_t)grid * 3 * p.ldo) {
        p.stats = st->buf;
        if (st->n_partials) *st->n_partials = grid;
    }
}

// the shape envelope of halo_conv_s1 (no side effects): callers that must prepare the filter first ask this
static bool halo_s1_accepts(int Cx, int OCx, int KH, int KW, int OH, int OW) {
    if (Cx % 32 != 0 || (OCx != 64 && OCx != 128) || KH > HALO_MAX_KH) return false;
    const int rows = HALO_TH - 1 + KH;
    const int stage = (rows * (HALO_TW + KW - 1) * A_ROW_BYTES + 1023) & ~1023;     // (the kwd form: the larger stage)
    if (halo_stages(KH * KW * (Cx / 32), OCx, stage) == 0 || rows > 256) return false;
    const int tiles_w = (OW + HALO_TW - 1) / HALO_TW, tiles_h = (OH + HALO_TH - 1) / HALO_TH;
    return (int64_t)OH * OW * 10 >= (int64_t)tiles_w * tiles_h * HALO_TW * HALO_TH * 8;
}

// Generic stride-1 path: conv of channels-last in[N,H,W,Cx] with flt[OCx, KH*KW*Cx] -> out[N,OH,OW,OCx].
// Returns XSB_ERR_UNSUPPORTED when the shape is outside the envelope (caller runs the im2col kernel).
static int halo_conv_s1(const float* in, int N, int H, int W, int Cx, const float* flt, int OCx, int KH, int KW, int pad_h,
                        int pad_w, int OH, int OW, const float* bias, float* out, int accumulate, cudaStream_t s,
                        StatsReq* st = nullptr) {
    const int cblocks = Cx / 32;
    if (Cx % 32 != 0 || (OCx != 64 && OCx != 128) || KH > HALO_MAX_KH) return XSB_ERR_UNSUPPORTED;
    const int n_btiles = KH * KW * cblocks;
    const int rows = HALO_TH - 1 + KH;
    static const int kwd_on = [] { const char* e = getenv("XSB_TC_HALO_KWD"); return (e && e[0] == '0') ? 0 : 1; }();
    const int kwd = (kwd_on && KW > 1) ? HALO_TW + KW - 1 : 0;
    const int tx_bytes = rows * (kwd ? kwd : HALO_TW) * A_ROW_BYTES;
    const int stage_bytes = (tx_bytes + 1023) & ~1023;
    const int stages = halo_stages(n_btiles, OCx, stage_bytes);
    if (stages == 0 || rows > 256) return XSB_ERR_UNSUPPORTED;
    // small images waste most of an 8x16 tile: leave them to the im2col kernel
    const int tiles_w = (OW + HALO_TW - 1) / HALO_TW, tiles_h = (OH + HALO_TH - 1) / HALO_TH;
    if ((int64_t)OH * OW * 10 < (int64_t)tiles_w * tiles_h * HALO_TW * HALO_TH * 8) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    CUtensorMap ta, tb;
    cuuint64_t dims[4] = {(cuuint64_t)Cx, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cx * 4, (cuuint64_t)W * Cx * 4, (cuuint64_t)H * W * Cx * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(kwd ? kwd : HALO_TW), (cuuint32_t)rows, 1};
    if (!make_map_tiled(&ta, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_map_2d(&tb, flt, (uint64_t)OCx, (uint64_t)KH * KW * Cx, (uint32_t)OCx))
        return XSB_ERR_UNSUPPORTED;
    HaloParams p{};
    p.out = out; p.bias = bias; p.N = N; p.OH = OH; p.OW = OW; p.ldo = OCx;
    p.tiles_w = tiles_w; p.tiles_h = tiles_h; p.n_tiles = N * tiles_w * tiles_h;
    p.cblocks = cblocks; p.KWl = KW; p.KH = KH; p.sh = 1; p.pad_w = pad_w; p.bt_kh_stride = KW * cblocks;
    p.h_off = -pad_h;
    p.stage_bytes = stage_bytes; p.n_stages = stages; p.n_btiles = n_btiles; p.accumulate = accumulate;
    p.kwd = kwd; p.tx_bytes = tx_bytes;
    halo_attach_stats(p, st);
    return OCx == 64 ? launch_halo<64>(ta, tb, p, s) : launch_halo<128>(ta, tb, p, s);
}

}  // namespace tc
