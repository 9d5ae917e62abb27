// Persistent tcgen05 weight-gradient kernel for stride-1 convolutions (XSB_MATH_TF32):
//   dW[oc, kh, kw, c] (+)= sum_{n,oh,ow} dY[n,oh,ow,oc] * x[n, oh+kh-pad, ow+kw-pad, c]        (xs/nn/grad_fn.py:186-194)
//
// conv_wgrad_tc_k gives every 128-row block of (tap, c) its own CTAs: each re-reads dY and fetches the im2col patch
// of its taps -- 22 FLOP per L2->SM byte at C = OC = 64, i.e. pinned to the ~12 TB/s L2->SM fabric. Here
//  * one CTA owns ALL filter rows of `cbp` channel blocks x KW filter columns x BN output channels: cbp*KW
//    accumulators of 128 x BN in TMEM (<= 512 columns), so a dY tile is loaded once for every (tap, c) it meets;
//  * the pixel (K) dimension walks 8 x 8 output tiles; for filter column kw and channel block cb ONE TMA halo block
//    {32 ch, 8 w, 8+KH-1 rows} lands as rows of 128 B (pixel-major = MN-major A operand, SWIZZLE_128B_ATOM_32B);
//    tile row r / filter row kh is the 1024-byte group at (r + kh)*1024, so the four 32-row M blocks of one MMA are
//    the filter rows kh = 0..3 (descriptor LBO = 1024 B; rows of a non-existent kh land in accumulator lanes nobody
//    reads) and the K = 8 pixels of tile row r start at r*1024: KH x less patch traffic than im2col, no re-read of dY;
//  * CTAs are persistent over a contiguous range of tiles (split-K across CTAs), accumulate in TMEM over the whole
//    range and leave with ONE TMA reduce-add of a {32 c, 32 oc} box per warp, accumulator and 32 output channels.
// grid = (CTAs per pass, passes); pass = (channel-block group, BN-wide output-channel tile).
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int WH_TW = 8, WH_TH = 8;     // output tile = 64 pixels = 8 MMAs of K = 8

struct WgradHaloParams {
    int N, OH, OW;
    int tiles_w, tiles_h, n_tiles;
    int KH, KW, pad, cblocks;
    int cbp;              // channel blocks per pass
    int n_oc_tiles;       // OC / BN
    int sa_slots, sb_slots, a_slot_bytes;
    int C;                // input channels (column stride of the dW map)
    // generalisation for the stem (C <= 4 pre-padded to 4 channels, filter columns folded into the 32-float row):
    int sh;               // conv stride along h: tile row r / filter row kh is the group at (r*sh + kh)*1024
    int groups;           // ceil(KH / 4) accumulators per load: M blocks kh = 4g .. 4g+3
    int KWl;              // halo loads per channel block (KW, or 1 when the row already holds all filter columns)
    int pad_h, pad_w;     // subtracted from the TMA coordinates (0 for a pre-padded image)
    int stem;             // epilogue: rows are (kh, kw*4 + c) -> scattered red.add into dw_raw[oc][kh][kw][c]
    int KW_real, C_real;
    float* dw_raw;
    // kwd > 0: one halo load {32 ch, kwd = 8 + KW - 1 columns, rows} per channel block feeds all KW filter columns -- the
    // operand view of column kw starts kw*128 B into the block, rows are kwd*128 B apart (LBO). As in conv_halo_tc.cuh
    // the MMA unit takes the swizzle phase (here SWIZZLE_128B_BASE32B) from the absolute shared-memory address.
    int kwd;
};

// SH = conv stride along h, G = accumulators (groups of four filter rows) per load: compile-time so that the issue
// loop is immediates only (N = 64 MMAs are issue-sensitive)
template <int BN, int SH, int G>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_halo_k(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                  const __grid_constant__ CUtensorMap tmW, const WgradHaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int B_BOX = WH_TW * WH_TH * A_ROW_BYTES;     // 64 pixels x 32 oc = 8 KB
    constexpr int B_BYTES = (BN / 32) * B_BOX;
    constexpr int MAXS = 16;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sB = sA + p.sa_slots * p.a_slot_bytes;
    // "full" barriers are per TILE (the dY load and every halo load of a tile arrive on the same one): the MMA thread
    // consumes one try_wait per tile instead of one per load -- each costs ~160 clk of idle tensor pipe (tc_common.cuh).
    // "empty" barriers stay per slot, so the producer refills slot by slot.
    constexpr int TS = 4;
    uint64_t* fullT = reinterpret_cast<uint64_t*>(sB + p.sb_slots * B_BYTES);
    uint64_t* emptyA = fullT + TS;
    uint64_t* emptyB = emptyA + MAXS;
    uint64_t* tmem_full = emptyB + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int SA = p.sa_slots, SB = p.sb_slots;
    const int pass = blockIdx.y;
    const int oc0 = (pass % p.n_oc_tiles) * BN;
    const int cb0 = (pass / p.n_oc_tiles) * p.cbp;
    int cbn = p.cblocks - cb0;                 // channel blocks of this pass
    if (cbn > p.cbp) cbn = p.cbp;
    const int kw_per_load = p.kwd ? p.KWl : 1;                  // filter columns served by one halo load
    const int n_loads = cbn * (p.kwd ? 1 : p.KWl);              // TMA halo loads per tile
    const int n_acc = n_loads * kw_per_load * G;
    const int wpx = p.kwd ? p.kwd : WH_TW;                      // pixels per halo row
    // contiguous tile range of this CTA
    const int t_begin = (int)((int64_t)blockIdx.x * p.n_tiles / gridDim.x);
    const int t_end = (int)((int64_t)(blockIdx.x + 1) * p.n_tiles / gridDim.x);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmDy);
        prefetch_tmap(&tmW);
        for (int i = 0; i < TS; ++i) mbar_init(&fullT[i], (uint32_t)n_loads + 1);   // one arrive.expect_tx per load
        for (int i = 0; i < SA; ++i) mbar_init(&emptyA[i], 1);
        for (int i = 0; i < SB; ++i) mbar_init(&emptyB[i], 1);
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int a_box_bytes = ((WH_TH - 1) * SH + p.KH) * wpx * A_ROW_BYTES;     // bytes the halo load writes

    if (warp == 0) {
        int sa = 0, sb = 0, ts = 0;
        uint32_t pha = 0, phb = 0;
        int n_img = t_begin / tiles_per_img, rem = t_begin - n_img * tiles_per_img;
        int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        for (int t = t_begin; t < t_end; ++t) {
            const int w0 = tw * WH_TW, h0 = th * WH_TH;
            mbar_wait(&emptyB[sb], phb ^ 1);
            uint64_t* full = &fullT[ts];
            ts = (ts + 1) & (TS - 1);
            if (elect_one()) {
                mbar_expect_tx(full, B_BYTES);
#pragma unroll
                for (int b = 0; b < BN / 32; ++b)
                    tma_load_4d(sB + sb * B_BYTES + b * B_BOX, &tmDy, full, oc0 + b * 32, w0, h0, n_img);
            }
            if (++sb == SB) { sb = 0; phb ^= 1; }
            for (int cb = cb0; cb < cb0 + cbn; ++cb)
                for (int kw = 0; kw < (p.kwd ? 1 : p.KWl); ++kw) {
                    mbar_wait(&emptyA[sa], pha ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(full, a_box_bytes);
                        tma_load_4d(sA + sa * p.a_slot_bytes, &tmX, full, cb * 32, w0 + kw - p.pad_w,
                                    h0 * SH - p.pad_h, n_img);
                    }
                    if (++sa == SA) { sa = 0; pha ^= 1; }
                }
            if (++tw == p.tiles_w) {
                tw = 0;
                if (++th == p.tiles_h) { th = 0; ++n_img; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, 1, 1);
        const uint32_t hi = desc_hi(512, kLayoutSW128Base32B);
        const uint32_t a_lo0 = desc_lo(smem_u32(sA), (uint32_t)wpx * 128), b_lo0 = desc_lo(smem_u32(sB), B_BOX);
        const uint32_t row_enc = (uint32_t)wpx * 8;     // one halo row in descriptor units (16 B)
        const uint32_t a_enc = (uint32_t)p.a_slot_bytes >> 4;
        int sa = 0, sb = 0, ts = 0;
        uint32_t pht = 0, a_lo = a_lo0, b_lo = b_lo0;
        uint32_t fresh = 0;     // 0 while this CTA's accumulators have not been written yet
        for (int t = t_begin; t < t_end; ++t) {
            mbar_wait(&fullT[ts], pht);
            tc_fence_after();
            if (++ts == TS) { ts = 0; pht ^= 1; }
            for (int a = 0; a < n_loads; ++a) {
                if (elect_one()) {
                    for (int kw = 0; kw < kw_per_load; ++kw) {
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const uint32_t tacc = tmem_base + (uint32_t)(((a * kw_per_load + kw) * G + g) * BN);
#pragma unroll
                            for (int r = 0; r < WH_TH; ++r)   // tile row r: K = 8 pixels = 1024 bytes of both operands
                                umma_tf32_split(tacc, a_lo + (uint32_t)(g * 4 + r * SH) * row_enc + (uint32_t)(kw * 8), hi,
                                                b_lo + (uint32_t)(r * 64), hi, idesc, r ? 1u : fresh);
                        }
                    }
                    umma_commit(&emptyA[sa]);
                    if (a == n_loads - 1) umma_commit(&emptyB[sb]);
                }
                __syncwarp();
                a_lo += a_enc;
                if (++sa == SA) { sa = 0; a_lo = a_lo0; }
            }
            fresh = 1;
            b_lo += (uint32_t)(B_BYTES >> 4);
            if (++sb == SB) { sb = 0; b_lo = b_lo0; }
        }
        if (elect_one()) umma_commit(tmem_full);
        __syncwarp();
    } else {
        // warp q reads TMEM lanes 32q..32q+31 = filter row kh = q, channel lane; lanes of kh >= KH hold garbage
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (t_end > t_begin && p.stem) {
            // rows of accumulator g: (kh = 4g + q, lane = kw*4 + c); scattered element-wise reduction into dW[oc][kh][kw][c]
            const int kw = lane >> 2, c = lane & 3;
            for (int g = 0; g < G; ++g) {
                const int kh = 4 * g + q;
                const bool valid = kh < p.KH && kw < p.KW_real && c < p.C_real;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN + c0), v);
                    tmem_wait_ld();
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            atomicAdd(p.dw_raw + (((int64_t)(oc0 + c0 + j) * p.KH + kh) * p.KW_real + kw) * p.C_real + c,
                                      __uint_as_float(v[j]));
                    }
                }
            }
        } else if (t_end > t_begin && q < p.KH) {
            float* stage = reinterpret_cast<float*>(sA + q * 4096);    // the rings are idle: [32 oc][32 c] per warp
            for (int a = 0; a < n_acc; ++a) {
                const int cb = cb0 + a / p.KW, kw = a - (a / p.KW) * p.KW;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c0), v);
                    tmem_wait_ld();
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) stage[j * 32 + lane] = __uint_as_float(v[j]);   // row = oc, column = c
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&tmW, stage, (q * p.KW + kw) * p.C + cb * 32, oc0 + c0);
                        bulk_commit();
                    }
                }
            }
            if (lane == 0) bulk_wait_all<0>();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// dw must hold the value to accumulate onto (zeros for "="): every CTA reduce-adds its partial sums.
// Returns XSB_ERR_UNSUPPORTED outside the envelope (caller runs conv_wgrad_tc_k).
static int wgrad_halo(const float* dy, const float* x, float* dw, const ConvGeom& g, cudaStream_t s) {
    if (g.sh != 1 || g.sw != 1 || g.C % 32 != 0 || g.OC % 64 != 0 || g.KH > 4 || g.KH < 2 || g.KW > 8) return XSB_ERR_UNSUPPORTED;
    const int BN = g.OC % 128 == 0 ? 128 : 64;
    const int cblocks = g.C / 32;
    int cbp = 512 / (g.KW * BN);               // accumulators of one pass: cbp * KW * BN <= 512 TMEM columns
    if (cbp < 1) return XSB_ERR_UNSUPPORTED;
    if (cbp > cblocks) cbp = cblocks;
    // measured (R224): with one pass (all channel blocks' accumulators resident, C = OC = 64) this beats the im2col
    // kernel (102 vs 108 us); with several passes every pass re-reads dY and it loses (84-95 vs 69 us): single pass only
    if (cbp < cblocks) return XSB_ERR_UNSUPPORTED;
    const int cb_groups = (cblocks + cbp - 1) / cbp;
    const int n_oc_tiles = g.OC / BN;
    const int passes = cb_groups * n_oc_tiles;
    const int S = xsb_sm_count_cached();
    if (passes > S) return XSB_ERR_UNSUPPORTED;
    const int tiles_w = (g.OW + WH_TW - 1) / WH_TW, tiles_h = (g.OH + WH_TH - 1) / WH_TH;
    // tiles mostly outside the image (small feature maps) waste the MMAs: leave those to the im2col kernel
    if ((int64_t)g.OH * g.OW * 10 < (int64_t)tiles_w * tiles_h * WH_TW * WH_TH * 7) return XSB_ERR_UNSUPPORTED;
    const int n_tiles = g.N * tiles_w * tiles_h;
    int ctas = S / passes;
    if (ctas > n_tiles) ctas = n_tiles;
    if (ctas < 1) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    const int a_rows = WH_TH + g.KH - 1;
    static const int kwd_on = [] { const char* e = getenv("XSB_TC_WGRAD_KWD"); return (e && e[0] == '0') ? 0 : 1; }();
    const int kwd = (kwd_on && g.KW > 1) ? WH_TW + g.KW - 1 : 0;
    const int wpx = kwd ? kwd : WH_TW;
    // room for the rows the (unused) M blocks kh < 4 touch
    const int a_slot_bytes = ((WH_TH + 4) * wpx * A_ROW_BYTES + 1023) & ~1023;
    const int b_bytes = (BN / 32) * WH_TW * WH_TH * A_ROW_BYTES;
    const int sb_slots = 3;
    int sa_slots = (227 * 1024 - 1024 - 512 - sb_slots * b_bytes) / a_slot_bytes;
    if (sa_slots > 16) sa_slots = 16;
    if (sa_slots < 4 || sa_slots < cbp * (kwd ? 1 : g.KW)) return XSB_ERR_UNSUPPORTED;   // a tile's loads must all fit the ring (one full barrier per tile)
    CUtensorMap tx, tdy, tw;
    cuuint64_t xd[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t xs_[3] = {(cuuint64_t)g.C * 4, (cuuint64_t)g.W * g.C * 4, (cuuint64_t)g.H * g.W * g.C * 4};
    cuuint32_t xb[4] = {32, (cuuint32_t)wpx, (cuuint32_t)a_rows, 1};
    cuuint64_t dd[4] = {(cuuint64_t)g.OC, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
    cuuint64_t ds[3] = {(cuuint64_t)g.OC * 4, (cuuint64_t)g.OW * g.OC * 4, (cuuint64_t)g.OH * g.OW * g.OC * 4};
    cuuint32_t db[4] = {32, WH_TW, WH_TH, 1};
    if (!make_map_tiled(&tx, x, 4, xd, xs_, xb, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map_tiled(&tdy, dy, 4, dd, ds, db, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map_2d(&tw, dw, (uint64_t)g.OC, (uint64_t)g.KH * g.KW * g.C, 32, CU_TENSOR_MAP_SWIZZLE_NONE))
        return XSB_ERR_UNSUPPORTED;
    WgradHaloParams p{g.N, g.OH, g.OW, tiles_w, tiles_h, n_tiles, g.KH, g.KW, g.pad, cblocks, cbp, n_oc_tiles,
                      sa_slots, sb_slots, a_slot_bytes, g.C, 1, 1, g.KW, g.pad, g.pad, 0, g.KW, g.C, dw, kwd};
    const int smem = 1024 + sa_slots * a_slot_bytes + sb_slots * b_bytes + 512;
    dim3 grid(ctas, passes);
    if (BN == 64) {
        static bool configured = false;
        if (!configured) {
            XSB_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_k<64, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            configured = true;
        }
        conv_wgrad_halo_k<64, 1, 1><<<grid, kThreads, smem, s>>>(tx, tdy, tw, p);
    } else {
        static bool configured = false;
        if (!configured) {
            XSB_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_k<128, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            configured = true;
        }
        conv_wgrad_halo_k<128, 1, 1><<<grid, kThreads, smem, s>>>(tx, tdy, tw, p);
    }
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// ------------------------------------------------------------------------------------------ row-strip wgrad (C, OC >= 128)
// conv_wgrad_tc_k needs 128 B/clk of L2 -> SM traffic per SM to keep the tensor pipe busy (a 64 KB stage per 8 MMAs of
// 128 x 128 x 8) against the ~55 B/clk an SM ingests: the wgrad of layers 2-4 of ResNet-18 sits at 415-430 TFLOP/s.
// Here the K (pixel) dimension of one MMA chain is a STRIP of TH whole image rows, stored exactly as TMA delivers it:
//   A = x halo strip  {32 ch, WX = W + 2*pad columns from w = -pad, TH rows from h0 + kh - pad}  -> TH*WX rows of 128 B
//   B = dY strip      {32 oc, WX columns from ow = 0 (columns >= OW are out of bounds = 0), TH rows from h0}
// so pixel (r, j) of the strip is row k = r*WX + j of BOTH blocks, and the operand of filter column kw is the A block viewed
// kw rows (kw * 128 B) further: x[h0+r+kh-pad, j+kw-pad] pairs with dY[h0+r, j]. Rows of a view that run past an image row
// (j + kw >= WX) or past the block meet dY's zero columns, so they contribute nothing (the A slack rows only have to hold
// finite numbers: shared memory is zeroed once). One load of A and B feeds KW accumulators (KW * BN <= 512 TMEM columns);
// M = 128 channels = four 32-channel blocks (descriptor LBO = block pitch), N = BN output channels.
// 2.7 KB of operands per MMA instead of 8 KB. A pass = (filter row kh, 128-channel tile, BN-wide oc tile); the CTAs of a pass
// split the strips (split-K) and leave with TMA reduce-adds, as conv_wgrad_halo_k does.
struct WgradStripParams {
    int n_tiles, tiles_h, TH, WX, K8;   // K8 = ceil(TH*WX / 8) MMAs (K = 8 pixels each) per filter column and strip
    int KH, KW, pad, C;
    int n_c_tiles, n_oc_tiles;          // C / 128, OC / BN
    int a_cb_bytes, b_box_bytes, stage_bytes, n_stages, tx_bytes;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_strip_k(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                   const __grid_constant__ CUtensorMap tmW, const WgradStripParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int MAXS = 8;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* ring = smem;                                             // [stage]: 4 A blocks, then BN/32 B blocks
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + p.n_stages * p.stage_bytes);
    uint64_t* empty = full + MAXS;
    uint64_t* tmem_full = empty + MAXS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.n_stages;
    // pass -> (kh, channel tile, oc tile)
    const int pass = blockIdx.y;
    const int oc_t = pass % p.n_oc_tiles, c_t = (pass / p.n_oc_tiles) % p.n_c_tiles, kh = pass / (p.n_oc_tiles * p.n_c_tiles);
    const int oc0 = oc_t * BN, cblk0 = c_t * 4;
    const int t_begin = (int)((int64_t)blockIdx.x * p.n_tiles / gridDim.x);
    const int t_end = (int)((int64_t)(blockIdx.x + 1) * p.n_tiles / gridDim.x);

    // zero the operand ring once: dY's slack rows must read as 0, the A slack rows as finite numbers
    for (int i = threadIdx.x; i < p.n_stages * p.stage_bytes / 16; i += kThreads)
        reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmDy);
        prefetch_tmap(&tmW);
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        int n_img = t_begin / p.tiles_h, th = t_begin - n_img * p.tiles_h;
        for (int t = t_begin; t < t_end; ++t) {
            const int h0 = th * p.TH;
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one()) {
                uint8_t* st = ring + s * p.stage_bytes;
                mbar_expect_tx(&full[s], (uint32_t)p.tx_bytes);
#pragma unroll
                for (int cb = 0; cb < 4; ++cb)
                    tma_load_4d(st + cb * p.a_cb_bytes, &tmX, &full[s], (cblk0 + cb) * 32, -p.pad, h0 + kh - p.pad, n_img);
#pragma unroll
                for (int b = 0; b < BN / 32; ++b)
                    tma_load_4d(st + 4 * p.a_cb_bytes + b * p.b_box_bytes, &tmDy, &full[s], oc0 + b * 32, 0, h0, n_img);
            }
            __syncwarp();
            if (++th == p.tiles_h) { th = 0; ++n_img; }
            if (++s == S) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, 1, 1);
        const uint32_t hi = desc_hi(512, kLayoutSW128Base32B);
        const uint32_t ring_u32 = smem_u32(ring);
        int s = 0;
        uint32_t ph = 0, fresh = 0;
        for (int t = t_begin; t < t_end; ++t) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = desc_lo(ring_u32 + s * p.stage_bytes, (uint32_t)p.a_cb_bytes);
                const uint32_t b_lo = desc_lo(ring_u32 + s * p.stage_bytes + 4 * p.a_cb_bytes, (uint32_t)p.b_box_bytes);
                for (int kw = 0; kw < p.KW; ++kw) {
                    const uint32_t tacc = tmem_base + (uint32_t)(kw * BN);
                    for (int k8 = 0; k8 < p.K8; ++k8)          // K = 8 pixels = 8 rows of 128 B = 1024 B of both operands
                        umma_tf32_split(tacc, a_lo + (uint32_t)(kw * 8 + k8 * 64), hi, b_lo + (uint32_t)(k8 * 64), hi, idesc,
                                        k8 ? 1u : fresh);
                }
                umma_commit(&empty[s]);
            }
            __syncwarp();
            fresh = 1;
            if (++s == S) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tmem_full);
        __syncwarp();
    } else {
        // warp q reads TMEM lanes 32q..32q+31 = input channels (cblk0 + q)*32 + lane of the pass
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (t_end > t_begin) {
            float* stage = reinterpret_cast<float*>(ring + q * 4096);    // the ring is idle: [32 oc][32 c] per warp
            for (int kw = 0; kw < p.KW; ++kw) {
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kw * BN + c0), v);
                    tmem_wait_ld();
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) stage[j * 32 + lane] = __uint_as_float(v[j]);   // row = oc, column = c
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&tmW, stage, (kh * p.KW + kw) * p.C + (cblk0 + q) * 32, oc0 + c0);
                        bulk_commit();
                    }
                }
            }
            if (lane == 0) bulk_wait_all<0>();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// dw must hold the value to accumulate onto (zeros for "="). XSB_ERR_UNSUPPORTED outside the envelope.
static int wgrad_strip(const float* dy, const float* x, float* dw, const ConvGeom& g, cudaStream_t s) {
    constexpr int BN = 128;
    if (g.sh != 1 || g.sw != 1 || g.C % 128 != 0 || g.OC % BN != 0 || g.KW < 1 || g.KW * BN > 512 || g.KH > 16)
        return XSB_ERR_UNSUPPORTED;
    const int WX = g.W + 2 * g.pad;
    if (WX != g.OW + g.KW - 1 || WX > 64 || WX < 1) return XSB_ERR_UNSUPPORTED;
    // rows per strip: K = TH*WX <= 64 pixels, little padding to the next multiple of 8, few rows past the image bottom
    int TH = 0;
    double best = -1.0;
    for (int th = 1; th * WX <= 64 && th <= g.OH; ++th) {
        const int k8 = (th * WX + 7) / 8, tiles = (g.OH + th - 1) / th;
        const double eff = (double)(th * WX) / (k8 * 8) * g.OH / ((double)tiles * th);
        if (eff > best + 1e-9 || (eff > best - 1e-9 && th > TH)) { best = eff; TH = th; }
    }
    if (TH == 0) return XSB_ERR_UNSUPPORTED;
    const int K8 = (TH * WX + 7) / 8;
    const int tiles_h = (g.OH + TH - 1) / TH, n_tiles = g.N * tiles_h;
    const int n_c_tiles = g.C / 128, n_oc_tiles = g.OC / BN;
    const int passes = g.KH * n_c_tiles * n_oc_tiles;
    const int S = xsb_sm_count_cached();
    if (passes > S) return XSB_ERR_UNSUPPORTED;
    int ctas = S / passes;
    if (ctas > n_tiles) ctas = n_tiles;
    if (ctas < 1) return XSB_ERR_UNSUPPORTED;
    const int a_cb_bytes = (K8 * 8 + 8) * A_ROW_BYTES;        // + 8 slack rows: the views of kw > 0 and the K padding
    const int b_box_bytes = K8 * 8 * A_ROW_BYTES;
    const int stage_bytes = 4 * a_cb_bytes + (BN / 32) * b_box_bytes;
    int stages = (227 * 1024 - 1024 - 512) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2 || stages * stage_bytes < 4 * 4096) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    CUtensorMap tx, tdy, tw;
    cuuint64_t xd[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t xs_[3] = {(cuuint64_t)g.C * 4, (cuuint64_t)g.W * g.C * 4, (cuuint64_t)g.H * g.W * g.C * 4};
    cuuint32_t xb[4] = {32, (cuuint32_t)WX, (cuuint32_t)TH, 1};
    cuuint64_t dd[4] = {(cuuint64_t)g.OC, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
    cuuint64_t ds[3] = {(cuuint64_t)g.OC * 4, (cuuint64_t)g.OW * g.OC * 4, (cuuint64_t)g.OH * g.OW * g.OC * 4};
    cuuint32_t db[4] = {32, (cuuint32_t)WX, (cuuint32_t)TH, 1};
    if (!make_map_tiled(&tx, x, 4, xd, xs_, xb, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map_tiled(&tdy, dy, 4, dd, ds, db, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map_2d(&tw, dw, (uint64_t)g.OC, (uint64_t)g.KH * g.KW * g.C, 32, CU_TENSOR_MAP_SWIZZLE_NONE))
        return XSB_ERR_UNSUPPORTED;
    WgradStripParams p{n_tiles, tiles_h, TH, WX, K8, g.KH, g.KW, g.pad, g.C, n_c_tiles, n_oc_tiles,
                       a_cb_bytes, b_box_bytes, stage_bytes, stages, (4 + BN / 32) * TH * WX * A_ROW_BYTES};
    const int smem = 1024 + stages * stage_bytes + 512;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(conv_wgrad_strip_k<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    conv_wgrad_strip_k<BN><<<dim3(ctas, passes), kThreads, smem, s>>>(tx, tdy, tw, p);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// Stem (C <= 4): xp is the zero-bordered 4-channel copy of the input ([N, Hp, Wp, 4], see stem_pad_c4_k); a row of the
// A operand is the 32 floats (8 pixels x 4 channels) starting at pixel (oh*sh + kh, ow*sw): the tensor map walks
// OVERLAPPING rows {32 floats, OW (stride sw pixels), Hp, N}. dw must already hold the value to accumulate onto.
static int wgrad_halo_stem(const float* dy, const float* xp, float* dw, const ConvGeom& g, int Hp, int Wp, cudaStream_t s) {
    if (g.OC != 64 || g.KH > 8 || g.KW > 8 || g.C > 4) return XSB_ERR_UNSUPPORTED;
    constexpr int BN = 64;
    const int groups = (g.KH + 3) / 4;
    const int tiles_w = (g.OW + WH_TW - 1) / WH_TW, tiles_h = (g.OH + WH_TH - 1) / WH_TH;
    if ((int64_t)g.OH * g.OW * 10 < (int64_t)tiles_w * tiles_h * WH_TW * WH_TH * 7) return XSB_ERR_UNSUPPORTED;
    const int n_tiles = g.N * tiles_w * tiles_h;
    int ctas = xsb_sm_count_cached();
    if (ctas > n_tiles) ctas = n_tiles;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    const int a_rows = (WH_TH - 1) * g.sh + g.KH;                 // rows the halo load brings
    const int a_slot_rows = (WH_TH - 1) * g.sh + 4 * groups;      // rows the M blocks kh < 4*groups may touch
    if (a_rows > 256) return XSB_ERR_UNSUPPORTED;
    const int a_slot_bytes = a_slot_rows * WH_TW * A_ROW_BYTES;
    const int b_bytes = (BN / 32) * WH_TW * WH_TH * A_ROW_BYTES;
    const int sb_slots = 3;
    int sa_slots = (227 * 1024 - 1024 - 512 - sb_slots * b_bytes) / a_slot_bytes;
    if (sa_slots > 16) sa_slots = 16;
    if (sa_slots < 3) return XSB_ERR_UNSUPPORTED;
    CUtensorMap tx, tdy, tw;
    cuuint64_t xd[4] = {32, (cuuint64_t)g.OW, (cuuint64_t)Hp, (cuuint64_t)g.N};
    cuuint64_t xs_[3] = {(cuuint64_t)g.sw * 16, (cuuint64_t)Wp * 16, (cuuint64_t)Hp * Wp * 16};
    cuuint32_t xb[4] = {32, WH_TW, (cuuint32_t)a_rows, 1};
    cuuint64_t dd[4] = {(cuuint64_t)g.OC, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
    cuuint64_t ds[3] = {(cuuint64_t)g.OC * 4, (cuuint64_t)g.OW * g.OC * 4, (cuuint64_t)g.OH * g.OW * g.OC * 4};
    cuuint32_t db[4] = {32, WH_TW, WH_TH, 1};
    if (!make_map_tiled(&tx, xp, 4, xd, xs_, xb, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map_tiled(&tdy, dy, 4, dd, ds, db, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return XSB_ERR_UNSUPPORTED;
    tw = tdy;   // unused by the stem epilogue
    WgradHaloParams p{g.N, g.OH, g.OW, tiles_w, tiles_h, n_tiles, g.KH, 1, 0, 1, 1, 1,
                      sa_slots, sb_slots, a_slot_bytes, 4, g.sh, groups, 1, 0, 0, 1, g.KW, g.C, dw, 0};
    const int smem = 1024 + sa_slots * a_slot_bytes + sb_slots * b_bytes + 512;
    auto go = [&](auto kern) -> int {
        XSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        kern<<<dim3(ctas, 1), kThreads, smem, s>>>(tx, tdy, tw, p);
        return XSB_OK;
    };
    int rc;
    if (g.sh == 1 && groups == 1) rc = go(conv_wgrad_halo_k<64, 1, 1>);
    else if (g.sh == 1 && groups == 2) rc = go(conv_wgrad_halo_k<64, 1, 2>);
    else if (g.sh == 2 && groups == 1) rc = go(conv_wgrad_halo_k<64, 2, 1>);
    else if (g.sh == 2 && groups == 2) rc = go(conv_wgrad_halo_k<64, 2, 2>);
    else return XSB_ERR_UNSUPPORTED;
    if (rc != XSB_OK) return rc;
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

}  // namespace tc
