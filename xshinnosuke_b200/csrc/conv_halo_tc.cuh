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
// 64->64 3x3: 6 loads (2 channel blocks x 3 filter columns, the column shift is the TMA start coordinate, zero padding
// by out-of-bounds fill) x 3 k-blocks. Stem (C <= 4): the image is pre-padded to 4 channels (stem_pad_c4_k) and the
// tensor map walks OVERLAPPING rows {32 floats = 8 pixels x 4 ch, stride sw pixels}: 1 load x KH k-blocks.
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_MAX_LOADS = 16, HALO_MAX_KH = 8;

struct HaloParams {
    float* out;              // [N, OH, OW, ldo]
    const float* bias;       // [ldo] or null
    int N, OH, OW, ldo;
    int tiles_w, tiles_h, n_tiles;
    int n_loads, KH, sh;
    int h_off;               // TMA row coordinate of output row 0, filter row 0 (= -pad, or 0 for a pre-padded image)
    int stage_bytes, n_stages, n_btiles;
    int accumulate;
    int load_c[HALO_MAX_LOADS];       // TMA coordinate 0 (k-element offset) of load l
    int load_wofs[HALO_MAX_LOADS];    // TMA coordinate 1 offset (filter column - pad) of load l
    uint8_t btile[HALO_MAX_LOADS][HALO_MAX_KH];   // resident filter tile multiplied with (load l, filter row kh)
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int B_TILE = BN * A_ROW_BYTES;
    constexpr int MAXS = 8;
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sB = smem;
    uint8_t* sA = smem + p.n_btiles * B_TILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(sA + p.n_stages * p.stage_bytes);
    uint64_t* empty = full + MAXS;
    uint64_t* bfull = empty + MAXS;
    uint64_t* tmem_full = bfull + 1;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.n_stages;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
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
            const int w0 = tw * HALO_TW, h0 = th * HALO_TH * p.sh + p.h_off;
            for (int l = 0; l < p.n_loads; ++l) {
                mbar_wait(&empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[s], p.stage_bytes);
                    tma_load_4d(sA + s * p.stage_bytes, &tmA, &full[s], p.load_c[l], w0 + p.load_wofs[l], h0, n_img);
                }
                if (++s == S) { s = 0; ph ^= 1; }
            }
            // next tile of this CTA: (n_img, th, tw) += gridDim.x tiles, without divisions
            rem += step_rem;
            n_img += step_img;
            if (rem >= tiles_per_img) { rem -= tiles_per_img; ++n_img; }
            th = rem / p.tiles_w;
            tw = rem - th * p.tiles_w;
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, 0, 0);
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        const uint32_t a_sbo = (uint32_t)p.sh * 1024;
        mbar_wait(bfull, 0);
        int s = 0, buf = 0;
        uint32_t ph = 0, tph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[buf], tph ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
            for (int l = 0; l < p.n_loads; ++l) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = a_base + s * p.stage_bytes;
                    for (int kh = 0; kh < p.KH; ++kh) {
                        const uint64_t da = make_desc(a0 + kh * 1024, 16, a_sbo);
                        const uint64_t db = make_desc(b_base + p.btile[l][kh] * B_TILE, 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_tf32(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (l | kh | k) != 0);
                    }
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&tmem_full[buf]);
            __syncwarp();
            buf ^= 1;
            if (buf == 0) tph ^= 1;
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int r = row / HALO_TW, jw = row % HALO_TW;
        int buf = 0;
        uint32_t tph = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int n_img = tile / tiles_per_img, rem = tile - n_img * tiles_per_img;
            const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
            const int oh = th * HALO_TH + r, ow = tw * HALO_TW + jw;
            const bool valid = oh < p.OH && ow < p.OW;
            float* orow = p.out + (((int64_t)n_img * p.OH + oh) * p.OW + ow) * p.ldo;
            mbar_wait(&tmem_full[buf], tph);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), v);
                tmem_wait_ld();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                               __uint_as_float(v[j + 3]));
                        if (p.bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        float4* dst = reinterpret_cast<float4*>(orow + c0 + j);
                        if (p.accumulate) {
                            const float4 e = *dst;
                            o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                        }
                        *dst = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            buf ^= 1;
            if (buf == 0) tph ^= 1;
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
    const int fixed = 1024 + n_btiles * BN * A_ROW_BYTES + 256;
    int s = (HALO_SMEM_MAX - fixed) / stage_bytes;
    if (s > 6) s = 6;
    return s >= 2 ? s : 0;
}

template <int BN>
static int launch_halo(const CUtensorMap& a, const CUtensorMap& b, const HaloParams& p, cudaStream_t s) {
    const int smem = 1024 + p.n_btiles * BN * A_ROW_BYTES + p.n_stages * p.stage_bytes + 256;
    static int configured = 0;
    if (configured < smem) {
        XSB_CUDA(cudaFuncSetAttribute(conv_halo_tc_k<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX));
        configured = HALO_SMEM_MAX;
    }
    int grid = xsb_sm_count_cached();
    if (grid > p.n_tiles) grid = p.n_tiles;
    conv_halo_tc_k<BN><<<grid, kThreads, smem, s>>>(a, b, p);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// Generic stride-1 path: conv of channels-last in[N,H,W,Cx] with flt[OCx, KH*KW*Cx] -> out[N,OH,OW,OCx].
// Returns XSB_ERR_UNSUPPORTED when the shape is outside the envelope (caller runs the im2col kernel).
static int halo_conv_s1(const float* in, int N, int H, int W, int Cx, const float* flt, int OCx, int KH, int KW, int pad_h,
                        int pad_w, int OH, int OW, const float* bias, float* out, int accumulate, cudaStream_t s) {
    const int cblocks = Cx / 32;
    if (Cx % 32 != 0 || (OCx != 64 && OCx != 128) || KH > HALO_MAX_KH || cblocks * KW > HALO_MAX_LOADS) return XSB_ERR_UNSUPPORTED;
    const int n_btiles = KH * KW * cblocks;
    if (n_btiles > 255) return XSB_ERR_UNSUPPORTED;
    const int rows = HALO_TH - 1 + KH;
    const int stage_bytes = rows * 1024;
    const int stages = halo_stages(n_btiles, OCx, stage_bytes);
    if (stages == 0 || rows > 256) return XSB_ERR_UNSUPPORTED;
    // small images waste most of an 8x16 tile: leave them to the im2col kernel
    const int tiles_w = (OW + HALO_TW - 1) / HALO_TW, tiles_h = (OH + HALO_TH - 1) / HALO_TH;
    if ((int64_t)OH * OW * 10 < (int64_t)tiles_w * tiles_h * HALO_TW * HALO_TH * 8) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    CUtensorMap ta, tb;
    cuuint64_t dims[4] = {(cuuint64_t)Cx, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cx * 4, (cuuint64_t)W * Cx * 4, (cuuint64_t)H * W * Cx * 4};
    cuuint32_t box[4] = {32, HALO_TW, (cuuint32_t)rows, 1};
    if (!make_map_tiled(&ta, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_map_2d(&tb, flt, (uint64_t)OCx, (uint64_t)KH * KW * Cx, (uint32_t)OCx))
        return XSB_ERR_UNSUPPORTED;
    HaloParams p{};
    p.out = out; p.bias = bias; p.N = N; p.OH = OH; p.OW = OW; p.ldo = OCx;
    p.tiles_w = tiles_w; p.tiles_h = tiles_h; p.n_tiles = N * tiles_w * tiles_h;
    p.n_loads = cblocks * KW; p.KH = KH; p.sh = 1; p.h_off = -pad_h;
    p.stage_bytes = stage_bytes; p.n_stages = stages; p.n_btiles = n_btiles; p.accumulate = accumulate;
    for (int cb = 0; cb < cblocks; ++cb)
        for (int kw = 0; kw < KW; ++kw) {
            const int l = cb * KW + kw;
            p.load_c[l] = cb * 32;
            p.load_wofs[l] = kw - pad_w;
            for (int kh = 0; kh < KH; ++kh) p.btile[l][kh] = (uint8_t)((kh * KW + kw) * cblocks + cb);
        }
    return OCx == 64 ? launch_halo<64>(ta, tb, p, s) : launch_halo<128>(ta, tb, p, s);
}

}  // namespace tc
