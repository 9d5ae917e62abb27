// tcgen05 GEMM for the Dense layers (XSB_MATH_TF32): C[M,N] (+)= op(A) . op(B) (+ bias[N]), fp32 row-major operands fed
// to kind::tf32 as they are. Reference arithmetic: xs/nn/functional.py:81-120 (mm / addmm: y = x . W + b with W stored
// [in, out], xs/layers/linear.py:32) and xs/nn/grad_fn.py:73-99 (dX = dY . W^T, dW = X^T . dY).
//
// The three products of a Dense layer differ only in which operand is contiguous along K:
//   forward  y  = x[M,K]  . W[K,N]      A K-major,  B MN-major (n contiguous)
//   dX          = dY[M,K] . W[N,K]^T    A K-major,  B K-major
//   dW          = x[K,M]^T . dY[K,N]    A MN-major, B MN-major
// K-major tiles are one TMA box {32 k, 128 rows} (SWIZZLE_128B, MMA K-advance 32 B); MN-major tiles are four boxes
// {32 rows-of-the-other-dim, 32 k} (SWIZZLE_128B_ATOM_32B, descriptor LBO = 4 KB between the 32-wide blocks, K-advance
// 1 KB) -- the same two operand forms the convolution kernels use. Out-of-range rows / columns / k are zero-filled by
// the tensor maps, so M, N, K need not be multiples of the tile; the epilogue leaves through clipped TMA stores
// (reduce-adds for "+=" and for split-K). Row pitches must be multiples of 16 bytes (Dense(10): 40-byte rows -> the
// caller runs the FFMA kernel).
#pragma once
#include "tc_common.cuh"

namespace tc {

struct GemmParams {
    const float* bias;     // [N] or null (added by split 0 only)
    int M, N, K;
    int kb_per_split;      // 32-wide k-blocks per grid.z slice
    int reduce;            // 1: TMA reduce-add into C (accumulate or split-K); 0: plain store
    int act;               // 1: fused activation -- also store max(0, C) through tmAct (single split, plain store only)
};

template <bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAct, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int BN = 128;
    constexpr int TILE = BM * A_ROW_BYTES;      // 16 KB per operand and stage
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * TILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * TILE);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int total_kb = (p.K + 31) / 32;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    int kb_end = kb_begin + p.kb_per_split;
    if (kb_end > total_kb) kb_end = total_kb;
    const int num_kb = kb_end - kb_begin;       // >= 1 by construction of the grid

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmC);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < num_kb; ++i) {
            const int k0 = (kb_begin + i) * 32;
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full[s], 2 * TILE);
                if (A_MN) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) tma_load_2d(sA + s * TILE + b * 4096, &tmA, &full[s], m0 + b * 32, k0);
                } else {
                    tma_load_2d(sA + s * TILE, &tmA, &full[s], k0, m0);
                }
                if (B_MN) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) tma_load_2d(sB + s * TILE + b * 4096, &tmB, &full[s], n0 + b * 32, k0);
                } else {
                    tma_load_2d(sB + s * TILE, &tmB, &full[s], k0, n0);
                }
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < num_kb; ++i) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da = A_MN ? make_desc(a_base + s * TILE, 4096, 512, kLayoutSW128Base32B)
                                         : make_desc(a_base + s * TILE, 16, 1024);
                const uint64_t db = B_MN ? make_desc(b_base + s * TILE, 4096, 512, kLayoutSW128Base32B)
                                         : make_desc(b_base + s * TILE, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)     // K = 8 per MMA: +32 B in a K-major tile, +8 rows (1 KB) in an MN-major one
                    umma_tf32(tmem_base, da + (uint64_t)(k * (A_MN ? 64 : 2)), db + (uint64_t)(k * (B_MN ? 64 : 2)), idesc,
                              (i | k) != 0);
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
        uint8_t* stage = sA + q * 4096;      // the operand ring is idle
        const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= p.N) break;       // warp-uniform
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_wait_ld();
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                o[j] = __uint_as_float(r[j]);
                if (add_bias && n0 + c0 + j < p.N) o[j] += __ldg(p.bias + n0 + c0 + j);
            }
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            stage_rows_sw128(stage, lane, o);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && m0 + q * 32 < p.M) {
                if (p.reduce)
                    tma_reduce_add_2d(&tmC, stage, n0 + c0, m0 + q * 32);
                else
                    tma_store_2d(&tmC, stage, n0 + c0, m0 + q * 32);
                bulk_commit();
            }
            if (p.act) {    // Dense(..., activation='relu'): the rectified copy leaves from the same registers
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                stage_rows_sw128(stage, lane, o);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && m0 + q * 32 < p.M) {
                    tma_store_2d(&tmAct, stage, n0 + c0, m0 + q * 32);
                    bulk_commit();
                }
            }
        }
        if (lane == 0) bulk_wait_all<0>();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

// matrix [rows, cols] row-major (cols contiguous) as a TMA source/destination with box {32 cols, box_rows}
static bool gemm_map(CUtensorMap* m, const float* base, int rows, int cols, int box_rows, CUtensorMapSwizzle swz) {
    return make_map_2d(m, base, (uint64_t)rows, (uint64_t)cols, (uint32_t)box_rows, swz);
}

template <bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& act,
                       const GemmParams& p, dim3 grid, cudaStream_t s) {
    constexpr int STAGES = 4;
    constexpr int smem = STAGES * 2 * BM * A_ROW_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        XSB_CUDA(cudaFuncSetAttribute(gemm_tc_k<A_MN, B_MN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    gemm_tc_k<A_MN, B_MN, STAGES><<<grid, kThreads, smem, s>>>(a, b, c, act, p);
    XSB_LAUNCH_CHECK();
    return XSB_OK;
}

// C[M,N] (+)= op(A) . op(B) (+ bias). transA: A stored [K,M]; transB: B stored [N,K]. XSB_ERR_UNSUPPORTED when a row
// pitch is not a multiple of 16 bytes / a pointer is unaligned.
// act (nullable): fused activation output, act[M,N] = max(0, C) -- then the whole K range is one split ("=" only).
static int gemm_tf32(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int transA, int transB,
                     int accumulate, cudaStream_t s, float* act = nullptr) {
    const int lda = transA ? M : K, ldb = transB ? K : N;
    if (lda % 4 || ldb % 4 || N % 4 || !aligned16(A) || !aligned16(B) || !aligned16(C)) return XSB_ERR_UNSUPPORTED;
    if (M < 1 || N < 4 || K < 1) return XSB_ERR_UNSUPPORTED;
    if (act && (accumulate || !aligned16(act))) return XSB_ERR_UNSUPPORTED;
    if (!load_driver_entry_points()) return XSB_ERR_UNSUPPORTED;
    CUtensorMap ta, tb, tc_, tact;
    const bool okA = transA ? gemm_map(&ta, A, K, M, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)     // [K rows][M cols]: MN-major
                            : gemm_map(&ta, A, M, K, BM, CU_TENSOR_MAP_SWIZZLE_128B);             // [M rows][K cols]: K-major
    const bool okB = transB ? gemm_map(&tb, B, N, K, 128, CU_TENSOR_MAP_SWIZZLE_128B)             // [N rows][K cols]: K-major
                            : gemm_map(&tb, B, K, N, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);    // [K rows][N cols]: MN-major
    if (!okA || !okB || !gemm_map(&tc_, C, M, N, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return XSB_ERR_UNSUPPORTED;
    tact = tc_;
    if (act && !gemm_map(&tact, act, M, N, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return XSB_ERR_UNSUPPORTED;
    const int m_tiles = (M + BM - 1) / BM, n_tiles = (N + 127) / 128, total_kb = (K + 31) / 32;
    int splits = xsb_sm_count_cached() / (m_tiles * n_tiles);
    if (splits > total_kb / 2) splits = total_kb / 2;
    if (splits < 1 || act) splits = 1;      // an activation needs the finished sum in ONE epilogue
    const int kb_per = (total_kb + splits - 1) / splits;
    splits = (total_kb + kb_per - 1) / kb_per;
    const int reduce = (splits > 1 || accumulate) ? 1 : 0;
    if (splits > 1 && !accumulate) XSB_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), s));
    GemmParams p{bias, M, N, K, kb_per, reduce, act ? 1 : 0};
    dim3 grid(m_tiles, n_tiles, splits);
    if (transA && !transB) return launch_gemm<true, true>(ta, tb, tc_, tact, p, grid, s);
    if (!transA && transB) return launch_gemm<false, false>(ta, tb, tc_, tact, p, grid, s);
    if (!transA && !transB) return launch_gemm<false, true>(ta, tb, tc_, tact, p, grid, s);
    return launch_gemm<true, false>(ta, tb, tc_, tact, p, grid, s);
}

}  // namespace tc
