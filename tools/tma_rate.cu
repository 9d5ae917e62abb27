// Micro-benchmark: per-SM rate of TMA loads into shared memory when the same 16 KB arrive as
//   (a) a tiled 2-D box {32 floats x 128 rows} (128-byte rows, SWIZZLE_128B: what every operand tile of the conv kernels is),
//   (b) one linear cp.async.bulk of 16 KB.
// One producer thread per CTA keeps 4 loads in flight; all 148 SMs run at once. Prints clk per 16 KB and B/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_rate tools/tma_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int MODE, int NF>
__global__ void __launch_bounds__(32, 1) rate_k(const __grid_constant__ CUtensorMap tm, const float* src, long long* out, int iters,
                                               int rows_total) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[NF];
    if (threadIdx.x == 0) {
        for (int i = 0; i < NF; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters + NF; ++it) {
            const int s = it % NF;
            if (it >= NF) mbar_wait(&bar[s], ((it / NF) - 1) & 1);
            if (it < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(16384) : "memory");
                const int row0 = (int)(((long long)blockIdx.x * 4099 + (long long)it * 128) % (rows_total - 128));
                if (MODE == 0) {
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                            smem_u32(smem + s * 16384)),
                        "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar[s])), "r"(0), "r"(row0)
                        : "memory");
                } else {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(smem + s * 16384)),
                                 "l"(src + (size_t)row0 * 32), "r"(16384), "r"(smem_u32(&bar[s]))
                                 : "memory");
                }
            }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
}

int main() {
    const int rows = 1 << 20, pitch = 2304;     // a [rows][2304] fp32 matrix for the 2-D case (rows 9216 B apart), 9.7 GB is too much:
    const int rows2d = 1 << 16;                 // 64 Ki rows x 9216 B = 604 MB
    float *m2d, *lin;
    cudaMalloc(&m2d, (size_t)rows2d * pitch * 4);
    cudaMalloc(&lin, (size_t)rows * 32 * 4);    // 128 MB linear
    cudaMemset(m2d, 0, (size_t)rows2d * pitch * 4);
    cudaMemset(lin, 0, (size_t)rows * 32 * 4);
    long long* d;
    cudaMalloc(&d, 8);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows2d};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, m2d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    const int iters = 2000;
    auto run = [&](auto kern, int nf, int grid, int mode) {
        const int smem = nf * 16384 + 1024;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int rep = 0; rep < 2; ++rep) kern<<<grid, 32, smem>>>(tm, lin, d, iters, mode == 0 ? rows2d : rows);
        long long h = 0;
        cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("grid %3d in flight %2d x 16 KB %-32s : %8.1f clk per 16 KB -> %6.1f B/clk/SM (%s)\n", grid, nf,
               mode == 0 ? "2-D box {32 x 128 rows}, SW128" : "linear cp.async.bulk 16 KB", (double)h / iters,
               16384.0 * iters / (double)h, cudaGetErrorString(e));
    };
    for (int grid : {1, 148}) {
        run(rate_k<0, 2>, 2, grid, 0);
        run(rate_k<0, 4>, 4, grid, 0);
        run(rate_k<0, 8>, 8, grid, 0);
        run(rate_k<0, 12>, 12, grid, 0);
        run(rate_k<1, 4>, 4, grid, 1);
        run(rate_k<1, 12>, 12, grid, 1);
    }
    return 0;
}
