// Micro-benchmark: rate of tcgen05.mma.kind::tf32 (cta_group::1, M=128, K=8) when an operand is MN-MAJOR
// (SWIZZLE_128B_BASE32B descriptors, the only layout tf32 MN-major operands may use) against the K-major rate:
// the weight-gradient kernels feed both operands MN-major (pixels = K). Operands static in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate_mn tools/mma_rate_mn.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// kind 0: tf32 (a/b fmt 2), kind 1: bf16 (a/b fmt 1)
// kind here: bit 0 = A MN-major, bit 1 = B MN-major (always tf32)
__host__ __device__ constexpr uint32_t make_idesc(int n, int kind) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kind & 1) << 15) | ((uint32_t)((kind >> 1) & 1) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int KIND>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

template <int KIND, int N, int NACC, int ADV>
__global__ void __launch_bounds__(128, 1) rate_k(long long* out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    // A: 128 rows x 128 B (16 KB) x 2 ; B: 256 rows x 128 B (32 KB) x 2
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc(N, KIND);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int st = ADV ? (i & 1) : 0;
            // MN-major: 32-element blocks of 32 k-rows x 128 B (4 KB, LBO), 4-row swizzle atoms (SBO 512), K advance 1 KB
            const uint64_t da = (KIND & 1) ? make_desc(a0 + st * 16384, 4096, 512, 1) : make_desc(a0 + st * 16384, 16, 1024);
            const uint64_t db = (KIND & 2) ? make_desc(b0 + st * 32768, 4096, 512, 1) : make_desc(b0 + st * 32768, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma<KIND>(tmem + (uint32_t)((i % NACC) * N), da + (uint64_t)(k * ((KIND & 1) ? 64 : 2)),
                           db + (uint64_t)(k * ((KIND & 2) ? 64 : 2)), idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile(
            "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)),
            "r"(0)
            : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int KIND, int N, int NACC, int ADV>
void run(const char* name, int grid) {
    long long* d;
    cudaMalloc(&d, 8);
    const int iters = 4000, smem = 97 * 1024 + 1024;
    cudaFuncSetAttribute(rate_k<KIND, N, NACC, ADV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_k<KIND, N, NACC, ADV><<<grid, 128, smem>>>(d, iters);
    rate_k<KIND, N, NACC, ADV><<<grid, 128, smem>>>(d, iters);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 4);
    const double k = 8;
    printf("%-28s grid %3d N=%3d acc=%d adv=%d : %7.1f clk/MMA  -> %6.0f MAC/clk/SM  (%s)\n", name, grid, N, NACC, ADV, per,
           128.0 * N * k / per, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<0, 64, 1, 0>("A K-major,  B K-major", grid);
        run<1, 64, 1, 0>("A MN-major, B K-major", grid);
        run<2, 64, 1, 0>("A K-major,  B MN-major", grid);
        run<3, 64, 1, 0>("A MN-major, B MN-major", grid);
        run<0, 128, 1, 0>("A K-major,  B K-major", grid);
        run<1, 128, 1, 0>("A MN-major, B K-major", grid);
        run<2, 128, 1, 0>("A K-major,  B MN-major", grid);
        run<3, 128, 1, 0>("A MN-major, B MN-major", grid);
        run<0, 256, 1, 0>("A K-major,  B K-major", grid);
        run<3, 256, 1, 0>("A MN-major, B MN-major", grid);
        run<3, 64, 2, 1>("MN/MN 2 acc, 2 stages", grid);
        run<3, 128, 2, 1>("MN/MN 2 acc, 2 stages", grid);
    }
    return 0;
}
