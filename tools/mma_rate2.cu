// Micro-benchmark 2: what slows tcgen05.mma (kind::tf32, M=128, N=64/128/256) down inside a real conv kernel?
// Same issue loop as mma_rate.cu plus optional concurrent activity in the same CTA:
//   DATA : random (non-zero) operands            BULK : a producer warp streams cp.async.bulk global->smem (24 KB / 4 MMAs worth)
//   STG  : 4 warps store float4 rows strided by 256 B to global (the conv epilogue pattern)
//   LDTM : 4 warps read the other accumulator with tcgen05.ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

enum { DATA = 1, BULK = 2, STG = 4, LDTM = 8, COMMIT8 = 16, WAIT8 = 32, ACC6 = 64, COMMIT4 = 128, NOFENCE = 256, NOWAIT = 512, WAIT16 = 1024, LDSPOLL = 2048, DUAL = 4096 };

template <int N, int FLAGS>
__global__ void __launch_bounds__(192, 1) rate_k(long long* out, int iters, const float* gsrc, float* gdst) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2, cbar[2];
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) {
        uint32_t h = (i * 2654435761u) ^ (i >> 3);
        reinterpret_cast<float*>(smem)[i] = (FLAGS & DATA) ? ((float)(h & 0xFFFF) / 65536.f - 0.5f) : 0.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 1 || ((FLAGS & DUAL) && warp == 2)) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(N);
            const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
            long long t0 = clock64();
            __shared__ uint64_t dummy[2];
            if ((FLAGS & (COMMIT8 | COMMIT4 | WAIT8)) && warp == 2) { for (volatile int z = 0; z < 2000; ++z) {} }
            if ((FLAGS & (COMMIT8 | COMMIT4 | WAIT8)) && warp == 1) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[0])) : "memory");
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[1])) : "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&dummy[1])) : "memory");   // phase 0 of [1] complete
            }
            for (int i = 0; i < iters; ++i) {
                const int st = i & 1;
                const uint64_t da = make_desc(a0 + st * 16384, 16, 1024);
                const uint64_t db = make_desc(b0 + st * 32768, 16, 1024);
                if ((FLAGS & WAIT8) && ((FLAGS & WAIT16) ? (i & 3) == 0 : (i & 1) == 0)) {
                    if (FLAGS & LDSPOLL) {
                        uint32_t v;
                        do {
                            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&dummy[1])) : "memory");
                        } while (v == 0xFFFFFFFFu);
                    } else if (!(FLAGS & NOWAIT)) mbar_wait(&dummy[1], 0);     // already complete: pure try_wait latency
                    if (!(FLAGS & NOFENCE)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t acc = (FLAGS & ACC6) ? tmem + (uint32_t)(((i >> 1) % 6) * 64) : tmem;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1);
                if (((FLAGS & COMMIT8) && (i & 1) == 1) || (FLAGS & COMMIT4))
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[0])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(warp == 1 ? &bar : &bar2)) : "memory");
            mbar_wait(warp == 1 ? &bar : &bar2, 0);
            long long t1 = clock64();
            if (warp == 1) stop = 1;
            if (blockIdx.x == 0 && warp == 1) out[0] = t1 - t0;
        }
    } else if (warp == 0) {
        if ((FLAGS & BULK) && lane == 0) {
            // stream 24 KB chunks into a scratch smem region (offset 96 KB.. not used by the MMAs), 2 in flight
            const uint8_t* src = reinterpret_cast<const uint8_t*>(gsrc) + (size_t)blockIdx.x * (1 << 20);
            uint8_t* dst = smem + 96 * 1024;
            int it = 0;
            while (!stop) {
                const int s = it & 1;
                if (it >= 2) mbar_wait(&cbar[s], ((it >> 1) - 1) & 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[s])), "r"(24576) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(dst + s * 24576)),
                             "l"(src + (size_t)(it & 31) * 24576), "r"(24576), "r"(smem_u32(&cbar[s]))
                             : "memory");
                ++it;
            }
            // drain
            if (it >= 1) mbar_wait(&cbar[(it - 1) & 1], ((it - 1) >> 1) & 1);
            if (it >= 2) mbar_wait(&cbar[(it - 2) & 1], ((it - 2) >> 1) & 1);
        }
    } else if (!((FLAGS & DUAL) && warp == 2)) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        float* orow = gdst + ((size_t)blockIdx.x * 4096 + row) * 64;
        int it = 0;
        while (!stop) {
            if (FLAGS & LDTM) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(tmem + ((uint32_t)(q * 32) << 16) + 256u));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (r[0] == 0x12345678u) orow[0] = 1.f;
            }
            if (FLAGS & STG) {
                float* o = orow + (size_t)(it & 31) * 128 * 64;
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(1.f, 2.f, 3.f, (float)it);
            }
            ++it;
            if (!(FLAGS & (STG | LDTM))) break;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

static float* g_src;
static float* g_dst;

template <int N, int FLAGS>
void run(const char* name) {
    long long* d;
    cudaMalloc(&d, 8);
    const int iters = 4000, smem = 96 * 1024 + 48 * 1024 + 1024;
    cudaFuncSetAttribute(rate_k<N, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_k<N, FLAGS><<<148, 192, smem>>>(d, iters, g_src, g_dst);
    rate_k<N, FLAGS><<<148, 192, smem>>>(d, iters, g_src, g_dst);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 4);
    printf("N=%3d %-34s : %7.1f clk/MMA -> %6.0f MAC/clk/SM (%s)\n", N, name, per, 128.0 * N * 8 / per, cudaGetErrorString(e));
    cudaFree(d);
}

template <int N>
void sweep() {
    run<N, 0>("zeros");
    run<N, DATA>("random data");
    run<N, DATA | BULK>("data + bulk copies into smem");
    run<N, DATA | STG>("data + strided STG.128");
    run<N, DATA | LDTM>("data + tcgen05.ld");
    run<N, DATA | BULK | STG | LDTM>("data + bulk + STG + ld");
    run<N, DATA | COMMIT8>("commit every 8 MMAs");
    run<N, DATA | COMMIT4>("commit every 4 MMAs");
    run<N, DATA | COMMIT8 | WAIT8>("try_wait + commit every 8");
    run<N, DATA | COMMIT8 | WAIT8 | ACC6>("try_wait + commit /8, 6 accumulators");
    run<N, DATA | COMMIT8 | WAIT8 | ACC6 | BULK>("same + bulk copies");
    run<N, DATA | COMMIT8 | WAIT8 | NOFENCE>("try_wait (no fence) + commit /8");
    run<N, DATA | COMMIT8 | WAIT8 | NOWAIT>("fence only + commit /8");
    run<N, DATA | COMMIT8 | WAIT8 | WAIT16>("try_wait+fence every 16, commit /8");
    run<N, DATA | WAIT8 | NOFENCE>("try_wait every 8, no commit");
    run<N, DATA | COMMIT8 | WAIT8 | LDSPOLL>("ld.shared poll + fence + commit /8");
    run<N, DATA | COMMIT8 | WAIT8 | DUAL>("DUAL issuers: try_wait+commit /8 each");
    run<N, DATA | COMMIT4 | WAIT8 | DUAL>("DUAL issuers: try_wait/8 commit/4");
}

int main() {
    cudaMalloc(&g_src, (size_t)148 << 20);
    cudaMemset(g_src, 0, (size_t)148 << 20);
    cudaMalloc(&g_dst, (size_t)148 * 4096 * 64 * 4);
    sweep<64>();
    sweep<128>();
    return 0;
}
