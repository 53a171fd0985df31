// tc_mma_rate.cu — issue rate of small tcgen05.mma kind::tf32 (M=128, K=8) as a function of N and of the
// A operand's source (shared memory descriptor vs TMEM): cycles per MMA from clock64 around NREP back-to-back MMAs.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// mode 0: A smem, 1: A tmem ; kind 0: tf32 (K=8), 1: bf16 (K=16)
__global__ void __launch_bounds__(128) k_rate(int N, int mode, int kind, int nrep, int nacc, long long *out)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid / 32;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) ((float *)smem)[i] = 0.001f * (i % 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | ((kind ? 1u : 2u) << 7) | ((kind ? 1u : 2u) << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t ad = desc_k128(smem_u32(smem)), bd = desc_k128(smem_u32(smem + 16384));
        const uint32_t a_t = tmem + 480; // 32 columns at the top
        long long t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < nrep; r += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint32_t d = tmem + (u % nacc) * N; // nacc independent accumulators
                    if (mode == 0) {
                        if (kind)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                    } else {
                        if (kind)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                    }
                }
            }
        }
        __syncwarp();
        long long t1 = clock64();
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (tid == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main()
{
    long long *d, h[2];
    CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int nrep = 2000;
    printf("kind  A-src  N  nacc : issue cyc/MMA   complete cyc/MMA   (floor N/2)\n");
    for (int kind = 0; kind < 2; kind++)
        for (int mode = 0; mode < 2; mode++)
            for (int N : {16, 32, 64, 96, 128, 256})
                for (int nacc : {1, 4}) {
                    if (nacc * N > 448) continue;
                    k_rate<<<1, 128, 64 * 1024>>>(N, mode, kind, nrep, nacc, d);
                    CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                    printf("%s  %s  %3d  %d : %8.1f %8.1f   (%d)\n", kind ? "bf16" : "tf32", mode ? "tmem" : "smem", N, nacc, (double)h[0] / nrep,
                           (double)h[1] / nrep, N / 2);
                }
    return 0;
}
