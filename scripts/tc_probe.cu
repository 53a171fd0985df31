// tc_probe.cu — stand-alone probe for the tcgen05 pieces K4 (the convolution-matrix GEMM) is built from:
// MN-major SWIZZLE_128B tf32 operands written by threads or by TMA, M=128 x N=32 x K=8 UMMAs
// accumulating in TMEM, tcgen05.ld drain, and the numerics of 1xTF32 / 3xTF32 over K ~ 30000.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_probe scripts/tc_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

static constexpr int KC = 32; // K rows per chunk

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int n)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46; // version 1 (Blackwell)
    d |= (uint64_t)2 << 61; // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

struct ProbeArgs {
    const float *At; // [K][128]
    const float *Bt; // [K][32]
    float *D;        // [128][32]
    int K;
    int tma;         // 1: operands staged by TMA (SWIZZLE_128B tensor maps), 0: written by threads
    int split;       // 0: 1xTF32 raw, 1: 3xTF32 hi = raw bits / lo = x - trunc, 2: 3xTF32 hi = rna(x) stored, lo = rna(x - hi)
    int swap;        // swap LBO / SBO in the A descriptor
    int drain_every; // chunks between TMEM drains (0: only at the end)
    int kmajor;      // 1: K-major SWIZZLE_128B operands (thread fill only)
    int prefill;     // 1: tcgen05.st a pattern (lane + col/64) into D first and accumulate onto it
};

__global__ void __launch_bounds__(128) k_probe(ProbeArgs a, const __grid_constant__ CUtensorMap tmA,
                                                const __grid_constant__ CUtensorMap tmB)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float *Ahi = (float *)smem;                 // 4 column blocks x [KC rows][32 floats]
    float *Alo = (float *)(smem + 16384);
    float *Bhi = (float *)(smem + 32768);       // [KC rows][32 floats]
    float *Blo = (float *)(smem + 32768 + 4096);
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid / 32;

    if (tid == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    // instruction descriptor: D f32, A/B tf32, both MN-major, N = 32, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a.kmajor ? 0u : ((1u << 15) | (1u << 16))) | ((32u >> 3) << 17) |
                           ((128u >> 4) << 24);
    if (a.prefill) {
        for (int c = 0; c < 32; c++) {
            float v = (float)tid + (float)c / 64.f;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + c),
                         "r"(__float_as_uint(v))
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0.f;
    uint32_t ph_tma = 0, ph_mma = 0;
    bool fresh = !a.prefill; // next MMA starts a new accumulation
    const int nchunks = a.K / KC;
    for (int ch = 0; ch < nchunks; ch++) {
        const int k0 = ch * KC;
        if (a.tma) {
            if (tid == 0) {
                mbar_expect_tx(&bar_tma, KC * 128 * 4 + KC * 32 * 4);
                if (a.kmajor) {
                    tma_load_2d(Ahi, &tmA, k0, 0, &bar_tma);
                    tma_load_2d(Bhi, &tmB, k0, 0, &bar_tma);
                } else {
                    for (int mc = 0; mc < 4; mc++) tma_load_2d(Ahi + mc * KC * 32, &tmA, mc * 32, k0, &bar_tma);
                    tma_load_2d(Bhi, &tmB, 0, k0, &bar_tma);
                }
            }
            mbar_wait(&bar_tma, ph_tma);
            ph_tma ^= 1;
        } else if (a.kmajor) {
            for (int e = tid; e < KC * 128; e += 128) {
                int k = e / 128, m = e % 128;
                Ahi[m * 32 + (((k / 4) ^ (m & 7)) << 2) + (k % 4)] = a.At[(size_t)(k0 + k) * 128 + m];
            }
            for (int e = tid; e < KC * 32; e += 128) {
                int k = e / 32, n = e % 32;
                Bhi[n * 32 + (((k / 4) ^ (n & 7)) << 2) + (k % 4)] = a.Bt[(size_t)(k0 + k) * 32 + n];
            }
            __syncthreads();
        } else {
            for (int e = tid; e < KC * 128; e += 128) {
                int k = e / 128, m = e % 128;
                int mc = m / 32, c = (m % 32) / 4, t = m % 4;
                Ahi[mc * KC * 32 + k * 32 + ((c ^ (k & 7)) << 2) + t] = a.At[(size_t)(k0 + k) * 128 + m];
            }
            for (int e = tid; e < KC * 32; e += 128) {
                int k = e / 32, n = e % 32;
                int c = n / 4, t = n % 4;
                Bhi[k * 32 + ((c ^ (k & 7)) << 2) + t] = a.Bt[(size_t)(k0 + k) * 32 + n];
            }
            __syncthreads();
        }
        if (a.split) {
            for (int e = tid; e < KC * 128; e += 128) {
                float x = Ahi[e], hi, lo;
                if (a.split == 1) {
                    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
                    lo = x - hi;
                } else {
                    uint32_t h, l;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
                    hi = __uint_as_float(h);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
                    lo = __uint_as_float(l);
                    Ahi[e] = hi;
                }
                Alo[e] = lo;
            }
            for (int e = tid; e < KC * 32; e += 128) {
                float x = Bhi[e], hi, lo;
                if (a.split == 1) {
                    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
                    lo = x - hi;
                } else {
                    uint32_t h, l;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
                    hi = __uint_as_float(h);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
                    lo = __uint_as_float(l);
                    Bhi[e] = hi;
                }
                Blo[e] = lo;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lboA = KC * 128, sbo = 1024;
            for (int ks = 0; ks < KC / 8; ks++) {
                if (a.kmajor) {
                    uint64_t ah = make_desc(smem_u32(Ahi) + ks * 32, 16, 1024), bh = make_desc(smem_u32(Bhi) + ks * 32, 16, 1024);
                    uint64_t al = make_desc(smem_u32(Alo) + ks * 32, 16, 1024), bl = make_desc(smem_u32(Blo) + ks * 32, 16, 1024);
                    if (a.split) {
                        umma_tf32(tmem, al, bh, idesc, (fresh && ks == 0) ? 0u : 1u);
                        umma_tf32(tmem, ah, bl, idesc, 1u);
                        umma_tf32(tmem, ah, bh, idesc, 1u);
                    } else {
                        umma_tf32(tmem, ah, bh, idesc, (fresh && ks == 0) ? 0u : 1u);
                    }
                    continue;
                }
                uint64_t ah = make_desc(smem_u32(Ahi) + ks * 1024, a.swap ? sbo : lboA, a.swap ? lboA : sbo);
                uint64_t al = make_desc(smem_u32(Alo) + ks * 1024, a.swap ? sbo : lboA, a.swap ? lboA : sbo);
                uint64_t bh = make_desc(smem_u32(Bhi) + ks * 1024, 1024, sbo);
                uint64_t bl = make_desc(smem_u32(Blo) + ks * 1024, 1024, sbo);
                if (a.split) {
                    umma_tf32(tmem, al, bh, idesc, (fresh && ks == 0) ? 0u : 1u);
                    umma_tf32(tmem, ah, bl, idesc, 1u);
                    umma_tf32(tmem, ah, bh, idesc, 1u);
                } else {
                    umma_tf32(tmem, ah, bh, idesc, (fresh && ks == 0) ? 0u : 1u);
                }
            }
            umma_commit(&bar_mma);
        }
        fresh = false;
        mbar_wait(&bar_mma, ph_mma);
        ph_mma ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool drain = (ch == nchunks - 1) || (a.drain_every && (ch + 1) % a.drain_every == 0);
        if (drain) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] += v[i];
            fresh = true;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        __syncthreads();
    }
    if (nchunks == 0) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
        for (int i = 0; i < 32; i++) acc[i] = v[i];
    }
    for (int i = 0; i < 32; i++) a.D[tid * 32 + i] = acc[i];
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled get_encode()
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) {
        printf("cuTensorMapEncodeTiled not found\n");
        exit(1);
    }
    return (EncodeTiled)fn;
}

static CUtensorMap make_map(EncodeTiled enc, float *base, int cols, int rows, int box_cols, int box_rows)
{
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
        exit(1);
    }
    return tm;
}

static double lcg(uint64_t &s)
{
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return ((s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

int main()
{
    EncodeTiled enc = get_encode();
    const int KMAX = 30016;
    std::vector<float> At((size_t)KMAX * 128), Bt((size_t)KMAX * 32);
    uint64_t s = 12345;
    for (auto &v : At) v = (float)lcg(s);
    for (auto &v : Bt) v = (float)lcg(s);
    float *dA, *dB, *dD, *dAk, *dBk;
    std::vector<float> Ak((size_t)KMAX * 128), Bk((size_t)KMAX * 32);
    CK(cudaMalloc(&dA, At.size() * 4));
    CK(cudaMalloc(&dB, Bt.size() * 4));
    CK(cudaMalloc(&dD, 128 * 32 * 4));
    CK(cudaMemcpy(dA, At.data(), At.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dAk, At.size() * 4));
    CK(cudaMalloc(&dBk, Bt.size() * 4));
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));

    struct Case { int K, tma, split, swap, drain; const char *name; int kmajor = 0, prefill = 0; };
    Case cases[] = {
        {32, 0, 0, 0, 0, "K=32 K-major thread-fill 1xTF32", 1, 0},
        {32, 0, 0, 0, 0, "K=32 MN-major thread-fill 1xTF32 onto prefill", 0, 1},
        {64, 1, 0, 0, 0, "K=64 K-major TMA 1xTF32", 1, 0},
        {64, 1, 1, 0, 0, "K=64 K-major TMA 3xTF32 hi=raw lo=x-trunc(x)", 1, 0},
        {64, 1, 2, 0, 0, "K=64 K-major TMA 3xTF32 rna split", 1, 0},
        {KMAX, 1, 0, 0, 0, "K=30016 1xTF32 single accumulation", 1, 0},
        {KMAX, 1, 1, 0, 0, "K=30016 3xTF32 trunc, single accumulation", 1, 0},
        {KMAX, 1, 2, 0, 0, "K=30016 3xTF32 rna, single accumulation", 1, 0},
        {KMAX, 1, 2, 0, 64, "K=30016 3xTF32 rna, drain every 64 chunks (K=2048)", 1, 0},
        {KMAX, 1, 2, 0, 16, "K=30016 3xTF32 rna, drain every 16 chunks (K=512)", 1, 0},
        {KMAX, 1, 2, 0, 4, "K=30016 3xTF32 rna, drain every 4 chunks (K=128)", 1, 0},
        {KMAX, 1, 1, 0, 16, "K=30016 3xTF32 trunc, drain every 16 chunks", 1, 0},
        {KMAX, 1, 1, 0, 4, "K=30016 3xTF32 trunc, drain every 4 chunks", 1, 0},
    };
    std::vector<float> D(128 * 32);
    std::vector<double> ref(128 * 32), ref32(128 * 32);
    int lastK = -1;
    double rms = 0;
    for (auto &c : cases) {
        if (c.K != lastK) {
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < 32; n++) {
                    double acc = 0;
                    float f = 0.f;
                    for (int k = 0; k < c.K; k++) {
                        acc += (double)At[(size_t)k * 128 + m] * (double)Bt[(size_t)k * 32 + n];
                        f += At[(size_t)k * 128 + m] * Bt[(size_t)k * 32 + n];
                    }
                    ref[m * 32 + n] = acc;
                    ref32[m * 32 + n] = f;
                }
            rms = 0;
            for (double v : ref) rms += v * v;
            rms = std::sqrt(rms / ref.size());
            double e32 = 0;
            for (int i = 0; i < 128 * 32; i++) e32 = std::max(e32, std::fabs(ref32[i] - ref[i]));
            printf("K=%d: ref rms %.4f; sequential f32 FMA-free CPU sum max err / rms = %.3e\n", c.K, rms, e32 / rms);
            lastK = c.K;
        }
        const int Kc = c.K ? c.K : 32;
        CUtensorMap tmA = make_map(enc, dA, 128, Kc, 32, KC), tmB = make_map(enc, dB, 32, Kc, 32, KC);
        if (c.kmajor) {
            for (int k = 0; k < Kc; k++) {
                for (int m = 0; m < 128; m++) Ak[(size_t)m * Kc + k] = At[(size_t)k * 128 + m];
                for (int n = 0; n < 32; n++) Bk[(size_t)n * Kc + k] = Bt[(size_t)k * 32 + n];
            }
            CK(cudaMemcpy(dAk, Ak.data(), (size_t)Kc * 128 * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dBk, Bk.data(), (size_t)Kc * 32 * 4, cudaMemcpyHostToDevice));
            tmA = make_map(enc, dAk, Kc, 128, KC, 128);
            tmB = make_map(enc, dBk, Kc, 32, KC, 32);
        }
        ProbeArgs a{dA, dB, dD, c.K, c.tma, c.split, c.swap, c.drain, c.kmajor, c.prefill};
        CK(cudaMemset(dD, 0xFF, 128 * 32 * 4));
        k_probe<<<1, 128, 48 * 1024>>>(a, tmA, tmB);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double emax = 0, bias = 0;
        for (int i = 0; i < 128 * 32; i++) {
            emax = std::max(emax, std::fabs((double)D[i] - ref[i]));
            bias += ((double)D[i] - ref[i]) * (ref[i] >= 0 ? 1 : -1);
        }
        printf("   D[0][0..3] = %g %g %g %g   D[5][0..1] = %g %g   D[127][31] = %g   ref[0][0..3] = %g %g %g %g\n", D[0], D[1], D[2], D[3],
               D[5 * 32], D[5 * 32 + 1], D[127 * 32 + 31], ref[0], ref[1], ref[2], ref[3]);
        printf("%-58s max err / rms = %.3e   mean signed err (toward larger |x|) / rms = %+.3e  D[0]=%.5f ref=%.5f\n", c.name,
               emax / rms, bias / (128 * 32) / rms, D[0], ref[0]);
    }
    return 0;
}
