// tc_k4_test.cu — stand-alone harness for k_mimo_tc (mimo_tc.cuh): random K-major operands, CPU f64 check.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Ifft_convolution_b200/csrc -o build/tc_k4_test scripts/tc_k4_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mimo_tc.cuh"

using namespace fcb;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t p1, uint64_t p2, uint64_t p3,
                            uint32_t b1)
{
    static EncodeTiled enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        enc = (EncodeTiled)fn;
    }
    CUtensorMap tm;
    cuuint64_t dims[4] = {d0, d1, d2, d3}, strides[3] = {p1, p2, p3};
    cuuint32_t box[4] = {2 * TC_KSEG, b1, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d3 ? 4 : 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        exit(1);
    }
    return tm;
}
static double lcg(uint64_t &s)
{
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return ((s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

int main(int argc, char **argv)
{
    const int B = argc > 1 ? atoi(argv[1]) : 4, IN = argc > 2 ? atoi(argv[2]) : 2, S = argc > 3 ? atoi(argv[3]) : 38;
    const int NS = argc > 4 ? atoi(argv[4]) : 5, cur = argc > 5 ? atoi(argv[5]) : 7, groups = argc > 6 ? atoi(argv[6]) : 1;
    const int seg_lo = argc > 7 ? atoi(argv[7]) : 0, seg_hi = argc > 8 ? atoi(argv[8]) : S;
    const int OUT = 16, rows = seg_hi - seg_lo;
    const bool perf = getenv("TC_PERF") != nullptr; // device-filled operands, timing only
    const size_t nblk = (S + TC_KSEG - 1) / TC_KSEG, rowsP = (TC_LEAD + 1 + rows + 1) & ~1, NSP = (NS + 7) & ~7;
    auto ring_at = [&](int b, int in, int s, int sl, int p) {
        return (((((size_t)b * IN + in) * nblk + sl / TC_KSEG) * NSP + s) * TC_KSEG + sl % TC_KSEG) * 2 + p;
    };
    printf("B=%d IN=%d S=%d NS=%d cur=%d groups=%d segs [%d,%d)\n", B, IN, S, NS, cur, groups, seg_lo, seg_hi);
    const size_t copy = (size_t)B * IN * 2 * OUT * 2 * rowsP;
    const size_t ring_n = (size_t)B * IN * nblk * NSP * TC_KSEG * 2;
    std::vector<float> ring(perf ? 1 : ring_n, 0.f), ir(perf ? 1 : 2 * copy, 0.f);
    uint64_t seed = 99;
    for (int b = 0; b < (perf ? 0 : B); b++)
        for (int in = 0; in < IN; in++) {
            for (int s = 0; s < NS; s++)
                for (int sl = 0; sl < S; sl++)
                    for (int p = 0; p < 2; p++) ring[ring_at(b, in, s, sl, p)] = (float)lcg(seed);
            for (int n = 0; n < 2 * OUT; n++)
                for (int r = 0; r < rows; r++)
                    for (int p = 0; p < 2; p++) {
                        const float v = (float)lcg(seed);
                        ir[(((size_t)b * IN + in) * 2 * OUT + n) * 2 * rowsP + 2 * (TC_LEAD + r) + p] = v;
                        ir[copy + (((size_t)b * IN + in) * 2 * OUT + n) * 2 * rowsP + 2 * (TC_LEAD + 1 + r) + p] = v;
                    }
        }
    float *d_ring, *d_ir;
    float2 *d_part;
    CK(cudaMalloc(&d_ring, ring_n * 4));
    CK(cudaMalloc(&d_ir, 2 * copy * 4));
    if (perf) {
        CK(cudaMemset(d_ring, 0x3c, ring_n * 4)); // 0x3c3c3c3c = 0.0115
        CK(cudaMemset(d_ir, 0x3c, 2 * copy * 4));
    }
    CK(cudaMalloc(&d_part, (size_t)groups * NS * OUT * B * 8));
    if (!perf) {
        CK(cudaMemcpy(d_ring, ring.data(), ring.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_ir, ir.data(), ir.size() * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(d_part, 0xFF, (size_t)groups * NS * OUT * B * 8));
    const uint64_t tile = (uint64_t)NSP * TC_KSEG * 8;
    CUtensorMap tmr = make_map(d_ring, 2 * TC_KSEG, NSP, nblk, (uint64_t)B * IN, TC_KSEG * 8, tile, nblk * tile, (uint32_t)NSP);
    CUtensorMap tmi = make_map(d_ir, 2 * (TC_LEAD + rows), 2 * OUT, (uint64_t)B * IN, 0, 2 * rowsP * 4, 2 * OUT * 2 * rowsP * 4, 0, 2 * OUT);
    CUtensorMap tmi1 =
        make_map(d_ir + copy, 2 * (TC_LEAD + 1 + rows), 2 * OUT, (uint64_t)B * IN, 0, 2 * rowsP * 4, 2 * OUT * 2 * rowsP * 4, 0, 2 * OUT);
    CK(cudaFuncSetAttribute(k_mimo_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<16>::SMEM));
    TcArgs a{};
    a.part = d_part;
    a.B = B;
    a.n_in = IN;
    a.n_streams = NS;
    a.rows_pad = (int)NSP;
    a.n_out = OUT;
    a.out_groups = 1;
    a.stream_groups = 1;
    a.nblk = (int)nblk;
    a.S = S;
    a.current = cur;
    a.seg_lo = seg_lo;
    a.seg_hi = seg_hi;
    a.groups = groups;
    k_mimo_tc<16><<<B * groups, TC_THREADS, TcCfg<16>::SMEM>>>(a, tmr, tmi, tmi1);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (perf) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        const int reps = 10;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; r++) {
            a.current = (cur + 37 * r) % S;
            k_mimo_tc<16><<<B * groups, TC_THREADS, TcCfg<16>::SMEM>>>(a, tmr, tmi, tmi1);
        }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        const double bytes = (double)B * IN * ((double)nblk * NSP * TC_KSEG * 8 + 2.0 * OUT * rows * 8);
        printf("PERF %.4f ms per launch, %.0f GB/s of operand bytes, %.2f T cMAC/s (%d rows)\n", ms, bytes / ms / 1e6,
               (double)B * IN * rows * NS * OUT / ms / 1e9, NS);
        return 0;
    }
    std::vector<float2> part((size_t)groups * NS * OUT * B);
    CK(cudaMemcpy(part.data(), d_part, part.size() * 8, cudaMemcpyDeviceToHost));
    double emax = 0, rms = 0;
    size_t cnt = 0;
    for (int b = 0; b < B; b++)
        for (int s = 0; s < NS; s++)
            for (int n = 0; n < 2 * OUT; n++) {
                double ref = 0;
                for (int in = 0; in < IN; in++)
                    for (int i = seg_lo; i < seg_hi; i++) {
                        const int sl = (cur + i) % S;
                        for (int p = 0; p < 2; p++)
                            ref += (double)ring[ring_at(b, in, s, sl, p)] *
                                   (double)ir[(((size_t)b * IN + in) * 2 * OUT + n) * 2 * rowsP + 2 * (TC_LEAD + i - seg_lo) + p];
                    }
                double got = 0;
                for (int g = 0; g < groups; g++) {
                    float2 v = part[(((size_t)g * NS + s) * OUT + (n % OUT)) * B + b];
                    got += n < OUT ? v.x : v.y;
                }
                emax = std::max(emax, std::fabs(got - ref));
                rms += ref * ref;
                cnt++;
                if (b == 0 && s == 0 && n < 3) printf("  D[0][0][%d] = %.6f ref %.6f\n", n, got, ref);
            }
    rms = std::sqrt(rms / cnt);
    printf("max err / rms = %.3e (rms %.3f)\n", emax / rms, rms);
    return emax / rms < 1e-5 ? 0 : 2;
}
