// mac_kernels.cuh — K2: the frequency-domain delay-line complex multiply-accumulate, fused with
// the segment-ring indexing.  This is the HBM-bound kernel that carries >99 % of the bytes.
//
// Replaces src/fft_convolver.rs:244-255 (and complex_multiply_accumulate, :62-74):
//     pre_multiplied[k] = sum_{i=1}^{active-1} ir[i][k] * ring[(current+i) % active][k]
// accumulated in ascending i with every multiply / subtract / add rounded separately in f32
// (Rust does not contract to FMA), so for identical spectra the result is bit-identical to the
// reference loop.  Bin 0 of a packed row holds {DC.re, Nyquist.re}: two independent real
// products, which is exactly what the reference's complex product of two purely real bins gives.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fcb {

// streaming 16-byte load: every IR / ring byte is touched once per block, keep it out of L1
__device__ __forceinline__ float4 ld_stream4(const float4 *p)
{
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ld_stream2(const float2 *p)
{
    float2 r;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// acc += a * b for one complex bin, reference rounding; `packed` selects the {DC, Nyquist} rule
__device__ __forceinline__ void cmac_ref(float &ar, float &ai, float xr, float xi, float hr, float hi, bool packed)
{
    float rr = __fmul_rn(xr, hr), ii = __fmul_rn(xi, hi);
    float ri = __fmul_rn(xr, hi), ir = __fmul_rn(xi, hr);
    float pr = packed ? rr : __fsub_rn(rr, ii);
    float pi = packed ? ii : __fadd_rn(ri, ir);
    ar = __fadd_rn(ar, pr);
    ai = __fadd_rn(ai, pi);
}

struct MacArgs {
    const float2 *ir;   // packed rows [ir channel][rows][B]; row r holds IR segment ir_seg0 + r
    long long ir_stride;   // per IR channel (0 when one IR is shared by every channel)
    const float2 *ring; // [ring channel][S][B]
    long long ring_stride;
    float2 *premul;     // [C][B]
    int current, active;
    long long nchan;
    // segments accumulated: i in [seg_lo, seg_hi).  FFTConvolver: [1, active).  An IR-partition
    // shard of the MIMO matrix owns a sub-range and stores only those rows (ir_seg0 = first stored).
    int seg_lo, seg_hi, ir_seg0;
    // work channel c -> IR channel (c % ir_mod, 0 = identity) and ring channel
    // ((c / ring_div) * ring_mul + c % ring_mod, ring_div == 0 = identity).  The MIMO matrix uses
    // c = (stream*OUT + out)*IN + in  ->  IR (out, in), ring (stream, in).
    long long ir_mod, ring_div, ring_mul, ring_mod;

    __host__ __device__ long long ir_chan(long long c) const { return ir_mod ? c % ir_mod : c; }
    __host__ __device__ long long ring_chan(long long c) const
    {
        return ring_div ? (c / ring_div) * ring_mul + c % ring_mod : c;
    }
};

// One thread owns V = 2 adjacent bins (one float4) of one channel and walks all segments.
// blockDim.x = threads along the row (<= 256), blockDim.y = channels per CTA;
// grid.x = row tiles * channel groups.  U = segments in flight per thread (2*U 16-byte loads).
template <int B, int U>
__global__ void __launch_bounds__(256)
k_mac_v4(MacArgs a)
{
    constexpr int ROW4 = B / 2;                       // float4 per row
    constexpr int TX = ROW4 < 256 ? ROW4 : 256;       // threads along the row
    constexpr int TILES = ROW4 / TX;                  // row tiles per channel
    constexpr int CPB = 256 / TX;                     // channels per CTA
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long grp = blockIdx.x / TILES;
    const int tile = blockIdx.x % TILES;
    const long long c = grp * CPB + ty;
    if (c >= a.nchan) return;
    const int t4 = tile * TX + tx; // float4 index within the row
    // IR row r holds segment ir_seg0 + r: bias the base so that `ir + i*ROW4` is segment i
    const float4 *ir = reinterpret_cast<const float4 *>(a.ir + a.ir_chan(c) * a.ir_stride) + t4 - (long long)a.ir_seg0 * ROW4;
    const float4 *rg = reinterpret_cast<const float4 *>(a.ring + a.ring_chan(c) * a.ring_stride) + t4;
    const bool packed = (t4 == 0);
    const int cur = a.current, act = a.active, hi = a.seg_hi;

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = a.seg_lo;
    for (; i + U <= hi; i += U) {
        float4 h[U], x[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            int j = (cur + i + u) % act;
            h[u] = ld_stream4(ir + (long long)(i + u) * ROW4);
            x[u] = ld_stream4(rg + (long long)j * ROW4);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            cmac_ref(acc.x, acc.y, h[u].x, h[u].y, x[u].x, x[u].y, packed);
            cmac_ref(acc.z, acc.w, h[u].z, h[u].w, x[u].z, x[u].w, false);
        }
    }
    for (; i < hi; i++) {
        int j = (cur + i) % act;
        float4 h = ld_stream4(ir + (long long)i * ROW4);
        float4 x = ld_stream4(rg + (long long)j * ROW4);
        cmac_ref(acc.x, acc.y, h.x, h.y, x.x, x.y, packed);
        cmac_ref(acc.z, acc.w, h.z, h.w, x.z, x.w, false);
    }
    reinterpret_cast<float4 *>(a.premul + c * B)[t4] = acc;
}

// ---------------------------------------------------------------------------------------------
// TMA variant (the default for B >= 16): IR and ring rows are staged into shared memory by
// cp.async.bulk copies (SASS: UBLKCP) completing on mbarriers, NST stages deep, so the bytes in
// flight per SM are set by the pipeline (NST * 32 KB per CTA), not by what ptxas keeps in
// registers.  Per channel the [S][B] spectra are contiguous, so one stage = R = 4 consecutive IR
// rows (one copy) and 4 consecutive ring rows starting at (current+i) % active (one copy, two
// when the ring wraps inside the stage).  A CTA is CPB channels x TILE bins with CPB*TILE = 512:
// 256 threads, each owning one float4 (2 bins) of one channel, accumulating in ascending segment
// order with the reference's unfused rounding — same bits as k_mac_v4.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy, completion counted in bytes on `bar`; bytes % 16 == 0, 16 B aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int B>
struct MacBulkCfg {
    static constexpr int TILE = B < 512 ? B : 512;   // bins per CTA row chunk
    static constexpr int CPB = 512 / TILE;           // channels per CTA
    static constexpr int TILES = B / TILE;           // row chunks per channel
    static constexpr int R = 4;                      // rows (segments) per stage
    static constexpr int ARR = CPB * R * TILE;       // float2 per array per stage (= 2048 -> 16 KB)
    static constexpr size_t STAGE_BYTES = 2 * (size_t)ARR * sizeof(float2); // 32 KB
    __host__ __device__ static constexpr size_t smem_bytes(int nst) { return nst * STAGE_BYTES + 64; }
};

template <int B, int NST>
__global__ void __launch_bounds__(256)
k_mac_bulk(MacArgs a)
{
    using Cfg = MacBulkCfg<B>;
    constexpr int TILE = Cfg::TILE, CPB = Cfg::CPB, TILES = Cfg::TILES, R = Cfg::R, ARR = Cfg::ARR;
    constexpr int TX = TILE / 2; // threads (float4) along the row chunk
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *stages = reinterpret_cast<float2 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NST * Cfg::STAGE_BYTES);

    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long grp = blockIdx.x / TILES;
    const int tile = blockIdx.x % TILES;
    const long long c0 = grp * CPB;
    const int nlive = (int)((a.nchan - c0) < CPB ? (a.nchan - c0) : CPB); // channels of this CTA that exist
    const int cur = a.current, act = a.active, lo = a.seg_lo, hi = a.seg_hi;
    const int nrows = hi - lo;                 // segments lo .. hi-1
    const int niter = (nrows + R - 1) / R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer: fill stage (it % NST) with the rows of iteration `it`
    auto issue = [&](int it) {
        const int s = it % NST;
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        float2 *ir_s = stages + (size_t)s * 2 * ARR;
        float2 *rg_s = ir_s + ARR;
        mbar_expect_tx(&full[s], (uint32_t)(2 * nlive * cnt * TILE * sizeof(float2)));
        for (int ch = 0; ch < nlive; ch++) {
            const float2 *irc = a.ir + a.ir_chan(c0 + ch) * a.ir_stride + tile * TILE - (long long)a.ir_seg0 * B;
            const float2 *rgc = a.ring + a.ring_chan(c0 + ch) * a.ring_stride + tile * TILE;
            float2 *ir_d = ir_s + ch * R * TILE, *rg_d = rg_s + ch * R * TILE;
            const int j0 = (cur + i0) % act; // `current` may exceed `active` after a shrinking update()
            if (TILES == 1) {
                bulk_g2s(ir_d, irc + (long long)i0 * B, cnt * TILE * sizeof(float2), &full[s]);
                int first = (act - j0) < cnt ? (act - j0) : cnt; // rows before the ring wraps
                bulk_g2s(rg_d, rgc + (long long)j0 * B, first * TILE * sizeof(float2), &full[s]);
                if (first < cnt)
                    bulk_g2s(rg_d + first * TILE, rgc, (cnt - first) * TILE * sizeof(float2), &full[s]);
            } else {
                for (int r = 0; r < cnt; r++) {
                    int j = j0 + r;
                    j = j >= act ? j - act : j;
                    bulk_g2s(ir_d + r * TILE, irc + (long long)(i0 + r) * B, TILE * sizeof(float2), &full[s]);
                    bulk_g2s(rg_d + r * TILE, rgc + (long long)j * B, TILE * sizeof(float2), &full[s]);
                }
            }
        }
    };

    if (threadIdx.x == 0)
        for (int it = 0; it < NST && it < niter; it++) issue(it);

    const bool packed = (tile == 0 && tx == 0);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < niter; it++) {
        const int s = it % NST;
        mbar_wait(&full[s], (it / NST) & 1);
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        const float4 *ir_s = reinterpret_cast<const float4 *>(stages + (size_t)s * 2 * ARR + ty * R * TILE) + tx;
        const float4 *rg_s = ir_s + ARR / 2;
        if (ty < nlive) {
            if (cnt == R) {
                float4 h[R], x[R];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    h[r] = ir_s[r * TX];
                    x[r] = rg_s[r * TX];
                }
#pragma unroll
                for (int r = 0; r < R; r++) {
                    cmac_ref(acc.x, acc.y, h[r].x, h[r].y, x[r].x, x[r].y, packed);
                    cmac_ref(acc.z, acc.w, h[r].z, h[r].w, x[r].z, x[r].w, false);
                }
            } else {
                for (int r = 0; r < cnt; r++) {
                    float4 h = ir_s[r * TX], x = rg_s[r * TX];
                    cmac_ref(acc.x, acc.y, h.x, h.y, x.x, x.y, packed);
                    cmac_ref(acc.z, acc.w, h.z, h.w, x.z, x.w, false);
                }
            }
        }
        __syncthreads(); // every reader is done with stage s
        if (threadIdx.x == 0 && it + NST < niter) issue(it + NST);
    }
    if (ty < nlive)
        reinterpret_cast<float4 *>(a.premul + (c0 + ty) * B)[tile * TX + tx] = acc;
}

// ---------------------------------------------------------------------------------------------
// Matrix variant of K2 (convolution matrix, BASELINE configs[4]): one CTA owns OT outputs x ST
// streams of ONE input channel and one bin tile, for one chunk of the segment range.  Per stage
// the producer stages OT IR tiles (h[out][in], shared by the ST streams) and ST ring tiles
// (x[stream][in], shared by the OT outputs) with cp.async.bulk; every thread keeps OT*ST float4
// accumulators.  Traffic per complex MAC drops from 16 B to 8/ST + 8/OT bytes, so the IR matrix is
// read from HBM once per block however many outputs/streams reuse it.  FMA arithmetic: the sum
// over inputs/shards is re-associated anyway (parity is by tolerance here, not bit-exact).
// Partial results go to part[((z*NS + s)*OUT + o)*IN + in][B]; k_mimo_reduce sums over z and in.
// ---------------------------------------------------------------------------------------------
struct MacTileArgs {
    const float2 *ir;   // [OUT*IN][rows][B], row r = IR segment ir_seg0 + r
    long long ir_stride;
    const float2 *ring; // [NS*IN][S][B]
    long long ring_stride;
    float2 *part;       // [Z*NS*OUT*IN][B]
    int current, active; // ring slot of the current block, ring length S
    int seg_lo, seg_hi, ir_seg0;
    int n_in, n_out, n_streams, zchunks, zlen; // zlen = segments per z chunk
};

__device__ __forceinline__ void cmac_fma2(float4 &acc, const float4 &h, const float4 &x, bool packed)
{
    // bins (h.x,h.y)*(x.x,x.y) and (h.z,h.w)*(x.z,x.w); bin 0 of a row is {DC, Nyquist}: two real products
    if (packed) {
        acc.x = fmaf(h.x, x.x, acc.x);
        acc.y = fmaf(h.y, x.y, acc.y);
    } else {
        acc.x = fmaf(h.x, x.x, fmaf(-h.y, x.y, acc.x));
        acc.y = fmaf(h.x, x.y, fmaf(h.y, x.x, acc.y));
    }
    acc.z = fmaf(h.z, x.z, fmaf(-h.w, x.w, acc.z));
    acc.w = fmaf(h.z, x.w, fmaf(h.w, x.z, acc.w));
}

template <int B, int OT, int ST>
struct MacTileCfg {
    static constexpr int TILE = B < 512 ? B : 512;
    static constexpr int TILES = B / TILE;
    static constexpr int TX = TILE / 2;              // threads per CTA
    static constexpr int R = 2;                      // segments per stage
    static constexpr int ARR = R * TILE;             // float2 per (output or stream) per stage
    static constexpr size_t STAGE_BYTES = (size_t)(OT + ST) * ARR * sizeof(float2);
    static constexpr int NST = (3 * STAGE_BYTES + 64 <= 220 * 1024) ? 3 : 2;
    static constexpr size_t SMEM_BYTES = NST * STAGE_BYTES + 64;
};

template <int B, int OT, int ST>
__global__ void __launch_bounds__(MacTileCfg<B, OT, ST>::TX)
k_mac_tile(MacTileArgs a)
{
    using Cfg = MacTileCfg<B, OT, ST>;
    constexpr int TILE = Cfg::TILE, TILES = Cfg::TILES, TX = Cfg::TX, R = Cfg::R, ARR = Cfg::ARR, NST = Cfg::NST;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *stages = reinterpret_cast<float2 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NST * Cfg::STAGE_BYTES);

    // blockIdx.x = tile + TILES*(z + Z*(og + OG*(in + IN*sg)))
    const int OG = (a.n_out + OT - 1) / OT;
    long long bid = blockIdx.x;
    const int tile = (int)(bid % TILES); bid /= TILES;
    const int z = (int)(bid % a.zchunks); bid /= a.zchunks;
    const int og = (int)(bid % OG); bid /= OG;
    const int in = (int)(bid % a.n_in); bid /= a.n_in;
    const int sg = (int)bid;
    const int o0 = og * OT, s0 = sg * ST;
    const int no = (a.n_out - o0) < OT ? (a.n_out - o0) : OT;
    const int nsl = (a.n_streams - s0) < ST ? (a.n_streams - s0) : ST;
    const int lo = a.seg_lo + z * a.zlen;
    const int hi = (lo + a.zlen) < a.seg_hi ? (lo + a.zlen) : a.seg_hi;
    const int niter = hi > lo ? (hi - lo + R - 1) / R : 0;
    const int cur = a.current, act = a.active;
    const int tx = threadIdx.x;

    if (tx == 0) {
        for (int s = 0; s < NST; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int it) {
        const int s = it % NST;
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        float2 *h_s = stages + (size_t)s * (OT + ST) * ARR;
        float2 *x_s = h_s + OT * ARR;
        mbar_expect_tx(&full[s], (uint32_t)((no + nsl) * cnt * TILE * sizeof(float2)));
        const int j0 = (cur + i0) % act;
        const int first = (act - j0) < cnt ? (act - j0) : cnt;
        for (int o = 0; o < no; o++) {
            const float2 *src = a.ir + ((long long)(o0 + o) * a.n_in + in) * a.ir_stride + tile * TILE - (long long)a.ir_seg0 * B;
            if (TILES == 1) {
                bulk_g2s(h_s + o * ARR, src + (long long)i0 * B, cnt * TILE * sizeof(float2), &full[s]);
            } else {
                for (int r = 0; r < cnt; r++)
                    bulk_g2s(h_s + o * ARR + r * TILE, src + (long long)(i0 + r) * B, TILE * sizeof(float2), &full[s]);
            }
        }
        for (int st = 0; st < nsl; st++) {
            const float2 *src = a.ring + ((long long)(s0 + st) * a.n_in + in) * a.ring_stride + tile * TILE;
            if (TILES == 1) {
                bulk_g2s(x_s + st * ARR, src + (long long)j0 * B, first * TILE * sizeof(float2), &full[s]);
                if (first < cnt)
                    bulk_g2s(x_s + st * ARR + first * TILE, src, (cnt - first) * TILE * sizeof(float2), &full[s]);
            } else {
                for (int r = 0; r < cnt; r++) {
                    int j = j0 + r;
                    j = j >= act ? j - act : j;
                    bulk_g2s(x_s + st * ARR + r * TILE, src + (long long)j * B, TILE * sizeof(float2), &full[s]);
                }
            }
        }
    };

    if (tx == 0)
        for (int it = 0; it < NST && it < niter; it++) issue(it);

    const bool packed = (tile == 0 && tx == 0);
    float4 acc[OT][ST];
#pragma unroll
    for (int o = 0; o < OT; o++)
#pragma unroll
        for (int st = 0; st < ST; st++) acc[o][st] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int it = 0; it < niter; it++) {
        const int s = it % NST;
        mbar_wait(&full[s], (it / NST) & 1);
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        const float4 *h_s = reinterpret_cast<const float4 *>(stages + (size_t)s * (OT + ST) * ARR) + tx;
        const float4 *x_s = h_s + OT * ARR / 2;
        for (int r = 0; r < cnt; r++) {
            float4 x[ST];
#pragma unroll
            for (int st = 0; st < ST; st++) x[st] = st < nsl ? x_s[(st * ARR + r * TILE) / 2] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < OT; o++) {
                if (o < no) {
                    const float4 h = h_s[(o * ARR + r * TILE) / 2];
#pragma unroll
                    for (int st = 0; st < ST; st++) cmac_fma2(acc[o][st], h, x[st], packed);
                }
            }
        }
        __syncthreads();
        if (tx == 0 && it + NST < niter) issue(it + NST);
    }
#pragma unroll
    for (int o = 0; o < OT; o++)
#pragma unroll
        for (int st = 0; st < ST; st++)
            if (o < no && st < nsl) {
                const long long row = (((long long)z * a.n_streams + (s0 + st)) * a.n_out + (o0 + o)) * a.n_in + in;
                reinterpret_cast<float4 *>(a.part + row * B)[tile * TX + tx] = acc[o][st];
            }
}

// B == 1: a row is a single packed bin
static __global__ void k_mac_b1(MacArgs a)
{
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nchan) return;
    const float2 *ir = a.ir + a.ir_chan(c) * a.ir_stride - a.ir_seg0;
    const float2 *rg = a.ring + a.ring_chan(c) * a.ring_stride;
    float ar = 0.f, ai = 0.f;
    for (int i = a.seg_lo; i < a.seg_hi; i++) {
        int j = (a.current + i) % a.active;
        float2 h = ld_stream2(ir + i), x = ld_stream2(rg + j);
        cmac_ref(ar, ai, h.x, h.y, x.x, x.y, true);
    }
    a.premul[c] = make_float2(ar, ai);
}

} // namespace fcb
