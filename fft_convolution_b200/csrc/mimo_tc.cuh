// mimo_tc.cuh — K4: the convolution-matrix delay-line MAC as per-bin complex GEMMs on the
// sm_100a tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA).
//
// Replaces, for the OUT x IN matrix with NS streams sharing the IR matrix, the loops at
// src/fft_convolver.rs:244-261 of all OUT*IN reference convolvers at once.  Per bin k
//     D[s][o] = sum_in sum_i  X[s][in][(current+i) % S][k] * H[o][in][i][k]          (complex)
// is a dense contraction over j = (in, i): M = streams, N = outputs, K = IN*S — the one place on
// this path where the work really is a GEMM.  As a real GEMM (K doubles, N doubles):
//     [D.re | D.im][s][n] = sum_{j,p}  A[s][(j,p)] * Bm[n][(j,p)],   A[s][(j,0)] = X.re, A[s][(j,1)] = X.im
//     Bm[o][(j,0)] = H.re,  Bm[o][(j,1)] = -H.im        (n = o:        real part)
//     Bm[OUT+o][(j,0)] = H.im, Bm[OUT+o][(j,1)] = H.re  (n = OUT + o:  imaginary part)
// and for the packed bin 0 ({DC, Nyquist}: two real products) Bm[o] = {H.re, 0}, Bm[OUT+o] = {0, H.im}.
//
// HBM layout (K-major operands, so a tile is one TMA box and no transposition happens on chip):
//     ring_t [bin][in][stream group][slot block][NSP rows][16 slots] float2   A: row = stream, K = (slot, re/im)
//     ir_t   [2][bin][in][32 rows per 16 outputs][2*posP]              float    B: row = n,      K = (segment, plane)
// (more than 128 streams / 16 outputs run as further CTAs over stream groups of 128 / output groups of 16)
// ring_t is tile-major: the 16 slots x NSP streams one pipeline stage consumes are ONE contiguous
// run of NSP * 128 bytes (DRAM-page friendly; a [stream][slot] matrix would be NSP separate 128-byte
// pieces).  NSP = streams rounded up to 8: the UMMA always works on 128 rows, but rows >= NSP of the
// shared-memory tile are simply never loaded — their (stale) contents only reach accumulator rows
// that no stream owns.
// The new block's spectrum is scattered into slot `current` of ring_t by K1's epilogue kernel;
// ir_t is built once per set_ir.  Slot s holds segment i = (s - current) mod S, so the K loop walks
// the slot blocks of two slot ranges and reads IR positions that start wherever `current` puts
// them.  A TMA box must start on a 16-byte boundary, i.e. on an even IR position, so ir_t is kept
// twice — copy 1 shifted by one position — and each range reads the copy whose parity matches.
// Segment i sits at position TC_LEAD + copy + (i - seg_lo): the TC_LEAD zero positions in front
// and TMA's out-of-bounds zero fill behind (the tensor extent is exact) blank the slots of a block
// that belong to segments outside the range; ring slots >= S are never written and stay zero.
//
// Precision.  north_star asks for 1e-5 * RMS (f32); a tf32 product keeps 11 bits.  Every operand is
// split on chip x = hi + lo (hi = the bits the tensor core reads, lo = x - hi, both exact) and
// hi*hi, hi*lo, lo*hi are accumulated (3xTF32: error ~2^-22 per product).  TMEM accumulation
// truncates (measured: a single K = 30016 accumulation drifts by 9e-4 * RMS toward zero), so the
// accumulator is drained into f32 registers every 128 K-elements and the small cross terms go to
// their own accumulator columns (scripts/tc_probe.cu, profiles/r01_tc_probe.txt).
//
// One CTA = one (bin, input group): 14 warps — TMA producer, MMA issuer, 8 splitter warps
// (hi/lo), 4 drain warps (TMEM -> registers -> partial spectra).  A pipeline stage is TC_BOX = 2 boxes of 16
// segments (one hand-off per 32 segments: 1.47 -> 1.44 ms at 128 streams, 1.03 -> 0.94 ms at 16).  Shared memory:
// TC_NR raw stages (per box A 16 KB + B 4 KB, written by TMA) and TC_NL split stages of the small operand (per box
// [B_hi | B_lo], 8 KB).  The big operand never returns to shared memory: a splitter thread owns one stream row,
// reads its 128-byte swizzled row, and writes hi and lo straight into TMEM (tcgen05.st), from where
// the MMAs take A — with both operands in shared memory the one 128 B/clk port carried TMA writes, the split's
// LDS/STS and the UMMA operand reads (ncu on the first version: LSU wavefronts alone 41 %).  The other early limit
// was the issue path: see elect_one().  History and numbers: profiles/r01_k4_notes.txt.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mac_kernels.cuh"

namespace fcb {

constexpr int TC_M = 128;    // stream rows of one UMMA (streams are padded to this)
constexpr int TC_KSEG = 16;  // segments per stage: 16 (re,im) pairs = 32 tf32 = one 128-byte swizzle row
constexpr int TC_LEAD = 16;  // zero positions in front of every IR row
#ifndef TC_BOXES
#define TC_BOXES 2
#endif
constexpr int TC_BOX = TC_BOXES;              // 16-segment TMA boxes per pipeline stage (one hand-off per stage)
constexpr int TC_NR = 8 / TC_BOX;             // raw (TMA) stages
constexpr int TC_SETS = 1;   // accumulator sets the K steps alternate between (2 measured no faster: MMAs are issue-paced, not dependency-paced)
constexpr int TC_NL = TC_BOX == 1 ? 4 : 3;    // split-operand stages (A hi/lo in TMEM, [B_hi | B_lo] in shared memory)
constexpr int TC_DRAIN = 4 / TC_BOX;          // stages per TMEM accumulation interval (K = 128 per drain)
constexpr int TC_SPLIT_WARPS = 8;
constexpr int TC_THREADS = 32 * (2 + TC_SPLIT_WARPS + 4);

template <int NOUT>
struct TcCfg {
    static constexpr int N2 = 2 * NOUT;                    // GEMM N (re | im)
    static constexpr int A_BYTES = TC_M * 128;             // 16 KB: 128 rows x 128 B
    static constexpr int B_BYTES = N2 * 128;               // 4 KB at 16 outputs
    static constexpr int BOX_RAW = A_BYTES + B_BYTES;      // one box as loaded: A | B
    static constexpr int BOX_LO = 2 * B_BYTES;             // one box of the split small operand: B_hi | B_lo
    static constexpr int RAW_BYTES = TC_BOX * BOX_RAW;
    static constexpr int LO_BYTES = TC_BOX * BOX_LO;
    static constexpr int SET_COLS = 2 * N2;                // main | cross (hi*lo + lo*hi: both small, one accumulator)
    static constexpr int BUF_COLS = TC_SETS * SET_COLS;
    static constexpr int ACC_COLS = 2 * BUF_COLS;          // double-buffered against the drain
    static constexpr int BOX_COLS = 4 * TC_KSEG;           // per box: 32 columns of hi, 32 of lo
    static constexpr int A_COLS = TC_BOX * BOX_COLS;
    static constexpr int TMEM_NEED = ACC_COLS + TC_NL * A_COLS;
    static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
    static constexpr size_t SMEM = (size_t)TC_NR * RAW_BYTES + (size_t)TC_NL * LO_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(N2 % 16 == 0 && 2 * N2 <= 256, "UMMA M=128 needs N % 16 == 0, N <= 256");
    static_assert(TMEM_NEED <= 512, "TMEM has 512 columns");
    static_assert(RAW_BYTES % 1024 == 0 && LO_BYTES % 1024 == 0, "swizzled tiles need 1024-byte alignment");
};

struct TcArgs {
    float2 *part;        // [groups][NS][OUT][B] partial spectra (packed rows)
    int B, n_in, n_streams, n_out;
    int rows_pad;        // stream rows stored per ring tile (streams rounded up to 8, 128 with several stream groups)
    int out_groups;      // ceil(n_out / NOUT): outputs beyond NOUT run as further CTAs re-reading the ring
    int stream_groups;   // ceil(n_streams / 128)
    int nblk;            // slot blocks per ring row group
    int S;               // ring slots
    int current;
    int seg_lo, seg_hi;  // segments accumulated; IR position of segment i in copy sh = TC_LEAD + sh + i - seg_lo
    int groups;          // input groups per bin; grid = B * groups * out_groups * stream_groups
};

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// one lane of a converged warp; ptxas then issues the single-thread tcgen05 / TMA instructions
// directly instead of wrapping each one in an ELECT + R2UR + BRA.U.ANY lane-serialisation loop (~120 cycles per
// MMA with a plain `if (lane == 0)`: measured, scripts/tc_mma_rate.cu)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 B apart, version 1
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// D f32, A/B tf32, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
// A from TMEM (lanes = rows, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// chunk c of one input's K loop -> ring slot block, IR copy and IR position paired with the block's
// first slot.  Range 1: segments [seg_lo, min(seg_hi, S - current)) in slots current + i; range 2:
// segments [max(seg_lo, S - current), seg_hi) in slots i - (S - current).
struct TcSpan {
    int n1, blk1, sh1, pos1, n2, blk2, sh2, pos2;
    __host__ __device__ TcSpan(const TcArgs &a)
    {
        const int wrap = a.S - a.current;
        const int hi1 = a.seg_hi < wrap ? a.seg_hi : wrap;
        blk1 = (a.current + a.seg_lo) / TC_KSEG;
        n1 = hi1 > a.seg_lo ? (a.current + hi1 - 1) / TC_KSEG - blk1 + 1 : 0;
        sh1 = (a.current + a.seg_lo) & 1;
        pos1 = TC_LEAD + sh1 + blk1 * TC_KSEG - a.current - a.seg_lo; // even, >= 1
        const int lo2 = a.seg_lo > wrap ? a.seg_lo : wrap;
        blk2 = (lo2 - wrap) / TC_KSEG;
        n2 = a.seg_hi > lo2 ? (a.seg_hi - wrap - 1) / TC_KSEG - blk2 + 1 : 0;
        sh2 = (wrap + a.seg_lo) & 1;
        pos2 = TC_LEAD + sh2 + blk2 * TC_KSEG + wrap - a.seg_lo;
    }
    __host__ __device__ int per_input() const { return n1 + n2; }
    __host__ __device__ void chunk(int c, int &blk, int &copy, int &pos0) const
    {
        if (c < n1) {
            blk = blk1 + c;
            copy = sh1;
            pos0 = pos1 + c * TC_KSEG;
        } else {
            blk = blk2 + (c - n1);
            copy = sh2;
            pos0 = pos2 + (c - n1) * TC_KSEG;
        }
    }
};

template <int NOUT>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_mimo_tc(TcArgs a, const __grid_constant__ CUtensorMap tm_ring, const __grid_constant__ CUtensorMap tm_ir0,
          const __grid_constant__ CUtensorMap tm_ir1)
{
    using Cfg = TcCfg<NOUT>;
    constexpr int N2 = Cfg::N2;
    static_assert(TC_SPLIT_WARPS % 4 == 0 && Cfg::B_BYTES % (16 * 128) == 0, "splitter groups are 4 warps; B tile in 2 KB units");
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    // swizzled tiles need 1024-byte alignment; index the __shared__ array (not a uintptr_t round trip) so
    // every access below stays an LDS/STS instead of a generic LD/ST
    unsigned char *smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    unsigned char *smem_lo = smem + TC_NR * Cfg::RAW_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_lo + TC_NL * Cfg::LO_BYTES);
    uint64_t *full_raw = bars, *empty_raw = full_raw + TC_NR, *full_lo = empty_raw + TC_NR, *empty_lo = full_lo + TC_NL;
    uint64_t *acc_full = empty_lo + TC_NL, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int bid = blockIdx.x;
    const int sg = bid % a.stream_groups;
    bid /= a.stream_groups;
    const int og = bid % a.out_groups;
    bid /= a.out_groups;
    const int bin = bid / a.groups, g = bid % a.groups;
    const int in_lo = a.n_in * g / a.groups, in_hi = a.n_in * (g + 1) / a.groups;
    const TcSpan span(a);
    const int cpi = span.per_input();
    const int chunks = cpi * (in_hi - in_lo);             // 16-segment boxes of this CTA
    const int total = (chunks + TC_BOX - 1) / TC_BOX;     // pipeline stages

    if (tid == 0) {
        for (int s = 0; s < TC_NR; s++) {
            mbar_init(&full_raw[s], 1);
            mbar_init(&empty_raw[s], 128);
        }
        for (int s = 0; s < TC_NL; s++) {
            mbar_init(&full_lo[s], 128);
            mbar_init(&empty_lo[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer -------------------------------------------------------------------
        for (int t = 0; t < total; t++) {
            const int s = t % TC_NR;
            mbar_wait(&empty_raw[s], ((t / TC_NR) & 1) ^ 1);
            unsigned char *st = smem + s * Cfg::RAW_BYTES;
            if (elect_one()) {
                mbar_expect_tx(&full_raw[s], TC_BOX * (a.rows_pad * 128 + Cfg::B_BYTES));
#pragma unroll
                for (int bx = 0; bx < TC_BOX; bx++) {
                    const int q = t * TC_BOX + bx;
                    int in = in_lo, blk = a.stream_groups * a.nblk, copy = 0, pos0 = 1 << 20; // past both tensors: zero fill
                    if (q < chunks) {
                        in = in_lo + q / cpi;
                        span.chunk(q % cpi, blk, copy, pos0);
                        blk += sg * a.nblk;
                    }
                    tma_load_4d(st + bx * Cfg::BOX_RAW, &tm_ring, 0, 0, blk, bin * a.n_in + in, &full_raw[s]);
                    tma_load_3d(st + bx * Cfg::BOX_RAW + Cfg::A_BYTES, copy ? &tm_ir1 : &tm_ir0, 2 * pos0, og * N2, bin * a.n_in + in,
                                &full_raw[s]);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---- MMA issuer ---------------------------------------------------------------------
        constexpr uint32_t idesc_wide = umma_idesc_tf32(2 * N2), idesc_narrow = umma_idesc_tf32(N2);
        for (int t = 0; t < total; t++) {
            const int sl = t % TC_NL, iv = t / TC_DRAIN, b = iv & 1;
            const bool first = (t % TC_DRAIN) == 0, last = (t % TC_DRAIN) == TC_DRAIN - 1 || t == total - 1;
            if (first) mbar_wait(&acc_empty[b], ((iv >> 1) & 1) ^ 1);
            mbar_wait(&full_lo[sl], (t / TC_NL) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t lo = smem_u32(smem_lo + sl * Cfg::LO_BYTES);
                const uint32_t d_buf = tmem + b * Cfg::BUF_COLS;
#pragma unroll
                for (int bx = 0; bx < TC_BOX; bx++) {
                    const uint32_t a_hi = tmem + Cfg::ACC_COLS + sl * Cfg::A_COLS + bx * Cfg::BOX_COLS, a_lo = a_hi + 2 * TC_KSEG;
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        const uint64_t b_hl = umma_desc_k128(lo + bx * Cfg::BOX_LO + ks * 32);
                        // [main | cross] (+)= A_hi * [B_hi | B_lo]^T ;  cross += A_lo * B_hi^T
                        const uint32_t acc_on = (first && bx == 0 && ks == 0) ? 0u : 1u;
                        umma_tf32_ts(d_buf, a_hi + ks * 8, b_hl, idesc_wide, acc_on);
                        umma_tf32_ts(d_buf + N2, a_lo + ks * 8, b_hl, idesc_narrow, 1u);
                    }
                }
                umma_commit(&empty_lo[sl]);
                if (last) umma_commit(&acc_full[b]);
            }
            __syncwarp();
        }
    } else if (warp < 2 + TC_SPLIT_WARPS) {
        // ---- splitters: two groups of 4 warps take alternate stages (two stages in flight hide the
        // wait -> LDS -> convert -> STTM -> arrive chain); thread = one stream row.  hi = the 19 bits the
        // tensor core reads, lo = x - hi; A goes to TMEM, [B_hi | B_lo] to shared memory -----------
        const int grp = (warp - 2) >> 2, gtid = (tid - 64) & 127;
        const int row = (warp & 3) * 32 + lane; // TMEM lane quadrant of this warp
        const bool live = (warp & 3) * 32 < a.rows_pad; // quadrants above the stored stream rows hold nothing
        // lo = x - hi for two values at once: one LOP3 each for -hi = (x & 0xFFFFE000) ^ 0x80000000 and one packed
        // add.f32x2 (exact: hi shares x's exponent).  The tensor core truncates lo to 11 bits itself.  The splitter
        // warps' instruction count is what paces the kernel below 128 streams, hence 1.5 instead of 3 per value.
        auto lo2 = [](float x0, float x1, uint32_t &l0, uint32_t &l1) {
            const uint32_t n0 = (__float_as_uint(x0) & 0xFFFFE000u) ^ 0x80000000u, n1 = (__float_as_uint(x1) & 0xFFFFE000u) ^ 0x80000000u;
            asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
                : "=r"(l0), "=r"(l1)
                : "r"(__float_as_uint(x0)), "r"(__float_as_uint(x1)), "r"(n0), "r"(n1));
        };
        for (int t = grp; t < total; t += TC_SPLIT_WARPS / 4) {
            const int sr = t % TC_NR, sl = t % TC_NL;
            mbar_wait(&full_raw[sr], (t / TC_NR) & 1);
#pragma unroll
            for (int bx = 0; bx < TC_BOX; bx++) {
                const unsigned char *rawA = smem + sr * Cfg::RAW_BYTES + bx * Cfg::BOX_RAW + row * 128;
                const float4 *rawB = reinterpret_cast<const float4 *>(smem + sr * Cfg::RAW_BYTES + bx * Cfg::BOX_RAW + Cfg::A_BYTES);
                float4 *hiB = reinterpret_cast<float4 *>(smem_lo + sl * Cfg::LO_BYTES + bx * Cfg::BOX_LO);
                float4 *loB = reinterpret_cast<float4 *>(smem_lo + sl * Cfg::LO_BYTES + bx * Cfg::BOX_LO + Cfg::B_BYTES);
                float4 v[8], w[Cfg::B_BYTES / 16 / 128];
                if (live) {
#pragma unroll
                    for (int j = 0; j < 8; j++) // logical 16-byte chunk c of a row sits at chunk c ^ (row & 7)
                        v[j] = *reinterpret_cast<const float4 *>(rawA + ((j ^ (row & 7)) << 4));
                }
#pragma unroll
                for (int j = 0; j < Cfg::B_BYTES / 16 / 128; j++) w[j] = rawB[gtid + 128 * j];
                if (bx == TC_BOX - 1) mbar_arrive(&empty_raw[sr]); // the raw stage is in registers now
                if (bx == 0) {
                    mbar_wait(&empty_lo[sl], ((t / TC_NL) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + Cfg::ACC_COLS + sl * Cfg::A_COLS + bx * Cfg::BOX_COLS;
                if (live) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float4 x = v[4 * h + j];
                            hi[4 * j + 0] = __float_as_uint(x.x), hi[4 * j + 1] = __float_as_uint(x.y);
                            hi[4 * j + 2] = __float_as_uint(x.z), hi[4 * j + 3] = __float_as_uint(x.w);
                            lo2(x.x, x.y, lo[4 * j + 0], lo[4 * j + 1]);
                            lo2(x.z, x.w, lo[4 * j + 2], lo[4 * j + 3]);
                        }
                        tmem_st16(ta + 16 * h, hi);
                        tmem_st16(ta + 2 * TC_KSEG + 16 * h, lo);
                    }
                }
#pragma unroll
                for (int j = 0; j < Cfg::B_BYTES / 16 / 128; j++) {
                    hiB[gtid + 128 * j] = w[j];
                    uint4 l;
                    lo2(w[j].x, w[j].y, l.x, l.y);
                    lo2(w[j].z, w[j].w, l.z, l.w);
                    reinterpret_cast<uint4 *>(loB)[gtid + 128 * j] = l;
                }
            }
            if (live) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full_lo[sl]);
        }
    } else {
        // ---- drain warps: TMEM -> f32 registers every TC_DRAIN stages, then the partial spectra --
        const int q = warp & 3; // TMEM lane quadrant this warp may touch
        float acc[N2];
#pragma unroll
        for (int i = 0; i < N2; i++) acc[i] = 0.f;
        const int nint = (total + TC_DRAIN - 1) / TC_DRAIN;
        for (int iv = 0; iv < nint; iv++) {
            const int b = iv & 1;
            mbar_wait(&acc_full[b], (iv >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = tmem + ((uint32_t)(q * 32) << 16) + b * Cfg::BUF_COLS;
#pragma unroll
            for (int set = 0; set < TC_SETS; set++) {
#pragma unroll
                for (int c = 0; c < N2; c += 16) {
                    uint32_t m[16], x[16];
                    tmem_ld16(base + set * Cfg::SET_COLS + c, m);
                    tmem_ld16(base + set * Cfg::SET_COLS + N2 + c, x);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; i++) acc[c + i] += __uint_as_float(m[i]) + __uint_as_float(x[i]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[b]);
        }
        const int s = sg * TC_M + q * 32 + lane;
        if (s < a.n_streams) {
            float2 *dst = a.part + (((size_t)g * a.n_streams + s) * a.n_out + og * NOUT) * a.B + bin;
#pragma unroll
            for (int o = 0; o < NOUT; o++)
                if (og * NOUT + o < a.n_out) dst[(size_t)o * a.B] = make_float2(acc[o], acc[NOUT + o]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// K1 epilogue for the tensor-core layout: xcur [NS*IN][B] (packed spectra of the new block) -> slot
// `slot` of ring_t[bin][in][stream group][slot block][stream row][16]
__global__ void __launch_bounds__(256)
k_tc_scatter_ring(const float2 *__restrict__ xcur, float2 *__restrict__ ring_t, int B, int n_in, long long total, long long nblk, int slot,
                  int rows_pad, int stream_groups)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int bin = (int)(idx % B);
    const long long c = idx / B; // stream * IN + in
    const long long s = c / n_in, in = c % n_in;
    const long long tile = (((long long)bin * n_in + in) * stream_groups + s / TC_M) * nblk + slot / TC_KSEG;
    ring_t[(tile * rows_pad + s % TC_M) * TC_KSEG + slot % TC_KSEG] = xcur[idx];
}

// K5 epilogue: IR spectra of `npairs` (out, in) pairs starting at pair p0, src [npairs][rows][B] packed,
// -> the B operand rows of both copies of ir_t[copy][bin][in][2*OUT][2*posP]; segment row r sits at position
// TC_LEAD + copy + r
__global__ void __launch_bounds__(256)
k_tc_build_ir(const float2 *__restrict__ src, float *__restrict__ ir_t, int B, int n_in, int n_out, int rows, long long rowsP,
              long long p0, long long total, long long copy_stride)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int bin = (int)(idx % B);
    const long long r = idx / B;
    const int row = (int)(r % rows);
    const long long p = p0 + r / rows;
    const int out = (int)(p / n_in), in = (int)(p % n_in);
    const float2 h = src[idx];
    float2 re_row, im_row;
    if (bin == 0) { // {DC, Nyquist}: two real products
        re_row = make_float2(h.x, 0.f);
        im_row = make_float2(0.f, h.y);
    } else {
        re_row = make_float2(h.x, -h.y);
        im_row = make_float2(h.y, h.x);
    }
    // operand rows per (bin, in): out group og = out / 16 owns rows [32 og, 32 og + 32): 16 real-part rows, 16 imaginary
    const int rows_all = 32 * ((n_out + 15) / 16);
    const long long r_re = (out / 16) * 32 + out % 16, r_im = r_re + 16;
    float *base = ir_t + ((long long)bin * n_in + in) * rows_all * (2 * rowsP);
    *reinterpret_cast<float2 *>(base + r_re * (2 * rowsP) + 2 * (TC_LEAD + row)) = re_row;
    *reinterpret_cast<float2 *>(base + r_im * (2 * rowsP) + 2 * (TC_LEAD + row)) = im_row;
    base += copy_stride;
    *reinterpret_cast<float2 *>(base + r_re * (2 * rowsP) + 2 * (TC_LEAD + 1 + row)) = re_row;
    *reinterpret_cast<float2 *>(base + r_im * (2 * rowsP) + 2 * (TC_LEAD + 1 + row)) = im_row;
}

// conv[so][k] = sum over input groups of part[g][so][k], ascending g
__global__ void __launch_bounds__(256)
k_tc_reduce(const float2 *__restrict__ part, float2 *__restrict__ conv, long long n, int groups)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    float2 t = part[idx];
    for (int g = 1; g < groups; g++) {
        float2 q = part[(long long)g * n + idx];
        t.x += q.x;
        t.y += q.y;
    }
    conv[idx] = t;
}

} // namespace fcb
