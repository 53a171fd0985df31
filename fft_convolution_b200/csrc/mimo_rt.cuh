// mimo_rt.cuh — the convolution-matrix delay-line MAC for a FEW streams (2 .. ~48) sharing the IR matrix:
// a register-tiled per-bin complex GEMM on the FP32 pipes.
//
// Same contraction as K4 (mimo_tc.cuh), i.e. the loops at src/fft_convolver.rs:244-255 of all OUT*IN reference
// convolvers of every stream at once, per bin k
//     D[s][o] = sum_in sum_{i >= 1}  X[s][in][(current+i) % S][k] * H[o][in][i][k]          (complex)
// Below ~48 streams the tensor-core kernel sits on a floor set by its hand-off chain (profiles/r01_k4_notes.txt:
// 0.94 ms from 8 to 32 streams) while the arithmetic itself is small: 16 streams x 16 x 16 x 937 segments x 513 bins
// = 2.0 G complex MACs = 0.21 ms of the chip's FFMA rate, next to 0.30 ms of HBM time for the operands (IR matrix
// 985 MB + 16 rings of 61.5 MB).  Bins are independent GEMMs, so no operand is ever shared BETWEEN lanes:
//     lane  = one bin;  thread = 8 outputs x 4 streams of it (32 complex accumulators in registers);
//     warp  = 32 adjacent bins x (8 outputs x 4 streams);  CTA = WO x WS warps = 8*WO outputs x 4*WS streams
// and a K step (one segment of one input) costs a thread 12 eight-byte shared-memory loads for 128 FFMAs.  The
// operands use the plain layouts of the other kernels (ir [OUT*IN][rows][B], ring [NS*IN][S][B], no transposed
// copies): one pipeline stage is two TMA boxes, [8*WO outputs][RT_R segments][32 bins] of the IR matrix and
// [4*WS streams][RT_R slots][32 bins] of the rings, out-of-range outputs / streams / bins zero-filled by the TMA
// unit.  The ring wrap never falls inside a box: the segment range of a CTA is cut at the wrap into two runs and
// the last box of a run is consumed only up to the run's end.
// The CTA walks every input and its chunk of the segments, so the partial spectra carry no input dimension:
//     part[z][stream][out][B]      (k_mimo_reduce adds the z chunks and the segment-0 products)
// FMA arithmetic, re-associated sums: parity with the CPU restatement is by tolerance (1e-5 x RMS) as for every matrix kernel.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mimo_tc.cuh"

namespace fcb {

// Packed FP32 (fma.rn.f32x2, FFMA2) halves the issue slots but not the time: measured on the B200 the FFMA2 form of this
// loop is math-pipe throttled at one FFMA2 per 4 cycles per scheduler — the same FLOP rate the scalar FFMAs reach — and
// 16 streams take 0.420 ms instead of 0.406 (profiles/r02_k_mac_rt_notes.txt).  Kept for the record, off.
#ifndef RT_FFMA2
#define RT_FFMA2 0
#endif
constexpr int RT_BINS = 32; // bins per warp: one lane each
// segments per pipeline stage R and stages NST = 12 / R: the same bytes in flight as 3 stages of 4 or 6 stages of 2

template <int WO, int WS, int WB, int R_>
struct RtCfg {
    static constexpr int R = R_, NST = 12 / R_;
    static constexpr int OUTS = 8 * WO, STREAMS = 4 * WS, WARPS = WO * WS * WB, THREADS = 32 * WARPS;
    static constexpr int BINS = RT_BINS * WB;                     // bins per CTA: WB warps side by side
    static constexpr int ROW = 2 * BINS;                          // floats of one (row, segment) piece
    static constexpr int H_FLOATS = OUTS * R * ROW, X_FLOATS = STREAMS * R * ROW;
    static constexpr int STAGE_BYTES = (H_FLOATS + X_FLOATS) * 4;
    static constexpr size_t SMEM = (size_t)NST * STAGE_BYTES + NST * (sizeof(uint64_t) + sizeof(unsigned int));
    static constexpr int MIN_CTAS = 512 / THREADS > 8 ? 8 : 512 / THREADS; // <= 128 registers per thread
    static_assert(STAGE_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");
};

struct RtArgs {
    float2 *part;          // [Z][NS*OUT][B]
    int B, n_in, n_out, n_streams;
    int S, current;        // ring length and the slot of the current block
    int seg_lo, seg_hi;    // segments accumulated, seg_lo >= 1
    int seg_base;          // the segment at coordinate 0 of the IR tensor map
    int zchunks;           // segment chunks, cut in units of R segments (see the kernel)
    int out_groups, stream_groups;
};

// acc[o][s] += h[o] * x[s] for RT_R (or cnt) segments of one stage.  PACKED: the CTA holds bin 0 = {DC, Nyquist},
// two real products for the one lane `p0` (x.y is replaced by 0 in both cross terms and x.x by x.y in the last one).
template <bool PACKED, int ROW, int RT_R>
__device__ __forceinline__ void rt_row(float2 (&acc)[8][4], const float *hs, const float *xs, int r, bool p0)
{
    float2 x[4], h[8];
#pragma unroll
    for (int s = 0; s < 4; s++) x[s] = *reinterpret_cast<const float2 *>(xs + (s * RT_R + r) * ROW);
#pragma unroll
    for (int o = 0; o < 8; o++) h[o] = *reinterpret_cast<const float2 *>(hs + (o * RT_R + r) * ROW);
#if RT_FFMA2
    // packed FP32 (fma.rn.f32x2: one issue slot for the two FMAs of an accumulator's (re, im) pair):
    //     acc += (x.x, x.x) * (h.x, h.y);   acc += (x.y, x.y) * (-h.y, h.x)
    // FFMA2 takes a broadcast scalar, a swapped pair and a per-half negation as operand forms, so no operand needs
    // preparing.  Bin-0 lane: (x.x, x.y) * (h.x, h.y) and nothing else.
    float2 xa[4], xb[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        xa[s] = make_float2(x[s].x, (PACKED && p0) ? x[s].y : x[s].x);
        const float t = (PACKED && p0) ? 0.f : x[s].y;
        xb[s] = make_float2(t, t);
    }
#pragma unroll
    for (int o = 0; o < 8; o++) {
        const float2 hj = make_float2(-h[o].y, h[o].x);
#pragma unroll
        for (int s = 0; s < 4; s++) {
            acc[o][s] = __ffma2_rn(xa[s], h[o], acc[o][s]);
            acc[o][s] = __ffma2_rn(xb[s], hj, acc[o][s]);
        }
    }
#else
    float xt[4], xq[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        xt[s] = (PACKED && p0) ? 0.f : x[s].y;
        xq[s] = (PACKED && p0) ? x[s].y : x[s].x;
    }
#pragma unroll
    for (int o = 0; o < 8; o++)
#pragma unroll
        for (int s = 0; s < 4; s++) {
            acc[o][s].x = fmaf(h[o].x, x[s].x, fmaf(-h[o].y, xt[s], acc[o][s].x));
            acc[o][s].y = fmaf(h[o].x, xt[s], fmaf(h[o].y, xq[s], acc[o][s].y));
        }
#endif
}

template <int WO, int WS, int WB, int R>
__global__ void __launch_bounds__(RtCfg<WO, WS, WB, R>::THREADS, RtCfg<WO, WS, WB, R>::MIN_CTAS)
k_mac_rt(RtArgs a, const __grid_constant__ CUtensorMap tm_ir, const __grid_constant__ CUtensorMap tm_ring)
{
    using Cfg = RtCfg<WO, WS, WB, R>;
    extern __shared__ __align__(1024) unsigned char rt_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(rt_smem + Cfg::NST * Cfg::STAGE_BYTES);
    unsigned int *left = reinterpret_cast<unsigned int *>(full + Cfg::NST); // warps that have left a stage

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = sg + SG*(og + OG*(tile + TILES*z)): the CTAs that read the same IR tiles (stream groups) and the same
    // ring tiles (output groups) are neighbours in launch order, so the second reader finds them in L2
    const int tiles = (a.B + Cfg::BINS - 1) / Cfg::BINS;
    int bid = blockIdx.x;
    const int sg = bid % a.stream_groups;
    bid /= a.stream_groups;
    const int og = bid % a.out_groups;
    bid /= a.out_groups;
    const int tile = bid % tiles, z = bid / tiles;
    // Chunk z owns the R-segment units [U*z/Z, U*(z+1)/Z) of the range: every chunk but the last is a whole number of
    // boxes long, and the last one ends where the IR tensor ends.
    const int units = (a.seg_hi - a.seg_lo + R - 1) / R;
    const int lo = a.seg_lo + R * (int)((long long)units * z / a.zchunks);
    int hi = a.seg_lo + R * (int)((long long)units * (z + 1) / a.zchunks);
    hi = hi < a.seg_hi ? hi : a.seg_hi;
    // Slots run upwards with the segments and wrap once inside [lo, hi) at most: run A = [lo, w) takes its boxes from
    // lo upwards, run B = [w, hi) from hi downwards.  Every box is a full R rows; the rows a box has too many are rows
    // the TMA unit fills with zeros in one of the two operands, so they add nothing: past the end of run A the ring
    // slots are >= S, before the start of run B they are < 0, and past the end of the last chunk the IR rows are past
    // the tensor.
    int w = lo + (a.S - (a.current + lo) % a.S);
    w = w < hi ? w : hi;
    const int nA = hi > lo ? (w - lo + R - 1) / R : 0, nB = hi > w ? (hi - w + R - 1) / R : 0;
    const int per_in = nA + nB, total = per_in * a.n_in;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::NST; s++) {
            mbar_init(&full[s], 1);
            left[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // fill stage s with box c of input `in` (one thread)
    auto issue = [&](int s, int in, int c) {
        const int i0 = c < nA ? lo + c * R : hi - (per_in - c) * R;
        const int j0 = c < nA ? (a.current + lo) % a.S + c * R : i0 - w;
        unsigned char *st = rt_smem + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        tma_load_4d(st, &tm_ir, 2 * tile * Cfg::BINS, i0 - a.seg_base, in, og * Cfg::OUTS, &full[s]);
        tma_load_4d(st + Cfg::H_FLOATS * 4, &tm_ring, 2 * tile * Cfg::BINS, j0, in, sg * Cfg::STREAMS, &full[s]);
    };
    // the box NST iterations ahead of the current one (kept by every thread; whoever refills a stage uses it)
    int n_in = 0, n_c = 0;
    auto advance = [&]() {
        if (++n_c == per_in) {
            n_c = 0;
            n_in++;
        }
    };
    for (int s = 0; s < Cfg::NST; s++) {
        if (threadIdx.x == 0 && s < total) issue(s, n_in, n_c);
        advance();
    }

    const int wb = warp % WB, wo = (warp / WB) / WS, ws = (warp / WB) % WS;
    const bool p0 = tile == 0 && wb == 0 && lane == 0;
    const float *hs0 = reinterpret_cast<const float *>(rt_smem) + wo * 8 * R * Cfg::ROW + 2 * (wb * RT_BINS + lane);
    const float *xs0 = reinterpret_cast<const float *>(rt_smem) + Cfg::H_FLOATS + ws * 4 * R * Cfg::ROW + 2 * (wb * RT_BINS + lane);
    float2 acc[8][4];
#pragma unroll
    for (int o = 0; o < 8; o++)
#pragma unroll
        for (int s = 0; s < 4; s++) acc[o][s] = make_float2(0.f, 0.f);

    // No producer warp and no warp waiting for the others: the LAST warp to leave a stage refills it (a counter in shared
    // memory, one atomic per warp and stage), so the load for iteration it + NST starts the moment stage it % NST is free.
    int s = 0, ph = 0;
    for (int it = 0; it < total; it++) {
        mbar_wait(&full[s], ph);
        const float *hs = hs0 + s * (Cfg::STAGE_BYTES / 4), *xs = xs0 + s * (Cfg::STAGE_BYTES / 4);
        if (tile == 0 && wb == 0) {
#pragma unroll
            for (int r = 0; r < R; r++) rt_row<true, Cfg::ROW, R>(acc, hs, xs, r, p0);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) rt_row<false, Cfg::ROW, R>(acc, hs, xs, r, false);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block(); // this warp's reads of the stage are done before the count says so
            if (atomicAdd(&left[s], 1u) == Cfg::WARPS - 1) {
                left[s] = 0; // nobody counts on this stage again before the refill below has landed
                __threadfence_block();
                if (it + Cfg::NST < total) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(s, n_in, n_c);
                }
            }
        }
        advance();
        if (++s == Cfg::NST) {
            s = 0;
            ph ^= 1;
        }
    }

    const int k = tile * Cfg::BINS + wb * RT_BINS + lane;
    if (k < a.B) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int st = sg * Cfg::STREAMS + ws * 4 + s;
            if (st >= a.n_streams) continue;
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const int out = og * Cfg::OUTS + wo * 8 + o;
                if (out < a.n_out)
                    a.part[(((long long)z * a.n_streams + st) * a.n_out + out) * a.B + k] = acc[o][s];
            }
        }
    }
}

} // namespace fcb
