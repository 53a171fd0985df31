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

constexpr int RT_BINS = 32; // bins per CTA: one lane each
constexpr int RT_R = 4;     // segments per pipeline stage
constexpr int RT_NST = 3;   // stages

template <int WO, int WS>
struct RtCfg {
    static constexpr int OUTS = 8 * WO, STREAMS = 4 * WS, WARPS = WO * WS, THREADS = 32 * WARPS;
    static constexpr int ROW = 2 * RT_BINS;                       // floats of one (row, segment) piece
    static constexpr int H_FLOATS = OUTS * RT_R * ROW, X_FLOATS = STREAMS * RT_R * ROW;
    static constexpr int STAGE_BYTES = (H_FLOATS + X_FLOATS) * 4;
    static constexpr size_t SMEM = (size_t)RT_NST * STAGE_BYTES + 2 * RT_NST * sizeof(uint64_t);
    static constexpr int MIN_CTAS = 512 / THREADS > 8 ? 8 : 512 / THREADS; // <= 128 registers per thread
    static_assert(STAGE_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");
};

struct RtArgs {
    float2 *part;          // [Z][NS*OUT][B]
    int B, n_in, n_out, n_streams;
    int S, current;        // ring length and the slot of the current block
    int seg_lo, seg_hi;    // segments accumulated, seg_lo >= 1
    int seg_base;          // the segment at coordinate 0 of the IR tensor map
    int zchunks, zlen;     // segment chunks and their length
    int out_groups, stream_groups;
};

// acc[o][s] += h[o] * x[s] for RT_R (or cnt) segments of one stage.  PACKED: the CTA holds bin 0 = {DC, Nyquist},
// two real products for the one lane `p0` (x.y is replaced by 0 in both cross terms and x.x by x.y in the last one).
template <bool PACKED>
__device__ __forceinline__ void rt_row(float2 (&acc)[8][4], const float *hs, const float *xs, int r, bool p0)
{
    float2 x[4], h[8];
#pragma unroll
    for (int s = 0; s < 4; s++) x[s] = *reinterpret_cast<const float2 *>(xs + (s * RT_R + r) * (2 * RT_BINS));
#pragma unroll
    for (int o = 0; o < 8; o++) h[o] = *reinterpret_cast<const float2 *>(hs + (o * RT_R + r) * (2 * RT_BINS));
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
}

template <bool PACKED>
__device__ __forceinline__ void rt_stage(float2 (&acc)[8][4], const float *hs, const float *xs, int cnt, bool p0)
{
    if (cnt == RT_R) {
#pragma unroll
        for (int r = 0; r < RT_R; r++) rt_row<PACKED>(acc, hs, xs, r, p0);
    } else {
        for (int r = 0; r < cnt; r++) rt_row<PACKED>(acc, hs, xs, r, p0);
    }
}

template <int WO, int WS>
__global__ void __launch_bounds__(RtCfg<WO, WS>::THREADS, RtCfg<WO, WS>::MIN_CTAS)
k_mac_rt(RtArgs a, const __grid_constant__ CUtensorMap tm_ir, const __grid_constant__ CUtensorMap tm_ring)
{
    using Cfg = RtCfg<WO, WS>;
    extern __shared__ __align__(1024) unsigned char rt_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(rt_smem + RT_NST * Cfg::STAGE_BYTES), *empty = full + RT_NST;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = tile + TILES*(z + Z*(og + OG*sg))
    const int tiles = (a.B + RT_BINS - 1) / RT_BINS;
    int bid = blockIdx.x;
    const int tile = bid % tiles;
    bid /= tiles;
    const int z = bid % a.zchunks;
    bid /= a.zchunks;
    const int og = bid % a.out_groups, sg = bid / a.out_groups;
    const int lo = a.seg_lo + z * a.zlen;
    const int hi = (lo + a.zlen) < a.seg_hi ? (lo + a.zlen) : a.seg_hi;
    // slots run upwards with the segments and wrap once: runs [lo, w) and [w, hi)
    int w = lo + (a.S - (a.current + lo) % a.S);
    w = w < hi ? w : hi;
    const int nA = hi > lo ? (w - lo + RT_R - 1) / RT_R : 0, nB = hi > w ? (hi - w + RT_R - 1) / RT_R : 0;
    const int per_in = nA + nB, total = per_in * a.n_in;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RT_NST; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto first_seg = [&](int c) { return c < nA ? lo + c * RT_R : w + (c - nA) * RT_R; };
    // producer state (warp 0, kept by every lane; one elected lane issues): the next stage to fill is (p_in, p_c)
    int p_in = 0, p_c = 0, p_left = total;
    auto issue = [&](int s) {
        const int i0 = first_seg(p_c);
        if (elect_one()) {
            unsigned char *st = rt_smem + s * Cfg::STAGE_BYTES;
            mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
            tma_load_4d(st, &tm_ir, 2 * tile * RT_BINS, i0 - a.seg_base, p_in, og * Cfg::OUTS, &full[s]);
            int j0 = a.current + i0;
            j0 = j0 >= a.S ? j0 - a.S : j0; // current < S and i0 < S
            tma_load_4d(st + Cfg::H_FLOATS * 4, &tm_ring, 2 * tile * RT_BINS, j0, p_in, sg * Cfg::STREAMS, &full[s]);
        }
        __syncwarp();
        if (++p_c == per_in) {
            p_c = 0;
            p_in++;
        }
        p_left--;
    };
    if (warp == 0)
        for (int s = 0; s < RT_NST && p_left > 0; s++) issue(s);

    const int wo = warp / WS, ws = warp % WS;
    const bool p0 = tile == 0 && lane == 0;
    const float *hs0 = reinterpret_cast<const float *>(rt_smem) + wo * 8 * RT_R * Cfg::ROW + 2 * lane;
    const float *xs0 = reinterpret_cast<const float *>(rt_smem) + Cfg::H_FLOATS + ws * 4 * RT_R * Cfg::ROW + 2 * lane;
    float2 acc[8][4];
#pragma unroll
    for (int o = 0; o < 8; o++)
#pragma unroll
        for (int s = 0; s < 4; s++) acc[o][s] = make_float2(0.f, 0.f);

    int s = 0, ph = 0, c = 0, prev_s = 0, prev_ph = 0; // stage and phase of this iteration and of the one before
    for (int it = 0; it < total; it++) {
        if (warp == 0 && it >= 1 && p_left > 0) { // refill the stage every warp left at it - 1
            mbar_wait(&empty[prev_s], prev_ph);
            issue(prev_s);
        }
        mbar_wait(&full[s], ph);
        const int i0 = first_seg(c), end = c < nA ? w : hi;
        const int cnt = (end - i0) < RT_R ? (end - i0) : RT_R;
        const float *hs = hs0 + s * (Cfg::STAGE_BYTES / 4), *xs = xs0 + s * (Cfg::STAGE_BYTES / 4);
        if (tile == 0) rt_stage<true>(acc, hs, xs, cnt, p0);
        else rt_stage<false>(acc, hs, xs, cnt, false);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        prev_s = s;
        prev_ph = ph;
        if (++s == RT_NST) {
            s = 0;
            ph ^= 1;
        }
        if (++c == per_in) c = 0;
    }

    const int k = tile * RT_BINS + lane;
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
