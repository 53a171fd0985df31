// offline_kernels.cuh — multi-block calls: NB whole blocks of every channel in one pass.
//
// FFTConvolver::process accepts any call length (src/fft_convolver.rs:222); a caller that hands over
// several blocks at once (offline rendering, large host buffers) lets the engine look at time, too.
// For output block d the delay-line sum (:244-255) is
//     pre_multiplied_d = sum_{i=1}^{A-1} H_i * X_{t+d-i}
// — the same IR row H_i meets T consecutive input spectra for T consecutive output blocks.  A thread of
// k_mac_time keeps that window of T spectra bins in registers (it slides by one row per segment), so per
// segment it loads ONE IR row and ONE new spectrum row and does T complex MACs: T blocks for the HBM
// traffic of one.  Every accumulator still sums i = 1, 2, ... in ascending order with separately rounded
// multiplies and adds, so the output is bit-identical to NB single-block calls.
//
// Spectra of the call's own blocks (time index m >= 0) come from `xnew` [C][NB][B]; older ones (m < 0)
// from the ring slot (current - m) mod A, which this pass does not modify — the ring is updated by
// k_ring_update only after the MACs.  The NB inverse FFTs are independent (K3 in raw mode writes all 2B
// samples); k_ola_time then forms out_d = y_d[0..B) + y_{d-1}[B..2B) (overlap-add, :270-274, :283-284)
// with the epilogue, in parallel over d.
//
// Measured (B200, 4096 channels x 2 s IR x block 512, device buffers, profiles/r01_offline_multi_block.jsonl):
// 1 / 2 / 4 / 8 / 16 blocks per call = 49.4 / 82.7 / 143 / 165 / 173 k channel-seconds per second; a window of
// T = 4 (79 registers) with the second group of 4 re-reading IR rows from L2 beats T = 8 (165 registers: 144 k; 127
// registers with the prefetch bounded to 4 segments: 156 k) — at T = 4 the kernel is still HBM-bound (ncu: 6.6 TB/s,
// FMA pipe 40 %), at T = 8 the arithmetic takes over.
// The reference example's shape (mono, 64-sample blocks, 128 000 taps, 1000 blocks): 0.58 ms in one call vs
// 278 ms block by block, identical bits.
#pragma once

#include "fft_kernels.cuh"
#include "mac_kernels.cuh"

namespace fcb {

struct MacTimeArgs {
    const float2 *ir;    // [ir channel][S][B]
    long long ir_stride; // 0 when one IR is shared
    const float2 *ring;  // [C][S][B]
    long long ring_stride;
    const float2 *xnew;  // [C][NB][B]
    float2 *premul;      // [C][NB][B]
    int current, active, nblocks;
    long long nchan;
};

// spectrum of time index m (relative to the call's first block) for channel c, float4 index t4
template <int B>
__device__ __forceinline__ float4 time_src(const MacTimeArgs &a, long long c, int m, int t4)
{
    constexpr int ROW4 = B / 2;
    if (m >= a.nblocks) return make_float4(0.f, 0.f, 0.f, 0.f); // ragged last group: feeds no stored output
    if (m >= 0) return ld_stream4(reinterpret_cast<const float4 *>(a.xnew + (c * a.nblocks + m) * B) + t4);
    const int slot = (a.current - m) % a.active; // -m in [1, A-1]
    return ld_stream4(reinterpret_cast<const float4 *>(a.ring + c * a.ring_stride) + (long long)slot * ROW4 + t4);
}

// One thread owns 2 adjacent bins (one float4) of one channel for T consecutive output blocks; P = segments
// whose IR / spectrum rows are loaded ahead of their MACs (2*P 16-byte loads in flight per thread).
// grid.x = tiles * block groups * channel groups, block groups of one channel adjacent (IR rows re-read from L2).
template <int B, int T, int P>
__global__ void __launch_bounds__(256)
k_mac_time(MacTimeArgs a)
{
    static_assert(T % P == 0, "the window rotation is unrolled over T segments in steps of P");
    constexpr int ROW4 = B / 2;
    constexpr int TX = ROW4 < 256 ? ROW4 : 256;
    constexpr int TILES = ROW4 / TX;
    constexpr int CPB = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int ngq = (a.nblocks + T - 1) / T;
    long long bid = blockIdx.x;
    const int tile = (int)(bid % TILES);
    bid /= TILES;
    const int gq = (int)(bid % ngq);
    bid /= ngq;
    const long long c = bid * CPB + ty;
    if (c >= a.nchan) return;
    const int t4 = tile * TX + tx;
    const int d0 = gq * T;
    const bool packed = (t4 == 0);
    const float4 *ir = reinterpret_cast<const float4 *>(a.ir + c * a.ir_stride) + t4;
    const int A = a.active;

    float4 acc[T], w[T];
#pragma unroll
    for (int d = 0; d < T; d++) acc[d] = make_float4(0.f, 0.f, 0.f, 0.f);
    // window for segment i: output d needs X_{d0+d-i}, kept in w[(d - i) mod T]; (i0 - 1) % T == 0 below
#pragma unroll
    for (int d = 1; d < T; d++) w[d - 1] = time_src<B>(a, c, d0 + d - 1, t4);
    for (int i0 = 1; i0 < A; i0 += T) {
#pragma unroll
        for (int u0 = 0; u0 < T; u0 += P) {
            float4 h[P], e[P];
#pragma unroll
            for (int p = 0; p < P; p++) {
                const int i = i0 + u0 + p;
                if (i < A) {
                    h[p] = ld_stream4(ir + (long long)i * ROW4);
                    e[p] = time_src<B>(a, c, d0 - i, t4); // the spectrum that enters the window at segment i
                }
            }
#pragma unroll
            for (int p = 0; p < P; p++) {
                const int u = u0 + p, i = i0 + u;
                if (i < A) {
                    w[(T - 1 - u) % T] = e[p]; // (0 - i) mod T with i = i0 + u, (i0 - 1) % T == 0
#pragma unroll
                    for (int d = 0; d < T; d++) {
                        const float4 x = w[(d + T - 1 - u) % T]; // (d - i) mod T
                        cmac_ref(acc[d].x, acc[d].y, h[p].x, h[p].y, x.x, x.y, packed);
                        cmac_ref(acc[d].z, acc[d].w, h[p].z, h[p].w, x.z, x.w, false);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int d = 0; d < T; d++)
        if (d0 + d < a.nblocks) reinterpret_cast<float4 *>(a.premul + (c * a.nblocks + d0 + d) * B)[t4] = acc[d];
}

// ring slot of block d <- xnew[c][d] for the last min(NB, A) blocks (older ones would be overwritten anyway);
// slot_0 = current, slot_{d+1} = slot_d > 0 ? slot_d - 1 : A - 1  (src/fft_convolver.rs:287-291)
__global__ void __launch_bounds__(256)
k_ring_update(const float2 *__restrict__ xnew, float2 *__restrict__ ring, long long ring_stride, int B, int nblocks, int current,
              int active, int first_block, long long total)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x; // over [C][nblocks - first_block][B/2] float4
    if (idx >= total) return;
    const int row4 = B / 2;
    const int t4 = (int)(idx % row4);
    const long long r = idx / row4;
    const int nb = nblocks - first_block;
    const int d = first_block + (int)(r % nb);
    const long long c = r / nb;
    int slot = (current - d) % active;
    if (slot < 0) slot += active;
    reinterpret_cast<float4 *>(ring + c * ring_stride + (long long)slot * B)[t4] =
        reinterpret_cast<const float4 *>(xnew + (c * nblocks + d) * B)[t4];
}

// out[c][d*B + i] = epilogue(y_d[i] + (d == 0 ? overlap[c][i] : y_{d-1}[B + i]))
// y: [C][NB][2B] normalised inverse-FFT samples, one thread per output sample.  The new overlap
// (y_{NB-1}[B..2B)) is copied by the caller afterwards: here the old one is still being read.
__global__ void __launch_bounds__(256)
k_ola_time(const float *__restrict__ y, const float *__restrict__ overlap, float *__restrict__ out, long long out_stride, int B,
           int nblocks, long long total, fcb_epilogue epi)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x; // over [C][NB][B]
    if (idx >= total) return;
    const int i = (int)(idx % B);
    const long long r = idx / B;
    const int d = (int)(r % nblocks);
    const long long c = r / nblocks;
    const float *yd = y + (c * nblocks + d) * 2 * B;
    const float prev = d == 0 ? overlap[c * B + i] : yd[i - B]; // y_{d-1}[B + i] sits B floats before y_d[i]
    const float v = apply_epilogue(__fadd_rn(yd[i], prev), epi, c, d * B + i);
    out[c * out_stride + (long long)d * B + i] = v;
}

} // namespace fcb
