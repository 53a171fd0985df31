// fused_kernel.cuh — one kernel per whole block (n == B, fill == 0) for B <= 512:
// K1 (forward FFT of the new block into ring[current]), K2 (delay-line MAC, TMA pipeline) and K3
// (segment-0 MAC, inverse FFT, /N, overlap-add, epilogue, overlap save) in the same CTA.
//
// Why: K1 and K3 are latency-bound (a dependent chain of FFT passes) and, launched on their own,
// cost ~5 % of the step.  Here a CTA starts its TMA pipeline first, does the forward FFT while the
// first stages land, and runs the inverse FFT after its stream ends while the co-resident CTA keeps
// the HBM pipes busy; the step becomes one launch and `pre_multiplied` never leaves the chip.
// Same arithmetic, same order as the separate kernels (src/fft_convolver.rs:234-284): outputs are
// bit-identical to the K1 -> K2 -> K3 sequence.
#pragma once

#include "fft_kernels.cuh"
#include "mac_kernels.cuh"

namespace fcb {

// Small batches: with fewer channel groups than SMs one CTA per group would stream a whole delay line by itself
// (770 KB for a 2 s response) while most of the machine idles.  `split` cuts the segment range into zsplit slices of
// zlen segments, one CTA each (grid = groups * zsplit).  Slice 0 also runs K1 and writes ring[current]; every CTA
// publishes its partial pre_multiplied (one float4 per thread) and the LAST one to arrive (one atomic per CTA) adds the
// zsplit partials in slice order — deterministic — and runs K3.  zsplit == 1 is the unsplit kernel, bit for bit.
struct SplitArgs {
    int zsplit, zlen;        // slices per channel group, segments per slice (a multiple of the stage rows)
    float4 *part;            // [groups][zsplit][arrays][256] partial sums
    unsigned int *count;     // [groups] arrival counters, zero between launches
};

struct FusedArgs {
    const float *in;      // [C][B] new block, channel stride in_stride
    long long in_stride;
    MacArgs mac;          // ir / ring / strides / current / active / nchan (premul unused)
    IfftArgs ifft;        // overlap / out / out_stride / epilogue (fill = 0, n = B, complete)
    SplitArgs split;
    int k1_late;          // TMA_IO only: run K1 after the MAC stream (the input block is still crossing PCIe)
};

// publish this CTA's partial accumulator(s); returns true in the one CTA of the group that arrives last
template <int NACC>
__device__ __forceinline__ bool split_arrive(const SplitArgs &sp, long long group, int z, const float4 (&acc)[NACC])
{
    __shared__ int s_last;
    float4 *mine = sp.part + ((size_t)(group * sp.zsplit + z) * NACC) * 256;
#pragma unroll
    for (int n = 0; n < NACC; n++) mine[n * 256 + threadIdx.x] = acc[n];
    __threadfence(); // partials (and slice 0's ring row) before the arrival
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(sp.count + group, 1u);
        s_last = prev == (unsigned int)(sp.zsplit - 1);
        if (s_last) sp.count[group] = 0; // ready for the next launch
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// sum of the zsplit partials in slice order (every CTA of the group has arrived)
template <int NACC>
__device__ __forceinline__ void split_gather(const SplitArgs &sp, long long group, float4 (&acc)[NACC])
{
    const float4 *base = sp.part + (size_t)group * sp.zsplit * NACC * 256;
#pragma unroll
    for (int n = 0; n < NACC; n++) acc[n] = __ldcg(base + n * 256 + threadIdx.x);
    for (int z = 1; z < sp.zsplit; z++) {
#pragma unroll
        for (int n = 0; n < NACC; n++) {
            const float4 p = __ldcg(base + ((size_t)z * NACC + n) * 256 + threadIdx.x);
            acc[n].x = __fadd_rn(acc[n].x, p.x);
            acc[n].y = __fadd_rn(acc[n].y, p.y);
            acc[n].z = __fadd_rn(acc[n].z, p.z);
            acc[n].w = __fadd_rn(acc[n].w, p.w);
        }
    }
}

template <int LOGB, int ROWS = 4>
struct FusedCfg {
    static constexpr int B = 1 << LOGB;
    static constexpr int CPB = 512 / B;              // channels per CTA (B <= 512)
    static constexpr int R = ROWS;                   // rows (segments) per stage
    static constexpr int ARR = CPB * R * B;          // float2 per array per stage (= 2048)
    static constexpr size_t STAGE_BYTES = 2 * (size_t)ARR * sizeof(float2);
    static constexpr int FFT_PER = (sidx(B) + 2) & ~1; // float2 per transform buffer (even: rows stay 16-byte aligned)
    static constexpr size_t FFT_BYTES = (size_t)CPB * FFT_PER * sizeof(float2);
    static constexpr size_t FFT_BYTES_AL = ((FFT_BYTES + 15) / 16) * 16;
    static constexpr size_t IO_BYTES = (size_t)CPB * B * sizeof(float); // one block of samples per channel
    __host__ __device__ static constexpr size_t smem_bytes(int nst, bool tma_io = false)
    {
        return nst * STAGE_BYTES + 64 + FFT_BYTES_AL + (tma_io ? 2 * IO_BYTES : 0);
    }
};

// shared -> global bulk store (async proxy); the caller fences its generic-proxy smem writes first
__device__ __forceinline__ void bulk_s2g(void *dst_global, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}

// TMA_IO: the input block comes in and the output block goes out as ONE bulk copy per channel
// (cp.async.bulk global<->shared).  Used when the caller's buffers are pinned HOST memory: a 2 KB
// burst per channel crosses PCIe instead of hundreds of 4-byte accesses, so the kernel can work on
// host buffers directly (no copy engines, no staging, one launch per step).  Same arithmetic.
template <int LOGB, int NST, int ROWS = 4, bool TMA_IO = false>
__global__ void __launch_bounds__(256)
k_block_fused(FusedArgs fa, const float2 *__restrict__ tw)
{
    using Cfg = FusedCfg<LOGB, ROWS>;
    using P = FftPlan<LOGB>;
    constexpr int B = Cfg::B, CPB = Cfg::CPB, R = Cfg::R, ARR = Cfg::ARR;
    constexpr int TX = B / 2;             // MAC threads along a row (float4 each)
    constexpr int T = P::T, E = P::E;     // FFT threads per transform, points per thread
    static_assert(CPB * T <= 256, "FFT lanes must fit the CTA");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *stages = reinterpret_cast<float2 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NST * Cfg::STAGE_BYTES); // [NST] + in_bar at [7]
    float2 *fbuf = reinterpret_cast<float2 *>(smem_raw + NST * Cfg::STAGE_BYTES + 64);
    float *in_s = reinterpret_cast<float *>(smem_raw + NST * Cfg::STAGE_BYTES + 64 + Cfg::FFT_BYTES_AL); // TMA_IO only
    float *out_s = in_s + CPB * B;
    uint64_t *in_bar = full + 7;
    static_assert(NST <= 7, "barrier slots");

    const MacArgs &a = fa.mac;
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;                 // MAC role
    const int fslot = tid / T, flane = tid % T;             // FFT role (threads 0 .. CPB*T-1)
    const int Z = fa.split.zsplit > 1 ? fa.split.zsplit : 1;
    const long long group = Z > 1 ? blockIdx.x / Z : blockIdx.x;
    const int zs = Z > 1 ? (int)(blockIdx.x % Z) : 0;
    const bool k1 = zs == 0;                                // this CTA transforms the new block
    const bool fwork = tid < CPB * T && k1;
    const long long c0 = group * CPB;
    const int nlive = (int)((a.nchan - c0) < CPB ? (a.nchan - c0) : CPB);
    const int cur = a.current, act = a.active;
    const int lo = Z > 1 ? a.seg_lo + zs * fa.split.zlen : a.seg_lo;
    const int hi = Z > 1 ? (lo + fa.split.zlen < a.seg_hi ? lo + fa.split.zlen : a.seg_hi) : a.seg_hi;
    const int niter = hi > lo ? (hi - lo + R - 1) / R : 0;

    if (tid == 0) {
        for (int s = 0; s < NST; s++) mbar_init(&full[s], 1);
        if (TMA_IO) mbar_init(in_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (TMA_IO && tid == 0 && k1) { // the new block of every live channel: one bulk copy each (host or device memory)
        mbar_expect_tx(in_bar, (uint32_t)(nlive * B * sizeof(float)));
        for (int ch = 0; ch < nlive; ch++)
            bulk_g2s(in_s + ch * B, fa.in + (c0 + ch) * fa.in_stride, B * sizeof(float), in_bar);
    }

    auto issue = [&](int it) {
        const int s = it % NST;
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        float2 *ir_s = stages + (size_t)s * 2 * ARR;
        float2 *rg_s = ir_s + ARR;
        mbar_expect_tx(&full[s], (uint32_t)(2 * nlive * cnt * B * sizeof(float2)));
        const int j0 = (cur + i0) % act;
        const int first = (act - j0) < cnt ? (act - j0) : cnt;
        for (int ch = 0; ch < nlive; ch++) {
            const float2 *irc = a.ir + a.ir_chan(c0 + ch) * a.ir_stride - (long long)a.ir_seg0 * B;
            const float2 *rgc = a.ring + a.ring_chan(c0 + ch) * a.ring_stride;
            bulk_g2s(ir_s + ch * R * B, irc + (long long)i0 * B, cnt * B * sizeof(float2), &full[s]);
            bulk_g2s(rg_s + ch * R * B, rgc + (long long)j0 * B, first * B * sizeof(float2), &full[s]);
            if (first < cnt) bulk_g2s(rg_s + ch * R * B + first * B, rgc, (cnt - first) * B * sizeof(float2), &full[s]);
        }
    };
    // the MAC stream never touches ring[current]: start it before the forward FFT
    if (tid == 0)
        for (int it = 0; it < NST && it < niter; it++) issue(it);

    // segment-0 IR bins of this thread, fetched now so the latency hides under the whole stream
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool mlive = ty < nlive;
    if (mlive) h0 = __ldg(reinterpret_cast<const float4 *>(a.ir + a.ir_chan(c0 + ty) * a.ir_stride + (long long)(0 - a.ir_seg0) * B) + tx);

    // ---- K1: forward real FFT of the new block (src/fft_convolver.rs:234-241) -------------------
    // K2 below only reads ring slots OLDER than the current block, so K1 may run before or after it.  Before is the
    // default (its latency hides under the first stages' flight).  With the block still crossing PCIe (zero-copy on
    // pinned host buffers, fa.k1_late) K1 runs AFTER the MAC stream: the pull that was issued at the top of the kernel has
    // the whole stream to land instead of stalling every CTA at its start.
    float2 *fs = fbuf + (fwork ? fslot : 0) * Cfg::FFT_PER;
    bool flive = fwork && fslot < nlive;
    auto k1_stage = [&]() {
        if (TMA_IO && k1) mbar_wait(in_bar, 0);
        if (fwork) {
            const float *x = TMA_IO ? in_s + fslot * B : fa.in + (c0 + fslot) * fa.in_stride;
            if (TMA_IO) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    int j = flane + e * T;
                    float2 z = make_float2(0.f, 0.f);
                    if (flive && 2 * j + 1 < B) z = *reinterpret_cast<const float2 *>(x + 2 * j);
                    fs[sidx(j)] = z;
                }
            } else {
                load_block_as_complex<LOGB>(fs, flane, x, flive ? B : 0); // 16-byte loads when the row is aligned
            }
        }
        __syncthreads();
        stockham_all<LOGB, -1, 0, 1>(fs, flane, tw, fwork);
        // split -> packed spectrum; written to ring[current] and kept in shared memory for segment 0
        float2 xk[E];
        if (fwork) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                int k = flane + e * T;
                xk[e] = rfft_split_bin<LOGB>(fs, k, tw);
            }
        }
        __syncthreads(); // every Z[k], Z[B-k] has been read
        if (fwork) {
            float2 *row = flive ? const_cast<float2 *>(a.ring) + a.ring_chan(c0 + fslot) * a.ring_stride + (long long)cur * B : nullptr;
#pragma unroll
            for (int e = 0; e < E; e++) {
                int k = flane + e * T;
                fs[k] = xk[e]; // unpadded packed row: the MAC threads read it as float4
                if (flive) row[k] = xk[e];
            }
        }
        __syncthreads();
    };
    const bool k1_late = TMA_IO && fa.k1_late != 0;
    if (!k1_late) k1_stage();

    // ---- K2: delay-line MAC over segments lo..hi-1 (src/fft_convolver.rs:244-255) ---------------
    const bool packed = (tx == 0);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < niter; it++) {
        const int s = it % NST;
        mbar_wait(&full[s], (it / NST) & 1);
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        const float4 *ir_s = reinterpret_cast<const float4 *>(stages + (size_t)s * 2 * ARR + ty * R * B) + tx;
        const float4 *rg_s = ir_s + ARR / 2;
        if (mlive) {
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
        __syncthreads();
        if (tid == 0 && it + NST < niter) issue(it + NST);
    }

    if (k1_late) k1_stage();

    // ---- small batches: the delay line was cut over Z CTAs; the last one to arrive finishes the block ----
    bool fwork3 = fwork;
    if (Z > 1) {
        float4 part[1] = {acc};
        if (!split_arrive<1>(fa.split, group, zs, part)) return;
        split_gather<1>(fa.split, group, part);
        acc = part[0];
        fwork3 = tid < CPB * T;
        flive = fwork3 && fslot < nlive;
        fs = fbuf + (fwork3 ? fslot : 0) * Cfg::FFT_PER;
    }

    // ---- K3: conv = pre_multiplied + X[current] * H[0] (:256-261), inverse FFT, overlap-add ------
    float4 conv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mlive) {
        // X[current]: this CTA's own transform, or (split, another CTA did K1) the ring row slice 0 wrote before it arrived
        const float4 x = Z > 1 && !k1
                             ? __ldcg(reinterpret_cast<const float4 *>(a.ring + a.ring_chan(c0 + ty) * a.ring_stride + (long long)cur * B) + tx)
                             : reinterpret_cast<const float4 *>(fbuf + ty * Cfg::FFT_PER)[tx];
        conv = acc;
        // the reference multiplies segments[current] (a) by segments_ir[0] (b): im = a.re*b.im + a.im*b.re
        cmac_ref(conv.x, conv.y, x.x, x.y, h0.x, h0.y, packed);
        cmac_ref(conv.z, conv.w, x.z, x.w, h0.z, h0.w, false);
    }
    __syncthreads(); // all X[current] rows consumed before the buffer is reused
    if (mlive) {
        float2 *d = fbuf + ty * Cfg::FFT_PER;
        d[sidx(2 * tx)] = make_float2(conv.x, conv.y);
        d[sidx(2 * tx + 1)] = make_float2(conv.z, conv.w);
    } else if (ty < CPB) {
        float2 *d = fbuf + ty * Cfg::FFT_PER;
        d[sidx(2 * tx)] = make_float2(0.f, 0.f);
        d[sidx(2 * tx + 1)] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    if (fwork3) irfft_presplit<LOGB>(fs, flane, tw);
    __syncthreads();
    stockham_all<LOGB, +1, 0, 1>(fs, flane, tw, fwork3);

    const IfftArgs &o = fa.ifft;
    const float inv_n = 1.0f / (float)(2 * B);
    const long long c = c0 + fslot;
    if (flive) {
#pragma unroll
        for (int e = 0; e < E; e++) {
            int j = flane + e * T;
            if (2 * j >= B) continue;
            float2 z = fs[sidx(j)];
            float y[2] = {z.x * inv_n, z.y * inv_n};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int i = 2 * j + h;
                float v = apply_epilogue(__fadd_rn(y[h], o.overlap[c * B + i]), o.epi, c, i);
                if (TMA_IO) out_s[fslot * B + i] = v;
                else o.out[c * o.out_stride + i] = v;
            }
        }
    }
    if (TMA_IO) asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // smem writes -> async proxy
    __syncthreads(); // every reader of the old overlap is done (and every output sample is in out_s)
    if (TMA_IO && tid == 0) {
        for (int ch = 0; ch < nlive; ch++)
            bulk_s2g(o.out + (c0 + ch) * o.out_stride, out_s + ch * B, B * sizeof(float));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (flive) {
#pragma unroll
        for (int e = 0; e < E; e++) {
            int j = flane + e * T;
            if (2 * j >= B) {
                float2 z = fs[sidx(j)];
                *reinterpret_cast<float2 *>(o.overlap + c * B + (2 * j - B)) = make_float2(z.x * inv_n, z.y * inv_n);
            }
        }
    }
    if (TMA_IO && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // smem stays alive until read
}

// ---------------------------------------------------------------------------------------------
// Pair variant: TWO convolvers that are fed the same input — TwoStageFFTConvolver's head and
// tail_convolver0 (src/fft_convolver.rs:417, :464-473) or CrossfadeConvolver's A and B
// (src/crossfade_convolver.rs:72-73) — hold identical input-spectrum rings, so one CTA does ONE forward FFT,
// streams the ring rows ONCE next to both IR row sets, and finishes with both inverse FFTs side by side
// (threads 0..63 convolver A, 64..127 convolver B).  Per channel and block the stream is 24*S*K bytes instead
// of 32*S*K and the pair is one launch instead of two.  Each convolver's sums are formed exactly as in
// k_block_fused (same operands, same order): outputs are bit-identical to two separate launches.  The new
// spectrum is written into BOTH rings, so either convolver can continue on its own afterwards.
// B's epilogue may mix in A's output of this very block (crossfade): A's samples are stored first, and B reads
// them back with ordinary (coherent) loads after a barrier.
// ---------------------------------------------------------------------------------------------
struct FusedPairArgs {
    const float *in;
    long long in_stride;
    MacArgs mac;            // convolver A: ir / ring / strides / current / active / nchan
    const float2 *ir_b;     // convolver B's IR rows, channel stride ir_b_stride (0 when shared)
    long long ir_b_stride;
    float2 *ring_b;         // convolver B's ring (same geometry as A's): receives the new spectrum as well
    IfftArgs ifft_a, ifft_b; // overlap / out / out_stride / epilogue of each
    SplitArgs split;        // small batches: the delay line cut over zsplit CTAs (see FusedArgs)
    int mix_from_a;         // ifft_b.epi.mix_other is out_a of this very launch (crossfade): TMA_IO reads it from shared memory
    float *copy_in;         // optional: the input block is also stored here (TwoStage's tail_input, :459-461), stride copy_stride
    long long copy_stride;
};

template <int LOGB, int ROWS>
struct FusedPairCfg {
    static constexpr int B = 1 << LOGB;
    static constexpr int CPB = 512 / B;
    static constexpr int R = ROWS;
    static constexpr int ARR = CPB * R * B;                                   // float2 per array per stage
    static constexpr size_t STAGE_BYTES = 3 * (size_t)ARR * sizeof(float2);   // IR_A | IR_B | ring
    static constexpr int NST = 2;
    static constexpr int FFT_PER = (sidx(B) + 2) & ~1;
    // one transform buffer per channel: X[current] during the stream, then convolver A's inverse transform;
    // convolver B's inverse transform reuses stage memory (the stream has ended by then) — 4 CTAs per SM at B = 512
    static constexpr size_t FFT_BYTES = (((size_t)CPB * FFT_PER * sizeof(float2)) + 15) / 16 * 16;
    static constexpr size_t SMEM_BYTES = NST * STAGE_BYTES + 64 + FFT_BYTES;
    static_assert((size_t)CPB * FFT_PER * sizeof(float2) <= STAGE_BYTES, "B's transform buffer must fit one stage");
};

// like apply_epilogue, but the mixed-in samples were written by this CTA a barrier ago: coherent loads
__device__ __forceinline__ float apply_epilogue_pair(float v, const fcb_epilogue &epi, long long c, int i)
{
    if (epi.add0) v = __fadd_rn(v, __ldg(epi.add0 + c * (long long)epi.add_stride + i));
    if (epi.add1) v = __fadd_rn(v, __ldg(epi.add1 + c * (long long)epi.add_stride + i));
    if (epi.mix_other) {
        float2 g = __ldg(reinterpret_cast<const float2 *>(epi.gains) + i);
        float o = *(reinterpret_cast<const volatile float *>(epi.mix_other) + c * (long long)epi.mix_stride + i);
        if (g.x == 1.f && g.y == 0.f) {
            /* mine, untouched */
        } else if (g.x == 0.f && g.y == 1.f) {
            v = o;
        } else {
            v = __fadd_rn(__fmul_rn(v, g.x), __fmul_rn(o, g.y));
        }
    }
    return v;
}

// TMA_IO (16-byte aligned rows): the input block comes in as ONE bulk copy per channel and both output blocks leave as
// one bulk store per channel — the form that works on pinned HOST buffers (small-batch host calls go through mapped
// staging; a 512 B .. 2 KB burst per channel crosses PCIe instead of hundreds of 4-byte accesses).  Same arithmetic.
template <int LOGB, int ROWS, bool TMA_IO = false>
__global__ void __launch_bounds__(256)
k_block_fused_pair(FusedPairArgs fa, const float2 *__restrict__ tw)
{
    using Cfg = FusedPairCfg<LOGB, ROWS>;
    using P = FftPlan<LOGB>;
    constexpr int B = Cfg::B, CPB = Cfg::CPB, R = Cfg::R, ARR = Cfg::ARR, NST = Cfg::NST;
    constexpr int TX = B / 2;
    constexpr int T = P::T, E = P::E;
    constexpr int FT = CPB * T; // FFT threads per convolver (64)
    static_assert(2 * FT <= 256, "both inverse transforms must fit the CTA");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *stages = reinterpret_cast<float2 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NST * Cfg::STAGE_BYTES);
    float2 *fbuf = reinterpret_cast<float2 *>(smem_raw + NST * Cfg::STAGE_BYTES + 64); // [CPB][FFT_PER]
    float2 *fbuf_b = stages;                                                            // [CPB][FFT_PER], after the stream
    // TMA_IO: output staging of A and B, in stage memory behind fbuf_b (the stream has ended by then)
    float *out_sa = reinterpret_cast<float *>(stages + CPB * Cfg::FFT_PER), *out_sb = out_sa + CPB * B;
    uint64_t *in_bar = full + 7;
    static_assert((size_t)CPB * Cfg::FFT_PER * sizeof(float2) + 2 * (size_t)CPB * B * sizeof(float) <= Cfg::STAGE_BYTES, "output staging must fit a stage");

    const MacArgs &a = fa.mac;
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int which = tid / FT;                              // FFT role: 0 = convolver A, 1 = B (threads 0 .. 2*FT-1)
    const int fslot = (tid % FT) / T, flane = tid % T;
    const int Z = fa.split.zsplit > 1 ? fa.split.zsplit : 1;
    const long long group = Z > 1 ? blockIdx.x / Z : blockIdx.x;
    const int zs = Z > 1 ? (int)(blockIdx.x % Z) : 0;
    const bool k1 = zs == 0;
    const bool fwork1 = tid < FT && k1, fwork2 = tid < 2 * FT;
    const long long c0 = group * CPB;
    const int nlive = (int)((a.nchan - c0) < CPB ? (a.nchan - c0) : CPB);
    const int cur = a.current, act = a.active;
    const int lo = Z > 1 ? a.seg_lo + zs * fa.split.zlen : a.seg_lo;
    const int hi = Z > 1 ? (lo + fa.split.zlen < a.seg_hi ? lo + fa.split.zlen : a.seg_hi) : a.seg_hi;
    const int niter = hi > lo ? (hi - lo + R - 1) / R : 0;

    if (tid == 0) {
        for (int s = 0; s < NST; s++) mbar_init(&full[s], 1);
        if (TMA_IO) mbar_init(in_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (TMA_IO && tid == 0 && k1) { // the new block of every live channel lands (raw) at the head of its transform buffer
        mbar_expect_tx(in_bar, (uint32_t)(nlive * B * sizeof(float)));
        for (int ch = 0; ch < nlive; ch++)
            bulk_g2s(fbuf + ch * Cfg::FFT_PER, fa.in + (c0 + ch) * fa.in_stride, B * sizeof(float), in_bar);
    }

    auto issue = [&](int it) {
        const int s = it % NST;
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        float2 *ira_s = stages + (size_t)s * 3 * ARR;
        float2 *irb_s = ira_s + ARR, *rg_s = ira_s + 2 * ARR;
        mbar_expect_tx(&full[s], (uint32_t)(3 * nlive * cnt * B * sizeof(float2)));
        const int j0 = (cur + i0) % act;
        const int first = (act - j0) < cnt ? (act - j0) : cnt;
        for (int ch = 0; ch < nlive; ch++) {
            const float2 *ira = a.ir + a.ir_chan(c0 + ch) * a.ir_stride;
            const float2 *irb = fa.ir_b + (c0 + ch) * fa.ir_b_stride;
            const float2 *rgc = a.ring + (c0 + ch) * a.ring_stride;
            bulk_g2s(ira_s + ch * R * B, ira + (long long)i0 * B, cnt * B * sizeof(float2), &full[s]);
            bulk_g2s(irb_s + ch * R * B, irb + (long long)i0 * B, cnt * B * sizeof(float2), &full[s]);
            bulk_g2s(rg_s + ch * R * B, rgc + (long long)j0 * B, first * B * sizeof(float2), &full[s]);
            if (first < cnt) bulk_g2s(rg_s + ch * R * B + first * B, rgc, (cnt - first) * B * sizeof(float2), &full[s]);
        }
    };
    if (tid == 0)
        for (int it = 0; it < NST && it < niter; it++) issue(it);

    float4 h0a = make_float4(0.f, 0.f, 0.f, 0.f), h0b = h0a;
    const bool mlive = ty < nlive;
    if (mlive) {
        h0a = __ldg(reinterpret_cast<const float4 *>(a.ir + a.ir_chan(c0 + ty) * a.ir_stride) + tx);
        h0b = __ldg(reinterpret_cast<const float4 *>(fa.ir_b + (c0 + ty) * fa.ir_b_stride) + tx);
    }

    // ---- K1 once: forward real FFT of the new block, into both rings (src/fft_convolver.rs:234-241) ----
    float2 *fs = fbuf + (fwork1 ? fslot : 0) * Cfg::FFT_PER;
    const bool flive1 = fwork1 && fslot < nlive;
    if (TMA_IO) {
        if (k1) mbar_wait(in_bar, 0);
        float2 z[E]; // raw block -> padded transform layout, in place through registers
        if (fwork1) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                const int j = flane + e * T;
                z[e] = flive1 && j < B / 2 ? fs[j] : make_float2(0.f, 0.f);
            }
        }
        __syncthreads();
        if (fwork1) {
#pragma unroll
            for (int e = 0; e < E; e++) fs[sidx(flane + e * T)] = z[e];
        }
    } else if (fwork1) {
        load_block_as_complex<LOGB>(fs, flane, fa.in + (c0 + fslot) * fa.in_stride, flive1 ? B : 0);
    }
    __syncthreads();
    if (fa.copy_in && flive1) { // the block as it was fed, for the caller's own buffer (z[j] = x[2j] + i x[2j+1])
        float2 *dst = reinterpret_cast<float2 *>(fa.copy_in + (c0 + fslot) * fa.copy_stride);
#pragma unroll
        for (int e = 0; e < E; e++) // the block is the first B/2 complex points; the rest of the transform is zero padding
            if (flane + e * T < B / 2) dst[flane + e * T] = fs[sidx(flane + e * T)];
    }
    stockham_all<LOGB, -1, 0, 1>(fs, flane, tw, fwork1);
    float2 xk[E];
    if (fwork1) {
#pragma unroll
        for (int e = 0; e < E; e++) xk[e] = rfft_split_bin<LOGB>(fs, flane + e * T, tw);
    }
    __syncthreads();
    if (fwork1) {
        float2 *row_a = flive1 ? const_cast<float2 *>(a.ring) + (c0 + fslot) * a.ring_stride + (long long)cur * B : nullptr;
        float2 *row_b = flive1 ? fa.ring_b + (c0 + fslot) * a.ring_stride + (long long)cur * B : nullptr;
#pragma unroll
        for (int e = 0; e < E; e++) {
            int k = flane + e * T;
            fs[k] = xk[e];
            if (flive1) {
                row_a[k] = xk[e];
                row_b[k] = xk[e];
            }
        }
    }
    __syncthreads();

    // ---- K2 for both convolvers on one ring stream (src/fft_convolver.rs:244-255) ----------------
    const bool packed = (tx == 0);
    float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = acc_a;
    for (int it = 0; it < niter; it++) {
        const int s = it % NST;
        mbar_wait(&full[s], (it / NST) & 1);
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        const float4 *ira_s = reinterpret_cast<const float4 *>(stages + (size_t)s * 3 * ARR + ty * R * B) + tx;
        const float4 *irb_s = ira_s + ARR / 2, *rg_s = ira_s + ARR;
        if (mlive) {
            for (int r = 0; r < cnt; r++) {
                const float4 ha = ira_s[r * TX], hb = irb_s[r * TX], x = rg_s[r * TX];
                cmac_ref(acc_a.x, acc_a.y, ha.x, ha.y, x.x, x.y, packed);
                cmac_ref(acc_a.z, acc_a.w, ha.z, ha.w, x.z, x.w, false);
                cmac_ref(acc_b.x, acc_b.y, hb.x, hb.y, x.x, x.y, packed);
                cmac_ref(acc_b.z, acc_b.w, hb.z, hb.w, x.z, x.w, false);
            }
        }
        __syncthreads();
        if (tid == 0 && it + NST < niter) issue(it + NST);
    }

    if (Z > 1) { // small batches: the last CTA of the group to arrive adds the partials in slice order and finishes
        float4 part[2] = {acc_a, acc_b};
        if (!split_arrive<2>(fa.split, group, zs, part)) return;
        split_gather<2>(fa.split, group, part);
        acc_a = part[0];
        acc_b = part[1];
    }

    // ---- K3 twice, side by side: conv = pre_multiplied + X[current] * H[0] (:256-261), inverse FFT ------
    float4 conv_a = make_float4(0.f, 0.f, 0.f, 0.f), conv_b = conv_a;
    if (mlive) {
        const float4 x = Z > 1 && !k1 ? __ldcg(reinterpret_cast<const float4 *>(a.ring + (c0 + ty) * a.ring_stride + (long long)cur * B) + tx)
                                      : reinterpret_cast<const float4 *>(fbuf + ty * Cfg::FFT_PER)[tx];
        conv_a = acc_a;
        cmac_ref(conv_a.x, conv_a.y, x.x, x.y, h0a.x, h0a.y, packed);
        cmac_ref(conv_a.z, conv_a.w, x.z, x.w, h0a.z, h0a.w, false);
        conv_b = acc_b;
        cmac_ref(conv_b.x, conv_b.y, x.x, x.y, h0b.x, h0b.y, packed);
        cmac_ref(conv_b.z, conv_b.w, x.z, x.w, h0b.z, h0b.w, false);
    }
    __syncthreads();
    if (ty < CPB) {
        float2 *da = fbuf + ty * Cfg::FFT_PER, *db = fbuf_b + ty * Cfg::FFT_PER;
        da[sidx(2 * tx)] = make_float2(conv_a.x, conv_a.y);
        da[sidx(2 * tx + 1)] = make_float2(conv_a.z, conv_a.w);
        db[sidx(2 * tx)] = make_float2(conv_b.x, conv_b.y);
        db[sidx(2 * tx + 1)] = make_float2(conv_b.z, conv_b.w);
    }
    __syncthreads();
    float2 *fs2 = (fwork2 && which ? fbuf_b : fbuf) + (fwork2 ? fslot : 0) * Cfg::FFT_PER;
    if (fwork2) irfft_presplit<LOGB>(fs2, flane, tw);
    __syncthreads();
    stockham_all<LOGB, +1, 0, 1>(fs2, flane, tw, fwork2);

    const float inv_n = 1.0f / (float)(2 * B);
    const long long c = c0 + fslot;
    const bool flive2 = fwork2 && fslot < nlive;
    // A's samples first, then (after a barrier) B's: B's epilogue may read A's output of this block
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        if (flive2 && which == pass) {
            const IfftArgs &o = pass == 0 ? fa.ifft_a : fa.ifft_b;
#pragma unroll
            for (int e = 0; e < E; e++) {
                int j = flane + e * T;
                if (2 * j >= B) continue;
                float2 z = fs2[sidx(j)];
                float y[2] = {z.x * inv_n, z.y * inv_n};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    int i = 2 * j + h;
                    float v = __fadd_rn(y[h], o.overlap[c * B + i]);
                    if (TMA_IO && pass == 1 && fa.mix_from_a && o.epi.mix_other) {
                        // B mixes in A's sample of this block: it sits in A's output staging (same rule as apply_epilogue)
                        const float2 g = __ldg(reinterpret_cast<const float2 *>(o.epi.gains) + i);
                        const float other = out_sa[fslot * B + i];
                        if (o.epi.add0) v = __fadd_rn(v, __ldg(o.epi.add0 + c * (long long)o.epi.add_stride + i));
                        if (o.epi.add1) v = __fadd_rn(v, __ldg(o.epi.add1 + c * (long long)o.epi.add_stride + i));
                        if (g.x == 1.f && g.y == 0.f) {
                        } else if (g.x == 0.f && g.y == 1.f) v = other;
                        else v = __fadd_rn(__fmul_rn(v, g.x), __fmul_rn(other, g.y));
                    } else {
                        v = apply_epilogue_pair(v, o.epi, c, i);
                    }
                    if (TMA_IO) (pass == 0 ? out_sa : out_sb)[fslot * B + i] = v;
                    else o.out[c * o.out_stride + i] = v;
                }
            }
        }
        if (TMA_IO && pass == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // staging -> async proxy
        __syncthreads(); // pass 0: A's output visible to B's mix; pass 1: every reader of the old overlaps is done
    }
    if (TMA_IO && tid == 0) {
        for (int ch = 0; ch < nlive; ch++) {
            bulk_s2g(fa.ifft_a.out + (c0 + ch) * fa.ifft_a.out_stride, out_sa + ch * B, B * sizeof(float));
            bulk_s2g(fa.ifft_b.out + (c0 + ch) * fa.ifft_b.out_stride, out_sb + ch * B, B * sizeof(float));
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (flive2) {
        const IfftArgs &o = which == 0 ? fa.ifft_a : fa.ifft_b;
#pragma unroll
        for (int e = 0; e < E; e++) {
            int j = flane + e * T;
            if (2 * j >= B) {
                float2 z = fs2[sidx(j)];
                *reinterpret_cast<float2 *>(o.overlap + c * B + (2 * j - B)) = make_float2(z.x * inv_n, z.y * inv_n);
            }
        }
    }
    if (TMA_IO && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // staging stays alive until read
}


// ---------------------------------------------------------------------------------------------
// Shared-IR variant: all channels of the engine use ONE impulse response (fcb_engine_desc.shared_ir).
// A CTA owns G channels per thread group (CPB*G channels) and stages each IR tile ONCE per CTA —
// the IR spectra are "staged by TMA and reused across channels that share an IR": per complex MAC
// the CTA moves 8 + 8/(CPB*G) bytes instead of 16.  Arithmetic and order per channel are those of
// k_block_fused, so results are bit-identical to the unshared kernels.
// ---------------------------------------------------------------------------------------------
template <int LOGB, int G>
struct FusedSharedCfg {
    static constexpr int B = 1 << LOGB;
    static constexpr int CPB = 512 / B;
    static constexpr int NSLOT = CPB * G;            // channels per CTA
    static constexpr int R = 4;
    static constexpr int ROWS = R * B;               // float2 per tile
    static constexpr size_t STAGE_BYTES = (size_t)(1 + NSLOT) * ROWS * sizeof(float2);
    static constexpr int NST = 2;
    static constexpr int FFT_PER = (sidx(B) + 2) & ~1;
    static constexpr size_t FFT_BYTES = (size_t)NSLOT * FFT_PER * sizeof(float2);
    static constexpr size_t SMEM_BYTES = NST * STAGE_BYTES + 64 + ((FFT_BYTES + 15) / 16) * 16;
};

template <int LOGB, int G>
__global__ void __launch_bounds__(256)
k_block_fused_shared(FusedArgs fa, const float2 *__restrict__ tw)
{
    using Cfg = FusedSharedCfg<LOGB, G>;
    using P = FftPlan<LOGB>;
    constexpr int B = Cfg::B, NSLOT = Cfg::NSLOT, R = Cfg::R, ROWS = Cfg::ROWS, NST = Cfg::NST;
    constexpr int TX = B / 2;
    constexpr int T = P::T, E = P::E;
    constexpr int FPASS = (NSLOT * T + 255) / 256; // FFT rounds when the transforms need more than 256 threads
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *stages = reinterpret_cast<float2 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NST * Cfg::STAGE_BYTES);
    float2 *fbuf = reinterpret_cast<float2 *>(smem_raw + NST * Cfg::STAGE_BYTES + 64);

    const MacArgs &a = fa.mac;
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX; // MAC role: thread group ty owns slots ty*G .. ty*G+G-1
    const long long c0 = (long long)blockIdx.x * NSLOT;
    const int nlive = (int)((a.nchan - c0) < NSLOT ? (a.nchan - c0) : NSLOT);
    const int cur = a.current, act = a.active, lo = a.seg_lo, hi = a.seg_hi;
    const int niter = hi > lo ? (hi - lo + R - 1) / R : 0;

    if (tid == 0) {
        for (int s = 0; s < NST; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int it) {
        const int s = it % NST;
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        float2 *ir_s = stages + (size_t)s * (1 + NSLOT) * ROWS;
        float2 *rg_s = ir_s + ROWS;
        mbar_expect_tx(&full[s], (uint32_t)((1 + nlive) * cnt * B * sizeof(float2)));
        const int j0 = (cur + i0) % act;
        const int first = (act - j0) < cnt ? (act - j0) : cnt;
        bulk_g2s(ir_s, a.ir + (long long)(i0 - a.ir_seg0) * B, cnt * B * sizeof(float2), &full[s]); // ONE IR tile
        for (int ch = 0; ch < nlive; ch++) {
            const float2 *rgc = a.ring + a.ring_chan(c0 + ch) * a.ring_stride;
            bulk_g2s(rg_s + ch * ROWS, rgc + (long long)j0 * B, first * B * sizeof(float2), &full[s]);
            if (first < cnt) bulk_g2s(rg_s + ch * ROWS + first * B, rgc, (cnt - first) * B * sizeof(float2), &full[s]);
        }
    };
    if (tid == 0)
        for (int it = 0; it < NST && it < niter; it++) issue(it);

    const float4 h0 = __ldg(reinterpret_cast<const float4 *>(a.ir + (long long)(0 - a.ir_seg0) * B) + tx);

    // ---- K1 for the CTA's NSLOT channels (FPASS rounds of up to 256/T transforms) ----------------
#pragma unroll
    for (int fp = 0; fp < FPASS; fp++) {
        const int fslot = fp * (256 / T) + tid / T, flane = tid % T;
        const bool fwork = fslot < NSLOT;
        const bool flive = fwork && fslot < nlive;
        float2 *fs = fbuf + (fwork ? fslot : 0) * Cfg::FFT_PER;
        if (fwork) {
            const float *x = fa.in + (c0 + fslot) * fa.in_stride;
#pragma unroll
            for (int e = 0; e < E; e++) {
                int j = flane + e * T;
                float2 z = make_float2(0.f, 0.f);
                if (flive && 2 * j < B) z.x = __ldg(x + 2 * j);
                if (flive && 2 * j + 1 < B) z.y = __ldg(x + 2 * j + 1);
                fs[sidx(j)] = z;
            }
        }
        __syncthreads();
        stockham_all<LOGB, -1, 0, 1>(fs, flane, tw, fwork);
        float2 xk[E];
        if (fwork) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                int k = flane + e * T;
                xk[e] = rfft_split_bin<LOGB>(fs, k, tw);
            }
        }
        __syncthreads();
        if (fwork) {
            float2 *row = flive ? const_cast<float2 *>(a.ring) + a.ring_chan(c0 + fslot) * a.ring_stride + (long long)cur * B : nullptr;
#pragma unroll
            for (int e = 0; e < E; e++) {
                int k = flane + e * T;
                fs[k] = xk[e];
                if (flive) row[k] = xk[e];
            }
        }
        __syncthreads();
    }

    // ---- K2: G accumulators per thread, one IR tile per stage ------------------------------------
    const bool packed = (tx == 0);
    float4 acc[G];
#pragma unroll
    for (int g = 0; g < G; g++) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < niter; it++) {
        const int s = it % NST;
        mbar_wait(&full[s], (it / NST) & 1);
        const int i0 = lo + it * R;
        const int cnt = (hi - i0) < R ? (hi - i0) : R;
        const float4 *ir_s = reinterpret_cast<const float4 *>(stages + (size_t)s * (1 + NSLOT) * ROWS) + tx;
        const float4 *rg_s = ir_s + ROWS / 2;
        for (int r = 0; r < cnt; r++) {
            const float4 h = ir_s[r * TX];
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int slot = ty * G + g;
                if (slot < nlive) {
                    const float4 x = rg_s[(slot * ROWS) / 2 + r * TX];
                    cmac_ref(acc[g].x, acc[g].y, h.x, h.y, x.x, x.y, packed);
                    cmac_ref(acc[g].z, acc[g].w, h.z, h.w, x.z, x.w, false);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && it + NST < niter) issue(it + NST);
    }

    // ---- K3 --------------------------------------------------------------------------------------
    float4 conv[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int slot = ty * G + g;
        conv[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (slot < nlive) {
            const float4 x = reinterpret_cast<const float4 *>(fbuf + slot * Cfg::FFT_PER)[tx];
            conv[g] = acc[g];
            cmac_ref(conv[g].x, conv[g].y, x.x, x.y, h0.x, h0.y, packed);
            cmac_ref(conv[g].z, conv[g].w, x.z, x.w, h0.z, h0.w, false);
        }
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int slot = ty * G + g;
        float2 *d = fbuf + slot * Cfg::FFT_PER;
        d[sidx(2 * tx)] = make_float2(conv[g].x, conv[g].y);
        d[sidx(2 * tx + 1)] = make_float2(conv[g].z, conv[g].w);
    }
    __syncthreads();
    const IfftArgs &o = fa.ifft;
    const float inv_n = 1.0f / (float)(2 * B);
#pragma unroll
    for (int fp = 0; fp < FPASS; fp++) {
        const int fslot = fp * (256 / T) + tid / T, flane = tid % T;
        const bool fwork = fslot < NSLOT;
        const bool flive = fwork && fslot < nlive;
        float2 *fs = fbuf + (fwork ? fslot : 0) * Cfg::FFT_PER;
        if (fwork) irfft_presplit<LOGB>(fs, flane, tw);
        __syncthreads();
        stockham_all<LOGB, +1, 0, 1>(fs, flane, tw, fwork);
        const long long c = c0 + fslot;
        if (flive) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                int j = flane + e * T;
                if (2 * j >= B) continue;
                float2 z = fs[sidx(j)];
                float y[2] = {z.x * inv_n, z.y * inv_n};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    int i = 2 * j + h;
                    float v = apply_epilogue(__fadd_rn(y[h], o.overlap[c * B + i]), o.epi, c, i);
                    o.out[c * o.out_stride + i] = v;
                }
            }
        }
        __syncthreads();
        if (flive) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                int j = flane + e * T;
                if (2 * j >= B) {
                    float2 z = fs[sidx(j)];
                    *reinterpret_cast<float2 *>(o.overlap + c * B + (2 * j - B)) = make_float2(z.x * inv_n, z.y * inv_n);
                }
            }
        }
    }
}

} // namespace fcb
