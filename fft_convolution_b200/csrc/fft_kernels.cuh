// fft_kernels.cuh — K1 / K5 (real-to-complex forward FFT) and K3 (complex-to-real inverse FFT
// fused with the segment-0 MAC, 1/N normalisation, overlap-add, overlap save and the
// two-stage / crossfade epilogues) for sm_100a.
//
// Replaces the reference's `Fft` adapter over realfft/rustfft (src/fft_convolver.rs:7-50) and
// the per-chunk arithmetic of FFTConvolver::process (:234-241, :256-274, :283-284).
//
// Transform: a real FFT of N = 2B points is one B-point complex FFT plus a split pass
// (z[j] = x[2j] + i x[2j+1]).  The complex FFT is a shared-memory Stockham autosort with
// radix-8/4/2 register butterflies: every thread holds E = 8 (16 for B = 16384) points, so one
// transform uses B/E threads and several transforms share a CTA when B is small.  Twiddles
// come from a table rounded from f64 (tw[t] = exp(-2 pi i t / N), t < N), never from
// fast-math sincos — the 1e-5*RMS parity budget leaves only ~6x headroom.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fftconv_b200.h"

namespace fcb {

// ---- small complex helpers -------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (DIR < 0) or +i (DIR > 0)
template <int DIR>
__device__ __forceinline__ float2 mul_dir_i(float2 a)
{
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

// shared-memory index padding: one float2 of padding per 16 keeps the strided Stockham
// scatter (stride R*Ns) off a single bank pair
__device__ __host__ __forceinline__ constexpr int sidx(int i) { return i + (i >> 4); }

// ---- compile-time plan ----------------------------------------------------------------
#ifndef FCB_FFT_E16_FROM
#define FCB_FFT_E16_FROM 14 // block sizes from 2^this on: 16 points per thread (two radix-8 butterflies in flight)
#endif
// Large transforms (B >= 4096: the two-stage tail, K5 of long segments) use the "wide" plan: 32 points per thread — two
// ADJACENT radix-16 butterflies — so that every shared-memory access of a Stockham pass is 16 bytes (LDS.128 / STS.128:
// butterflies j and j+1 read neighbouring points and, from the second pass on, write neighbouring points; in the first
// pass a butterfly's own 16 outputs are neighbours), three radix-16 passes plus one radix-2 / radix-4 pass instead of
// five radix-8/4 passes, and half the threads.  ncu on the round-1 kernel (profiles/r02_tailfft_before_*): the top stall
// was mio_throttle — the shared-memory instruction queue — at one 1024-thread CTA per SM; the wide plan issues 2.5x
// fewer shared-memory instructions per transform and fits two CTAs per SM.
template <int LOGB>
struct FftPlan {
    static constexpr int B = 1 << LOGB;
    static constexpr bool WIDE = LOGB >= 12;
    static constexpr int E = WIDE ? 32 : (LOGB >= FCB_FFT_E16_FROM ? 16 : (LOGB >= 3 ? 8 : B)); // points per thread
    static constexpr int T = B / E;                                 // threads per transform
    static constexpr int CTA = T >= 256 ? T : 256;                  // threads per CTA
    static constexpr int TPB = CTA / T;                             // transforms per CTA
    // padded shared-memory index of point i: one float2 per 16 (narrow plan), two per 32 (wide plan: even indices stay
    // even, so pairs of points stay 16-byte aligned)
    __host__ __device__ static constexpr int pidx(int i) { return WIDE ? i + 2 * (i >> 5) : sidx(i); }
    static constexpr int SMEM_PER = WIDE ? pidx(B) + 2 : sidx(B) + 1; // float2 per transform
    static constexpr size_t SMEM_BYTES = (size_t)TPB * SMEM_PER * sizeof(float2);
    // (tried in round 1: capping at 32 registers for 8 CTAs/SM — spills made K1/K5 slower, K3 only
    // 12 % faster; left at the natural 40 registers / 6 CTAs per SM)
#ifndef FCB_WIDE_MIN_CTAS
#define FCB_WIDE_MIN_CTAS 2
#endif
    static constexpr int MIN_CTAS = WIDE && LOGB < 14 ? FCB_WIDE_MIN_CTAS : 1; // B = 16384: one 139 KB transform per SM anyway
};

// radix of pass `p` for a 2^LOGB-point transform, 0 when past the last pass
__host__ __device__ constexpr int radix_at(int logb, int p)
{
    int n8 = logb / 3, rem = logb % 3;
    if (logb >= 12) return p < 3 ? 16 : (p == 3 ? (logb == 12 ? 0 : (logb == 13 ? 2 : 4)) : 0); // wide plan
    if (logb == 0) return 0;
    if (logb == 1) return p == 0 ? 2 : 0;
    if (rem == 0) return p < n8 ? 8 : 0;
    if (rem == 2) return p < n8 ? 8 : (p == n8 ? 4 : 0);
    /* rem == 1, logb >= 4: (n8-1) radix-8 passes then two radix-4 */
    return p < n8 - 1 ? 8 : (p < n8 + 1 ? 4 : 0);
}

// offset (in entries, from the end of the N-entry base table) of the pass-ordered twiddles of the pass with
// stride `ns` (see get_twiddles): passes with ns > 1 store (R-1)*ns entries each
__host__ __device__ constexpr int pass_tw_offset(int logb, int ns)
{
    int off = 0, n = 1;
    for (int p = 0; n < ns; p++) {
        const int r = radix_at(logb, p);
        if (n > 1) off += (r - 1) * n;
        n *= r;
    }
    return off;
}

// ---- in-register DFTs, natural order out ------------------------------------------------
template <int DIR>
__device__ __forceinline__ void dft2(float2 &a, float2 &b)
{
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <int DIR>
__device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_dir_i<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

template <int DIR>
__device__ __forceinline__ void dft8(float2 (&v)[8])
{
    // evens / odds
    dft4<DIR>(v[0], v[2], v[4], v[6]);
    dft4<DIR>(v[1], v[3], v[5], v[7]);
    // odd outputs times w8^k, k = 0..3 (w8 = exp(DIR * 2 pi i / 8))
    const float h = 0.70710678118654752440f;
    float2 o1 = v[3], o2 = v[5], o3 = v[7];
    // after dft4 on (v1,v3,v5,v7): v1 = O[0], v3 = O[1], v5 = O[2], v7 = O[3]
    if (DIR < 0) {
        o1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));  // * (1 - i)/sqrt2
        o2 = make_float2(o2.y, -o2.x);                           // * -i
        o3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y)); // * (-1 - i)/sqrt2
    } else {
        o1 = make_float2(h * (o1.x - o1.y), h * (o1.x + o1.y));  // * (1 + i)/sqrt2
        o2 = make_float2(-o2.y, o2.x);                           // * +i
        o3 = make_float2(-h * (o3.x + o3.y), h * (o3.x - o3.y)); // * (-1 + i)/sqrt2
    }
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// 16-point DFT as 4 x 4: n = n1 + 4 n2, k = 4 k1 + k2:  X[4 k1 + k2] = sum_n1 w4^(n1 k1) [ w16^(n1 k2) sum_n2 x[n1 + 4 n2] w4^(n2 k2) ].
// In place; output X[r] ends up in v[(r >> 2) + 4 (r & 3)] (dft16_out).
template <int DIR>
__device__ __forceinline__ float2 mul_w16(float2 a, int m)
{
    // a * exp(DIR * 2 pi i m / 16) for the exponents that occur (m = n1 * k2: 1, 2, 3, 4, 6, 9)
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    float wr = 1.f, wi = 0.f;
    if (m == 1) { wr = c1; wi = s1; }
    else if (m == 2) { wr = h; wi = h; }
    else if (m == 3) { wr = s1; wi = c1; }
    else if (m == 4) { wr = 0.f; wi = 1.f; }
    else if (m == 6) { wr = -h; wi = h; }
    else if (m == 9) { wr = -c1; wi = -s1; }
    if (DIR < 0) wi = -wi;
    return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
}
template <int DIR>
__device__ __forceinline__ void dft16(float2 (&v)[16])
{
#pragma unroll
    for (int n1 = 0; n1 < 4; n1++) dft4<DIR>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]); // v[n1 + 4 k2] = y[n1][k2]
#pragma unroll
    for (int n1 = 1; n1 < 4; n1++)
#pragma unroll
        for (int k2 = 1; k2 < 4; k2++) v[n1 + 4 * k2] = mul_w16<DIR>(v[n1 + 4 * k2], n1 * k2);
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) dft4<DIR>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]); // v[k1 + 4 k2] = X[4 k1 + k2]
}

template <int R, int DIR>
__device__ __forceinline__ void dftR(float2 (&v)[R])
{
    if constexpr (R == 2) dft2<DIR>(v[0], v[1]);
    else if constexpr (R == 4) dft4<DIR>(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 16) dft16<DIR>(v);
    else dft8<DIR>(v);
}
// where output r of dftR sits in the array
template <int R>
__device__ __forceinline__ constexpr int dft_out(int r)
{
    return R == 16 ? (r >> 2) + 4 * (r & 3) : r;
}

// ---- one Stockham pass over a transform resident in shared memory ------------------------
// s: this transform's padded buffer; tid in [0, T); tw: table of N = 2B entries.
template <int LOGB, int R, int DIR, int NS>
__device__ __forceinline__ void stockham_pass(float2 *s, int tid, const float2 *__restrict__ tw, bool work = true)
{
    using P = FftPlan<LOGB>;
    constexpr int B = P::B, T = P::T, NB = P::E / R, Q = B / R; // butterflies per thread, stride
    float2 v[NB][R];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        int j = tid + b * T;
#pragma unroll
        for (int r = 0; r < R; r++) v[b][r] = work ? s[sidx(j + r * Q)] : make_float2(0.f, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; b++) {
        if (!work) break;
        int j = tid + b * T;
        int k = j & (NS - 1);
        if constexpr (NS > 1) {
            // twiddle v[r] *= w_{NS*R}^{r k} = tw[r * k * (N / (NS*R))], N = 2B — read from the pass-ordered copy
            // (coalesced over k) that get_twiddles appends to the base table
            constexpr int OFF = 2 * B + pass_tw_offset(LOGB, NS);
#pragma unroll
            for (int r = 1; r < R; r++) {
                float2 w = __ldg(&tw[OFF + (r - 1) * NS + k]);
                if (DIR > 0) w.y = -w.y;
                v[b][r] = cmul(v[b][r], w);
            }
        }
        dftR<R, DIR>(v[b]);
        int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) s[sidx(j0 + r * NS)] = v[b][r];
    }
    __syncthreads();
}

// the same pass in the wide plan: a thread owns pairs of adjacent butterflies (j, j + 1), j even; every shared-memory
// access and every twiddle load is 16 bytes
template <int LOGB, int R, int DIR, int NS>
__device__ __forceinline__ void stockham_pass_wide(float2 *s, int tid, const float2 *__restrict__ tw)
{
    using P = FftPlan<LOGB>;
    constexpr int B = P::B, T = P::T, NP = P::E / (2 * R), Q = B / R; // butterfly pairs per thread, input stride
    float4 v[NP][R];
#pragma unroll
    for (int b = 0; b < NP; b++) {
        const int j = 2 * (tid + b * T);
#pragma unroll
        for (int r = 0; r < R; r++) v[b][r] = *reinterpret_cast<const float4 *>(&s[P::pidx(j + r * Q)]);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NP; b++) {
        const int j = 2 * (tid + b * T);
        const int k = j & (NS - 1);
        float2 a[R], c[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            a[r] = make_float2(v[b][r].x, v[b][r].y);
            c[r] = make_float2(v[b][r].z, v[b][r].w);
        }
        if constexpr (NS > 1) {
            constexpr int OFF = 2 * B + pass_tw_offset(LOGB, NS);
#pragma unroll
            for (int r = 1; r < R; r++) {
                float4 w = __ldg(reinterpret_cast<const float4 *>(&tw[OFF + (r - 1) * NS + k])); // factors of (r, k), (r, k + 1)
                if (DIR > 0) {
                    w.y = -w.y;
                    w.w = -w.w;
                }
                a[r] = cmul(a[r], make_float2(w.x, w.y));
                c[r] = cmul(c[r], make_float2(w.z, w.w));
            }
        }
        dftR<R, DIR>(a);
        dftR<R, DIR>(c);
        if constexpr (NS > 1) {
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float2 x = a[dft_out<R>(r)], y = c[dft_out<R>(r)];
                *reinterpret_cast<float4 *>(&s[P::pidx(j0 + r * NS)]) = make_float4(x.x, x.y, y.x, y.y);
            }
        } else { // first pass: a butterfly's own outputs are neighbours
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                const float2 x0 = a[dft_out<R>(r)], x1 = a[dft_out<R>(r + 1)], y0 = c[dft_out<R>(r)], y1 = c[dft_out<R>(r + 1)];
                *reinterpret_cast<float4 *>(&s[P::pidx(j * R + r)]) = make_float4(x0.x, x0.y, x1.x, x1.y);
                *reinterpret_cast<float4 *>(&s[P::pidx((j + 1) * R + r)]) = make_float4(y0.x, y0.y, y1.x, y1.y);
            }
        }
    }
    __syncthreads();
}

template <int LOGB, int DIR, int PASS, int NS>
__device__ __forceinline__ void stockham_all(float2 *s, int tid, const float2 *__restrict__ tw, bool work = true)
{
    constexpr int R = radix_at(LOGB, PASS);
    if constexpr (R > 0) {
        if constexpr (FftPlan<LOGB>::WIDE) stockham_pass_wide<LOGB, R, DIR, NS>(s, tid, tw);
        else stockham_pass<LOGB, R, DIR, NS>(s, tid, tw, work);
        stockham_all<LOGB, DIR, PASS + 1, NS * R>(s, tid, tw, work);
    }
}


// ---- pieces shared by k_rfft_forward / k_irfft_ola and the fused block kernels ---------------------

// forward split of bin k from the B-point complex FFT Z of z[j] = x[2j] + i x[2j+1] (padded smem):
// X[k] = Ev + w^k Od, Ev = (Z[k] + conj Z[B-k])/2, Od = (Z[k] - conj Z[B-k])/(2i); bin 0 is
// packed {DC.re, Nyquist.re}
template <int LOGB>
__device__ __forceinline__ float2 rfft_split_bin(const float2 *s, int k, const float2 *__restrict__ tw)
{
    using P = FftPlan<LOGB>;
    constexpr int B = 1 << LOGB;
    float2 a = s[P::pidx(k)];
    if (k == 0) return make_float2(a.x + a.y, a.x - a.y);
    float2 b = cconj(s[P::pidx(B - k)]);
    float2 ev = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
    float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y));
    float2 od = make_float2(d.y, -d.x);
    return cadd(ev, cmul(od, __ldg(&tw[k])));
}

// inverse pre-split, in place on pairs (k, B-k) of a packed spectrum held in padded smem:
// Z'[k] = (X[k] + conj X[B-k]) + i w^{-k} (X[k] - conj X[B-k]),  w = exp(-2 pi i / N).
// Thread `tid` of the transform's T threads; caller synchronises before and after.
template <int LOGB>
__device__ __forceinline__ void irfft_presplit(float2 *s, int tid, const float2 *__restrict__ tw)
{
    using P = FftPlan<LOGB>;
    constexpr int B = P::B, T = P::T, HALF = B / 2;
    constexpr int PAIRS_PER_THREAD = HALF >= T ? HALF / T : 1; // pairs p = 0..HALF-1 (p = 0 also does k = B/2)
#pragma unroll
    for (int e = 0; e < PAIRS_PER_THREAD; e++) {
        int k = tid + e * T;
        if (B == 1) {
            if (k == 0) {
                float2 x = s[0];
                s[0] = make_float2(x.x + x.y, x.x - x.y);
            }
        } else if (k < HALF) {
            if (k == 0) {
                float2 x = s[0]; // {DC, Nyquist}
                s[0] = make_float2(x.x + x.y, x.x - x.y);
                float2 m = s[P::pidx(HALF)];
                s[P::pidx(HALF)] = make_float2(2.f * m.x, -2.f * m.y);
            } else {
                float2 p = s[P::pidx(k)], q = s[P::pidx(B - k)];
                float2 w = __ldg(&tw[k]);
                w.y = -w.y; // w^{-k}
                // k:   (p + conj q) + i w^{-k} (p - conj q)
                float2 sm = make_float2(p.x + q.x, p.y - q.y);
                float2 df = make_float2(p.x - q.x, p.y + q.y);
                float2 t = cmul(df, w);
                s[P::pidx(k)] = make_float2(sm.x - t.y, sm.y + t.x);
                // B-k: (q + conj p) + i w^{-(B-k)} (q - conj p),  w^{-(B-k)} = -conj(w^{-k})
                float2 sm2 = make_float2(sm.x, -sm.y);
                float2 df2 = make_float2(-df.x, df.y);
                float2 w2 = make_float2(-w.x, w.y);
                float2 t2 = cmul(df2, w2);
                s[P::pidx(B - k)] = make_float2(sm2.x - t2.y, sm2.y + t2.x);
            }
        }
    }
}

// output sample i of channel c: (y + overlap) then the optional fused epilogue, in the reference's
// operation order — two-stage ((y+ov)+p0)+p1 (src/fft_convolver.rs:438-454), crossfade
// a*g1 + b*g2 with separately rounded products (src/crossfade_convolver.rs:160-169, 242-278)
__device__ __forceinline__ float apply_epilogue(float v, const fcb_epilogue &epi, long long c, int i)
{
    if (epi.add0) v = __fadd_rn(v, __ldg(epi.add0 + c * (long long)epi.add_stride + i));
    if (epi.add1) v = __fadd_rn(v, __ldg(epi.add1 + c * (long long)epi.add_stride + i));
    if (epi.mix_other) {
        float2 g = __ldg(reinterpret_cast<const float2 *>(epi.gains) + i);
        float o = __ldg(epi.mix_other + c * (long long)epi.mix_stride + i);
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


// global -> shared: z[j] = x'[2j] + i x'[2j+1], x' = [x[0..valid) | zeros], for the T threads of one transform.
// 16-byte loads (two complex points per thread and load) whenever the row is 16-byte aligned and the four
// samples are real data; 4-byte loads at the ragged end of a partially filled block and for unaligned rows.
template <int LOGB>
__device__ __forceinline__ void load_block_as_complex(float2 *s, int tid, const float *__restrict__ x, int valid)
{
    using P = FftPlan<LOGB>;
    constexpr int B = P::B, T = P::T, E = P::E;
    if constexpr (E >= 2) {
        if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
#pragma unroll
            for (int e = 0; e < E / 2; e++) {
                const int j2 = tid + e * T; // complex points 2*j2, 2*j2 + 1 = samples 4*j2 .. 4*j2 + 3
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (4 * j2 + 3 < valid) {
                    v = __ldg(reinterpret_cast<const float4 *>(x) + j2);
                } else {
                    if (4 * j2 < valid) v.x = __ldg(x + 4 * j2);
                    if (4 * j2 + 1 < valid) v.y = __ldg(x + 4 * j2 + 1);
                    if (4 * j2 + 2 < valid) v.z = __ldg(x + 4 * j2 + 2);
                }
                s[P::pidx(2 * j2)] = make_float2(v.x, v.y);
                s[P::pidx(2 * j2 + 1)] = make_float2(v.z, v.w);
            }
            return;
        }
    }
#pragma unroll
    for (int e = 0; e < E; e++) {
        int j = tid + e * T;
        if (j >= B) continue;
        float2 z = make_float2(0.f, 0.f);
        if (2 * j < valid) z.x = __ldg(x + 2 * j);
        if (2 * j + 1 < valid) z.y = __ldg(x + 2 * j + 1);
        s[P::pidx(j)] = z;
    }
}

// ---- wide plan, forward direction: the two ends of the transform without a trip through shared memory ----------
// First Stockham pass (Ns = 1, no twiddles) straight from global memory: butterflies (j, j + 1) read z[j + r Q], z[j + 1 + r Q]
// = four consecutive samples x'[2 (j + r Q) ..], one 16-byte load; inputs past B / 2 are the zero padding of the block
// (src/fft_convolver.rs:56-60) and cost nothing.  Saves the store + load of the whole input (128 KB per 16384-point transform).
template <int LOGB>
__device__ __forceinline__ void stockham_pass0_from_global(float2 *s, int tid, const float *__restrict__ x, int valid)
{
    using P = FftPlan<LOGB>;
    constexpr int B = P::B, R = 16, Q = B / R;
    static_assert(P::WIDE && P::E == 2 * R, "one pair of radix-16 butterflies per thread");
    const int j = 2 * tid;
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    float2 a[R], c[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < R / 2) { // points j + r Q < B / 2: real samples; the rest is zero padding
            const int n0 = 2 * (j + r * Q);
            if (aligned && n0 + 3 < valid) {
                v = __ldg(reinterpret_cast<const float4 *>(x + n0));
            } else {
                if (n0 < valid) v.x = __ldg(x + n0);
                if (n0 + 1 < valid) v.y = __ldg(x + n0 + 1);
                if (n0 + 2 < valid) v.z = __ldg(x + n0 + 2);
                if (n0 + 3 < valid) v.w = __ldg(x + n0 + 3);
            }
        }
        a[r] = make_float2(v.x, v.y);
        c[r] = make_float2(v.z, v.w);
    }
    dft16<-1>(a);
    dft16<-1>(c);
#pragma unroll
    for (int r = 0; r < R; r += 2) {
        const float2 x0 = a[dft_out<R>(r)], x1 = a[dft_out<R>(r + 1)], y0 = c[dft_out<R>(r)], y1 = c[dft_out<R>(r + 1)];
        *reinterpret_cast<float4 *>(&s[P::pidx(j * R + r)]) = make_float4(x0.x, x0.y, x1.x, x1.y);
        *reinterpret_cast<float4 *>(&s[P::pidx((j + 1) * R + r)]) = make_float4(y0.x, y0.y, y1.x, y1.y);
    }
    __syncthreads();
}

// B = 8192: the last pass is radix 2 (Z[k] = u[k] + W^k u[k + B/2], Z[k + B/2] = u[k] - W^k u[k + B/2], W = exp(-2 pi i / B)).
// The real-FFT split of bins k and B - k needs exactly Z[k] and Z[B - k]: both come from four points of u, so the pass and
// the split run as one step from shared memory straight to the spectrum row — no store + load of Z (256 KB per transform).
template <int LOGB>
__device__ __forceinline__ void radix2_split_to_global(const float2 *s, int tid, const float2 *__restrict__ tw, float2 *__restrict__ row)
{
    using P = FftPlan<LOGB>;
    constexpr int B = P::B, T = P::T, H = B / 2;
    constexpr int OFF = 2 * B + pass_tw_offset(LOGB, H); // pass-ordered twiddles of the radix-2 pass: W^k at [OFF + k]
    auto split = [](float2 zk, float2 zm, float2 w) { // X[k] from Z[k], Z[B - k] and w = exp(-2 pi i k / 2B)
        const float2 b = cconj(zm);
        const float2 ev = make_float2(0.5f * (zk.x + b.x), 0.5f * (zk.y + b.y));
        const float2 d = make_float2(0.5f * (zk.x - b.x), 0.5f * (zk.y - b.y));
        return cadd(ev, cmul(make_float2(d.y, -d.x), w));
    };
#pragma unroll
    for (int e = 0; e < H / T; e++) {
        const int k = tid + e * T; // 0 .. B/2 - 1
        const float2 u0 = s[P::pidx(k)], u1 = s[P::pidx(k + H)];
        if (k == 0) {
            const float2 z0 = cadd(u0, u1), zh = csub(u0, u1);        // Z[0], Z[B/2]
            row[0] = make_float2(z0.x + z0.y, z0.x - z0.y);          // packed {DC, Nyquist}
            row[H] = split(zh, zh, __ldg(&tw[H]));                   // bin B/2 is its own mirror
        } else {
            const int m = H - k;
            const float2 v0 = s[P::pidx(m)], v1 = s[P::pidx(m + H)];
            const float2 wk = __ldg(&tw[OFF + k]);
            const float2 wm = make_float2(-wk.x, wk.y);              // W^(B/2 - k) = -conj(W^k)
            const float2 zk = cadd(u0, cmul(u1, wk));                // Z[k]
            const float2 zm = csub(v0, cmul(v1, wm));                // Z[B - k] = Z[(B/2 - k) + B/2]
            const float2 tk = __ldg(&tw[k]);
            row[k] = split(zk, zm, tk);
            row[B - k] = split(zm, zk, make_float2(-tk.x, tk.y));    // exp(-2 pi i (B - k) / 2B) = -conj(exp(-2 pi i k / 2B))
        }
    }
}

// ========================================================================================
// K1 / K5: batched forward real FFT.
// transform q -> (channel c = q / nseg, segment i = q % nseg); source = src + c*src_stride + i*B,
// of which only the first clamp(len - i*B, 0, B) samples are real data (copy_and_pad,
// src/fft_convolver.rs:56-60); destination row = dst + c*dst_stride + i*B (packed bins).
// K1 uses nseg = 1, len = fill + n, dst = ring + current*B  (:234-241);
// K5 uses nseg = S,  len = IR length                         (:131-142, :193-212).
// ========================================================================================
template <int LOGB>
__global__ void __launch_bounds__(FftPlan<LOGB>::CTA, FftPlan<LOGB>::MIN_CTAS)
k_rfft_forward(const float *__restrict__ src, long long src_stride, int len, float2 *__restrict__ dst,
               long long dst_stride, int nseg, long long ntransforms, const float2 *__restrict__ tw)
{
    using P = FftPlan<LOGB>;
    constexpr int B = P::B, T = P::T, E = P::E;
    extern __shared__ float2 smem[];
    const int slot = threadIdx.x / T, tid = threadIdx.x % T;
    float2 *s = smem + slot * P::SMEM_PER;
    const long long q = (long long)blockIdx.x * P::TPB + slot;
    const bool live = q < ntransforms;
    const long long c = live ? q / nseg : 0;
    const int i = live ? (int)(q % nseg) : 0;
    const float *x = src + c * src_stride + (long long)i * B;
    int valid = len - i * B;
    valid = valid < 0 ? 0 : (valid > B ? B : valid);
    if (!live) valid = 0;

    // z[j] = x'[2j] + i x'[2j+1], x' = [x[0..valid) | zeros]
    if constexpr (P::WIDE) {
        stockham_pass0_from_global<LOGB>(s, tid, x, valid);
        if constexpr (LOGB == 13) {
            stockham_pass_wide<LOGB, 16, -1, 16>(s, tid, tw);
            stockham_pass_wide<LOGB, 16, -1, 256>(s, tid, tw);
            if (live) radix2_split_to_global<LOGB>(s, tid, tw, dst + c * dst_stride + (long long)i * B);
            return;
        } else {
            stockham_all<LOGB, -1, 1, 16>(s, tid, tw);
        }
    } else {
        load_block_as_complex<LOGB>(s, tid, x, valid);
        __syncthreads();
        stockham_all<LOGB, -1, 0, 1>(s, tid, tw);
    }

    // split: X[k] = Ev + w^k Od, Ev = (Z[k] + conj Z[B-k])/2, Od = (Z[k] - conj Z[B-k])/(2i)
    if (live) {
        float2 *row = dst + c * dst_stride + (long long)i * B;
        if constexpr (P::WIDE) { // two bins per 16-byte store
#pragma unroll
            for (int e = 0; e < E / 2; e++) {
                const int k = 2 * (tid + e * T);
                const float2 o0 = rfft_split_bin<LOGB>(s, k, tw), o1 = rfft_split_bin<LOGB>(s, k + 1, tw);
                *reinterpret_cast<float4 *>(row + k) = make_float4(o0.x, o0.y, o1.x, o1.y);
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; e++) {
                int k = tid + e * T;
                float2 out = rfft_split_bin<LOGB>(s, k, tw);
                row[k] = out;
            }
        }
    }
}

// ========================================================================================
// K3: conv = pre_multiplied + ring[current] * ir[0]  (src/fft_convolver.rs:256-261, unfused
// f32 like the reference), inverse real FFT, /N (:44-46), overlap-add into the output
// (:270-274) with the two-stage (:438-454) or crossfade (src/crossfade_convolver.rs:75-77)
// epilogue, and overlap save on block completion (:297-298).
// ========================================================================================
struct IfftArgs {
    const float2 *ring_cur; // ring + current*B, channel stride ring_stride
    long long ring_stride;
    const float2 *ir0;      // IR segment 0, channel stride ir_stride (0 when shared); NULL = conv is `premul` as is
    long long ir_stride;
    const float2 *premul;   // [C][B]
    float *overlap;         // [C][B]
    float *out;             // [C][n], channel stride out_stride
    long long out_stride;
    int fill, n, block_complete;
    long long nchan;
    fcb_epilogue epi;
    // peer exchange (IR-partition shards on one NVLink node): conv = sum over g < gather_n of
    // gather[g * gather_stride + c*B + k], ascending g, once gather_flags[g] == gather_seq for every g
    // (written by the peers' reduce kernels with release/system scope)
    const float2 *gather;
    long long gather_stride;
    const unsigned int *gather_flags;
    unsigned int gather_seq;
    int gather_n;
    int *gather_err; // set to 1 when a flag never arrived (bounded spin)
    // multi-block calls (offline_kernels.cuh): `nchan` counts (channel, block) pairs, the IR channel is
    // c / ir_div, and all 2B normalised samples go to raw_out[c][2B] — no overlap-add here
    long long ir_div;
    float *raw_out;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int LOGB>
__global__ void __launch_bounds__(FftPlan<LOGB>::CTA, FftPlan<LOGB>::MIN_CTAS)
k_irfft_ola(IfftArgs a, const float2 *__restrict__ tw)
{
    using P = FftPlan<LOGB>;
    constexpr int B = P::B, T = P::T, E = P::E;
    extern __shared__ float2 smem[];
    const int slot = threadIdx.x / T, tid = threadIdx.x % T;
    float2 *s = smem + slot * P::SMEM_PER;
    const long long c = (long long)blockIdx.x * P::TPB + slot;
    const bool live = c < a.nchan;

    if (a.gather_n > 0) { // wait for every shard's partial spectra (peer stores + release flag)
        __shared__ int s_timed_out;
        if (threadIdx.x == 0) s_timed_out = 0;
        __syncthreads();
        if (threadIdx.x < a.gather_n) {
            long long spins = 0;
            while (ld_acquire_sys(a.gather_flags + threadIdx.x) != a.gather_seq) {
                if (++spins > (1ll << 26)) {
                    *reinterpret_cast<volatile int *>(a.gather_err) = 1; // mapped host word: the next call fails loudly
                    s_timed_out = 1;
                    break;
                }
            }
        }
        __syncthreads();
        if (s_timed_out) {
            // a shard never published: the inbox holds a stale or partial sum.  Never turn that into audio —
            // this block is silence, the overlap is left alone, and the host sees the error word.
            if (live)
                for (int i = tid; i < a.n; i += T) a.out[c * a.out_stride + i] = 0.f;
            return;
        }
    }

    // 1. conv, packed layout (bin 0 = {DC, Nyquist}: two real products)
    bool conv_done = false;
    if constexpr (P::WIDE) {
        if (a.gather_n == 0 && a.ir0) {
            // wide plan, plain case: two bins per 16-byte load, four loads of each array in flight per thread — with 32
            // points per thread and two CTAs per SM the latency has to be covered by loads in flight, not by warps
            const float4 *xr = reinterpret_cast<const float4 *>(a.ring_cur + c * a.ring_stride);
            const float4 *hr = reinterpret_cast<const float4 *>(a.ir0 + (a.ir_div ? c / a.ir_div : c) * a.ir_stride);
            const float4 *pr4 = reinterpret_cast<const float4 *>(a.premul + c * B);
            constexpr int U = 4;
#pragma unroll
            for (int e0 = 0; e0 < E / 2; e0 += U) {
                float4 x[U], h[U], q[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k2 = tid + (e0 + u) * T; // bins 2*k2, 2*k2 + 1
                    x[u] = live ? xr[k2] : make_float4(0.f, 0.f, 0.f, 0.f);
                    h[u] = live ? __ldg(hr + k2) : make_float4(0.f, 0.f, 0.f, 0.f);
                    q[u] = live ? pr4[k2] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k2 = tid + (e0 + u) * T;
                    float r0, i0;
                    if (k2 == 0) {
                        r0 = __fmul_rn(x[u].x, h[u].x);
                        i0 = __fmul_rn(x[u].y, h[u].y);
                    } else {
                        r0 = __fsub_rn(__fmul_rn(x[u].x, h[u].x), __fmul_rn(x[u].y, h[u].y));
                        i0 = __fadd_rn(__fmul_rn(x[u].x, h[u].y), __fmul_rn(x[u].y, h[u].x));
                    }
                    const float r1 = __fsub_rn(__fmul_rn(x[u].z, h[u].z), __fmul_rn(x[u].w, h[u].w));
                    const float i1 = __fadd_rn(__fmul_rn(x[u].z, h[u].w), __fmul_rn(x[u].w, h[u].z));
                    *reinterpret_cast<float4 *>(&s[P::pidx(2 * k2)]) =
                        make_float4(__fadd_rn(q[u].x, r0), __fadd_rn(q[u].y, i0), __fadd_rn(q[u].z, r1), __fadd_rn(q[u].w, i1));
                }
            }
            conv_done = true;
        }
    }
#pragma unroll
    for (int e = 0; e < E; e++) {
        if (conv_done) break;
        int k = tid + e * T;
        float2 v = make_float2(0.f, 0.f);
        if (live && a.gather_n > 0) {
            v = a.gather[c * B + k];
            for (int g = 1; g < a.gather_n; g++) {
                const float2 q = a.gather[g * a.gather_stride + c * B + k];
                v.x = __fadd_rn(v.x, q.x);
                v.y = __fadd_rn(v.y, q.y);
            }
        } else if (live && !a.ir0) {
            v = a.premul[c * B + k]; // conv already complete (MIMO: summed over inputs and shards)
        } else if (live) {
            float2 x = a.ring_cur[c * a.ring_stride + k];
            float2 h = __ldg(&a.ir0[(a.ir_div ? c / a.ir_div : c) * a.ir_stride + k]);
            float2 p = a.premul[c * B + k];
            float pr, pi;
            if (k == 0) {
                pr = __fmul_rn(x.x, h.x);
                pi = __fmul_rn(x.y, h.y);
            } else {
                pr = __fsub_rn(__fmul_rn(x.x, h.x), __fmul_rn(x.y, h.y));
                pi = __fadd_rn(__fmul_rn(x.x, h.y), __fmul_rn(x.y, h.x));
            }
            v = make_float2(__fadd_rn(p.x, pr), __fadd_rn(p.y, pi));
        }
        s[P::pidx(k)] = v;
    }
    __syncthreads();

    // 2. pre-split, in place on pairs (k, B-k)
    irfft_presplit<LOGB>(s, tid, tw);
    __syncthreads();

    // 3. unnormalised inverse complex FFT
    const float inv_n = 1.0f / (float)(2 * B);
    if constexpr (LOGB == 13) {
        // B = 8192, whole block, plain overlap-add: the last pass is a radix 2 — y[k] = u[k] + conj(W^k) u[k + B/2] is sample
        // pair (2k, 2k+1) of the FIRST half of the result, y[k + B/2] = u[k] - ... the same pair of the SECOND half — so it
        // is folded into the overlap-add: a thread adds the old overlap to the first-half samples and writes the
        // second-half samples over the very overlap samples it has just read (no barrier, no store + load of y: 256 KB of
        // shared-memory traffic per transform less)
        const bool whole = !a.raw_out && a.fill == 0 && a.n == B && a.block_complete && !a.epi.add0 && !a.epi.add1 && !a.epi.mix_other &&
                           (a.out_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
        if (whole) {
            constexpr int H = B / 2, OFF = 2 * B + pass_tw_offset(LOGB, H);
            stockham_pass_wide<LOGB, 16, +1, 1>(s, tid, tw);
            stockham_pass_wide<LOGB, 16, +1, 16>(s, tid, tw);
            stockham_pass_wide<LOGB, 16, +1, 256>(s, tid, tw);
            if (live) {
#pragma unroll
                for (int e = 0; e < H / 2 / T; e++) {
                    const int k = 2 * (tid + e * T); // points k, k + 1 and their partners k + H, k + 1 + H
                    const float4 u0 = *reinterpret_cast<const float4 *>(&s[P::pidx(k)]);
                    const float4 u1 = *reinterpret_cast<const float4 *>(&s[P::pidx(k + H)]);
                    const float4 w = __ldg(reinterpret_cast<const float4 *>(&tw[OFF + k]));
                    const float2 t0 = cmul(make_float2(u1.x, u1.y), make_float2(w.x, -w.y));
                    const float2 t1 = cmul(make_float2(u1.z, u1.w), make_float2(w.z, -w.w));
                    float *ov = a.overlap + c * B + 2 * k;
                    const float4 old = *reinterpret_cast<const float4 *>(ov);
                    *reinterpret_cast<float4 *>(a.out + c * a.out_stride + 2 * k) =
                        make_float4(__fadd_rn((u0.x + t0.x) * inv_n, old.x), __fadd_rn((u0.y + t0.y) * inv_n, old.y),
                                    __fadd_rn((u0.z + t1.x) * inv_n, old.z), __fadd_rn((u0.w + t1.y) * inv_n, old.w));
                    *reinterpret_cast<float4 *>(ov) =
                        make_float4((u0.x - t0.x) * inv_n, (u0.y - t0.y) * inv_n, (u0.z - t1.x) * inv_n, (u0.w - t1.y) * inv_n);
                }
            }
            return;
        }
    }
    stockham_all<LOGB, +1, 0, 1>(s, tid, tw);

    // 4. epilogue.  y[2j] = Re z[j] / N, y[2j+1] = Im z[j] / N (N = 2B: exact scaling).
    if (a.raw_out) {
        if (live) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                int j = tid + e * T;
                float2 z = s[P::pidx(j)];
                *reinterpret_cast<float2 *>(a.raw_out + c * 2 * B + 2 * j) = make_float2(z.x * inv_n, z.y * inv_n);
            }
        }
        return;
    }
    const int lo = a.fill, hi = a.fill + a.n;
    // whole block, plain overlap-add, 16-byte aligned rows: two complex points = four samples per access
    const bool vec = E >= 2 && B >= 4 && a.fill == 0 && a.n == B && a.block_complete && !a.epi.add0 && !a.epi.add1 && !a.epi.mix_other &&
                     (a.out_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
    if (vec) {
        if (live) {
#pragma unroll
            for (int e = 0; e < E / 2; e++) {
                const int j2 = tid + e * T; // samples 4*j2 .. 4*j2 + 3 of the 2B-sample result
                if (4 * j2 < B) {
                    const float2 z0 = s[P::pidx(2 * j2)], z1 = s[P::pidx(2 * j2 + 1)];
                    const float4 y = make_float4(z0.x * inv_n, z0.y * inv_n, z1.x * inv_n, z1.y * inv_n);
                    const float4 ov = *reinterpret_cast<const float4 *>(a.overlap + c * B + 4 * j2);
                    *reinterpret_cast<float4 *>(a.out + c * a.out_stride + 4 * j2) =
                        make_float4(__fadd_rn(y.x, ov.x), __fadd_rn(y.y, ov.y), __fadd_rn(y.z, ov.z), __fadd_rn(y.w, ov.w));
                }
            }
        }
        __syncthreads(); // every reader of the old overlap is done
        if (live) {
#pragma unroll
            for (int e = 0; e < E / 2; e++) {
                const int j2 = tid + e * T;
                if (4 * j2 >= B) { // second half of the result (still in shared memory) -> the new overlap
                    const float2 z0 = s[P::pidx(2 * j2)], z1 = s[P::pidx(2 * j2 + 1)];
                    *reinterpret_cast<float4 *>(a.overlap + c * B + (4 * j2 - B)) =
                        make_float4(z0.x * inv_n, z0.y * inv_n, z1.x * inv_n, z1.y * inv_n);
                }
            }
        }
        return;
    }
    // 4a. first half -> output (+ overlap, + epilogue)
#pragma unroll
    for (int e = 0; e < E; e++) {
        int j = tid + e * T;
        if (!live || 2 * j >= B) continue;
        float2 z = s[P::pidx(j)];
        float y[2] = {z.x * inv_n, z.y * inv_n};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int m = 2 * j + h;
            if (m >= lo && m < hi && m < B) {
                int i = m - lo;
                float v = apply_epilogue(__fadd_rn(y[h], a.overlap[c * B + m]), a.epi, c, i);
                a.out[c * a.out_stride + i] = v;
            }
        }
    }
    // 4b. second half -> new overlap, only once every reader of the old overlap is done
    if (a.block_complete) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; e++) {
            int j = tid + e * T;
            if (!live) continue;
            if (B == 1) {
                a.overlap[c] = s[0].y * inv_n; // y[1]
            } else if (2 * j >= B) {
                float2 z = s[P::pidx(j)];
                *reinterpret_cast<float2 *>(a.overlap + c * B + (2 * j - B)) = make_float2(z.x * inv_n, z.y * inv_n);
            }
        }
    }
}

} // namespace fcb
