/*
 * fftconv_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see fftconv_oracle.h).
 *
 * Plain-C restatement of the reference crate's algorithm; every function cites the
 * reference file:line (relative to /root/reference) it follows.  Build with
 *   gcc -O3 -ffp-contract=off   (Rust/LLVM does not contract a*b+c into FMA, so neither do we)
 *
 * The FFT here is NOT the reference's (realfft/rustfft are third-party and absent):
 * it is an ordinary f32 radix-2 real FFT with twiddles rounded from f64 — "any
 * mathematically correct f32 real DFT" per SURVEY.md §8(c).
 */
#include "fftconv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <time.h>
#ifdef __AVX2__
#include <immintrin.h>
#endif

/* ------------------------------------------------------------------------------------------
 * real FFT stand-in.  Contract mirrored from the call sites src/fft_convolver.rs:36-49:
 * forward = unnormalised R2C (n reals -> n/2+1 bins, DC/Nyquist imaginary parts exactly 0),
 * inverse = unnormalised C2R then every sample divided by n.
 * ---------------------------------------------------------------------------------------- */
struct orc_plan {
    size_t n, m;       /* real length, complex half length */
    int log2m;
    float *tw_re, *tw_im; /* stage twiddles of the m-point complex FFT, stage s at offset (1<<s)-1 */
    float *sp_re, *sp_im; /* e^{-2 pi i k/n}, k = 0..m */
    uint32_t *rev;        /* bit reversal of 0..m-1 */
    int refs;
};

orc_plan *orc_plan_new(size_t n)
{
    orc_plan *p = (orc_plan *)calloc(1, sizeof *p);
    p->n = n;
    p->m = n / 2;
    p->refs = 1;
    if (n == 0) return p; /* Fft::default() plans length 0 (src/fft_convolver.rs:13-21) */
    size_t m = p->m;
    int lg = 0;
    while (((size_t)1 << lg) < m) lg++;
    p->log2m = lg;
    p->tw_re = (float *)malloc(sizeof(float) * (m ? m : 1));
    p->tw_im = (float *)malloc(sizeof(float) * (m ? m : 1));
    for (int s = 0; s < lg; s++) {
        size_t half = (size_t)1 << s;
        for (size_t j = 0; j < half; j++) {
            double a = -M_PI * (double)j / (double)half;
            p->tw_re[half - 1 + j] = (float)cos(a);
            p->tw_im[half - 1 + j] = (float)sin(a);
        }
    }
    p->sp_re = (float *)malloc(sizeof(float) * (m + 1));
    p->sp_im = (float *)malloc(sizeof(float) * (m + 1));
    for (size_t k = 0; k <= m; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        p->sp_re[k] = (float)cos(a);
        p->sp_im[k] = (float)sin(a);
    }
    p->rev = (uint32_t *)malloc(sizeof(uint32_t) * (m ? m : 1));
    for (size_t i = 0; i < m; i++) {
        uint32_t r = 0;
        for (int b = 0; b < lg; b++)
            if (i & ((size_t)1 << b)) r |= 1u << (lg - 1 - b);
        p->rev[i] = r;
    }
    return p;
}

static orc_plan *plan_ref(orc_plan *p)
{
    __atomic_add_fetch(&p->refs, 1, __ATOMIC_RELAXED);
    return p;
}

void orc_plan_free(orc_plan *p)
{
    if (!p) return;
    if (__atomic_sub_fetch(&p->refs, 1, __ATOMIC_ACQ_REL) > 0) return;
    free(p->tw_re); free(p->tw_im); free(p->sp_re); free(p->sp_im); free(p->rev);
    free(p);
}

/* in-place m-point complex FFT on split arrays already in bit-reversed order;
 * sign = -1 forward, +1 inverse (unnormalised) */
static void cfft_stages(const orc_plan *p, float *restrict re, float *restrict im, float sign)
{
    size_t m = p->m;
    int s = 0;
    if (p->log2m >= 2) {
        /* first two radix-2 stages fused (twiddles 1 and -+i): plain radix-4 on bit-reversed input */
        for (size_t b = 0; b < m; b += 4) {
            float a0r = re[b], a0i = im[b], a1r = re[b + 1], a1i = im[b + 1];
            float a2r = re[b + 2], a2i = im[b + 2], a3r = re[b + 3], a3i = im[b + 3];
            float s0r = a0r + a1r, s0i = a0i + a1i, d0r = a0r - a1r, d0i = a0i - a1i;
            float s1r = a2r + a3r, s1i = a2i + a3i, d1r = a2r - a3r, d1i = a2i - a3i;
            /* second stage twiddle for j=1: w = exp(sign*i*pi/2) = sign*i */
            float tr = sign < 0 ? d1i : -d1i, ti = sign < 0 ? -d1r : d1r;
            re[b] = s0r + s1r; im[b] = s0i + s1i;
            re[b + 2] = s0r - s1r; im[b + 2] = s0i - s1i;
            re[b + 1] = d0r + tr; im[b + 1] = d0i + ti;
            re[b + 3] = d0r - tr; im[b + 3] = d0i - ti;
        }
        s = 2;
    }
    for (; s < p->log2m; s++) {
        size_t half = (size_t)1 << s;
        const float *restrict wr = p->tw_re + (half - 1), *restrict wi = p->tw_im + (half - 1);
        for (size_t b = 0; b < m; b += 2 * half) {
            float *restrict ar = re + b, *restrict ai = im + b, *restrict br = re + b + half, *restrict bi = im + b + half;
            if (sign < 0) {
                for (size_t j = 0; j < half; j++) {
                    float c = wr[j], d = wi[j];
                    float vr = br[j] * c - bi[j] * d;
                    float vi = br[j] * d + bi[j] * c;
                    float ur = ar[j], ui = ai[j];
                    ar[j] = ur + vr; ai[j] = ui + vi;
                    br[j] = ur - vr; bi[j] = ui - vi;
                }
            } else {
                for (size_t j = 0; j < half; j++) {
                    float c = wr[j], d = -wi[j];
                    float vr = br[j] * c - bi[j] * d;
                    float vi = br[j] * d + bi[j] * c;
                    float ur = ar[j], ui = ai[j];
                    ar[j] = ur + vr; ai[j] = ui + vi;
                    br[j] = ur - vr; bi[j] = ui - vi;
                }
            }
        }
    }
}

#define ORC_STACK_M 4096

void orc_rfft_forward(const orc_plan *p, const float *in, orc_cpx *out)
{
    size_t m = p->m;
    if (p->n == 0) return;
    float sre[ORC_STACK_M], sim[ORC_STACK_M];
    float *re = sre, *im = sim;
    if (m > ORC_STACK_M) { re = (float *)malloc(sizeof(float) * 2 * m); im = re + m; }
    /* z[j] = x[2j] + i x[2j+1], stored bit-reversed */
    for (size_t j = 0; j < m; j++) {
        uint32_t r = p->rev[j];
        re[r] = in[2 * j];
        im[r] = in[2 * j + 1];
    }
    cfft_stages(p, re, im, -1.0f);
    /* split: X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[m-k])/2, O = (Z[k] - conj Z[m-k])/(2i) */
    out[0].re = re[0] + im[0]; out[0].im = 0.0f;
    out[m].re = re[0] - im[0]; out[m].im = 0.0f;
    for (size_t k = 1; k < m; k++) {
        float ar = re[k], ai = im[k], br = re[m - k], bi = -im[m - k];
        float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
        float dr = 0.5f * (ar - br), di = 0.5f * (ai - bi); /* (Z - conj Z')/2 */
        float orr = di, oi = -dr;                           /* divide by i */
        float c = p->sp_re[k], s = p->sp_im[k];
        out[k].re = er + (orr * c - oi * s);
        out[k].im = ei + (orr * s + oi * c);
    }
    if (re != sre) free(re);
}

void orc_rfft_inverse(const orc_plan *p, const orc_cpx *in, float *out)
{
    size_t m = p->m, n = p->n;
    if (n == 0) return;
    float sre[ORC_STACK_M], sim[ORC_STACK_M];
    float *re = sre, *im = sim;
    if (m > ORC_STACK_M) { re = (float *)malloc(sizeof(float) * 2 * m); im = re + m; }
    /* Z'[k] = (X[k] + conj X[m-k]) + i w^{-k} (X[k] - conj X[m-k]); DC/Nyquist imaginary
     * parts are ignored (realfft zeroes them before transforming) */
    for (size_t k = 0; k < m; k++) {
        float ar = in[k].re, ai = (k == 0) ? 0.0f : in[k].im;
        float br = in[m - k].re, bi = (k == 0) ? 0.0f : -in[m - k].im;
        float er = ar + br, ei = ai + bi;
        float dr = ar - br, di = ai - bi;
        float c = p->sp_re[k], s = -p->sp_im[k]; /* w^{-k} */
        float tr = dr * c - di * s, ti = dr * s + di * c;
        uint32_t r = p->rev[k];
        re[r] = er - ti; /* + i*(tr + i ti) = -ti + i tr */
        im[r] = ei + tr;
    }
    cfft_stages(p, re, im, +1.0f);
    /* FFT normalisation, src/fft_convolver.rs:44-46: `*bin /= len as f32` */
    float len = (float)n;
    for (size_t j = 0; j < m; j++) {
        out[2 * j] = re[j] / len;
        out[2 * j + 1] = im[j] / len;
    }
    if (re != sre) free(re);
}

/* src/fft_convolver.rs:52-54 */
size_t orc_complex_size(size_t n) { return n / 2 + 1; }

/* src/fft_convolver.rs:62-74: result[i] += a[i] * b[i]; Complex<f32> multiply =
 * (ar*br - ai*bi, ar*bi + ai*br), every operation rounded separately */
void orc_complex_multiply_accumulate(orc_cpx *result, const orc_cpx *a, const orc_cpx *b, size_t len)
{
    size_t i = 0;
#ifdef __AVX2__
    /* same operations as the scalar loop, 4 complex bins per step: two rounded products per
     * component, one rounded subtract (even lanes) / add (odd lanes) via addsub, one rounded
     * accumulate — no FMA, so the bits equal the reference's scalar arithmetic */
    float *r = (float *)result;
    const float *pa = (const float *)a, *pb = (const float *)b;
    for (; i + 4 <= len; i += 4) {
        __m256 va = _mm256_loadu_ps(pa + 2 * i), vb = _mm256_loadu_ps(pb + 2 * i);
        __m256 are = _mm256_moveldup_ps(va), aim = _mm256_movehdup_ps(va);
        __m256 bsw = _mm256_permute_ps(vb, 0xB1);           /* (b.im, b.re) */
        __m256 t0 = _mm256_mul_ps(are, vb);                 /* a.re*b.re , a.re*b.im */
        __m256 t1 = _mm256_mul_ps(aim, bsw);                /* a.im*b.im , a.im*b.re */
        __m256 prod = _mm256_addsub_ps(t0, t1);             /* re: t0 - t1 ; im: t0 + t1 */
        _mm256_storeu_ps(r + 2 * i, _mm256_add_ps(_mm256_loadu_ps(r + 2 * i), prod));
    }
#endif
    for (; i < len; i++) {
        float pr = a[i].re * b[i].re - a[i].im * b[i].im;
        float pi = a[i].re * b[i].im + a[i].im * b[i].re;
        result[i].re += pr;
        result[i].im += pi;
    }
}

static size_t next_power_of_two(size_t v)
{
    size_t p = 1; /* usize::next_power_of_two(0) == 1 */
    while (p < v) p <<= 1;
    return p;
}

/* ------------------------------------------------------------------------------------------
 * FFTConvolver — src/fft_convolver.rs:86-307
 * ---------------------------------------------------------------------------------------- */
struct orc_fftconv {
    size_t ir_len, block_size, seg_count, active_seg_count; /* :88-91 */
    size_t fft_complex_size;
    orc_cpx *segments;    /* seg_count x K  (input spectra ring, :92) */
    orc_cpx *segments_ir; /* seg_count x K  (:93) */
    float *fft_buffer;    /* 2B (:94) */
    orc_plan *fft;        /* :95 */
    orc_cpx *pre_multiplied, *conv; /* K each (:96-97) */
    float *overlap;       /* B (:98) */
    size_t current;       /* :99 */
    float *input_buffer;  /* B (:100) */
    size_t input_buffer_fill; /* :101 */
};

/* #[derive(Default)]: everything empty / zero, Fft::default() */
orc_fftconv *orc_fftconv_default(void)
{
    orc_fftconv *c = (orc_fftconv *)calloc(1, sizeof *c);
    c->fft = orc_plan_new(0);
    return c;
}

/* src/fft_convolver.rs:105-172 */
orc_fftconv *orc_fftconv_init(const float *ir, size_t n_ir, size_t block_size, size_t max_response_length)
{
    if (max_response_length < n_ir) return NULL; /* panic! :106-110 */
    size_t ir_len = max_response_length;          /* padded_ir.resize(max_response_length, 0.) :111-113 */
    float *padded = (float *)calloc(ir_len ? ir_len : 1, sizeof(float));
    if (n_ir) memcpy(padded, ir, n_ir * sizeof(float));

    orc_fftconv *c = (orc_fftconv *)calloc(1, sizeof *c);
    c->ir_len = ir_len;
    c->block_size = next_power_of_two(block_size); /* :115 */
    size_t B = c->block_size, seg_size = 2 * B;    /* :116 */
    c->seg_count = (size_t)ceil((double)ir_len / (double)B); /* :117 */
    c->active_seg_count = c->seg_count;            /* :118 */
    size_t K = c->fft_complex_size = orc_complex_size(seg_size); /* :119 */
    c->fft = orc_plan_new(seg_size);               /* :122-123 */
    c->fft_buffer = (float *)calloc(seg_size, sizeof(float));
    c->segments = (orc_cpx *)calloc(c->seg_count * K + 1, sizeof(orc_cpx));    /* :127 */
    c->segments_ir = (orc_cpx *)calloc(c->seg_count * K + 1, sizeof(orc_cpx)); /* :128 */
    for (size_t i = 0; i < c->seg_count; i++) {    /* :131-142 */
        size_t remaining = ir_len - i * B;
        size_t size_copy = remaining >= B ? B : remaining;
        memcpy(c->fft_buffer, padded + i * B, size_copy * sizeof(float)); /* copy_and_pad :56-60 */
        memset(c->fft_buffer + size_copy, 0, (seg_size - size_copy) * sizeof(float));
        orc_rfft_forward(c->fft, c->fft_buffer, c->segments_ir + i * K);
    }
    c->pre_multiplied = (orc_cpx *)calloc(K, sizeof(orc_cpx)); /* :145-147 */
    c->conv = (orc_cpx *)calloc(K, sizeof(orc_cpx));
    c->overlap = (float *)calloc(B, sizeof(float));
    c->input_buffer = (float *)calloc(B, sizeof(float));       /* :150-154 */
    c->input_buffer_fill = 0;
    c->current = 0;
    free(padded);
    return c;
}

static void *dup_mem(const void *src, size_t bytes)
{
    if (!src) return NULL;
    void *d = malloc(bytes ? bytes : 1);
    memcpy(d, src, bytes);
    return d;
}

/* #[derive(Clone)] (:86): deep copy of all Vec state, Arc-shared plans (:9-10) */
orc_fftconv *orc_fftconv_clone(const orc_fftconv *s)
{
    orc_fftconv *c = (orc_fftconv *)malloc(sizeof *c);
    *c = *s;
    size_t K = s->fft_complex_size, B = s->block_size;
    c->fft = plan_ref(s->fft);
    c->segments = (orc_cpx *)dup_mem(s->segments, (s->seg_count * K + 1) * sizeof(orc_cpx));
    c->segments_ir = (orc_cpx *)dup_mem(s->segments_ir, (s->seg_count * K + 1) * sizeof(orc_cpx));
    c->fft_buffer = (float *)dup_mem(s->fft_buffer, 2 * B * sizeof(float));
    c->pre_multiplied = (orc_cpx *)dup_mem(s->pre_multiplied, K * sizeof(orc_cpx));
    c->conv = (orc_cpx *)dup_mem(s->conv, K * sizeof(orc_cpx));
    c->overlap = (float *)dup_mem(s->overlap, B * sizeof(float));
    c->input_buffer = (float *)dup_mem(s->input_buffer, B * sizeof(float));
    return c;
}

void orc_fftconv_free(orc_fftconv *c)
{
    if (!c) return;
    orc_plan_free(c->fft);
    free(c->segments); free(c->segments_ir); free(c->fft_buffer); free(c->pre_multiplied);
    free(c->conv); free(c->overlap); free(c->input_buffer);
    free(c);
}

/* src/fft_convolver.rs:174-213 */
int orc_fftconv_update(orc_fftconv *c, const float *response, size_t new_ir_len)
{
    if (new_ir_len > c->ir_len) return ORC_PANIC; /* :177-179 */
    if (c->ir_len == 0) return ORC_OK;            /* :181-183 */
    size_t B = c->block_size, K = c->fft_complex_size;
    memset(c->fft_buffer, 0, 2 * B * sizeof(float));  /* :185-188 */
    memset(c->conv, 0, K * sizeof(orc_cpx));
    memset(c->pre_multiplied, 0, K * sizeof(orc_cpx));
    memset(c->overlap, 0, B * sizeof(float));
    c->active_seg_count = (size_t)ceil((double)new_ir_len / (double)B); /* :190 */
    for (size_t i = 0; i < c->active_seg_count; i++) { /* :193-207 */
        size_t remaining = new_ir_len - i * B;
        size_t size_copy = remaining >= B ? B : remaining;
        memcpy(c->fft_buffer, response + i * B, size_copy * sizeof(float));
        memset(c->fft_buffer + size_copy, 0, (2 * B - size_copy) * sizeof(float));
        orc_rfft_forward(c->fft, c->fft_buffer, c->segments_ir + i * K);
    }
    for (size_t i = c->active_seg_count; i < c->seg_count; i++) /* :210-212 */
        memset(c->segments_ir + i * K, 0, K * sizeof(orc_cpx));
    return ORC_OK;
}

/* src/fft_convolver.rs:215-295 */
int orc_fftconv_process(orc_fftconv *c, const float *input, size_t in_len, float *output, size_t out_len)
{
    if (c->active_seg_count == 0) { /* :216-219 */
        memset(output, 0, out_len * sizeof(float));
        return ORC_OK;
    }
    if (in_len < out_len) return ORC_PANIC; /* slice index at :231 would panic */
    size_t B = c->block_size, K = c->fft_complex_size;
    size_t processed = 0;
    while (processed < out_len) { /* :222 */
        int input_buffer_was_empty = c->input_buffer_fill == 0; /* :223 */
        size_t processing = out_len - processed;                  /* :224-227 */
        if (B - c->input_buffer_fill < processing) processing = B - c->input_buffer_fill;
        size_t pos = c->input_buffer_fill;                        /* :229-231 */
        memcpy(c->input_buffer + pos, input + processed, processing * sizeof(float));

        /* forward FFT of [input_buffer | zeros] into segments[current] (:234-241) */
        memcpy(c->fft_buffer, c->input_buffer, B * sizeof(float));
        memset(c->fft_buffer + B, 0, B * sizeof(float));
        orc_rfft_forward(c->fft, c->fft_buffer, c->segments + c->current * K);

        if (input_buffer_was_empty) { /* :244-255 */
            memset(c->pre_multiplied, 0, K * sizeof(orc_cpx));
            for (size_t i = 1; i < c->active_seg_count; i++) {
                size_t index_ir = i;
                size_t index_audio = (c->current + i) % c->active_seg_count;
                orc_complex_multiply_accumulate(c->pre_multiplied, c->segments_ir + index_ir * K,
                                                c->segments + index_audio * K, K);
            }
        }
        memcpy(c->conv, c->pre_multiplied, K * sizeof(orc_cpx)); /* :256-261 */
        orc_complex_multiply_accumulate(c->conv, c->segments + c->current * K, c->segments_ir, K);

        orc_rfft_inverse(c->fft, c->conv, c->fft_buffer);        /* :264-267 */

        for (size_t i = 0; i < processing; i++)                  /* sum(), :270-274 and :76-84 */
            output[processed + i] = c->fft_buffer[pos + i] + c->overlap[pos + i];

        c->input_buffer_fill += processing;                      /* :277-292 */
        if (c->input_buffer_fill == B) {
            memset(c->input_buffer, 0, B * sizeof(float));
            c->input_buffer_fill = 0;
            memcpy(c->overlap, c->fft_buffer + B, B * sizeof(float));
            c->current = c->current > 0 ? c->current - 1 : c->active_seg_count - 1;
        }
        processed += processing;
    }
    return ORC_OK;
}

/* src/fft_convolver.rs:296-306 */
void orc_fftconv_reset(orc_fftconv *c)
{
    size_t B = c->block_size, K = c->fft_complex_size;
    if (c->overlap) memset(c->overlap, 0, B * sizeof(float));
    if (c->segments) memset(c->segments, 0, c->seg_count * K * sizeof(orc_cpx));
    c->current = 0;
    if (c->input_buffer) memset(c->input_buffer, 0, B * sizeof(float));
    if (c->pre_multiplied) memset(c->pre_multiplied, 0, K * sizeof(orc_cpx));
    if (c->conv) memset(c->conv, 0, K * sizeof(orc_cpx));
    c->input_buffer_fill = 0;
}

size_t orc_fftconv_block_size(const orc_fftconv *c) { return c->block_size; }
size_t orc_fftconv_seg_count(const orc_fftconv *c) { return c->seg_count; }
size_t orc_fftconv_active_seg_count(const orc_fftconv *c) { return c->active_seg_count; }
size_t orc_fftconv_current(const orc_fftconv *c) { return c->current; }
size_t orc_fftconv_fill(const orc_fftconv *c) { return c->input_buffer_fill; }
const orc_cpx *orc_fftconv_segment_ir(const orc_fftconv *c, size_t i) { return c->segments_ir + i * c->fft_complex_size; }
const orc_cpx *orc_fftconv_segment(const orc_fftconv *c, size_t i) { return c->segments + i * c->fft_complex_size; }
const orc_cpx *orc_fftconv_premul(const orc_fftconv *c) { return c->pre_multiplied; }
const float *orc_fftconv_overlap(const orc_fftconv *c) { return c->overlap; }

/* ------------------------------------------------------------------------------------------
 * TwoStageFFTConvolver — src/fft_convolver.rs:323-526
 * ---------------------------------------------------------------------------------------- */
/* :514-526, all arithmetic in f32 */
size_t orc_compute_tail_block_size(size_t head_len, size_t response_len)
{
    const float FFT_K = 1.5f;
    float kn = (FFT_K * (float)head_len) / (2.0f * logf(2.0f));
    float b = -kn + sqrtf(kn * kn + (float)response_len * (float)head_len);
    b = fmaxf(b, (float)head_len);
    return next_power_of_two((size_t)b); /* `b as usize` truncates */
}

struct orc_twostage {
    size_t head_block_size, tail_block_size;
    orc_fftconv *head_convolver, *tail_convolver0, *tail_convolver;
    float *tail_output0, *tail_precalculated0, *tail_output, *tail_precalculated, *tail_input;
    size_t tail_input_fill, precalculated_pos;
    size_t max_response_length; /* not a field of the reference struct; only orc_twostage_update_ext reads it */
    /* EXTENSION (orc_twostage_init_multi): when not NULL the tail [2T, L) is itself a two-stage convolver whose head
     * block is T — the reference's partition applied recursively (Gardner-style non-uniform partition, SURVEY.md
     * §8(f)4); tail_convolver is then a Default convolver and unused */
    struct orc_twostage *tail_nested;
};

static orc_twostage *twostage_init_impl(const float *ir, size_t n_ir, size_t block_size, size_t max_response_length,
                                        size_t forced_tail, size_t stages, size_t max_block);

/* :340-406 */
orc_twostage *orc_twostage_init_tail(const float *ir, size_t n_ir, size_t block_size,
                                     size_t max_response_length, size_t forced_tail)
{
    return twostage_init_impl(ir, n_ir, block_size, max_response_length, forced_tail, 2, 0);
}

/* EXTENSION: `stages` > 2 nests the partition — the tail of every level but the last is again a two-stage convolver
 * (head block = this level's T, its own T from the same formula on what is left, capped at max_block when that is not
 * 0).  stages <= 2 is the reference's TwoStageFFTConvolver. */
orc_twostage *orc_twostage_init_multi(const float *ir, size_t n_ir, size_t block_size, size_t max_response_length,
                                      size_t stages, size_t max_block)
{
    return twostage_init_impl(ir, n_ir, block_size, max_response_length, 0, stages < 2 ? 2 : stages, max_block);
}

static orc_twostage *twostage_init_impl(const float *ir, size_t n_ir, size_t block_size, size_t max_response_length,
                                        size_t forced_tail, size_t stages, size_t max_block)
{
    size_t head = block_size;
    size_t T = forced_tail ? forced_tail : orc_compute_tail_block_size(block_size, max_response_length);
    if (!forced_tail && max_block && T > max_block) T = max_block;
    if (max_response_length < n_ir) return NULL; /* panic! :344-348 */
    size_t L = max_response_length;
    float *padded = (float *)calloc(L ? L : 1, sizeof(float));
    if (n_ir) memcpy(padded, ir, n_ir * sizeof(float));

    orc_twostage *c = (orc_twostage *)calloc(1, sizeof *c);
    c->head_block_size = head;
    c->tail_block_size = T;
    c->max_response_length = L;
    size_t head_ir_len = L < T ? L : T; /* :352-354 */
    c->head_convolver = orc_fftconv_init(padded, head_ir_len, head, head_ir_len);
    if (L > T) { /* :356-368 */
        size_t tail_ir_len = (L - T) < T ? (L - T) : T;
        c->tail_convolver0 = orc_fftconv_init(padded + T, tail_ir_len, head, tail_ir_len);
    } else {
        c->tail_convolver0 = orc_fftconv_default();
    }
    if (L > 2 * T && stages > 2) { /* EXTENSION: the tail is again a two-stage convolver fed T samples per call */
        size_t tail_ir_len = L - 2 * T;
        c->tail_nested = twostage_init_impl(padded + 2 * T, tail_ir_len, T, tail_ir_len, 0, stages - 1, max_block);
        c->tail_convolver = orc_fftconv_default();
    } else if (L > 2 * T) { /* :373-384 */
        size_t tail_ir_len = L - 2 * T;
        c->tail_convolver = orc_fftconv_init(padded + 2 * T, tail_ir_len, T, tail_ir_len);
    } else {
        c->tail_convolver = orc_fftconv_default();
    }
    c->tail_output0 = (float *)calloc(T, sizeof(float));        /* :370-371 */
    c->tail_precalculated0 = (float *)calloc(T, sizeof(float));
    c->tail_output = (float *)calloc(T, sizeof(float));         /* :386-388 */
    c->tail_precalculated = (float *)calloc(T, sizeof(float));
    c->tail_input = (float *)calloc(T, sizeof(float));
    free(padded);
    return c;
}

orc_twostage *orc_twostage_init(const float *ir, size_t n_ir, size_t block_size, size_t max_response_length)
{
    return orc_twostage_init_tail(ir, n_ir, block_size, max_response_length, 0);
}

orc_twostage *orc_twostage_clone(const orc_twostage *s)
{
    orc_twostage *c = (orc_twostage *)malloc(sizeof *c);
    *c = *s;
    size_t T = s->tail_block_size;
    c->head_convolver = orc_fftconv_clone(s->head_convolver);
    c->tail_convolver0 = orc_fftconv_clone(s->tail_convolver0);
    c->tail_convolver = orc_fftconv_clone(s->tail_convolver);
    c->tail_nested = s->tail_nested ? orc_twostage_clone(s->tail_nested) : NULL;
    c->tail_output0 = (float *)dup_mem(s->tail_output0, T * sizeof(float));
    c->tail_precalculated0 = (float *)dup_mem(s->tail_precalculated0, T * sizeof(float));
    c->tail_output = (float *)dup_mem(s->tail_output, T * sizeof(float));
    c->tail_precalculated = (float *)dup_mem(s->tail_precalculated, T * sizeof(float));
    c->tail_input = (float *)dup_mem(s->tail_input, T * sizeof(float));
    return c;
}

void orc_twostage_free(orc_twostage *c)
{
    if (!c) return;
    orc_fftconv_free(c->head_convolver); orc_fftconv_free(c->tail_convolver0); orc_fftconv_free(c->tail_convolver);
    orc_twostage_free(c->tail_nested);
    free(c->tail_output0); free(c->tail_precalculated0); free(c->tail_output);
    free(c->tail_precalculated); free(c->tail_input);
    free(c);
}

/* :408-410 — todo!() */
int orc_twostage_update(orc_twostage *c, const float *ir, size_t len)
{
    (void)c; (void)ir; (void)len;
    return ORC_PANIC;
}

/* EXTENSION — not in the reference (its method is todo!(), :408-410; SURVEY.md §8(f)2 asks for it).
 * Defined as the per-stage FFTConvolver::update (:174-213) on the re-sliced response, sliced exactly
 * like init slices it (:349-384): the response is zero-padded to max_response_length, head gets
 * [0, min(L,T)), tail0 [T, T + min(L-T,T)) when L > T, tail [2T, L) when L > 2T.  Every stage is
 * handed its FULL slice length, so active_seg_count stays seg_count and no ring slot is re-read
 * modulo a new count.  Like FFTConvolver::update it keeps everything already heard: the input
 * rings, the partially filled blocks, tail_input and the tail outputs computed with the old
 * response (tail_output*, tail_precalculated* — audio already in flight); it zeroes each stage's
 * overlap and pre_multiplied (:185-188).  Panics like FFTConvolver::update (:177-179) when the
 * response is longer than max_response_length. */
int orc_twostage_update_ext(orc_twostage *c, const float *ir, size_t len)
{
    size_t L = c->max_response_length, T = c->tail_block_size;
    if (len > L) return ORC_PANIC;
    float *padded = (float *)calloc(L ? L : 1, sizeof(float));
    if (len) memcpy(padded, ir, len * sizeof(float));
    int rc = orc_fftconv_update(c->head_convolver, padded, L < T ? L : T);
    if (!rc && L > T) rc = orc_fftconv_update(c->tail_convolver0, padded + T, (L - T) < T ? (L - T) : T);
    if (!rc && L > 2 * T)
        rc = c->tail_nested ? orc_twostage_update_ext(c->tail_nested, padded + 2 * T, L - 2 * T)
                            : orc_fftconv_update(c->tail_convolver, padded + 2 * T, L - 2 * T);
    free(padded);
    return rc;
}

static void swap_ptr(float **a, float **b) { float *t = *a; *a = *b; *b = t; }

/* :412-495 */
int orc_twostage_process(orc_twostage *c, const float *input, size_t in_len, float *output, size_t out_len)
{
    if (!(in_len <= c->head_block_size)) return ORC_PANIC; /* assert! :414 */
    /* head.process slices input[..output.len()] (needs in_len >= out_len) and the tail loop
     * indexes output[..input.len()] (needs out_len >= in_len): anything else panics */
    if (in_len != out_len) return ORC_PANIC;
    size_t H = c->head_block_size, T = c->tail_block_size;

    orc_fftconv_process(c->head_convolver, input, in_len, output, out_len); /* :417 */
    if (T == 0) return ORC_OK; /* tail_input.is_empty() :420-422 */

    size_t len = in_len, processed = 0;
    while (processed < len) { /* :427 */
        size_t remaining = len - processed;
        size_t processing = H - (c->tail_input_fill % H); /* :429-432 */
        if (remaining < processing) processing = remaining;
        size_t sum_begin = processed, sum_end = processed + processing;

        if (c->precalculated_pos + processing > T || c->tail_input_fill + processing > T)
            return ORC_PANIC; /* index / slice panic at :442 / :459 (head size not dividing T) */
        { /* :439-445 */
            size_t pp = c->precalculated_pos;
            for (size_t i = sum_begin; i < sum_end; i++) output[i] += c->tail_precalculated0[pp++];
        }
        { /* :448-454 */
            size_t pp = c->precalculated_pos;
            for (size_t i = sum_begin; i < sum_end; i++) output[i] += c->tail_precalculated[pp++];
        }
        c->precalculated_pos += processing; /* :456 */

        memcpy(c->tail_input + c->tail_input_fill, input + processed, processing * sizeof(float)); /* :459-461 */
        c->tail_input_fill += processing;

        if (c->tail_input_fill % H == 0) { /* :464-476 */
            size_t block_offset = c->tail_input_fill - H;
            orc_fftconv_process(c->tail_convolver0, c->tail_input + block_offset, H,
                                c->tail_output0 + block_offset, H);
            if (c->tail_input_fill == T) swap_ptr(&c->tail_precalculated0, &c->tail_output0);
        }
        if (c->tail_input_fill == T) { /* :479-486 */
            swap_ptr(&c->tail_precalculated, &c->tail_output);
            if (c->tail_nested) orc_twostage_process(c->tail_nested, c->tail_input, T, c->tail_output, T);
            else orc_fftconv_process(c->tail_convolver, c->tail_input, T, c->tail_output, T);
        }
        if (c->tail_input_fill == T) { /* :488-491 */
            c->tail_input_fill = 0;
            c->precalculated_pos = 0;
        }
        processed += processing;
    }
    return ORC_OK;
}

/* :497-511 */
void orc_twostage_reset(orc_twostage *c)
{
    size_t T = c->tail_block_size;
    orc_fftconv_reset(c->head_convolver);
    orc_fftconv_reset(c->tail_convolver0);
    memset(c->tail_output0, 0, T * sizeof(float));
    memset(c->tail_precalculated0, 0, T * sizeof(float));
    orc_fftconv_reset(c->tail_convolver);
    if (c->tail_nested) orc_twostage_reset(c->tail_nested);
    memset(c->tail_output, 0, T * sizeof(float));
    memset(c->tail_precalculated, 0, T * sizeof(float));
    memset(c->tail_input, 0, T * sizeof(float));
    c->tail_input_fill = 0;
    c->precalculated_pos = 0;
}

size_t orc_twostage_tail_block_size(const orc_twostage *c) { return c->tail_block_size; }
/* block sizes of the nested partition, outermost first: head, T1, T2, ...; returns how many */
size_t orc_twostage_stage_blocks(const orc_twostage *c, size_t *out, size_t cap)
{
    size_t n = 0;
    if (n < cap) out[n] = c->head_block_size;
    n++;
    for (; c; c = c->tail_nested) {
        if (n < cap) out[n] = c->tail_block_size;
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Crossfader<RaisedCosineMixer> — src/crossfade_convolver.rs:160-279
 * ---------------------------------------------------------------------------------------- */
/* :147, :160-169 */
float orc_raised_cosine_mix(float a, float b, float value)
{
    const float PI_HALF = 3.14159265358979323846f * 0.5f;
    float rad = PI_HALF * value;
    float cs = cosf(rad);
    float gain1 = cs * cs; /* powi(2) */
    float gain2 = 1.0f - gain1;
    return a * gain1 + b * gain2;
}

/* :204-214 */
void orc_crossfader_new(orc_crossfader *x, size_t fading_samples, size_t hold_samples)
{
    x->fading_samples = (int64_t)fading_samples;
    x->hold_samples = (int64_t)hold_samples;
    x->counter = 0;
    x->mix_value_step = 1.0f / (float)fading_samples;
    x->mix_value = 0.0f;
    x->approaching = 0;
    x->target = 0;
}

/* :216-240 */
void orc_crossfader_fade_into(orc_crossfader *x, int target)
{
    if (x->target == target) return;
    if (!x->approaching) {
        x->counter = -x->hold_samples;
        x->approaching = 1;
        x->target = target;
        x->mix_value_step = -x->mix_value_step;
    } else if (x->counter >= 0) {
        x->counter = x->fading_samples - x->counter;
        x->target = target;
        x->mix_value_step = -x->mix_value_step;
    } else {
        x->approaching = 0;
        x->target = target;
    }
}

/* :242-278 */
float orc_crossfader_mix(orc_crossfader *x, float a, float b)
{
    if (!x->approaching) return x->target == 0 ? a : b;
    x->counter += 1;
    if (x->counter <= 0) return x->target == 0 ? b : a; /* holding the previous target */
    x->mix_value += x->mix_value_step;
    if (x->counter == x->fading_samples) {
        x->approaching = 0;
        if (x->target == 0) { x->mix_value = 0.0f; return a; }
        x->mix_value = 1.0f;
        return b;
    }
    return orc_raised_cosine_mix(a, b, x->mix_value);
}

/* ------------------------------------------------------------------------------------------
 * CrossfadeConvolver<FFTConvolver> — src/crossfade_convolver.rs:3-105
 * ---------------------------------------------------------------------------------------- */
struct orc_crossfade {
    orc_fftconv *convolver_a, *convolver_b;
    orc_crossfader crossfader;
    float *buffer_a, *buffer_b; /* max_buffer_size each */
    size_t max_buffer_size;
    float *stored_response;
    size_t stored_len;
    int response_pending;
};

/* :19-43 */
orc_crossfade *orc_crossfade_new(orc_fftconv *convolver, size_t max_response_length,
                                 size_t max_buffer_size, size_t crossfade_samples)
{
    orc_crossfade *c = (orc_crossfade *)calloc(1, sizeof *c);
    c->convolver_a = orc_fftconv_clone(convolver);
    c->convolver_b = convolver;
    size_t hold = max_buffer_size < max_response_length ? max_buffer_size : max_response_length;
    orc_crossfader_new(&c->crossfader, crossfade_samples, hold);
    c->buffer_a = (float *)calloc(max_buffer_size ? max_buffer_size : 1, sizeof(float));
    c->buffer_b = (float *)calloc(max_buffer_size ? max_buffer_size : 1, sizeof(float));
    c->max_buffer_size = max_buffer_size;
    c->stored_response = (float *)calloc(max_response_length ? max_response_length : 1, sizeof(float));
    c->stored_len = max_response_length;
    c->response_pending = 0;
    return c;
}

/* :46-49 — note response.len() is passed for both the stored capacity and crossfade_samples */
orc_crossfade *orc_crossfade_init(const float *ir, size_t n_ir, size_t max_block_size, size_t max_response_length)
{
    orc_fftconv *conv = orc_fftconv_init(ir, n_ir, max_block_size, max_response_length);
    if (!conv) return NULL;
    return orc_crossfade_new(conv, n_ir, max_block_size, n_ir);
}

void orc_crossfade_free(orc_crossfade *c)
{
    if (!c) return;
    orc_fftconv_free(c->convolver_a); orc_fftconv_free(c->convolver_b);
    free(c->buffer_a); free(c->buffer_b); free(c->stored_response);
    free(c);
}

int orc_crossfade_is_crossfading(const orc_crossfade *c) { return c->crossfader.approaching; } /* :85-92 */
const orc_crossfader *orc_crossfade_crossfader(const orc_crossfade *c) { return &c->crossfader; }

/* :94-105 */
static int crossfade_swap(orc_crossfade *c, const float *response, size_t len)
{
    int rc;
    if (c->crossfader.target == 0) {
        rc = orc_fftconv_update(c->convolver_b, response, len);
        orc_crossfader_fade_into(&c->crossfader, 1);
    } else {
        rc = orc_fftconv_update(c->convolver_a, response, len);
        orc_crossfader_fade_into(&c->crossfader, 0);
    }
    return rc;
}

/* :51-64 */
int orc_crossfade_update(orc_crossfade *c, const float *response, size_t len)
{
    if (!orc_crossfade_is_crossfading(c)) {
        int rc = crossfade_swap(c, response, len);
        c->response_pending = 0;
        return rc;
    }
    if (!(len <= c->stored_len)) return ORC_PANIC; /* assert! :59 */
    memcpy(c->stored_response, response, len * sizeof(float));
    memset(c->stored_response + len, 0, (c->stored_len - len) * sizeof(float));
    c->response_pending = 1;
    return ORC_OK;
}

/* :66-78 */
int orc_crossfade_process(orc_crossfade *c, const float *input, size_t in_len, float *output, size_t out_len)
{
    if (!orc_crossfade_is_crossfading(c) && c->response_pending) {
        if (crossfade_swap(c, c->stored_response, c->stored_len)) return ORC_PANIC;
        c->response_pending = 0;
    }
    /* both inner calls are sized by buffer_a/b.len() == max_buffer_size (:72-73) */
    if (out_len > c->max_buffer_size) return ORC_PANIC; /* buffer_a[i] index at :76 */
    if (orc_fftconv_process(c->convolver_a, input, in_len, c->buffer_a, c->max_buffer_size)) return ORC_PANIC;
    if (orc_fftconv_process(c->convolver_b, input, in_len, c->buffer_b, c->max_buffer_size)) return ORC_PANIC;
    for (size_t i = 0; i < out_len; i++)
        output[i] = orc_crossfader_mix(&c->crossfader, c->buffer_a[i], c->buffer_b[i]);
    return ORC_OK;
}

/* :80-82 — todo!() */
int orc_crossfade_reset(orc_crossfade *c) { (void)c; return ORC_PANIC; }

/* EXTENSION — not in the reference (its method is todo!(), :80-82; SURVEY.md §8(f)2 asks for it).
 * Defined by analogy with FFTConvolver::reset (src/fft_convolver.rs:296-306: forget all audio,
 * keep the responses): both convolvers are reset, buffer_a / buffer_b are zeroed, and a fade in
 * progress is completed at once — the crossfader lands where `mix` would have left it when
 * counter == fading_samples (:261-273: Reached(target), mix_value 0.0 for A / 1.0 for B; the step
 * keeps the sign fade_into gave it), so the most recently requested response is the one heard.
 * A response still pending (stored_response, :58-63) stays pending and is swapped in by the next
 * process() exactly as :67-70 does when no fade is running. */
int orc_crossfade_reset_ext(orc_crossfade *c)
{
    orc_fftconv_reset(c->convolver_a);
    orc_fftconv_reset(c->convolver_b);
    memset(c->buffer_a, 0, (c->max_buffer_size ? c->max_buffer_size : 1) * sizeof(float));
    memset(c->buffer_b, 0, (c->max_buffer_size ? c->max_buffer_size : 1) * sizeof(float));
    if (c->crossfader.approaching) {
        c->crossfader.approaching = 0;
        c->crossfader.mix_value = c->crossfader.target == 0 ? 0.0f : 1.0f;
    }
    c->crossfader.counter = 0;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * synthetic data
 * ---------------------------------------------------------------------------------------- */
uint64_t orc_mix64(uint64_t v)
{
    uint64_t z = v + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_gen_noise(float *x, uint64_t channel, size_t first_sample, size_t n)
{
    const uint64_t seed_x = 0x5EED0001ull;
    for (size_t i = 0; i < n; i++) {
        uint64_t r = orc_mix64(seed_x + (channel << 32) + (uint64_t)(first_sample + i));
        float u = (float)(r >> 40) / 16777216.0f;
        x[i] = 2.0f * u - 1.0f;
    }
}

void orc_gen_ir(float *h, uint64_t channel, uint64_t update_index, size_t len)
{
    const uint64_t seed_h = 0x5EED0002ull + (update_index << 48);
    double energy = 0.0;
    double *v = (double *)malloc(sizeof(double) * (len ? len : 1));
    for (size_t i = 0; i < len; i++) {
        uint64_t r = orc_mix64(seed_h + (channel << 32) + (uint64_t)i);
        double u = (double)(r >> 40) / 16777216.0;
        v[i] = (2.0 * u - 1.0) * exp(-6.9078 * (double)i / (double)len);
        energy += v[i] * v[i];
    }
    double s = energy > 0.0 ? 1.0 / sqrt(energy) : 0.0;
    for (size_t i = 0; i < len; i++) h[i] = (float)(v[i] * s);
    free(v);
}

void orc_direct_conv_f64(const float *x, size_t nx, const float *h, size_t nh, double *y)
{
    for (size_t n = 0; n < nx; n++) {
        double acc = 0.0;
        size_t kmax = n + 1 < nh ? n + 1 : nh;
        for (size_t k = 0; k < kmax; k++) acc += (double)h[k] * (double)x[n - k];
        y[n] = acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * multi-threaded CPU baseline: one FFTConvolver per channel, channels
 * statically partitioned over OpenMP threads, wall clock around the block loop only, the
 * shape of examples/compare_partitioned.rs:28-39.
 * ---------------------------------------------------------------------------------------- */
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double orc_batch_fftconv_run(size_t channels, size_t block_size, size_t ir_len, const float *irs,
                             const float *in, float *out, size_t n_per_call, size_t calls, int threads)
{
    orc_fftconv **cv = (orc_fftconv **)malloc(sizeof(*cv) * channels);
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long c = 0; c < (long)channels; c++)
        cv[c] = orc_fftconv_init(irs + (size_t)c * ir_len, ir_len, block_size, ir_len);
    size_t total = n_per_call * calls;
    double t0 = now_s();
#pragma omp parallel num_threads(threads)
    {
        /* real-time order: every thread walks its own contiguous channel range once per block
         * (an audio callback serves all of its channels each block period), so per-channel
         * state is streamed, not kept cache-hot across blocks */
#ifdef _OPENMP
        size_t t = (size_t)omp_get_thread_num(), nt = (size_t)omp_get_num_threads();
#else
        size_t t = 0, nt = 1;
#endif
        size_t c0 = channels * t / nt, c1 = channels * (t + 1) / nt;
        for (size_t k = 0; k < calls; k++)
            for (size_t c = c0; c < c1; c++)
                orc_fftconv_process(cv[c], in + c * total + k * n_per_call, n_per_call,
                                    out + c * total + k * n_per_call, n_per_call);
    }
    double t1 = now_s();
    for (size_t c = 0; c < channels; c++) orc_fftconv_free(cv[c]);
    free(cv);
    return t1 - t0;
}

/* C independent TwoStageFFTConvolvers (BASELINE configs[1]); calls of n_per_call <= head samples.
 * update_every > 0: orc_twostage_update_ext with irs_upd[(k / update_every - 1) % n_upd] before call k */
double orc_batch_twostage_run(size_t channels, size_t head_block, size_t ir_len, size_t forced_tail,
                              const float *irs, const float *irs_upd, size_t n_upd, size_t update_every,
                              const float *in, float *out, size_t n_per_call, size_t calls, int threads)
{
    orc_twostage **cv = (orc_twostage **)malloc(sizeof(*cv) * channels);
    if (threads < 1) threads = 1;
    size_t total = n_per_call * calls;
    double t0 = now_s();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long c = 0; c < (long)channels; c++) {
        cv[c] = orc_twostage_init_tail(irs + (size_t)c * ir_len, ir_len, head_block, ir_len, forced_tail);
        for (size_t k = 0; k < calls; k++) {
            if (update_every && n_upd && k && k % update_every == 0)
                orc_twostage_update_ext(cv[c], irs_upd + (((k / update_every - 1) % n_upd) * channels + (size_t)c) * ir_len, ir_len);
            orc_twostage_process(cv[c], in + (size_t)c * total + k * n_per_call, n_per_call,
                                 out + (size_t)c * total + k * n_per_call, n_per_call);
        }
        orc_twostage_free(cv[c]);
    }
    double t1 = now_s();
    free(cv);
    return t1 - t0;
}

/* C independent CrossfadeConvolver::init(h, block, ir_len) (BASELINE configs[2]); every call is one
 * whole block; update(irs_upd[(k / update_every - 1) % n_upd]) before call k for k = update_every, 2*update_every, ... */
double orc_batch_crossfade_run(size_t channels, size_t block, size_t ir_len, const float *irs,
                               const float *irs_upd, size_t n_upd, size_t update_every, const float *in,
                               float *out, size_t calls, int threads)
{
    if (threads < 1) threads = 1;
    size_t total = block * calls;
    double t0 = now_s();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long c = 0; c < (long)channels; c++) {
        orc_crossfade *x = orc_crossfade_init(irs + (size_t)c * ir_len, ir_len, block, ir_len);
        for (size_t k = 0; k < calls; k++) {
            if (update_every && n_upd && k && k % update_every == 0)
                orc_crossfade_update(x, irs_upd + (((k / update_every - 1) % n_upd) * channels + (size_t)c) * ir_len, ir_len);
            orc_crossfade_process(x, in + (size_t)c * total + k * block, block, out + (size_t)c * total + k * block, block);
        }
        orc_crossfade_free(x);
    }
    return now_s() - t0;
}
