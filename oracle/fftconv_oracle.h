/*
 * fftconv_oracle.h — CPU oracle for the partitioned-FFT-convolution hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fft_convolution_b200/ may include, link or
 * call this.  Allowed users: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs (as the checker / the reported CPU baseline, never as the product).
 *
 * It restates, in plain C, the algorithm of the reference crate Sin-tel/fft-convolution:
 *   FFTConvolver          src/fft_convolver.rs:86-307
 *   TwoStageFFTConvolver  src/fft_convolver.rs:323-526
 *   CrossfadeConvolver    src/crossfade_convolver.rs:3-105
 *   Crossfader / mixer    src/crossfade_convolver.rs:126-279
 *
 * Parity status.  The reference's FFT arithmetic lives in the third-party crates
 * realfft 3.3 / rustfft 6.1 (Cargo.toml:7-8, unpinned, sources not on this machine, no
 * Rust toolchain) so the reference itself cannot run here: FFT bit patterns are
 * PARITY UNPINNED.  Everything the reference's own tests pin IS checked against this
 * oracle (tests/test_oracle_reference_tests.py restates all 10 reference tests: the three
 * delta-IR known answers, test_crossfader's exact equalities and the six behavioural tests
 * of src/tests.rs), and it is cross-checked against an independent numpy restatement
 * (oracle/oracle_np.py, pocketfft) and an f64 direct convolution.
 */
#ifndef FFTCONV_ORACLE_H
#define FFTCONV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } orc_cpx;

/* return codes: 0 ok; ORC_PANIC = the reference would panic!/assert!/todo!() here */
#define ORC_OK 0
#define ORC_PANIC 1

/* ---- real FFT stand-in for realfft/rustfft (src/fft_convolver.rs:7-50) ---- */
typedef struct orc_plan orc_plan;
orc_plan *orc_plan_new(size_t n);            /* n = real length, power of two >= 2, or 0 */
void orc_plan_free(orc_plan *p);
/* unnormalised forward: n reals -> n/2+1 complex.  `in` is NOT clobbered here. */
void orc_rfft_forward(const orc_plan *p, const float *in, orc_cpx *out);
/* Fft::inverse: unnormalised C2R followed by division of every sample by n (:41-49) */
void orc_rfft_inverse(const orc_plan *p, const orc_cpx *in, float *out);

/* free helpers (src/fft_convolver.rs:52-84) */
size_t orc_complex_size(size_t n);
void orc_complex_multiply_accumulate(orc_cpx *result, const orc_cpx *a, const orc_cpx *b, size_t len);

/* ---- FFTConvolver (src/fft_convolver.rs:86-307) ---- */
typedef struct orc_fftconv orc_fftconv;
orc_fftconv *orc_fftconv_init(const float *ir, size_t ir_len, size_t block_size, size_t max_response_length);
orc_fftconv *orc_fftconv_default(void);
orc_fftconv *orc_fftconv_clone(const orc_fftconv *c);
void orc_fftconv_free(orc_fftconv *c);
int orc_fftconv_update(orc_fftconv *c, const float *ir, size_t len);
void orc_fftconv_reset(orc_fftconv *c);
int orc_fftconv_process(orc_fftconv *c, const float *in, size_t in_len, float *out, size_t out_len);
/* introspection (tests compare device state against these) */
size_t orc_fftconv_block_size(const orc_fftconv *c);
size_t orc_fftconv_seg_count(const orc_fftconv *c);
size_t orc_fftconv_active_seg_count(const orc_fftconv *c);
size_t orc_fftconv_current(const orc_fftconv *c);
size_t orc_fftconv_fill(const orc_fftconv *c);
const orc_cpx *orc_fftconv_segment_ir(const orc_fftconv *c, size_t i);
const orc_cpx *orc_fftconv_segment(const orc_fftconv *c, size_t i);
const orc_cpx *orc_fftconv_premul(const orc_fftconv *c);
const float *orc_fftconv_overlap(const orc_fftconv *c);

/* ---- TwoStageFFTConvolver (src/fft_convolver.rs:323-526) ---- */
size_t orc_compute_tail_block_size(size_t head_len, size_t response_len);
typedef struct orc_twostage orc_twostage;
orc_twostage *orc_twostage_init(const float *ir, size_t ir_len, size_t block_size, size_t max_response_length);
/* same, but with the tail block forced (0 = derive as the reference does) — only for the
 * BASELINE.json "tail 4096" wording, see SURVEY.md §8 config-2 note */
orc_twostage *orc_twostage_init_tail(const float *ir, size_t ir_len, size_t block_size,
                                     size_t max_response_length, size_t forced_tail);
/* EXTENSION beyond the reference: `stages` > 2 nests the two-stage partition (the tail of every level but the last is
 * again a two-stage convolver with head block T; each level's T from the reference's own formula, capped at
 * max_block when non-zero) — a Gardner-style non-uniform partition built from the reference's scheme */
orc_twostage *orc_twostage_init_multi(const float *ir, size_t ir_len, size_t block_size, size_t max_response_length,
                                      size_t stages, size_t max_block);
size_t orc_twostage_stage_blocks(const orc_twostage *c, size_t *out, size_t cap);
orc_twostage *orc_twostage_clone(const orc_twostage *c);
void orc_twostage_free(orc_twostage *c);
int orc_twostage_update(orc_twostage *c, const float *ir, size_t len); /* todo!() => ORC_PANIC */
/* EXTENSION beyond the reference (which leaves it todo!()): per-stage FFTConvolver::update on the
 * response re-sliced like init does; semantics written out above its definition */
int orc_twostage_update_ext(orc_twostage *c, const float *ir, size_t len);
void orc_twostage_reset(orc_twostage *c);
int orc_twostage_process(orc_twostage *c, const float *in, size_t in_len, float *out, size_t out_len);
size_t orc_twostage_tail_block_size(const orc_twostage *c);

/* ---- Crossfader<RaisedCosineMixer> (src/crossfade_convolver.rs:160-279) ---- */
typedef struct {
    int64_t fading_samples, hold_samples, counter;
    float mix_value_step, mix_value;
    int approaching; /* 0 = Reached, 1 = Approaching */
    int target;      /* 0 = A, 1 = B */
} orc_crossfader;
void orc_crossfader_new(orc_crossfader *x, size_t fading_samples, size_t hold_samples);
void orc_crossfader_fade_into(orc_crossfader *x, int target);
float orc_crossfader_mix(orc_crossfader *x, float a, float b);
float orc_raised_cosine_mix(float a, float b, float value);

/* ---- CrossfadeConvolver<FFTConvolver> (src/crossfade_convolver.rs:3-105) ---- */
typedef struct orc_crossfade orc_crossfade;
/* CrossfadeConvolver::new(convolver, max_response_length, max_buffer_size, crossfade_samples);
 * takes ownership of `convolver` */
orc_crossfade *orc_crossfade_new(orc_fftconv *convolver, size_t max_response_length,
                                 size_t max_buffer_size, size_t crossfade_samples);
/* <CrossfadeConvolver as Convolution>::init */
orc_crossfade *orc_crossfade_init(const float *ir, size_t ir_len, size_t max_block_size,
                                  size_t max_response_length);
void orc_crossfade_free(orc_crossfade *c);
int orc_crossfade_update(orc_crossfade *c, const float *ir, size_t len);
int orc_crossfade_process(orc_crossfade *c, const float *in, size_t in_len, float *out, size_t out_len);
int orc_crossfade_reset(orc_crossfade *c); /* todo!() => ORC_PANIC */
/* EXTENSION beyond the reference (which leaves it todo!()): forget all audio, finish a running fade
 * at once, keep the responses and a pending update; semantics written out above its definition */
int orc_crossfade_reset_ext(orc_crossfade *c);
int orc_crossfade_is_crossfading(const orc_crossfade *c);
const orc_crossfader *orc_crossfade_crossfader(const orc_crossfade *c);

/* ---- synthetic data of SURVEY.md §8(d) (stateless splitmix64 hash) ---- */
uint64_t orc_mix64(uint64_t v);
void orc_gen_noise(float *x, uint64_t channel, size_t first_sample, size_t n);
void orc_gen_ir(float *h, uint64_t channel, uint64_t update_index, size_t len);

/* ---- f64 truth: y[n] = sum_k h[k] x[n-k], n < nx ---- */
void orc_direct_conv_f64(const float *x, size_t nx, const float *h, size_t nh, double *y);

/* ---- multi-threaded CPU baseline: C independent FFTConvolvers, one per channel,
 * channels statically partitioned over `threads` OpenMP threads.
 * Returns seconds spent in the block loop only (init excluded).  in/out: [C][blocks*n]. */
double orc_batch_fftconv_run(size_t channels, size_t block_size, size_t ir_len,
                             const float *irs /* [C][ir_len] */, const float *in, float *out,
                             size_t n_per_call, size_t calls, int threads);
int orc_max_threads(void);
/* the same for C TwoStageFFTConvolvers / C CrossfadeConvolver::init (full-count parity of BASELINE
 * configs[1] and configs[2]); irs_upd: [n_upd][C][ir_len], applied every `update_every` calls */
double orc_batch_twostage_run(size_t channels, size_t head_block, size_t ir_len, size_t forced_tail,
                              const float *irs, const float *irs_upd, size_t n_upd, size_t update_every,
                              const float *in, float *out, size_t n_per_call, size_t calls, int threads);
double orc_batch_crossfade_run(size_t channels, size_t block, size_t ir_len, const float *irs,
                               const float *irs_upd, size_t n_upd, size_t update_every, const float *in,
                               float *out, size_t calls, int threads);

#ifdef __cplusplus
}
#endif
#endif
