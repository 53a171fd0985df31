/*
 * fftconv_b200.h — C ABI of the B200-native partitioned-FFT-convolution engine.
 *
 * This is the drop-in boundary for the hot path of Sin-tel/fft-convolution
 * (/root/reference): the per-block inner loop of FFTConvolver::process
 * (src/fft_convolver.rs:215-295) as reused by TwoStageFFTConvolver (:412-495) and
 * CrossfadeConvolver (src/crossfade_convolver.rs:66-78).  Plain pointers and sizes only.
 *
 * Two layers are exported from libfftconv_b200.so:
 *
 *  1. fcb_engine_*   — the device stages.  This is what a Rust `extern "C"` FFI crate binds
 *     when the Rust host keeps the block scheduler, the input-buffer fill and the segment
 *     ring rotation (INTEGRATION.md shows the binding).  One engine = C lock-step channels
 *     of one FFTConvolver shape (block size B, S segments); the caller passes the scalars
 *     the reference keeps in `current`, `input_buffer_fill`, `active_seg_count`.
 *
 *  2. fcb_fftconv_* / fcb_twostage_* / fcb_crossfade_*   — a C++ host mirror of the
 *     reference's three `Convolution` implementors (scheduler included), batched over C
 *     channels, for C/C++/Python callers that have no Rust host.  With C = 1 the semantics,
 *     argument meaning and failure conditions are those of the reference types.
 *
 * There is no CPU fallback: every entry point needs a CUDA device (sm_100a build).
 *
 * Layouts.  Host/device sample buffers are planar [channel][sample] f32 with an explicit
 * channel stride in samples.  Spectra live only on the device, packed B complex per row
 * (bin 0 = {DC.re, Nyquist.re}; bins 1..B-1 interleaved re,im) — see DESIGN.md.
 */
#ifndef FFTCONV_B200_H
#define FFTCONV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------- */
#define FCB_OK 0
#define FCB_ERR_PANIC 1       /* contract violation: the reference panic!s / assert!s here */
#define FCB_ERR_TODO 2        /* the reference method is todo!() (TwoStage::update, Crossfade::reset) */
#define FCB_ERR_CUDA 3        /* CUDA runtime / launch failure (sticky text in fcb_last_error) */
#define FCB_ERR_UNSUPPORTED 4 /* valid in the reference, outside this engine's limits (e.g. B > 16384) */
#define FCB_ERR_ARG 5         /* NULL / out-of-range argument at the C boundary */

/* thread-local, never NULL */
const char *fcb_last_error(void);
/* "fftconv_b200 <version> sm_100a" */
const char *fcb_version(void);
/* kernels launched by this library in this process (bench.py's gpu_launches) */
uint64_t fcb_launch_count(void);
/* number of visible CUDA devices, or -1 */
int fcb_device_count(void);
/* test hook: device / pinned allocations, stream and event creations and their releases made by this library so far.
 * update() and process() must leave it unchanged in the steady state (the reference's real-time rule, src/lib.rs:8);
 * tests/test_gpu_realtime.py pins that. */
uint64_t fcb_debug_alloc_count(void);

/* tuning knobs for benchmarking sweeps: "mac_impl" (0 auto, 1 LDG kernel, 2 TMA pipeline),
 * "mac_stages" (2, 3, 4, 6 pipeline stages of 32 KB), "pipe_group" (channels per group of the
 * end-to-end copy/compute pipeline, default 512), "mimo_tile" (1 = matrix K2 with in-CTA reuse of
 * IR and ring tiles, 0 = the per-channel K2), "mimo_tc" (the tensor-core matrix MAC K4: 0 = never, 1 = always,
 * 2 = when at least "mimo_tc_min" streams share the matrix; read by fcb_mimo_create), "mimo_rt" (1 = fewer streams than that run
 * the register-tiled matrix MAC k_mac_rt, 0 = the shared-memory tile kernel; sweeps: "mimo_rt_min" = streams from which it is used, "mimo_rt_waves" = waves
 * of CTAs its segment chunking aims at (0 = by stream count), "mimo_rt_r" = segments per pipeline stage (4 or 2), "mimo_rt_wb" = warps side by side along the bins), "fused_block" (1 = whole blocks with B in 32..512 run as
 * one fused K1+K2+K3 kernel, 0 = three launches; outputs are bit-identical), "fused_stages" (2 or 3), "fused_pair" (1 = convolvers fed the same input share one launch), "fused_short" (delay lines of up to this many segments run
 * the fused kernel with 2-row stages so that a fourth CTA per SM hides the FFT latency; default 40, 0 = off), "mapped_io" (1 = small-batch host
 * calls go through mapped pinned memory instead of the copy engines), "zero_copy" (1 = when the caller's
 * buffers are pinned, large-batch whole-block calls let the kernel read/write them over PCIe directly), "split" (1 = small
 * batches cut each delay line of a whole-block launch over several CTAs, partial sums added in slice order by the last
 * CTA to arrive: deterministic, within tolerance, not bit-equal to the unsplit order; 0 = never; sweeps: "split_slots" = CTAs
 * the split aims at, "split_min_stages" = pipeline stages per CTA at least), "k1_late" (1 = the fused kernel transforms a block
 * that sits in host memory AFTER its MAC stream; measured without effect, default 0), "tma_io" (1 = the whole-block
 * kernels move their input / output blocks with bulk copies when the buffers allow it, 0 = through registers),
 * "shared_reuse" (1 = engines with one IR for all channels stage each IR tile once per CTA and reuse it for the CTA's
 * channels, 0 = the per-channel kernel with stride 0), "strict_todo"
 * (1 = fcb_twostage_update and fcb_crossfade_reset answer FCB_ERR_TODO like the reference's todo!(); default 0 = the
 * extensions documented at those entry points), "xf_speculate" (1 = the synchronous crossfade host call computes the
 * NEXT block's per-sample gains while the GPU works on the current one — same gains, same state, off the critical
 * path; 0 = every block computes its own before its launch) */
int fcb_tune(const char *key, int value);

/* live timing of the K2 launches: while enabled every K2 launch is bracketed by CUDA events on
 * its stream; read returns the summed device time and the number of launches since enable */
int fcb_profile_mac(int enable);
int fcb_profile_mac_read(double *total_ms, uint64_t *launches);

/* pinned host memory for the end-to-end path (cudaHostAlloc / cudaFreeHost) */
void *fcb_host_alloc(size_t bytes);
void fcb_host_free(void *p);

/* ============================================================================================
 * Layer 1: device stages (replaces the arithmetic of src/fft_convolver.rs:7-84 and the loop
 * bodies at :131-142, :193-212, :234-274, :283-284).
 * ========================================================================================== */
typedef struct fcb_engine fcb_engine;

typedef struct {
    size_t channels;            /* C >= 1 lock-step channels */
    size_t block_size;          /* rounded up to a power of two like :115; 1..16384 */
    size_t max_response_length; /* ir_len of :111-113; seg_count = ceil(len / B) (:117) */
    int shared_ir;              /* 1: all channels use one IR (spectra stored once, reused on chip) */
    int device;                 /* CUDA device ordinal */
    void *stream;               /* cudaStream_t to run on, NULL = a private non-blocking stream */
} fcb_engine_desc;

int fcb_engine_create(const fcb_engine_desc *desc, fcb_engine **out);
void fcb_engine_destroy(fcb_engine *e);
/* #[derive(Clone)] (src/fft_convolver.rs:86): deep copy of ring, spectra, overlap, input buffer */
int fcb_engine_clone(const fcb_engine *e, fcb_engine **out);
int fcb_engine_set_stream(fcb_engine *e, void *stream);
void *fcb_engine_stream(const fcb_engine *e);
int fcb_engine_sync(fcb_engine *e);

size_t fcb_engine_channels(const fcb_engine *e);
size_t fcb_engine_block_size(const fcb_engine *e); /* rounded B */
size_t fcb_engine_seg_count(const fcb_engine *e);  /* S */

/* K5 — IR preparation (src/fft_convolver.rs:131-142 for init, :185-212 for update).
 * irs: [nchan][len] f32 with channel stride `stride` samples, host or device memory.
 * Rows >= ceil(len/B) are zeroed.  is_update != 0 additionally zeroes pre_multiplied and
 * overlap of those channels (:185-188).  With shared_ir the call must cover chan0 = 0, nchan = 1.
 * No allocation happens here (the staging buffer is created with the engine). */
int fcb_engine_set_ir(fcb_engine *e, size_t chan0, size_t nchan, const float *irs, size_t len,
                      size_t stride, int is_update);
int fcb_engine_set_ir_dev(fcb_engine *e, size_t chan0, size_t nchan, const float *irs_dev, size_t len,
                          size_t stride, int is_update);

/* Background IR update — update() that never makes a block wait.  A second copy of the IR spectra
 * (the "shadow") is filled by K5 on a low-priority side stream while the blocks keep reading the active copy; commit
 * swaps the two pointers and zeroes pre_multiplied / overlap exactly like update() (:185-188).
 *   reserve  once after create, outside the audio path (the only call here that allocates: S*B*8 bytes per IR channel)
 *   begin    irs: [ir_channels][len], stride in samples; host memory (on_device = 0) or device memory (1).  Returns at
 *            once when irs is page-locked or device memory — it must then stay untouched until ready() answers 1;
 *            pageable memory is staged by the driver during the call.  A second begin before commit overwrites the first.
 *   ready    1 = the shadow copy is complete, 0 = K5 still running, -1 = error; never blocks
 *   commit   swap.  Blocks queued afterwards wait ON THE DEVICE for K5 if it has not finished — the host never does.
 *   wait     host-blocking wait for the copy + K5 (to get a page-locked source buffer back)
 *   join     make `stream` wait on the device for the K5 queued so far */
int fcb_engine_update_reserve(fcb_engine *e);
int fcb_engine_update_reserved(const fcb_engine *e);
int fcb_engine_update_begin(fcb_engine *e, const float *irs, size_t len, size_t stride, int on_device);
int fcb_engine_update_ready(fcb_engine *e);
int fcb_engine_update_commit(fcb_engine *e);
int fcb_engine_update_wait(fcb_engine *e);
int fcb_engine_update_join(fcb_engine *e, void *stream);

/* reset() (:296-306): zero ring, overlap, input buffer, pre_multiplied.  IR spectra kept. */
int fcb_engine_reset(fcb_engine *e);

/* :229-231 — copy n new samples per channel into the device input buffer at [fill, fill+n) */
int fcb_engine_push_input(fcb_engine *e, const float *in, size_t stride, size_t fill, size_t n);
int fcb_engine_push_input_dev(fcb_engine *e, const float *in_dev, size_t stride, size_t fill, size_t n);

/* K1 (:234-241): forward real FFT of [input_buffer[0..valid) | zeros] into ring slot `current` */
int fcb_engine_fft_forward(fcb_engine *e, size_t current, size_t valid);

/* K2 (:244-255): pre_multiplied = sum_{i=1}^{active-1} ir[i] * ring[(current+i) % active],
 * ascending i, every multiply/add rounded separately like the reference.  Only touches ring
 * slots older than `current`, so it may be issued before the block's input arrives. */
int fcb_engine_mac(fcb_engine *e, size_t current, size_t active);

/* fused output epilogues for K3 */
typedef struct {
    /* two-stage head/tail sum (:438-454): out = ((y + overlap) + add0[i]) + add1[i]; device
     * pointers already offset to this call's first sample, channel stride add_stride; may be NULL */
    const float *add0, *add1;
    size_t add_stride;
    /* crossfade gain ramp (src/crossfade_convolver.rs:75-77, :160-169, :242-278):
     * out = mine*g[i].x + other[i]*g[i].y evaluated as two rounded products and one rounded add;
     * g = {1,0} returns mine and g = {0,1} returns other untouched.  `gains` = n device float2. */
    const float *mix_other;
    size_t mix_stride;
    const float *gains;
} fcb_epilogue;

/* K3 (:270-288, :297-298): conv = pre_multiplied + ring[current]*ir[0]; inverse real FFT, /N;
 * out[c][0..n) = y[fill..fill+n) + overlap[fill..fill+n) (+ epilogue); when block_complete the
 * overlap is replaced by y[B..2B).  out_dev: device [C][n] with channel stride out_stride. */
int fcb_engine_ifft_ola(fcb_engine *e, size_t current, size_t fill, size_t n, int block_complete,
                        float *out_dev, size_t out_stride, const fcb_epilogue *epi);

/* a device buffer [C][B] (channel stride B) owned by the engine, for callers that have no device
 * allocator of their own: pass it as out_dev to ifft_ola, then fetch() it */
float *fcb_engine_scratch(fcb_engine *e);
/* the device input buffer [C][B] that push_input fills (src/fft_convolver.rs:100) */
float *fcb_engine_input_buffer(fcb_engine *e);

/* device -> host copy of a planar result (stream-ordered, then synchronised) */
int fcb_engine_fetch(fcb_engine *e, float *out_host, size_t host_stride, const float *src_dev,
                     size_t dev_stride, size_t n);

/* whole-block fast path for full blocks (n == B, fill == 0): K1 -> K2 -> K3 on device buffers */
int fcb_engine_process_block_dev(fcb_engine *e, const float *in_dev, size_t in_stride, float *out_dev,
                                 size_t out_stride, size_t current, size_t active,
                                 const fcb_epilogue *epi);

/* Whole block for TWO engines that have been fed the same input since creation / reset (identical input-spectrum
 * rings, same geometry, same stream): TwoStage's head + tail_convolver0, Crossfade's A + B.  One launch does one forward
 * FFT (written into both rings), streams the ring rows once beside both IR row sets, and runs both inverse FFTs; each
 * engine's output is bit-identical to its own fcb_engine_process_block_dev.  epi_b may mix in out_a of this block.
 * fcb_tune("fused_pair", 0) disables the pairing in the host mirror. */
int fcb_engine_pair_ok(const fcb_engine *ea, const fcb_engine *eb, size_t active);
int fcb_engine_process_block_pair_dev(fcb_engine *ea, fcb_engine *eb, const float *in_dev, size_t in_stride, float *out_a,
                                      size_t stride_a, const fcb_epilogue *epi_a, float *out_b, size_t stride_b,
                                      const fcb_epilogue *epi_b, size_t current, size_t active);
/* the same; the input block is also stored to copy_to ([C][B], channel stride copy_stride, 8-byte aligned rows) — TwoStage's
 * append to tail_input (:459-461) without a copy of its own */
int fcb_engine_process_block_pair_copy_dev(fcb_engine *ea, fcb_engine *eb, const float *in_dev, size_t in_stride, float *out_a,
                                           size_t stride_a, const fcb_epilogue *epi_a, float *out_b, size_t stride_b,
                                           const fcb_epilogue *epi_b, size_t current, size_t active, float *copy_to,
                                           size_t copy_stride);
/* Multi-block calls (offline rendering, large host buffers): nblocks whole blocks of every channel in ONE
 * time-batched pass — K1 for all blocks, a MAC kernel whose threads keep a sliding window of T input spectra in
 * registers (one IR row + one spectrum row loaded per segment feed T output blocks: T blocks for the HBM traffic of
 * one), independent inverse FFTs and a parallel overlap-add.  Same arithmetic and summation order per block as
 * nblocks calls of fcb_engine_process_block_dev: bit-identical output.  in/out are device pointers, or host pointers
 * when host_io != 0; the caller rotates `current` nblocks times afterwards.  No process call allocates: the workspace
 * holds fcb_engine_multi_block_reserved blocks — fcb_engine_create reserves as many as fit 32 MB (when that is at
 * least 2), fcb_engine_multi_block_reserve resizes it outside the audio path, up to fcb_engine_multi_block_capacity —
 * and fcb_engine_process_blocks refuses more (the host mirror then splits the call, block by block if need be).
 * fcb_tune("multi_block", 0) makes the host mirror process block by block. */
int fcb_engine_multi_block_ok(const fcb_engine *e, size_t current, size_t active);
size_t fcb_engine_multi_block_capacity(fcb_engine *e);
size_t fcb_engine_multi_block_reserved(const fcb_engine *e);
int fcb_engine_multi_block_reserve(fcb_engine *e, size_t nblocks);
int fcb_engine_process_blocks(fcb_engine *e, const float *in, size_t in_stride, float *out, size_t out_stride,
                              size_t current, size_t active, size_t nblocks, const fcb_epilogue *epi, int host_io);
/* the same with HOST buffers (pinned for full overlap), pipelined over channel groups of
 * `group_channels` (0 = 512) on internal streams so the PCIe copies overlap K2; synchronous */
int fcb_engine_process_block_host(fcb_engine *e, const float *in, size_t in_stride, float *out,
                                  size_t out_stride, size_t current, size_t active, size_t group_channels);

/* debug / test readback of device state in the reference's layout (K = B+1 interleaved complex) */
int fcb_engine_read_ir_segment(fcb_engine *e, size_t chan, size_t seg, float *out_2k);
int fcb_engine_read_ring_segment(fcb_engine *e, size_t chan, size_t seg, float *out_2k);
int fcb_engine_read_premul(fcb_engine *e, size_t chan, float *out_2k);
int fcb_engine_read_overlap(fcb_engine *e, size_t chan, float *out_b);
/* test hook: overwrite one spectrum row from the reference layout (K interleaved complex) */
int fcb_engine_write_ir_segment(fcb_engine *e, size_t chan, size_t seg, const float *in_2k);
int fcb_engine_write_ring_segment(fcb_engine *e, size_t chan, size_t seg, const float *in_2k);

/* ============================================================================================
 * Layer 2: host mirror of the `Convolution` trait (src/lib.rs:5-14), batched over C channels.
 * `irs` is [C][ir_len] with channel stride ir_len (or a single IR when shared_ir != 0).
 * process(): in/out are HOST pointers, planar, strides in samples; *_dev: device pointers,
 * asynchronous on the convolver's stream.
 * ========================================================================================== */
typedef struct fcb_fftconv fcb_fftconv;
typedef struct fcb_twostage fcb_twostage;
typedef struct fcb_crossfade fcb_crossfade;

typedef struct {
    int device;     /* CUDA device ordinal */
    void *stream;   /* cudaStream_t or NULL */
    int shared_ir;  /* FFTConvolver only */
    int async_tail; /* TwoStage: run the big tail convolver on a second stream (the reference's
                       "might be done in some background thread", src/fft_convolver.rs:478) */
    size_t forced_tail_block; /* TwoStage: 0 = derive like the reference (:520-526) */
    size_t stages;  /* TwoStage: 0 or 2 = the reference's two stages.  N > 2 (EXTENSION): the partition
                       nested N-1 deep — the tail [2T, L) of every level but the last is again a two-stage convolver with
                       head block T and its own T from the same formula (capped at 16384): block sizes B, T1, T2, ...
                       grow geometrically with the distance into the response, Gardner-style */
} fcb_options;

/* ---- FFTConvolver (src/fft_convolver.rs:86-307) ---- */
int fcb_fftconv_init(fcb_fftconv **out, const float *irs, size_t channels, size_t ir_len,
                     size_t block_size, size_t max_response_length, const fcb_options *opt);
int fcb_fftconv_default(fcb_fftconv **out, size_t channels, const fcb_options *opt); /* Default::default() */
int fcb_fftconv_clone(const fcb_fftconv *c, fcb_fftconv **out);
void fcb_fftconv_free(fcb_fftconv *c);
int fcb_fftconv_update(fcb_fftconv *c, const float *irs, size_t ir_len);
/* update() for real-time callers (fcb_engine_update_*): reserve once after init; begin queues the copy + K5 in the
 * background and returns (irs page-locked: untouched until pending() answers 0; pageable: staged during the call);
 * the new response is swapped in between two process calls — with FCB_UPDATE_WAIT at the very next one (the device
 * waits for K5 if need be: same output as fcb_fftconv_update), otherwise at the first one that finds K5 finished
 * (no block ever waits; the old response plays until then). */
#define FCB_UPDATE_WAIT 1
int fcb_fftconv_update_reserve(fcb_fftconv *c);
int fcb_fftconv_update_begin(fcb_fftconv *c, const float *irs, size_t ir_len, int flags);
int fcb_fftconv_update_pending(const fcb_fftconv *c);
/* calls of several whole blocks run as one time-batched pass only within the reserved workspace (process never
 * allocates): reserve it for calls of up to max_call_samples samples, outside the audio path */
int fcb_fftconv_reserve(fcb_fftconv *c, size_t max_call_samples);
int fcb_fftconv_reset(fcb_fftconv *c);
int fcb_fftconv_process(fcb_fftconv *c, const float *in, size_t in_len, size_t in_stride, float *out,
                        size_t out_len, size_t out_stride);
int fcb_fftconv_process_dev(fcb_fftconv *c, const float *in_dev, size_t in_len, size_t in_stride,
                            float *out_dev, size_t out_len, size_t out_stride, const fcb_epilogue *epi);
int fcb_fftconv_sync(fcb_fftconv *c);
fcb_engine *fcb_fftconv_engine(fcb_fftconv *c); /* NULL for a default (empty) convolver */
/* scheduler scalars the host keeps (src/fft_convolver.rs:89-91, :99, :101) */
size_t fcb_fftconv_block_size(const fcb_fftconv *c);
size_t fcb_fftconv_seg_count(const fcb_fftconv *c);
size_t fcb_fftconv_active_seg_count(const fcb_fftconv *c);
size_t fcb_fftconv_current(const fcb_fftconv *c);
size_t fcb_fftconv_fill(const fcb_fftconv *c);

/* ---- TwoStageFFTConvolver (src/fft_convolver.rs:323-526) ---- */
size_t fcb_compute_tail_block_size(size_t head_len, size_t response_len); /* :520-526, f32 */
int fcb_twostage_init(fcb_twostage **out, const float *irs, size_t channels, size_t ir_len,
                      size_t block_size, size_t max_response_length, const fcb_options *opt);
int fcb_twostage_clone(const fcb_twostage *c, fcb_twostage **out);
void fcb_twostage_free(fcb_twostage *c);
/* todo!() in the reference (:408-410).  EXTENSION, default: the per-stage FFTConvolver::update (:174-213) on the
 * response zero-padded to max_response_length and re-sliced like init (:349-384); rings, partial blocks and tail
 * outputs already computed are kept, each stage's overlap and pre_multiplied are zeroed; panics (FCB_ERR_PANIC) when
 * ir_len > max_response_length; allocation-free.  With fcb_tune("strict_todo", 1): FCB_ERR_TODO. */
int fcb_twostage_update(fcb_twostage *c, const float *irs, size_t ir_len);
int fcb_twostage_reset(fcb_twostage *c);
int fcb_twostage_process(fcb_twostage *c, const float *in, size_t in_len, size_t in_stride, float *out,
                         size_t out_len, size_t out_stride);
int fcb_twostage_process_dev(fcb_twostage *c, const float *in_dev, size_t in_len, size_t in_stride,
                             float *out_dev, size_t out_len, size_t out_stride);
int fcb_twostage_sync(fcb_twostage *c);
size_t fcb_twostage_tail_block_size(const fcb_twostage *c);
/* block sizes of the (nested) partition, outermost first: head, T1, T2, ...; returns how many there are */
size_t fcb_twostage_stage_blocks(const fcb_twostage *c, size_t *out, size_t cap);

/* ---- CrossfadeConvolver<FFTConvolver> (src/crossfade_convolver.rs:3-105) ---- */
/* CrossfadeConvolver::new (:20-42): takes ownership of `convolver` */
int fcb_crossfade_new(fcb_crossfade **out, fcb_fftconv *convolver, size_t max_response_length,
                      size_t max_buffer_size, size_t crossfade_samples);
/* <CrossfadeConvolver as Convolution>::init (:46-49) */
int fcb_crossfade_init(fcb_crossfade **out, const float *irs, size_t channels, size_t ir_len,
                       size_t max_block_size, size_t max_response_length, const fcb_options *opt);
void fcb_crossfade_free(fcb_crossfade *c);
int fcb_crossfade_update(fcb_crossfade *c, const float *irs, size_t ir_len);
int fcb_crossfade_process(fcb_crossfade *c, const float *in, size_t in_len, size_t in_stride, float *out,
                          size_t out_len, size_t out_stride);
int fcb_crossfade_process_dev(fcb_crossfade *c, const float *in_dev, size_t in_len, size_t in_stride,
                              float *out_dev, size_t out_len, size_t out_stride);
/* todo!() in the reference (src/crossfade_convolver.rs:80-82).  EXTENSION, default: both convolvers reset
 * (src/fft_convolver.rs:296-306), buffer_a / buffer_b zeroed, a running fade finished at once (the crossfader lands
 * where `mix` leaves it when counter == fading_samples, :261-273), a pending response stays pending.
 * With fcb_tune("strict_todo", 1): FCB_ERR_TODO. */
int fcb_crossfade_reset(fcb_crossfade *c);
/* #[derive(Clone)] (src/crossfade_convolver.rs:10) */
int fcb_crossfade_clone(const fcb_crossfade *c, fcb_crossfade **out);
/* update() that never waits for a copy or a kernel: irs must be page-locked (else this is fcb_crossfade_update) and
 * stay untouched until fcb_crossfade_update_pending answers 0.  State changes and output are those of update(). */
int fcb_crossfade_update_begin(fcb_crossfade *c, const float *irs, size_t ir_len);
int fcb_crossfade_update_pending(fcb_crossfade *c);
int fcb_crossfade_is_crossfading(const fcb_crossfade *c);
int fcb_crossfade_sync(fcb_crossfade *c);
/* crossfader state for tests: counter, mix_value, approaching(0/1), target(0=A,1=B) */
int fcb_crossfade_state(const fcb_crossfade *c, int64_t *counter, float *mix_value, int *approaching,
                        int *target);

/* ============================================================================================
 * Convolution matrix (BASELINE configs[4]): y_out = sum_in FFTConvolver(h[out][in]).process(x_in),
 * OUT x IN reference convolvers sharing the IN input rings; NS independent streams may share the
 * one IR matrix; an object may be one IR-partition shard (a contiguous range of IR segments) of a
 * multi-GPU job, in which case the caller all-reduces (sum, f32) the partial spectra returned by
 * fcb_mimo_conv_buffer between partial and finish.  Full blocks only (B samples per call).
 * ========================================================================================== */
typedef struct fcb_mimo fcb_mimo;
typedef struct {
    size_t n_in, n_out, n_streams; /* n_streams 0 = 1 */
    size_t block_size, max_response_length;
    size_t shard_index, shard_count; /* shard g of G owns IR segments [S*g/G, S*(g+1)/G); 0,0 = all */
    int device;
    void *stream;
} fcb_mimo_desc;
int fcb_mimo_create(const fcb_mimo_desc *desc, fcb_mimo **out);
void fcb_mimo_destroy(fcb_mimo *m);
/* irs: host [OUT][IN][len]; every shard is given the whole IRs and keeps its segment rows */
int fcb_mimo_set_ir(fcb_mimo *m, const float *irs, size_t len);
int fcb_mimo_reset(fcb_mimo *m);
/* in_dev: device [NS*IN][B] (channel stride in_stride).  K1 + K2 over this shard's segments +
 * sum over inputs -> partial conv spectra in fcb_mimo_conv_buffer ([NS*OUT][B] packed complex) */
int fcb_mimo_partial_dev(fcb_mimo *m, const float *in_dev, size_t in_stride);
float *fcb_mimo_conv_buffer(fcb_mimo *m, size_t *n_floats);
/* K3 on the (all-reduced) conv buffer -> out_dev [NS*OUT][B]; advances the ring */
int fcb_mimo_finish_dev(fcb_mimo *m, float *out_dev, size_t out_stride);
/* single-shard convenience, host buffers: in [NS*IN][B] -> out [NS*OUT][B], synchronous */
int fcb_mimo_process(fcb_mimo *m, const float *in, float *out);
int fcb_mimo_sync(fcb_mimo *m);
void *fcb_mimo_stream(fcb_mimo *m);
/* Peer exchange for IR-partition shards on one NVLink node — replaces the all-reduce between partial and finish.
 * Once attached, fcb_mimo_partial_dev's reduce kernel stores this shard's partial spectra straight into every
 * shard's inbox (peer stores over NVLink) and raises a release flag; fcb_mimo_finish_dev's K3 acquires the G flags
 * and sums the G inbox slots in shard order (bit-identical on every shard).  Set-up: every shard exports its inbox
 * handle (cudaIpcMemHandle_t, FCB_PEER_HANDLE_BYTES bytes), the caller gathers the G handles in shard order and hands
 * them to every shard.  All shards must step in lock-step (same number of partial/finish calls). */
#define FCB_PEER_HANDLE_BYTES 64
#define FCB_MAX_PEERS 16
int fcb_mimo_peer_export(fcb_mimo *m, unsigned char *handle_out);
int fcb_mimo_peer_attach(fcb_mimo *m, const unsigned char *handles /* [shard_count][FCB_PEER_HANDLE_BYTES] */);
/* Failure behaviour: K3 waits a bounded time for the G flags.  If one never arrives it writes SILENCE for that block
 * (never a sum over a stale inbox), leaves the overlap alone and raises an error word in mapped host memory; every
 * later fcb_mimo_partial_dev / fcb_mimo_finish_dev / fcb_mimo_sync call returns FCB_ERR_CUDA.
 * same-process variant: this shard's inbox as a device pointer / attach with the G inbox pointers.  Shards that share
 * ONE GPU must all enqueue fcb_mimo_partial_dev before any of them enqueues fcb_mimo_finish_dev: a K3 spinning on a
 * flag can otherwise hold the SMs its peer's producer kernel is waiting for. */
void *fcb_mimo_peer_inbox(fcb_mimo *m);
/* Reduce-scatter form of the exchange (large payloads: many streams).  Shard g FINISHES rows [R*g/G, R*(g+1)/G) of the
 * R = NS*OUT output rows only (fcb_mimo_owned_rows), so a partial row travels to its owner alone — 1/G of the all-gather
 * form's NVLink bytes — and fcb_mimo_finish_dev writes just those rows of out_dev (the full-size buffer's other rows are
 * left alone: the output is sharded by row over the GPUs).  Set on EVERY shard before the first block.
 * NCCL callers get the same with a reduce_scatter of fcb_mimo_conv_buffer followed by fcb_mimo_finish_rows_dev. */
int fcb_mimo_peer_set_scatter(fcb_mimo *m, int on);
int fcb_mimo_owned_rows(const fcb_mimo *m, size_t *lo, size_t *hi);
/* Overlapped finish (peer exchange, one shard per GPU; throughput of back-to-back blocks).  K3 is the kernel that waits
 * for the peers' partial spectra; with this on it runs on a stream of its own beside K1 and the MAC of the NEXT block, and
 * only that block's reduce waits for it (the order the exchange protocol needs: a shard's reduce comes after its own
 * previous K3).  The output of a block is then complete, in the order of the convolver's stream, after fcb_mimo_join
 * (a device-side wait) and for the host after fcb_mimo_sync; without a join, work queued on the convolver's stream is NOT
 * ordered after the block's output.  Set on EVERY shard before the first block.  Shards sharing one GPU keep the rule above
 * (every partial, synchronised, before any finish). */
int fcb_mimo_set_overlap(fcb_mimo *m, int on);
int fcb_mimo_join(fcb_mimo *m);
int fcb_mimo_finish_rows_dev(fcb_mimo *m, float *out_dev, size_t out_stride, size_t row_lo, size_t row_hi);
int fcb_mimo_peer_attach_ptrs(fcb_mimo *m, void *const *inboxes);
/* test hook (host only): segment chunks (count, segments per chunk) the CUDA-core matrix kernel uses for a problem */
int fcb_debug_mac_tile_plan(int logb, int n_in, int n_out, int n_streams, int nsegs, int *zchunks, int *zlen);
/* test hook (host only): K4's pipeline stages for one input — ring slot block, IR copy (0 / 1 = shifted by one
 * position) and IR position paired with the block's first slot; returns the stage count */
int fcb_debug_tc_stages(int S, int current, int seg_lo, int seg_hi, int max_stages, int *blk, int *copy, int *pos0);
/* 1 when this object's delay-line MAC runs as per-bin complex GEMMs on the tensor cores (K4), 2 when it runs as the
 * register-tiled per-bin GEMM on the FP32 pipes (k_mac_rt: stream counts below the tensor-core threshold, blocks of 32+), else 0 */
int fcb_mimo_uses_tensor_cores(const fcb_mimo *m);
size_t fcb_mimo_block_size(const fcb_mimo *m);
size_t fcb_mimo_seg_count(const fcb_mimo *m);
int fcb_mimo_segment_range(const fcb_mimo *m, size_t *lo, size_t *hi);

#ifdef __cplusplus
}
#endif
#endif /* FFTCONV_B200_H */
