/*
 * replay.c — a C host that drives LAYER 1 of the C ABI (include/fftconv_b200.h) with exactly the call
 * sequence of the Rust binding's `impl Convolution for CudaFFTConvolver`
 * (rust/fft_convolution_b200_sys/src/lib.rs), i.e. of the reference's FFTConvolver
 * (/root/reference/src/fft_convolver.rs:105-307) with the arithmetic replaced by the stage calls:
 *
 *   init    -> fcb_engine_create + fcb_engine_set_ir(is_update = 0)                      (:105-172)
 *   update  -> active = ceil(len / B); fcb_engine_set_ir(is_update = 1)                  (:174-213)
 *   reset   -> fcb_engine_reset; current = fill = 0                                       (:296-306)
 *   process -> per chunk: push_input, fft_forward (K1), mac (K2, only when the block was empty),
 *              ifft_ola (K3), fetch; `current` / `input_buffer_fill` kept HERE, on the host (:215-295);
 *              a call spanning >= 2 whole blocks takes the fcb_engine_process_blocks branch
 *
 * It is test infrastructure: no Rust toolchain exists in the build image, so this is how the layer-1
 * sequence (fcb_engine_fetch included) is exercised by something that is not the C++ host mirror.
 *
 *   replay <dir> <block> <max_len> <update_before_call | -1> <reset_before_call | -1> <size> [<size> ...]
 * reads <dir>/h0.f32 (and h1.f32 when an update is asked for) and <dir>/x.f32 (raw little-endian f32),
 * feeds x in calls of the given sizes (cyclic) and writes <dir>/y.f32.  Exit code 0 = ok.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fftconv_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != FCB_OK) {                                                         \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fcb_last_error());         \
            exit(10 + rc_);                                                          \
        }                                                                            \
    } while (0)

static float *read_f32(const char *dir, const char *name, size_t *n)
{
    char path[4096];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    float *v = (float *)malloc(bytes > 0 ? (size_t)bytes : 4);
    if (fread(v, 1, (size_t)bytes, f) != (size_t)bytes) { perror(path); exit(2); }
    fclose(f);
    *n = (size_t)bytes / sizeof(float);
    return v;
}

/* the host-side state of the Rust struct CudaFFTConvolver */
typedef struct {
    fcb_engine *engine;
    size_t ir_len, block_size, seg_count, active_seg_count, current, input_buffer_fill;
} conv_t;

static void conv_init(conv_t *c, const float *ir, size_t n_ir, size_t block, size_t max_len)
{
    if (max_len < n_ir) { fprintf(stderr, "panic: max_response_length < impulse response\n"); exit(3); }
    fcb_engine_desc d;
    memset(&d, 0, sizeof d);
    d.channels = 1;
    d.block_size = block;
    d.max_response_length = max_len;
    CHECK(fcb_engine_create(&d, &c->engine));
    CHECK(fcb_engine_set_ir(c->engine, 0, 1, ir, n_ir, n_ir, 0));
    c->ir_len = max_len;
    c->block_size = fcb_engine_block_size(c->engine);
    c->seg_count = fcb_engine_seg_count(c->engine);
    c->active_seg_count = c->seg_count;
    c->current = 0;
    c->input_buffer_fill = 0;
}

static void conv_update(conv_t *c, const float *ir, size_t n_ir)
{
    if (n_ir > c->ir_len) { fprintf(stderr, "panic: new impulse response too long\n"); exit(3); }
    if (c->ir_len == 0) return;
    c->active_seg_count = (size_t)ceil((double)n_ir / (double)c->block_size);
    CHECK(fcb_engine_set_ir(c->engine, 0, 1, ir, n_ir, n_ir, 1));
}

static void conv_reset(conv_t *c)
{
    CHECK(fcb_engine_reset(c->engine));
    c->current = 0;
    c->input_buffer_fill = 0;
}

static void rotate(conv_t *c) { c->current = c->current > 0 ? c->current - 1 : c->active_seg_count - 1; }

static void conv_process(conv_t *c, const float *input, size_t in_len, float *output, size_t out_len)
{
    if (c->active_seg_count == 0) { memset(output, 0, out_len * sizeof(float)); return; }
    if (in_len < out_len) { fprintf(stderr, "panic: input shorter than output\n"); exit(3); }
    float *scratch = fcb_engine_scratch(c->engine); /* device [1][B] */
    const size_t B = c->block_size;
    size_t processed = 0;
    while (processed < out_len) {
        const int was_empty = c->input_buffer_fill == 0;
        const size_t whole = (out_len - processed) / B;
        if (was_empty && whole >= 2 && fcb_engine_multi_block_ok(c->engine, c->current, c->active_seg_count)) {
            size_t nb = whole, cap = fcb_engine_multi_block_reserved(c->engine); /* process never allocates */
            if (nb > cap) nb = cap;
            if (nb >= 2) {
                const size_t n = nb * B;
                CHECK(fcb_engine_process_blocks(c->engine, input + processed, n, output + processed, n, c->current,
                                                c->active_seg_count, nb, NULL, 1));
                for (size_t d = 0; d < nb; d++) rotate(c);
                processed += n;
                continue;
            }
        }
        size_t n = out_len - processed;
        if (B - c->input_buffer_fill < n) n = B - c->input_buffer_fill;
        const size_t pos = c->input_buffer_fill;
        CHECK(fcb_engine_push_input(c->engine, input + processed, n, pos, n));
        CHECK(fcb_engine_fft_forward(c->engine, c->current, pos + n));                         /* K1 */
        if (was_empty) CHECK(fcb_engine_mac(c->engine, c->current, c->active_seg_count));      /* K2 */
        CHECK(fcb_engine_ifft_ola(c->engine, c->current, pos, n, pos + n == B, scratch, B, NULL)); /* K3 */
        CHECK(fcb_engine_fetch(c->engine, output + processed, n, scratch, B, n));
        c->input_buffer_fill += n;
        if (c->input_buffer_fill == B) {
            c->input_buffer_fill = 0;
            rotate(c);
        }
        processed += n;
    }
}

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: %s <dir> <block> <max_len> <update_before_call|-1> <reset_before_call|-1> <size>...\n", argv[0]);
        return 1;
    }
    const char *dir = argv[1];
    const size_t block = (size_t)atol(argv[2]), max_len = (size_t)atol(argv[3]);
    const long update_at = atol(argv[4]), reset_at = atol(argv[5]);
    const int nsizes = argc - 6;
    size_t n_h0 = 0, n_h1 = 0, n_x = 0;
    float *h0 = read_f32(dir, "h0.f32", &n_h0), *h1 = NULL, *x = read_f32(dir, "x.f32", &n_x);
    if (update_at >= 0) h1 = read_f32(dir, "h1.f32", &n_h1);
    float *y = (float *)calloc(n_x ? n_x : 1, sizeof(float));

    conv_t c;
    conv_init(&c, h0, n_h0, block, max_len);
    CHECK(fcb_engine_multi_block_reserve(c.engine, 3)); /* CudaFFTConvolver::reserve_blocks(3): a 4-block call runs as 3 + 1 */
    size_t p = 0;
    long call = 0;
    while (p < n_x) {
        if (call == update_at) conv_update(&c, h1, n_h1);
        if (call == reset_at) conv_reset(&c);
        size_t n = (size_t)atol(argv[6 + call % nsizes]);
        if (n > n_x - p) n = n_x - p;
        conv_process(&c, x + p, n, y + p, n);
        p += n;
        call++;
    }
    fcb_engine_destroy(c.engine);

    char path[4096];
    snprintf(path, sizeof path, "%s/y.f32", dir);
    FILE *f = fopen(path, "wb");
    if (!f || fwrite(y, sizeof(float), n_x, f) != n_x) { perror(path); return 2; }
    fclose(f);
    printf("replay ok: %ld calls, %zu samples, launches %llu\n", call, n_x, (unsigned long long)fcb_launch_count());
    return 0;
}
