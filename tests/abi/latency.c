/*
 * latency.c — per-call latency of the C ABI as a C host sees it (no Python in the timed path), for the BASELINE
 * configs that are small batches: configs[0] FFTConvolver mono (block 256, 48 000 taps), configs[1] TwoStage x 64
 * (head 128, 240 000 taps), configs[2] CrossfadeConvolver::init x 256 (block 512, 96 000 taps).
 * One host-pointer process() call per block (copy in, kernels, copy out, synchronised), wall clock around each call.
 *   latency <config 0|1|2> <blocks> [pinned] [tune_key=value ...]     pinned: the caller's buffers come from fcb_host_alloc
 * prints one JSON line.  Test / bench infrastructure: links libfftconv_b200.so only.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fftconv_b200.h"

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != FCB_OK) {                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fcb_last_error()); \
            exit(10 + rc_);                                                  \
        }                                                                    \
    } while (0)

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e6 * (double)ts.tv_sec + 1e-3 * (double)ts.tv_nsec;
}
static int cmp(const void *a, const void *b) { return (*(const double *)a > *(const double *)b) - (*(const double *)a < *(const double *)b); }

static unsigned long long rng = 0x9E3779B97F4A7C15ull;
static float frand(void)
{
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((rng >> 40) & 0xFFFFFF) / 8388608.0f - 1.0f;
}

int main(int argc, char **argv)
{
    const int cfg = argc > 1 ? atoi(argv[1]) : 0;
    const int blocks = argc > 2 ? atoi(argv[2]) : 2000;
    for (int i = 3; i < argc; i++) {
        char *eq = strchr(argv[i], '=');
        if (eq) {
            *eq = 0;
            CHECK(fcb_tune(argv[i], atoi(eq + 1)));
            *eq = '=';
        }
    }
    int pinned = 0;
    for (int i = 3; i < argc; i++) pinned |= !strcmp(argv[i], "pinned");
    const size_t C = cfg == 0 ? 1 : cfg == 1 ? 64 : 256, B = cfg == 0 ? 256 : cfg == 1 ? 128 : 512;
    const size_t L = cfg == 0 ? 48000 : cfg == 1 ? 240000 : 96000;
    float *irs = (float *)malloc(C * L * sizeof(float));
    float *in = pinned ? (float *)fcb_host_alloc(C * B * sizeof(float)) : (float *)malloc(C * B * sizeof(float));
    float *out = pinned ? (float *)fcb_host_alloc(C * B * sizeof(float)) : (float *)malloc(C * B * sizeof(float));
    for (size_t i = 0; i < C * L; i++) irs[i] = frand() * 0.01f;
    for (size_t i = 0; i < C * B; i++) in[i] = frand();
    fcb_fftconv *u = NULL;
    fcb_twostage *t = NULL;
    fcb_crossfade *x = NULL;
    fcb_options opt;
    memset(&opt, 0, sizeof opt);
    opt.async_tail = 1;
    if (cfg == 0) CHECK(fcb_fftconv_init(&u, irs, C, L, B, L, &opt));
    if (cfg == 1) CHECK(fcb_twostage_init(&t, irs, C, L, B, L, &opt));
    if (cfg == 2) CHECK(fcb_crossfade_init(&x, irs, C, L, B, L, &opt));
    double *us = (double *)malloc(sizeof(double) * (size_t)blocks);
    const int warm = 200;
    for (int i = 0; i < warm + blocks; i++) {
        if (cfg == 2 && i % 50 == 49) CHECK(fcb_crossfade_update(x, irs, L)); /* outside the timed call */
        const double t0 = now_us();
        if (cfg == 0) CHECK(fcb_fftconv_process(u, in, B, B, out, B, B));
        if (cfg == 1) CHECK(fcb_twostage_process(t, in, B, B, out, B, B));
        if (cfg == 2) CHECK(fcb_crossfade_process(x, in, B, B, out, B, B));
        if (i >= warm) us[i - warm] = now_us() - t0;
    }
    double sum = 0;
    for (int i = 0; i < blocks; i++) sum += us[i];
    qsort(us, (size_t)blocks, sizeof(double), cmp);
    printf("{\"config\": %d, \"channels\": %zu, \"block\": %zu, \"ir_taps\": %zu, \"blocks\": %d, \"us_mean\": %.2f, \"us_p50\": %.2f, "
           "\"us_p99\": %.2f, \"us_max\": %.2f, \"block_period_us\": %.1f, \"caller_buffers\": \"%s\", \"through\": \"C ABI layer 2, host pointers, one synchronous call per block\"}\n",
           cfg, C, B, L, blocks, sum / blocks, us[blocks / 2], us[(int)(blocks * 0.99)], us[blocks - 1], 1e6 * (double)B / 48000.0,
           pinned ? "page-locked (fcb_host_alloc)" : "pageable (malloc)");
    if (u) fcb_fftconv_free(u);
    if (t) fcb_twostage_free(t);
    if (x) fcb_crossfade_free(x);
    return 0;
}
