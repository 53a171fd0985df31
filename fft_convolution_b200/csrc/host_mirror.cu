// host_mirror.cu — layer 2 of the C ABI: a C++ mirror of the reference's three `Convolution`
// implementors (src/lib.rs:5-14), batched over C lock-step channels.  It keeps exactly what
// north_star leaves on the host — the block scheduler, the input-buffer fill, the segment-ring
// rotation (src/fft_convolver.rs:222-231, :277-292), the two-stage bookkeeping (:424-494) and the
// crossfade state machine (src/crossfade_convolver.rs:51-105, 192-279) — and drives the device
// stages of engine.cu.  All arithmetic on samples happens in CUDA kernels; there is no CPU path.
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "common.cuh"

using namespace fcb;

// ---- small elementwise kernels for the unfused fall-backs of the K3 epilogues ----------------
namespace fcb {

// two-stage head/tail sum when a call does not map onto one K3 launch (src/fft_convolver.rs:438-454)
__global__ void k_add2(float *out, long long out_stride, const float *p0, const float *p1, long long p_stride, int n,
                       long long nchan)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nchan * n) return;
    long long c = idx / n;
    int i = (int)(idx % n);
    float v = out[c * out_stride + i];
    v = __fadd_rn(v, p0[c * p_stride + i]);
    v = __fadd_rn(v, p1[c * p_stride + i]);
    out[c * out_stride + i] = v;
}

// crossfade mix when it cannot ride on K3 (src/crossfade_convolver.rs:75-77): gains[i] = {gA, gB}
__global__ void k_mix(float *out, long long out_stride, const float *a, const float *b, long long ab_stride,
                      const float2 *gains, int n, long long nchan)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nchan * n) return;
    long long c = idx / n;
    int i = (int)(idx % n);
    float2 g = gains[i];
    float va = a[c * ab_stride + i], vb = b[c * ab_stride + i], v;
    if (g.x == 1.f && g.y == 0.f) v = va;
    else if (g.x == 0.f && g.y == 1.f) v = vb;
    else v = __fadd_rn(__fmul_rn(va, g.x), __fmul_rn(vb, g.y));
    out[c * out_stride + i] = v;
}

static int zero_planar(float *dst, size_t stride, size_t n, size_t C, cudaStream_t s)
{
    if (n == 0 || C == 0) return FCB_OK;
    FCB_CUDA(cudaMemset2DAsync(dst, stride * sizeof(float), 0, n * sizeof(float), C, s));
    return FCB_OK;
}

static fcb_options default_options()
{
    fcb_options o;
    memset(&o, 0, sizeof o);
    return o;
}

} // namespace fcb

// ==============================================================================================
// FFTConvolver — src/fft_convolver.rs:86-307
// ==============================================================================================
struct fcb_fftconv {
    fcb_engine *eng = nullptr; // NULL = Default::default() (no segments)
    size_t C = 0;
    size_t ir_len = 0, block_size = 0, seg_count = 0, active_seg_count = 0; // :88-91
    size_t current = 0, input_buffer_fill = 0;                              // :99, :101
    fcb_options opt{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    float *d_io = nullptr; // device staging for host-pointer process(): [C][B] (the engine's scratch)
    // low-latency path for small batches: pinned host buffers mapped into the device address space
    // ([C][B] each); the kernels read the input and write the output over PCIe themselves, so a
    // single-chunk call is: CPU copy in, one launch (whole block) or three, one sync, CPU copy out
    float *h_in = nullptr, *h_out = nullptr;   // host views
    float *m_in = nullptr, *m_out = nullptr;   // the same memory as seen from the device
    // fcb_fftconv_update_begin: the new spectra are being built in the engine's shadow buffer
    bool upd_pending = false, upd_wait = false;
    size_t upd_active = 0; // active_seg_count once the update is committed
};

static int fftconv_commit_update(fcb_fftconv *c);
static int fftconv_commit_wait(fcb_fftconv *c);

static void fftconv_alloc_mapped(fcb_fftconv *c)
{
    const size_t bytes = c->C * c->block_size * sizeof(float);
    if (!c->eng || bytes == 0 || bytes > ((size_t)256 << 10)) return; // small batches only
    if (cudaHostAlloc(&c->h_in, bytes, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostAlloc(&c->h_out, bytes, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&c->m_in, c->h_in, 0) != cudaSuccess ||
        cudaHostGetDevicePointer(&c->m_out, c->h_out, 0) != cudaSuccess) {
        cudaGetLastError();
        if (c->h_in) cudaFreeHost(c->h_in);
        if (c->h_out) cudaFreeHost(c->h_out);
        c->h_in = c->h_out = c->m_in = c->m_out = nullptr;
    }
}

static int fftconv_make_stream(fcb_fftconv *c)
{
    FCB_CUDA(cudaSetDevice(c->opt.device));
    if (c->opt.stream) {
        c->stream = (cudaStream_t)c->opt.stream;
    } else {
        FCB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return FCB_OK;
}

extern "C" int fcb_fftconv_default(fcb_fftconv **out, size_t channels, const fcb_options *opt)
{
    if (!out) return fail(FCB_ERR_ARG, "NULL out");
    fcb_fftconv *c = new fcb_fftconv();
    c->C = channels;
    c->opt = opt ? *opt : default_options();
    int rc = fftconv_make_stream(c);
    if (rc) {
        delete c;
        return rc;
    }
    *out = c;
    return FCB_OK;
}

extern "C" void fcb_fftconv_free(fcb_fftconv *c)
{
    if (!c) return;
    cudaSetDevice(c->opt.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    fcb_engine_destroy(c->eng);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

// :105-172
extern "C" int fcb_fftconv_init(fcb_fftconv **out, const float *irs, size_t channels, size_t ir_len,
                                size_t block_size, size_t max_response_length, const fcb_options *opt)
{
    if (!out) return fail(FCB_ERR_ARG, "NULL out");
    *out = nullptr;
    if (channels == 0) return fail(FCB_ERR_ARG, "channels must be >= 1");
    if (max_response_length < ir_len) // :106-110
        return fail(FCB_ERR_PANIC, "max_response_length must be at least the length of the initial impulse response");
    fcb_fftconv *c = nullptr;
    FCB_TRY(fcb_fftconv_default(&c, channels, opt));
    c->ir_len = max_response_length;                 // :111-113
    c->block_size = next_power_of_two(block_size);   // :115
    c->seg_count = (size_t)std::ceil((double)c->ir_len / (double)c->block_size); // :117
    c->active_seg_count = c->seg_count;              // :118
    fcb_engine_desc d{channels, c->block_size, c->ir_len, c->opt.shared_ir, c->opt.device, (void *)c->stream};
    int rc = fcb_engine_create(&d, &c->eng);
    // :131-142 — K5 over the zero-padded IR (rows past ir_len come out as zeros)
    if (rc == FCB_OK) rc = fcb_engine_set_ir(c->eng, 0, c->opt.shared_ir ? 1 : channels, irs, ir_len, ir_len, 0);
    if (rc == FCB_OK) c->d_io = fcb_engine_scratch(c->eng);
    if (rc == FCB_OK) fftconv_alloc_mapped(c);
    if (rc == FCB_OK) rc = fcb_engine_sync(c->eng);
    if (rc != FCB_OK) {
        fcb_fftconv_free(c);
        return rc;
    }
    *out = c;
    return FCB_OK;
}

// #[derive(Clone)] :86
extern "C" int fcb_fftconv_clone(const fcb_fftconv *s, fcb_fftconv **out)
{
    if (!s || !out) return fail(FCB_ERR_ARG, "NULL argument");
    fcb_fftconv *c = nullptr;
    FCB_TRY(fcb_fftconv_default(&c, s->C, &s->opt));
    c->ir_len = s->ir_len;
    c->block_size = s->block_size;
    c->seg_count = s->seg_count;
    c->active_seg_count = s->active_seg_count;
    c->current = s->current;
    c->input_buffer_fill = s->input_buffer_fill;
    int rc = FCB_OK;
    if (s->upd_pending) { // #[derive(Clone)] copies a value whose update() has returned: finish the background one first
        rc = fftconv_commit_wait(const_cast<fcb_fftconv *>(s));
        c->active_seg_count = s->active_seg_count;
    }
    if (rc == FCB_OK && s->eng) {
        rc = fcb_engine_clone(s->eng, &c->eng);
        if (rc == FCB_OK) rc = fcb_engine_set_stream(c->eng, (void *)c->stream);
        if (rc == FCB_OK) c->d_io = fcb_engine_scratch(c->eng);
        if (rc == FCB_OK) fftconv_alloc_mapped(c);
    }
    if (rc != FCB_OK) {
        fcb_fftconv_free(c);
        return rc;
    }
    *out = c;
    return FCB_OK;
}

// :174-213
extern "C" int fcb_fftconv_update(fcb_fftconv *c, const float *irs, size_t new_ir_len)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    if (new_ir_len > c->ir_len) return fail(FCB_ERR_PANIC, "New impulse response is longer than initialized length");
    if (c->ir_len == 0) return FCB_OK; // :181-183
    c->upd_pending = false; // supersedes an update still in the background
    c->active_seg_count = (size_t)std::ceil((double)new_ir_len / (double)c->block_size); // :190
    return fcb_engine_set_ir(c->eng, 0, c->opt.shared_ir ? 1 : c->C, irs, new_ir_len, new_ir_len, 1);
}

// the same from device memory (irs_dev: [C][stride]); used by the crossfade's deferred swap
static int fftconv_update_dev(fcb_fftconv *c, const float *irs_dev, size_t new_ir_len, size_t stride)
{
    if (new_ir_len > c->ir_len) return fail(FCB_ERR_PANIC, "New impulse response is longer than initialized length");
    if (c->ir_len == 0) return FCB_OK;
    c->upd_pending = false;
    c->active_seg_count = (size_t)std::ceil((double)new_ir_len / (double)c->block_size);
    return fcb_engine_set_ir_dev(c->eng, 0, c->opt.shared_ir ? 1 : c->C, irs_dev, new_ir_len, stride, 1);
}

// :296-306
extern "C" int fcb_fftconv_reset(fcb_fftconv *c)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    c->current = 0;
    c->input_buffer_fill = 0;
    return c->eng ? fcb_engine_reset(c->eng) : FCB_OK;
}

// update() with the new response zero-padded to `full_len` samples: `valid` samples are read from `irs`, the segment
// count follows full_len (used by the two-stage update, whose stages always keep their whole slice active)
static int fftconv_update_padded(fcb_fftconv *c, const float *irs, size_t valid, size_t stride, size_t full_len)
{
    if (full_len > c->ir_len) return fail(FCB_ERR_PANIC, "New impulse response is longer than initialized length");
    if (c->ir_len == 0) return FCB_OK;
    c->upd_pending = false; // a synchronous update supersedes one still in the background
    c->active_seg_count = (size_t)std::ceil((double)full_len / (double)c->block_size);
    return fcb_engine_set_ir(c->eng, 0, c->opt.shared_ir ? 1 : c->C, irs, valid, stride, 1);
}

// ---- real-time update: returns at once, the blocks keep running, the new response is swapped in between two calls ----
extern "C" int fcb_fftconv_update_reserve(fcb_fftconv *c)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    return c->eng ? fcb_engine_update_reserve(c->eng) : FCB_OK;
}

static int fftconv_update_begin(fcb_fftconv *c, const float *irs, size_t valid, size_t stride, size_t full_len, bool on_device,
                                bool wait)
{
    if (full_len > c->ir_len) return fail(FCB_ERR_PANIC, "New impulse response is longer than initialized length");
    if (c->ir_len == 0) return FCB_OK;
    FCB_TRY(fcb_engine_update_begin(c->eng, irs, valid, stride, on_device ? 1 : 0));
    c->upd_pending = true;
    c->upd_wait = wait;
    c->upd_active = (size_t)std::ceil((double)full_len / (double)c->block_size);
    return FCB_OK;
}

extern "C" int fcb_fftconv_update_begin(fcb_fftconv *c, const float *irs, size_t ir_len, int flags)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    if (c->eng && !fcb_engine_update_reserved(c->eng)) return fail(FCB_ERR_ARG, "update_begin: call fcb_fftconv_update_reserve first");
    return fftconv_update_begin(c, irs, ir_len, ir_len, ir_len, false, (flags & FCB_UPDATE_WAIT) != 0);
}

extern "C" int fcb_fftconv_update_pending(const fcb_fftconv *c) { return c && c->upd_pending ? 1 : 0; }

// called at the top of every process call: swap a finished (or, in wait mode, any) background update in
static int fftconv_commit_update(fcb_fftconv *c)
{
    if (!c->upd_pending) return FCB_OK;
    if (!c->upd_wait) {
        const int ready = fcb_engine_update_ready(c->eng);
        if (ready < 0) return fail(FCB_ERR_CUDA, "background IR update failed");
        if (!ready) return FCB_OK; // keep playing the old response; look again at the next call
    }
    FCB_TRY(fcb_engine_update_commit(c->eng));
    c->active_seg_count = c->upd_active;
    c->upd_pending = false;
    return FCB_OK;
}

// the same, but always: whatever update() has accepted becomes the active response now (device-side wait)
static int fftconv_commit_wait(fcb_fftconv *c)
{
    if (!c || !c->upd_pending) return FCB_OK;
    c->upd_wait = true;
    return fftconv_commit_update(c);
}

// multi-block calls never allocate: reserve the workspace for calls of up to `max_call_samples` samples ahead of time
extern "C" int fcb_fftconv_reserve(fcb_fftconv *c, size_t max_call_samples)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    if (!c->eng) return FCB_OK;
    const size_t nb = max_call_samples / c->block_size;
    return nb >= 2 ? fcb_engine_multi_block_reserve(c->eng, nb) : FCB_OK;
}

static bool g_strict_todo = false; // fcb_tune("strict_todo", 1): TwoStage::update / Crossfade::reset answer FCB_ERR_TODO like the reference
extern "C" void fcb_host_mirror_set_strict_todo(int on) { g_strict_todo = on != 0; }

static bool g_mapped_io = true; // fcb_tune("mapped_io", 0) forces the copy-engine path
static bool g_zero_copy = true; // fcb_tune("zero_copy", 0): never let kernels touch caller-pinned host buffers

extern "C" void fcb_host_mirror_set_mapped_io(int on) { g_mapped_io = on != 0; }
extern "C" void fcb_host_mirror_set_zero_copy(int on) { g_zero_copy = on != 0; }

// fcb_tune("xf_speculate", 0): the crossfade host call computes every block's gains on the critical path again
static bool g_xf_speculate = true;
extern "C" void fcb_host_mirror_set_xf_speculate(int on) { g_xf_speculate = on != 0; }

// device-visible alias of a pinned (page-locked) host pointer, or NULL for pageable / device memory
static float *pinned_alias(const float *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? static_cast<float *>(a.devicePointer) : nullptr;
}

static fcb_epilogue offset_epilogue(const fcb_epilogue *epi, size_t off)
{
    fcb_epilogue e;
    memset(&e, 0, sizeof e);
    if (!epi) return e;
    e = *epi;
    if (e.add0) e.add0 += off;
    if (e.add1) e.add1 += off;
    if (e.mix_other) {
        e.mix_other += off;
        e.gains += 2 * off;
    }
    return e;
}

// the block scheduler of :222-294; `host_in`/`host_out` select the H2D/D2H flavour of each chunk
static int fftconv_run(fcb_fftconv *c, const float *in, size_t in_len, size_t in_stride, float *out, size_t out_len,
                       size_t out_stride, const fcb_epilogue *epi, bool host)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    if (out_len && !out) return fail(FCB_ERR_ARG, "NULL output");
    FCB_CUDA(cudaSetDevice(c->opt.device));
    FCB_TRY(fftconv_commit_update(c));
    if (c->active_seg_count == 0) { // :216-219
        if (host) {
            for (size_t ch = 0; ch < c->C; ch++) memset(out + ch * out_stride, 0, out_len * sizeof(float));
            return FCB_OK;
        }
        return zero_planar(out, out_stride, out_len, c->C, c->stream);
    }
    if (in_len < out_len) return fail(FCB_ERR_PANIC, "range end index %zu out of range for slice of length %zu", out_len, in_len);
    if (out_len && !in) return fail(FCB_ERR_ARG, "NULL input");
    const size_t B = c->block_size;
    if (host && g_zero_copy && c->C < 1024 && out_len > 0 && out_len <= B - c->input_buffer_fill) {
        // one chunk, small batch, caller's buffers page-locked: the device path works on them in place
        float *din = pinned_alias(in), *dout = pinned_alias(out);
        if (din && dout) {
            FCB_TRY(fftconv_run(c, din, out_len, in_stride, dout, out_len, out_stride, nullptr, false));
            FCB_CUDA(cudaStreamSynchronize(c->stream));
            return FCB_OK;
        }
    }
    if (host && c->m_in && g_mapped_io && out_len > 0 && out_len <= B - c->input_buffer_fill) {
        // one chunk, small batch: stage through mapped pinned memory and run the device path on it
        const size_t n = out_len;
        for (size_t ch = 0; ch < c->C; ch++) memcpy(c->h_in + ch * B, in + ch * in_stride, n * sizeof(float));
        FCB_TRY(fftconv_run(c, c->m_in, n, B, c->m_out, n, B, nullptr, false));
        FCB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t ch = 0; ch < c->C; ch++) memcpy(out + ch * out_stride, c->h_out + ch * B, n * sizeof(float));
        return FCB_OK;
    }
    size_t processed = 0;
    while (processed < out_len) { // :222
        const bool was_empty = c->input_buffer_fill == 0; // :223
        size_t n = out_len - processed;                   // :224-227
        if (B - c->input_buffer_fill < n) n = B - c->input_buffer_fill;
        const size_t pos = c->input_buffer_fill;
        const bool complete = pos + n == B;
        fcb_epilogue e = offset_epilogue(epi, processed);
        if (was_empty && out_len - processed >= 2 * B && fcb_engine_multi_block_ok(c->eng, c->current, c->active_seg_count)) {
            // the call spans several whole blocks: one time-batched pass over as many as the workspace holds
            size_t nb = (out_len - processed) / B;
            const size_t cap = fcb_engine_multi_block_reserved(c->eng); // never allocates here (fcb_fftconv_reserve)
            if (cap >= 2) {
                if (nb > cap) nb = cap;
                FCB_TRY(fcb_engine_process_blocks(c->eng, in + processed, in_stride, out + processed, out_stride, c->current,
                                                  c->active_seg_count, nb, &e, host ? 1 : 0));
                for (size_t d = 0; d < nb; d++)
                    c->current = c->current > 0 ? c->current - 1 : c->active_seg_count - 1; // :287-291
                processed += nb * B;
                continue;
            }
        }
        float *dst = host ? c->d_io : out + processed;
        const size_t dst_stride = host ? B : out_stride;
        // (only where the whole-block kernel moves its I/O with bulk copies: 32 <= B <= 512, 16-byte
        // aligned rows — plain 4-byte accesses to host memory are 3x slower than staging)
        const bool bulk_io = B >= 32 && B <= 512 && in_stride % 4 == 0 && out_stride % 4 == 0 &&
                             (uintptr_t)(in + processed) % 16 == 0 && (uintptr_t)(out + processed) % 16 == 0;
        if (host && was_empty && complete && g_zero_copy && bulk_io && c->C >= 1024) {
            // caller's buffers are pinned: the whole-block kernel reads the input block and writes the
            // output block over PCIe itself (every CTA its own channels, spread over the kernel's
            // lifetime) — no copy engines, no staging, one launch per step
            float *din = pinned_alias(in + processed), *dout = pinned_alias(out + processed);
            if (din && dout) {
                FCB_TRY(fcb_engine_process_block_dev(c->eng, din, in_stride, dout, out_stride, c->current,
                                                     c->active_seg_count, &e));
                c->current = c->current > 0 ? c->current - 1 : c->active_seg_count - 1; // :287-291
                processed += n;
                continue;
            }
        }
        if (host && was_empty && complete && c->C >= 1024) {
            // many channels, whole block, host buffers: overlap the PCIe copies with K2
            FCB_TRY(fcb_engine_process_block_host(c->eng, in + processed, in_stride, out + processed, out_stride,
                                                  c->current, c->active_seg_count, 0));
            c->current = c->current > 0 ? c->current - 1 : c->active_seg_count - 1; // :287-291
            processed += n;
            continue;
        }
        if (host && was_empty && complete) {
            // whole block from host memory: one copy in, the whole-block device path (one fused
            // kernel for B <= 512), one copy out
            FCB_TRY(fcb_engine_push_input(c->eng, in + processed, in_stride, 0, B));
            FCB_TRY(fcb_engine_process_block_dev(c->eng, fcb_engine_input_buffer(c->eng), B, c->d_io, B, c->current,
                                                 c->active_seg_count, &e));
        } else if (!host && was_empty && complete) {
            // whole block resident on the device: K1 reads the caller's buffer directly
            FCB_TRY(fcb_engine_process_block_dev(c->eng, in + processed, in_stride, dst, dst_stride, c->current,
                                                 c->active_seg_count, &e));
        } else {
            if (host) FCB_TRY(fcb_engine_push_input(c->eng, in + processed, in_stride, pos, n)); // :229-231
            else FCB_TRY(fcb_engine_push_input_dev(c->eng, in + processed, in_stride, pos, n));
            FCB_TRY(fcb_engine_fft_forward(c->eng, c->current, pos + n));                        // :234-241
            if (was_empty) FCB_TRY(fcb_engine_mac(c->eng, c->current, c->active_seg_count));     // :244-255
            FCB_TRY(fcb_engine_ifft_ola(c->eng, c->current, pos, n, complete, dst, dst_stride, &e)); // :256-274
        }
        if (host)
            FCB_CUDA(cudaMemcpy2DAsync(out + processed, out_stride * sizeof(float), c->d_io, B * sizeof(float),
                                       n * sizeof(float), c->C, cudaMemcpyDeviceToHost, c->stream));
        c->input_buffer_fill += n; // :277-292
        if (c->input_buffer_fill == B) {
            c->input_buffer_fill = 0;
            c->current = c->current > 0 ? c->current - 1 : c->active_seg_count - 1;
        }
        processed += n;
    }
    if (host) FCB_CUDA(cudaStreamSynchronize(c->stream));
    return FCB_OK;
}


// Both convolvers at a block boundary with the same ring position, one whole block of device samples: they have
// seen the same input since creation / reset (head + tail0 of a two-stage convolver, A + B of a crossfade), so
// one paired launch serves both (fcb_engine_process_block_pair_dev).
static bool fftconv_pair_ok(const fcb_fftconv *a, const fcb_fftconv *b, size_t n)
{
    return a && b && a->eng && b->eng && n == a->block_size && a->block_size == b->block_size && a->input_buffer_fill == 0 &&
           b->input_buffer_fill == 0 && a->current == b->current && a->active_seg_count == b->active_seg_count &&
           a->active_seg_count >= 1 && a->current < a->active_seg_count && a->stream == b->stream &&
           fcb_engine_pair_ok(a->eng, b->eng, a->active_seg_count);
}
static int fftconv_process_pair_dev(fcb_fftconv *a, fcb_fftconv *b, const float *in, size_t in_stride, float *out_a,
                                    size_t stride_a, const fcb_epilogue *epi_a, float *out_b, size_t stride_b,
                                    const fcb_epilogue *epi_b, float *copy_to = nullptr, size_t copy_stride = 0)
{
    FCB_TRY(fcb_engine_process_block_pair_copy_dev(a->eng, b->eng, in, in_stride, out_a, stride_a, epi_a, out_b, stride_b, epi_b,
                                                   a->current, a->active_seg_count, copy_to, copy_stride));
    a->current = a->current > 0 ? a->current - 1 : a->active_seg_count - 1; // :287-291
    b->current = a->current;
    return FCB_OK;
}

extern "C" int fcb_fftconv_process(fcb_fftconv *c, const float *in, size_t in_len, size_t in_stride, float *out,
                                   size_t out_len, size_t out_stride)
{
    return fftconv_run(c, in, in_len, in_stride, out, out_len, out_stride, nullptr, true);
}
extern "C" int fcb_fftconv_process_dev(fcb_fftconv *c, const float *in, size_t in_len, size_t in_stride, float *out,
                                       size_t out_len, size_t out_stride, const fcb_epilogue *epi)
{
    return fftconv_run(c, in, in_len, in_stride, out, out_len, out_stride, epi, false);
}
extern "C" int fcb_fftconv_sync(fcb_fftconv *c)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    FCB_CUDA(cudaSetDevice(c->opt.device));
    FCB_CUDA(cudaStreamSynchronize(c->stream));
    return FCB_OK;
}
extern "C" fcb_engine *fcb_fftconv_engine(fcb_fftconv *c) { return c ? c->eng : nullptr; }
extern "C" size_t fcb_fftconv_block_size(const fcb_fftconv *c) { return c->block_size; }
extern "C" size_t fcb_fftconv_seg_count(const fcb_fftconv *c) { return c->seg_count; }
extern "C" size_t fcb_fftconv_active_seg_count(const fcb_fftconv *c) { return c->active_seg_count; }
extern "C" size_t fcb_fftconv_current(const fcb_fftconv *c) { return c->current; }
extern "C" size_t fcb_fftconv_fill(const fcb_fftconv *c) { return c->input_buffer_fill; }

// ==============================================================================================
// TwoStageFFTConvolver — src/fft_convolver.rs:323-526
// ==============================================================================================
// :514-526, f32 arithmetic throughout
extern "C" size_t fcb_compute_tail_block_size(size_t head_len, size_t response_len)
{
    const float FFT_K = 1.5f;
    volatile float kn = (FFT_K * (float)head_len) / (2.0f * logf(2.0f));
    volatile float prod = (float)response_len * (float)head_len;
    volatile float sq = kn * kn;
    volatile float sum = sq + prod;
    float b = -kn + sqrtf(sum);
    b = fmaxf(b, (float)head_len);
    return next_power_of_two((size_t)b);
}

struct fcb_twostage {
    size_t C = 0, head_block_size = 0, tail_block_size = 0; // :325-326
    size_t max_response_length = 0; // not a field of the reference struct; fcb_twostage_update re-slices with it
    // EXTENSION (fcb_options.stages > 2): the tail [2T, L) is itself a two-stage convolver whose head block is T — the
    // reference's partition applied recursively (Gardner-style non-uniform partition).  It lives on
    // the tail stream; `tail` is then a Default convolver.
    fcb_twostage *nested = nullptr;
    fcb_fftconv *head = nullptr, *tail0 = nullptr, *tail = nullptr;
    // device [C][T] each (:329-334); tail_in is double-buffered for the asynchronous tail
    float *tail_output0 = nullptr, *tail_precalculated0 = nullptr, *tail_output = nullptr,
          *tail_precalculated = nullptr, *tail_input[2] = {nullptr, nullptr};
    int tail_in_sel = 0;
    size_t tail_input_fill = 0, precalculated_pos = 0; // :335-336
    fcb_options opt{};
    cudaStream_t stream = nullptr, tail_stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev_in = nullptr, ev_tail_done = nullptr;
    bool tail_pending = false;
    float *d_in = nullptr, *d_out = nullptr; // host-call staging [C][head_block_size]
    // small batches: the staging is pinned host memory mapped into the device address space (h_* host views, m_* the
    // device aliases) — a host call is CPU copy in, the launches, one sync, CPU copy out; no copy-engine round trips
    float *h_in = nullptr, *h_out = nullptr, *m_in = nullptr, *m_out = nullptr;
};

extern "C" void fcb_twostage_free(fcb_twostage *c)
{
    if (!c) return;
    cudaSetDevice(c->opt.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->tail_stream) cudaStreamSynchronize(c->tail_stream);
    fcb_twostage_free(c->nested);
    fcb_fftconv_free(c->head);
    fcb_fftconv_free(c->tail0);
    fcb_fftconv_free(c->tail);
    cudaFree(c->tail_output0);
    cudaFree(c->tail_precalculated0);
    cudaFree(c->tail_output);
    cudaFree(c->tail_precalculated);
    cudaFree(c->tail_input[0]);
    cudaFree(c->tail_input[1]);
    cudaFree(c->d_in);
    cudaFree(c->d_out);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->ev_tail_done) cudaEventDestroy(c->ev_tail_done);
    if (c->tail_stream) cudaStreamDestroy(c->tail_stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int twostage_alloc(fcb_twostage *c)
{
    const size_t bytes = c->C * c->tail_block_size * sizeof(float);
    float **bufs[] = {&c->tail_output0, &c->tail_precalculated0, &c->tail_output, &c->tail_precalculated,
                      &c->tail_input[0], &c->tail_input[1]};
    for (float **b : bufs) {
        FCB_CUDA(cudaMalloc(b, bytes ? bytes : 16));
        FCB_CUDA(cudaMemsetAsync(*b, 0, bytes ? bytes : 16, c->stream));
    }
    const size_t io = c->C * (c->head_block_size ? c->head_block_size : 1) * sizeof(float);
    FCB_CUDA(cudaMalloc(&c->d_in, io));
    FCB_CUDA(cudaMalloc(&c->d_out, io));
    if (io <= ((size_t)256 << 10)) { // small batches only
        if (cudaHostAlloc(&c->h_in, io, cudaHostAllocMapped) != cudaSuccess || cudaHostAlloc(&c->h_out, io, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&c->m_in, c->h_in, 0) != cudaSuccess || cudaHostGetDevicePointer(&c->m_out, c->h_out, 0) != cudaSuccess) {
            cudaGetLastError();
            if (c->h_in) cudaFreeHost(c->h_in);
            if (c->h_out) cudaFreeHost(c->h_out);
            c->h_in = c->h_out = c->m_in = c->m_out = nullptr;
        }
    }
    FCB_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
    FCB_CUDA(cudaEventCreateWithFlags(&c->ev_tail_done, cudaEventDisableTiming));
    return FCB_OK;
}

static int twostage_shell(fcb_twostage **out, size_t channels, const fcb_options *opt)
{
    fcb_twostage *c = new fcb_twostage();
    c->C = channels;
    c->opt = opt ? *opt : default_options();
    cudaError_t err = cudaSetDevice(c->opt.device);
    if (err == cudaSuccess) {
        if (c->opt.stream) c->stream = (cudaStream_t)c->opt.stream;
        else {
            err = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
            c->own_stream = err == cudaSuccess;
        }
    }
    if (err == cudaSuccess && c->opt.async_tail) err = cudaStreamCreateWithFlags(&c->tail_stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) {
        delete c;
        return fail(FCB_ERR_CUDA, "two-stage stream setup failed: %s", cudaGetErrorString(err));
    }
    *out = c;
    return FCB_OK;
}

// :340-406
extern "C" int fcb_twostage_init(fcb_twostage **out, const float *irs, size_t channels, size_t ir_len,
                                 size_t block_size, size_t max_response_length, const fcb_options *opt)
{
    if (!out) return fail(FCB_ERR_ARG, "NULL out");
    *out = nullptr;
    if (channels == 0) return fail(FCB_ERR_ARG, "channels must be >= 1");
    fcb_options o = opt ? *opt : default_options();
    o.shared_ir = 0;
    const size_t head = block_size; // :341 (kept unrounded for the bookkeeping)
    size_t T = o.forced_tail_block ? o.forced_tail_block : fcb_compute_tail_block_size(block_size, max_response_length);
    const size_t stages = o.stages < 2 ? 2 : o.stages;
    if (stages > 2 && !o.forced_tail_block && T > 16384) T = 16384; // nested levels: stay within the engine's block limit
    if (max_response_length < ir_len) // :344-348
        return fail(FCB_ERR_PANIC, "max_response_length must be at least the length of the initial impulse response");
    if (head == 0) return fail(FCB_ERR_PANIC, "attempt to calculate the remainder with a divisor of zero"); // :431
    const size_t L = max_response_length;
    // padded_ir (:349-350), channel-major
    std::vector<float> padded(channels * (L ? L : 1), 0.f);
    for (size_t ch = 0; ch < channels; ch++)
        if (ir_len) memcpy(&padded[ch * L], irs + ch * ir_len, ir_len * sizeof(float));

    fcb_twostage *c = nullptr;
    FCB_TRY(twostage_shell(&c, channels, &o));
    c->head_block_size = head;
    c->tail_block_size = T;
    c->max_response_length = L;
    fcb_options sub = o;
    sub.stream = (void *)c->stream;
    int rc = FCB_OK;
    auto gather = [&](size_t off, size_t len) { // [C][len] slice of padded
        std::vector<float> v(channels * (len ? len : 1));
        for (size_t ch = 0; ch < channels; ch++) memcpy(&v[ch * len], &padded[ch * L + off], len * sizeof(float));
        return v;
    };
    {
        const size_t head_ir_len = L < T ? L : T; // :352-354
        auto v = gather(0, head_ir_len);
        rc = fcb_fftconv_init(&c->head, v.data(), channels, head_ir_len, head, head_ir_len, &sub);
    }
    if (rc == FCB_OK) {
        if (L > T) { // :356-368
            const size_t tl = (L - T) < T ? (L - T) : T;
            auto v = gather(T, tl);
            rc = fcb_fftconv_init(&c->tail0, v.data(), channels, tl, head, tl, &sub);
        } else {
            rc = fcb_fftconv_default(&c->tail0, channels, &sub);
        }
    }
    if (rc == FCB_OK) {
        fcb_options tsub = sub;
        if (c->tail_stream) tsub.stream = (void *)c->tail_stream;
        if (L > 2 * T && stages > 2) { // EXTENSION: the tail is again a two-stage convolver, fed T samples per call
            const size_t tl = L - 2 * T;
            auto v = gather(2 * T, tl);
            fcb_options nsub = tsub;
            nsub.stages = stages - 1;
            nsub.async_tail = 0; // everything of the tail stays in order on the tail stream
            // the nested level's own T: the reference's formula on what is left, within the engine's block limit
            const size_t tn = fcb_compute_tail_block_size(T, tl);
            nsub.forced_tail_block = tn > 16384 ? 16384 : tn;
            rc = fcb_twostage_init(&c->nested, v.data(), channels, tl, T, tl, &nsub);
            if (rc == FCB_OK) rc = fcb_fftconv_default(&c->tail, channels, &tsub);
        } else if (L > 2 * T) { // :373-384
            const size_t tl = L - 2 * T;
            auto v = gather(2 * T, tl);
            rc = fcb_fftconv_init(&c->tail, v.data(), channels, tl, T, tl, &tsub);
        } else {
            rc = fcb_fftconv_default(&c->tail, channels, &tsub);
        }
    }
    if (rc == FCB_OK) rc = twostage_alloc(c);
    if (rc == FCB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(FCB_ERR_CUDA, "sync failed");
    if (rc != FCB_OK) {
        fcb_twostage_free(c);
        return rc;
    }
    *out = c;
    return FCB_OK;
}

static int twostage_quiesce(const fcb_twostage *c)
{
    FCB_CUDA(cudaSetDevice(c->opt.device));
    FCB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->tail_stream) FCB_CUDA(cudaStreamSynchronize(c->tail_stream));
    return c->nested ? twostage_quiesce(c->nested) : FCB_OK;
}

// move a (nested) two-stage convolver and everything under it onto stream `st`
static int twostage_rehome(fcb_twostage *n, cudaStream_t st)
{
    FCB_TRY(twostage_quiesce(n));
    for (fcb_fftconv *f : {n->head, n->tail0, n->tail}) {
        if (f->eng) FCB_TRY(fcb_engine_set_stream(f->eng, (void *)st));
        if (f->own_stream) cudaStreamDestroy(f->stream);
        f->own_stream = false;
        f->stream = st;
        f->opt.stream = (void *)st;
    }
    if (n->own_stream && n->stream) cudaStreamDestroy(n->stream);
    n->own_stream = false;
    n->stream = st;
    n->opt.stream = (void *)st;
    return n->nested ? twostage_rehome(n->nested, st) : FCB_OK;
}

extern "C" int fcb_twostage_clone(const fcb_twostage *s, fcb_twostage **out)
{
    if (!s || !out) return fail(FCB_ERR_ARG, "NULL argument");
    FCB_TRY(twostage_quiesce(s));
    fcb_twostage *c = nullptr;
    FCB_TRY(twostage_shell(&c, s->C, &s->opt));
    c->head_block_size = s->head_block_size;
    c->tail_block_size = s->tail_block_size;
    c->max_response_length = s->max_response_length;
    c->tail_input_fill = s->tail_input_fill;
    c->precalculated_pos = s->precalculated_pos;
    c->tail_in_sel = s->tail_in_sel;
    int rc = fcb_fftconv_clone(s->head, &c->head);
    if (rc == FCB_OK) rc = fcb_fftconv_clone(s->tail0, &c->tail0);
    if (rc == FCB_OK) rc = fcb_fftconv_clone(s->tail, &c->tail);
    if (rc == FCB_OK && s->nested) {
        rc = fcb_twostage_clone(s->nested, &c->nested);
        if (rc == FCB_OK) rc = twostage_rehome(c->nested, c->tail_stream ? c->tail_stream : c->stream);
    }
    if (rc == FCB_OK) rc = twostage_alloc(c);
    if (rc == FCB_OK) {
        // the clones carry their own private streams; re-home them onto this object's streams
        for (fcb_fftconv *f : {c->head, c->tail0}) {
            if (f->eng) rc = rc ? rc : fcb_engine_set_stream(f->eng, (void *)c->stream);
            if (f->own_stream) cudaStreamDestroy(f->stream);
            f->own_stream = false;
            f->stream = c->stream;
            f->opt.stream = (void *)c->stream;
        }
        cudaStream_t ts = c->tail_stream ? c->tail_stream : c->stream;
        if (c->tail->eng) rc = rc ? rc : fcb_engine_set_stream(c->tail->eng, (void *)ts);
        if (c->tail->own_stream) cudaStreamDestroy(c->tail->stream);
        c->tail->own_stream = false;
        c->tail->stream = ts;
        c->tail->opt.stream = (void *)ts;
    }
    if (rc == FCB_OK) {
        const size_t bytes = s->C * s->tail_block_size * sizeof(float);
        const float *src[] = {s->tail_output0, s->tail_precalculated0, s->tail_output, s->tail_precalculated,
                              s->tail_input[0], s->tail_input[1]};
        float *dst[] = {c->tail_output0, c->tail_precalculated0, c->tail_output, c->tail_precalculated,
                        c->tail_input[0], c->tail_input[1]};
        for (int i = 0; i < 6 && rc == FCB_OK; i++)
            if (bytes && cudaMemcpyAsync(dst[i], src[i], bytes, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess)
                rc = fail(FCB_ERR_CUDA, "two-stage clone copy failed");
        if (rc == FCB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(FCB_ERR_CUDA, "sync failed");
    }
    if (rc != FCB_OK) {
        fcb_twostage_free(c);
        return rc;
    }
    *out = c;
    return FCB_OK;
}

// :408-410 is todo!() in the reference.  EXTENSION (semantics written down in DESIGN.md §9 and
// pinned by the CPU restatement the tests check against): the per-stage FFTConvolver::update (:174-213) on the response
// re-sliced the way init slices it (:349-384) — zero-padded to max_response_length, head [0, min(L,T)), tail0
// [T, T + min(L-T,T)), tail [2T, L); every stage keeps its whole slice active.  Input rings, partially filled blocks,
// tail_input and the tail outputs already computed with the old response are kept (audio in flight), each stage's
// overlap and pre_multiplied are zeroed (:185-188).  No allocation: the slices are read in place with the caller's
// stride.  fcb_tune("strict_todo", 1) restores the reference's answer (FCB_ERR_TODO).
extern "C" int fcb_twostage_update(fcb_twostage *c, const float *irs, size_t ir_len)
{
    if (g_strict_todo) return fail(FCB_ERR_TODO, "not yet implemented");
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    const size_t L = c->max_response_length, T = c->tail_block_size;
    if (ir_len > L) return fail(FCB_ERR_PANIC, "New impulse response is longer than initialized length");
    if (ir_len && !irs) return fail(FCB_ERR_ARG, "NULL impulse response");
    FCB_CUDA(cudaSetDevice(c->opt.device));
    auto stage = [&](fcb_fftconv *f, size_t off, size_t slice) { // `slice` samples of the padded response from `off`
        const size_t valid = ir_len > off ? (ir_len - off < slice ? ir_len - off : slice) : 0;
        return fftconv_update_padded(f, valid ? irs + off : irs, valid, ir_len, slice);
    };
    FCB_TRY(stage(c->head, 0, L < T ? L : T));
    if (L > T) FCB_TRY(stage(c->tail0, T, (L - T) < T ? (L - T) : T));
    if (L > 2 * T) {
        if (c->nested) {
            const size_t off = 2 * T, valid = ir_len > off ? ir_len - off : 0;
            if (c->C > 1 && valid) return fail(FCB_ERR_UNSUPPORTED, "update of a nested partition: batched responses need a contiguous slice per channel");
            FCB_TRY(fcb_twostage_update(c->nested, valid ? irs + off : irs, valid));
        } else {
            FCB_TRY(stage(c->tail, 2 * T, L - 2 * T));
        }
    }
    return FCB_OK;
}

// :497-511
extern "C" int fcb_twostage_reset(fcb_twostage *c)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    FCB_TRY(twostage_quiesce(c));
    c->tail_pending = false;
    FCB_TRY(fcb_fftconv_reset(c->head));
    FCB_TRY(fcb_fftconv_reset(c->tail0));
    FCB_TRY(fcb_fftconv_reset(c->tail));
    if (c->nested) FCB_TRY(fcb_twostage_reset(c->nested));
    const size_t bytes = c->C * c->tail_block_size * sizeof(float);
    float *bufs[] = {c->tail_output0, c->tail_precalculated0, c->tail_output, c->tail_precalculated,
                     c->tail_input[0], c->tail_input[1]};
    for (float *b : bufs)
        if (bytes) FCB_CUDA(cudaMemsetAsync(b, 0, bytes, c->stream));
    c->tail_input_fill = 0;
    c->precalculated_pos = 0;
    return twostage_quiesce(c);
}

// :479-486 — the big tail block: one FFTConvolver call of T samples, or (nested partition) one two-stage call
static int twostage_run_tail(fcb_twostage *c, const float *tin)
{
    const size_t T = c->tail_block_size;
    if (c->nested) return fcb_twostage_process_dev(c->nested, tin, T, T, c->tail_output, T, T);
    return fcb_fftconv_process_dev(c->tail, tin, T, T, c->tail_output, T, T, nullptr);
}

// :412-495 on device buffers
extern "C" int fcb_twostage_process_dev(fcb_twostage *c, const float *in, size_t in_len, size_t in_stride, float *out,
                                        size_t out_len, size_t out_stride)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    if (!(in_len <= c->head_block_size)) // assert! :414
        return fail(FCB_ERR_PANIC, "assertion failed: input.len() <= self.head_block_size");
    // head.process slices input[..output.len()], the tail loop indexes output[..input.len()]
    if (in_len != out_len) return fail(FCB_ERR_PANIC, "index out of bounds: input and output lengths differ");
    FCB_CUDA(cudaSetDevice(c->opt.device));
    const size_t H = c->head_block_size, T = c->tail_block_size, C = c->C;

    // Can the head/tail sum (:438-454) ride on the head's K3 launch?  Only when this call is one
    // tail piece and one head chunk; otherwise the sum runs as a separate elementwise kernel.
    const bool one_piece = in_len <= H - (c->tail_input_fill % H);
    const bool one_chunk = c->head->active_seg_count == 0 ? false
                                                          : in_len <= c->head->block_size - c->head->input_buffer_fill;
    // a head block size that does not divide T runs off the end of tail_input: the reference's
    // slice at :459 panics there (every non-power-of-two head size does, eventually)
    const bool overflow = c->tail_input_fill + (one_piece ? in_len : 0) > T;
    const bool fuse = T != 0 && in_len > 0 && one_piece && one_chunk && !overflow;
    fcb_epilogue epi;
    memset(&epi, 0, sizeof epi);
    if (fuse) {
        epi.add0 = c->tail_precalculated0 + c->precalculated_pos;
        epi.add1 = c->tail_precalculated + c->precalculated_pos;
        epi.add_stride = T;
    }
    // head and tail_convolver0 see the same head blocks (:417 and :464-473): when this call is one whole head block
    // both run in one paired launch, tail0 writing where :464-473 would have put its block
    const bool paired = fuse && in_len == H && c->tail_input_fill % H == 0 && fftconv_pair_ok(c->head, c->tail0, in_len);
    // ... and the append of the input block to tail_input (:459-461) rides on the same launch
    const bool copy_fused = paired && T % 2 == 0 && c->tail_input_fill % 2 == 0;
    if (paired)
        FCB_TRY(fftconv_process_pair_dev(c->head, c->tail0, in, in_stride, out, out_stride, &epi,
                                         c->tail_output0 + c->tail_input_fill, T, nullptr,
                                         copy_fused ? c->tail_input[c->tail_in_sel] + c->tail_input_fill : nullptr, T));
    else
        FCB_TRY(fcb_fftconv_process_dev(c->head, in, in_len, in_stride, out, out_len, out_stride, fuse ? &epi : nullptr)); // :417
    if (T == 0) return FCB_OK; // :420-422

    size_t processed = 0;
    while (processed < in_len) { // :427
        const size_t remaining = in_len - processed;
        size_t n = H - (c->tail_input_fill % H); // :429-432
        if (remaining < n) n = remaining;
        if (!fuse && c->precalculated_pos + n <= T) { // :438-454 (past T the reference panics at :442)
            long long total = (long long)C * (long long)n;
            k_add2<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(
                out + processed, (long long)out_stride, c->tail_precalculated0 + c->precalculated_pos,
                c->tail_precalculated + c->precalculated_pos, (long long)T, (int)n, (long long)C);
            g_launches++;
            FCB_CUDA(cudaGetLastError());
        }
        c->precalculated_pos += n; // :456
        if (c->tail_input_fill + n > T) // slice index panic at :459
            return fail(FCB_ERR_PANIC, "range end index %zu out of range for slice of length %zu", c->tail_input_fill + n, T);
        float *tin = c->tail_input[c->tail_in_sel];
        if (!copy_fused)
            FCB_CUDA(cudaMemcpy2DAsync(tin + c->tail_input_fill, T * sizeof(float), in + processed, in_stride * sizeof(float),
                                       n * sizeof(float), C, cudaMemcpyDefault, c->stream)); // :459-461
        c->tail_input_fill += n;

        if (c->tail_input_fill % H == 0) { // :464-476
            const size_t off = c->tail_input_fill - H;
            if (!paired) FCB_TRY(fcb_fftconv_process_dev(c->tail0, tin + off, H, T, c->tail_output0 + off, H, T, nullptr));
            if (c->tail_input_fill == T) std::swap(c->tail_precalculated0, c->tail_output0);
        }
        if (c->tail_input_fill == T) { // :479-486
            if (c->tail_stream) {
                // the previous background tail must have produced tail_output before it becomes
                // tail_precalculated, and must be done before we hand it new buffers
                if (c->tail_pending) FCB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_tail_done, 0));
                std::swap(c->tail_precalculated, c->tail_output);
                FCB_CUDA(cudaEventRecord(c->ev_in, c->stream));
                FCB_CUDA(cudaStreamWaitEvent(c->tail_stream, c->ev_in, 0));
                FCB_TRY(twostage_run_tail(c, tin));
                FCB_CUDA(cudaEventRecord(c->ev_tail_done, c->tail_stream));
                c->tail_pending = true;
                c->tail_in_sel ^= 1; // the tail stream may still be reading `tin`
            } else {
                std::swap(c->tail_precalculated, c->tail_output);
                FCB_TRY(twostage_run_tail(c, tin));
            }
        }
        if (c->tail_input_fill == T) { // :488-491
            c->tail_input_fill = 0;
            c->precalculated_pos = 0;
        }
        processed += n;
    }
    return FCB_OK;
}

extern "C" int fcb_twostage_process(fcb_twostage *c, const float *in, size_t in_len, size_t in_stride, float *out,
                                    size_t out_len, size_t out_stride)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    if (!(in_len <= c->head_block_size)) return fail(FCB_ERR_PANIC, "assertion failed: input.len() <= self.head_block_size");
    if (in_len != out_len) return fail(FCB_ERR_PANIC, "index out of bounds: input and output lengths differ");
    if (in_len == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(c->opt.device));
    if (g_zero_copy) { // the caller's own buffers are page-locked (fcb_host_alloc): the kernels read and write them in place
        float *din = pinned_alias(in), *dout = pinned_alias(out);
        if (din && dout) {
            FCB_TRY(fcb_twostage_process_dev(c, din, in_len, in_stride, dout, out_len, out_stride));
            FCB_CUDA(cudaStreamSynchronize(c->stream));
            return FCB_OK;
        }
    }
    if (c->m_in && g_mapped_io) { // small batch: the kernels read and write mapped pinned staging themselves
        const size_t H = c->head_block_size;
        for (size_t ch = 0; ch < c->C; ch++) memcpy(c->h_in + ch * H, in + ch * in_stride, in_len * sizeof(float));
        FCB_TRY(fcb_twostage_process_dev(c, c->m_in, in_len, H, c->m_out, out_len, H));
        FCB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t ch = 0; ch < c->C; ch++) memcpy(out + ch * out_stride, c->h_out + ch * H, out_len * sizeof(float));
        return FCB_OK;
    }
    FCB_CUDA(cudaMemcpy2DAsync(c->d_in, in_len * sizeof(float), in, in_stride * sizeof(float), in_len * sizeof(float),
                               c->C, cudaMemcpyHostToDevice, c->stream));
    FCB_TRY(fcb_twostage_process_dev(c, c->d_in, in_len, in_len, c->d_out, out_len, out_len));
    FCB_CUDA(cudaMemcpy2DAsync(out, out_stride * sizeof(float), c->d_out, out_len * sizeof(float),
                               out_len * sizeof(float), c->C, cudaMemcpyDeviceToHost, c->stream));
    FCB_CUDA(cudaStreamSynchronize(c->stream));
    return FCB_OK;
}

extern "C" int fcb_twostage_sync(fcb_twostage *c)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    return twostage_quiesce(c);
}
extern "C" size_t fcb_twostage_tail_block_size(const fcb_twostage *c) { return c->tail_block_size; }
// block sizes of the (nested) partition, outermost first: head, T1, T2, ...; returns how many there are
extern "C" size_t fcb_twostage_stage_blocks(const fcb_twostage *c, size_t *out, size_t cap)
{
    size_t n = 0;
    if (!c) return 0;
    if (n < cap) out[n] = c->head_block_size;
    n++;
    for (; c; c = c->nested) {
        if (n < cap) out[n] = c->tail_block_size;
        n++;
    }
    return n;
}

// ==============================================================================================
// Crossfader<RaisedCosineMixer> — src/crossfade_convolver.rs:160-279, run on the host one sample
// at a time exactly like the reference (sequential f32 accumulation of mix_value); it
// emits the per-sample gain pair that K3 / k_mix apply.
// ==============================================================================================
namespace {
struct Crossfader {
    int64_t fading_samples = 0, hold_samples = 0, counter = 0; // :195-197
    float mix_value_step = 0.f, mix_value = 0.f;               // :198-199
    bool approaching = false;                                   // FadingState :178-181
    int target = 0;                                             // Target::A = 0, B = 1

    void init(size_t fading, size_t hold) // :204-214
    {
        fading_samples = (int64_t)fading;
        hold_samples = (int64_t)hold;
        counter = 0;
        mix_value_step = 1.0f / (float)fading;
        mix_value = 0.f;
        approaching = false;
        target = 0;
    }
    void fade_into(int t) // :216-240
    {
        if (target == t) return;
        if (!approaching) {
            counter = -hold_samples;
            approaching = true;
            target = t;
            mix_value_step = -mix_value_step;
        } else if (counter >= 0) {
            counter = fading_samples - counter;
            target = t;
            mix_value_step = -mix_value_step;
        } else {
            approaching = false;
            target = t;
        }
    }
    // :242-278 with the sample values factored out: returns {gain on a, gain on b}
    float2 next_gains()
    {
        const float2 take_a = make_float2(1.f, 0.f), take_b = make_float2(0.f, 1.f);
        if (!approaching) return target == 0 ? take_a : take_b;
        counter += 1;
        if (counter <= 0) return target == 0 ? take_b : take_a; // holding the previous target
        volatile float mv = mix_value + mix_value_step;         // plain f32 add, never widened
        mix_value = mv;
        if (counter == fading_samples) {
            approaching = false;
            mix_value = target == 0 ? 0.f : 1.f;
            return target == 0 ? take_a : take_b;
        }
        // RaisedCosineMixer :160-169
        const float PI_HALF = 3.14159265358979323846f * 0.5f; // :147
        volatile float rad = PI_HALF * mix_value;
        float cs = cosf(rad);
        volatile float gain1 = cs * cs; // powi(2)
        volatile float gain2 = 1.0f - gain1;
        return make_float2(gain1, gain2);
    }
};
} // namespace

// ==============================================================================================
// CrossfadeConvolver<FFTConvolver> — src/crossfade_convolver.rs:3-105
// ==============================================================================================
struct fcb_crossfade {
    fcb_fftconv *a = nullptr, *b = nullptr; // convolver_a / convolver_b :5-6
    Crossfader crossfader;                  // :7
    size_t C = 0, max_buffer_size = 0;
    float *buffer_a = nullptr, *buffer_b = nullptr; // device [C][max_buffer_size] :13-14
    float *stored_response = nullptr;               // DEVICE [C][stored_len] (:15): uploaded when update()
                                                    // is called mid-fade, so the deferred swap inside
                                                    // process() is K5 only — no host copy, no PCIe
    size_t stored_len = 0;
    bool response_pending = false;                  // :16
    cudaStream_t stream = nullptr;
    int device = 0;
    float2 *h_gains = nullptr, *d_gains = nullptr;  // pinned / device, max_buffer_size each
    // Gains of the NEXT call, computed by the synchronous host call while the GPU works on the current block (the
    // per-sample state machine costs ~20 ns a sample: 10 us per 512-sample block, otherwise ahead of the launch).
    // `crossfader` itself is not advanced: the speculated copy replaces it when the next call really asks for
    // `spec_len` samples, and anything else that touches the crossfader (swap, reset, clone) drops the speculation.
    std::vector<float2> spec_gains;
    Crossfader spec_after;
    size_t spec_len = 0;
    bool spec_valid = false, spec_all_a = false, spec_all_b = false;
    bool gains_in_flight = false; // h_gains was handed to an upload nobody has waited for yet
    cudaEvent_t ev_gains = nullptr;
    float *d_in = nullptr, *d_out = nullptr;        // host-call staging, [C][max_buffer_size] each (allocated in new())
    float *h_in = nullptr, *h_out = nullptr, *m_in = nullptr, *m_out = nullptr; // small batches: mapped pinned staging
    size_t max_response_length = 0;                 // as given to new() (= stored_len)
    size_t crossfade_samples = 0;
    // update() while a fade runs (:58-63) also starts K5 of the stored response into the shadow buffer of the
    // convolver that will take it when the fade ends, so the deferred swap inside process() (:67-70) is a pointer flip
    bool stored_staged = false;
    bool update_nowait = false; // inside fcb_crossfade_update_begin: the caller keeps the source buffer alive
    // A is a deep clone of B (:29) and both are fed the same samples, so their input-spectrum rings hold the same
    // spectra in the same slots — until an update() with another segment count (:204) lets their `current` drift
    // apart at the next wrap (:301-305); a call that starts with unequal segment counts or positions turns the
    // paired launch off for good
    bool rings_same = true;
};

extern "C" void fcb_crossfade_free(fcb_crossfade *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->buffer_a);
    cudaFree(c->buffer_b);
    cudaFree(c->stored_response);
    cudaFree(c->d_gains);
    cudaFree(c->d_in);
    cudaFree(c->d_out);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_gains) cudaFreeHost(c->h_gains);
    if (c->ev_gains) cudaEventDestroy(c->ev_gains);
    fcb_fftconv_free(c->a); // b owns the shared stream when it is private, so free it last
    fcb_fftconv_free(c->b);
    delete c;
}

// :19-43
extern "C" int fcb_crossfade_new(fcb_crossfade **out, fcb_fftconv *convolver, size_t max_response_length,
                                 size_t max_buffer_size, size_t crossfade_samples)
{
    if (!out || !convolver) return fail(FCB_ERR_ARG, "NULL argument");
    *out = nullptr;
    fcb_crossfade *c = new fcb_crossfade();
    c->C = convolver->C;
    c->device = convolver->opt.device;
    c->stream = convolver->stream;
    c->b = convolver; // :30
    c->max_buffer_size = max_buffer_size;
    c->stored_len = max_response_length;
    c->max_response_length = max_response_length;
    c->crossfade_samples = crossfade_samples;
    c->crossfader.init(crossfade_samples, max_buffer_size < max_response_length ? max_buffer_size : max_response_length); // :31-35
    int rc = FCB_OK;
    {
        // convolver_a = convolver.clone() (:29), sharing b's stream so A, B and the mix stay ordered
        fcb_options o = convolver->opt;
        o.stream = (void *)convolver->stream;
        fcb_fftconv tmp = *convolver;
        tmp.opt = o;
        rc = fcb_fftconv_clone(&tmp, &c->a);
    }
    const size_t n = max_buffer_size ? max_buffer_size : 1;
    auto cu = [&](cudaError_t e) {
        if (rc == FCB_OK && e != cudaSuccess) rc = fail(FCB_ERR_CUDA, "crossfade setup failed: %s", cudaGetErrorString(e));
    };
    cu(cudaSetDevice(c->device));
    cu(cudaMalloc(&c->buffer_a, c->C * n * sizeof(float)));
    cu(cudaMalloc(&c->buffer_b, c->C * n * sizeof(float)));
    cu(cudaMalloc(&c->stored_response, c->C * (max_response_length ? max_response_length : 1) * sizeof(float))); // :26
    cu(cudaMalloc(&c->d_gains, n * sizeof(float2)));
    cu(cudaHostAlloc(&c->h_gains, n * sizeof(float2), cudaHostAllocDefault));
    c->spec_gains.resize(n); // sized here once: process() never allocates
    cu(cudaMalloc(&c->d_out, c->C * n * sizeof(float)));
    cu(cudaMalloc(&c->d_in, c->C * n * sizeof(float)));
    if (rc == FCB_OK && c->C * n * sizeof(float) <= ((size_t)1 << 20)) { // small batches: mapped pinned staging, read and
        // written by the paired kernel's bulk copies (one burst per channel); fcb_tune("mapped_io", 0) = copy engines
        const size_t io = c->C * n * sizeof(float);
        if (cudaHostAlloc(&c->h_in, io, cudaHostAllocMapped) != cudaSuccess || cudaHostAlloc(&c->h_out, io, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&c->m_in, c->h_in, 0) != cudaSuccess || cudaHostGetDevicePointer(&c->m_out, c->h_out, 0) != cudaSuccess) {
            cudaGetLastError();
            if (c->h_in) cudaFreeHost(c->h_in);
            if (c->h_out) cudaFreeHost(c->h_out);
            c->h_in = c->h_out = c->m_in = c->m_out = nullptr;
        }
    }
    cu(cudaEventCreateWithFlags(&c->ev_gains, cudaEventDisableTiming));
    // live response changes are what this type is for: both convolvers get a shadow copy of their spectra so that
    // update() never makes a block wait for K5 (fcb_engine_update_*)
    if (rc == FCB_OK && c->a && c->a->eng) rc = fcb_engine_update_reserve(c->a->eng);
    if (rc == FCB_OK && c->b->eng) rc = fcb_engine_update_reserve(c->b->eng);
    if (rc == FCB_OK) {
        cu(cudaMemsetAsync(c->buffer_a, 0, c->C * n * sizeof(float), c->stream));
        cu(cudaMemsetAsync(c->buffer_b, 0, c->C * n * sizeof(float), c->stream));
        cu(cudaEventRecord(c->ev_gains, c->stream));
        cu(cudaStreamSynchronize(c->stream));
    }
    if (rc != FCB_OK) {
        c->b = nullptr; // ownership stays with the caller on failure
        fcb_crossfade_free(c);
        return rc;
    }
    *out = c;
    return FCB_OK;
}

// :46-49 — response.len() is passed for both the stored capacity and crossfade_samples
extern "C" int fcb_crossfade_init(fcb_crossfade **out, const float *irs, size_t channels, size_t ir_len,
                                  size_t max_block_size, size_t max_response_length, const fcb_options *opt)
{
    if (!out) return fail(FCB_ERR_ARG, "NULL out");
    fcb_options o = opt ? *opt : default_options();
    o.shared_ir = 0;
    fcb_fftconv *conv = nullptr;
    FCB_TRY(fcb_fftconv_init(&conv, irs, channels, ir_len, max_block_size, max_response_length, &o));
    int rc = fcb_crossfade_new(out, conv, ir_len, max_block_size, ir_len);
    if (rc != FCB_OK) fcb_fftconv_free(conv);
    return rc;
}

extern "C" int fcb_crossfade_is_crossfading(const fcb_crossfade *c) { return c && c->crossfader.approaching; } // :85-92

// :94-105.  The idle convolver's update() runs in the background (K5 into its shadow spectra on the update stream);
// it is committed by that convolver's next process call, which waits for it ON THE DEVICE if it is still running —
// same result as the synchronous update, and the host never blocks on K5.
static int crossfade_swap(fcb_crossfade *c, const float *irs, size_t len, bool from_stored)
{
    int rc = FCB_OK;
    fcb_fftconv *idle = c->crossfader.target == 0 ? c->b : c->a;
    if (!idle->eng || !fcb_engine_update_reserved(idle->eng)) {
        rc = from_stored ? fftconv_update_dev(idle, irs, len, c->stored_len) : fcb_fftconv_update(idle, irs, len);
    } else if (from_stored) {
        // stored_response (:58-63), staged into the shadow buffer when update() arrived; `len` == stored_len
        if (!c->stored_staged) rc = fftconv_update_begin(idle, c->stored_response, len, c->stored_len, len, true, true);
        else {
            idle->upd_pending = true; // the K5 queued by fcb_crossfade_update: commit it at the next process call
            idle->upd_wait = true;
            idle->upd_active = (size_t)std::ceil((double)len / (double)idle->block_size);
        }
    } else {
        if (len > idle->ir_len) return fail(FCB_ERR_PANIC, "New impulse response is longer than initialized length");
        rc = fftconv_update_begin(idle, irs, len, len, len, false, true);
        // the caller may drop `irs` when update() returns: a DMA out of page-locked memory is still in flight, wait for
        // it (pageable sources were staged by the driver during the call); fcb_crossfade_update_begin skips this wait
        if (rc == FCB_OK && len && !c->update_nowait && pinned_alias(irs)) rc = fcb_engine_update_wait(idle->eng);
    }
    c->stored_staged = false;
    c->spec_valid = false;
    c->crossfader.fade_into(c->crossfader.target == 0 ? 1 : 0);
    return rc;
}

// :51-64
extern "C" int fcb_crossfade_update(fcb_crossfade *c, const float *irs, size_t len)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    if (!fcb_crossfade_is_crossfading(c)) {
        int rc = crossfade_swap(c, irs, len, false);
        c->response_pending = false;
        return rc;
    }
    if (!(len <= c->stored_len)) return fail(FCB_ERR_PANIC, "assertion failed: response_len <= self.stored_response.len()");
    // :61-62 — copy + zero-fill the rest, straight into the device copy (stream-ordered after any
    // kernel still reading a previous pending response)
    FCB_CUDA(cudaSetDevice(c->device));
    // a convolver built over ONE shared response (fcb_crossfade_new on a shared_ir FFTConvolver) is handed one row
    const size_t rows = c->b->opt.shared_ir ? 1 : c->C;
    fcb_fftconv *next = c->crossfader.target == 0 ? c->b : c->a; // takes the stored response when the fade ends (:67-70)
    const bool shadow = next->eng && fcb_engine_update_reserved(next->eng) && len <= next->ir_len && c->stored_len <= next->ir_len;
    cudaStream_t st = c->stream;
    if (shadow) FCB_TRY(fcb_engine_update_join(next->eng, (void *)st)); // an earlier pending response may still be read by its K5
    if (len)
        FCB_CUDA(cudaMemcpy2DAsync(c->stored_response, c->stored_len * sizeof(float), irs, len * sizeof(float),
                                   len * sizeof(float), rows, cudaMemcpyHostToDevice, st));
    if (c->stored_len > len)
        FCB_CUDA(cudaMemset2DAsync(c->stored_response + len, c->stored_len * sizeof(float), 0,
                                   (c->stored_len - len) * sizeof(float), rows, st));
    if (!c->update_nowait) FCB_CUDA(cudaStreamSynchronize(st)); // the caller's buffer is free to change on return
    c->stored_staged = false;
    if (shadow) {
        // K5 now, in the background, into the spectra that are not playing; latest update wins (same stream order)
        FCB_TRY(fcb_engine_update_begin(next->eng, c->stored_response, c->stored_len, c->stored_len, 1));
        c->stored_staged = true;
    }
    c->response_pending = true;
    return FCB_OK;
}

// update() for real-time callers: identical state changes, but never waits for a copy or a kernel.  `irs` must be
// page-locked (fcb_host_alloc) or the call degrades to fcb_crossfade_update, and must stay untouched until
// fcb_crossfade_update_pending() returns 0.
extern "C" int fcb_crossfade_update_begin(fcb_crossfade *c, const float *irs, size_t len)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    c->update_nowait = len == 0 || pinned_alias(irs) != nullptr;
    const int rc = fcb_crossfade_update(c, irs, len);
    c->update_nowait = false;
    return rc;
}
extern "C" int fcb_crossfade_update_pending(fcb_crossfade *c)
{
    if (!c) return 0;
    for (fcb_fftconv *f : {c->a, c->b})
        if (f && f->eng && fcb_engine_update_reserved(f->eng) && fcb_engine_update_ready(f->eng) == 0) return 1;
    return cudaStreamQuery(c->stream) == cudaErrorNotReady ? 1 : 0; // the copy into stored_response rides on the main stream
}

// :66-78 on device buffers
extern "C" int fcb_crossfade_process_dev(fcb_crossfade *c, const float *in, size_t in_len, size_t in_stride, float *out,
                                         size_t out_len, size_t out_stride)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    FCB_CUDA(cudaSetDevice(c->device));
    FCB_TRY(fftconv_commit_update(c->a)); // background updates accepted by update() take effect here (the paired launch
    FCB_TRY(fftconv_commit_update(c->b)); // below does not go through the convolvers' own process)
    if (!fcb_crossfade_is_crossfading(c) && c->response_pending) { // :67-70
        FCB_TRY(crossfade_swap(c, c->stored_response, c->stored_len, true));
        c->response_pending = false;
        FCB_TRY(fftconv_commit_update(c->a));
        FCB_TRY(fftconv_commit_update(c->b));
    }
    const size_t M = c->max_buffer_size;
    if (out_len > M) return fail(FCB_ERR_PANIC, "index out of bounds: the len is %zu but the index is %zu", M, M); // :76

    // gains for this call, by the reference's per-sample state machine (:75-77 -> :242-278)
    if (c->gains_in_flight) { // (a completed event still costs 1.4 us to ask: scripts/api_cost_probe.cu)
        FCB_CUDA(cudaEventSynchronize(c->ev_gains)); // previous upload has left the pinned buffer
        c->gains_in_flight = false;
    }
    bool all_a = true, all_b = true;
    if (c->spec_valid && c->spec_len == out_len) { // computed while the previous block was on the GPU
        memcpy(c->h_gains, c->spec_gains.data(), out_len * sizeof(float2));
        c->crossfader = c->spec_after;
        all_a = c->spec_all_a;
        all_b = c->spec_all_b;
    } else {
        for (size_t i = 0; i < out_len; i++) {
            float2 g = c->crossfader.next_gains();
            c->h_gains[i] = g;
            all_a = all_a && g.x == 1.f && g.y == 0.f;
            all_b = all_b && g.x == 0.f && g.y == 1.f;
        }
    }
    c->spec_valid = false;

    // both convolvers always run, each on max_buffer_size samples (:72-73); A and B are independent,
    // so the one whose result is needed last may write straight into `out`
    auto run = [&](fcb_fftconv *f, float *dst, size_t dst_stride, const fcb_epilogue *epi) {
        return fcb_fftconv_process_dev(f, in, in_len, in_stride, dst, M, dst_stride, epi);
    };
    const bool whole = out_len == M && M > 0; // the mix can ride on the second convolver's K3
    if (c->a->current != c->b->current || c->a->active_seg_count != c->b->active_seg_count) c->rings_same = false;
    if (whole && in_len >= M && c->rings_same && fftconv_pair_ok(c->a, c->b, M)) {
        // A and B are clones fed the same blocks (:29, :72-73): one paired launch, the mix fused into B's epilogue
        if (all_a) return fftconv_process_pair_dev(c->a, c->b, in, in_stride, out, out_stride, nullptr, c->buffer_b, M, nullptr);
        if (all_b) return fftconv_process_pair_dev(c->a, c->b, in, in_stride, c->buffer_a, M, nullptr, out, out_stride, nullptr);
        for (size_t i = 0; i < out_len; i++) c->h_gains[i] = make_float2(c->h_gains[i].y, c->h_gains[i].x); // mine = B
        FCB_CUDA(cudaMemcpyAsync(c->d_gains, c->h_gains, out_len * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        FCB_CUDA(cudaEventRecord(c->ev_gains, c->stream));
        c->gains_in_flight = true;
        fcb_epilogue epi;
        memset(&epi, 0, sizeof epi);
        epi.mix_other = c->buffer_a;
        epi.mix_stride = M;
        epi.gains = reinterpret_cast<const float *>(c->d_gains);
        return fftconv_process_pair_dev(c->a, c->b, in, in_stride, c->buffer_a, M, nullptr, out, out_stride, &epi);
    }
    if (whole && all_a) {
        FCB_TRY(run(c->b, c->buffer_b, M, nullptr));
        return run(c->a, out, out_stride, nullptr);
    }
    if (whole && all_b) {
        FCB_TRY(run(c->a, c->buffer_a, M, nullptr));
        return run(c->b, out, out_stride, nullptr);
    }
    FCB_TRY(run(c->a, c->buffer_a, M, nullptr));
    if (whole && c->b->active_seg_count != 0) {
        // gain ramp fused into convolver B's K3: mine = B, other = A  =>  gains {gB, gA}
        for (size_t i = 0; i < out_len; i++) c->h_gains[i] = make_float2(c->h_gains[i].y, c->h_gains[i].x);
        FCB_CUDA(cudaMemcpyAsync(c->d_gains, c->h_gains, out_len * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        FCB_CUDA(cudaEventRecord(c->ev_gains, c->stream));
        c->gains_in_flight = true;
        fcb_epilogue epi;
        memset(&epi, 0, sizeof epi);
        epi.mix_other = c->buffer_a;
        epi.mix_stride = M;
        epi.gains = reinterpret_cast<const float *>(c->d_gains);
        return run(c->b, out, out_stride, &epi);
    }
    FCB_TRY(run(c->b, c->buffer_b, M, nullptr));
    if (out_len == 0) return FCB_OK;
    if (all_a || all_b) {
        FCB_CUDA(cudaMemcpy2DAsync(out, out_stride * sizeof(float), all_a ? c->buffer_a : c->buffer_b, M * sizeof(float),
                                   out_len * sizeof(float), c->C, cudaMemcpyDeviceToDevice, c->stream));
    } else {
        FCB_CUDA(cudaMemcpyAsync(c->d_gains, c->h_gains, out_len * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        FCB_CUDA(cudaEventRecord(c->ev_gains, c->stream));
        c->gains_in_flight = true;
        long long total = (long long)c->C * (long long)out_len;
        k_mix<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(out, (long long)out_stride, c->buffer_a, c->buffer_b,
                                                                     (long long)M, c->d_gains, (int)out_len,
                                                                     (long long)c->C);
        g_launches++;
        FCB_CUDA(cudaGetLastError());
    }
    return FCB_OK;
}

// the gains the next call of `out_len` samples will need, from a copy of the crossfader (see fcb_crossfade::spec_gains)
static void crossfade_speculate(fcb_crossfade *c, size_t out_len)
{
    c->spec_valid = false;
    // only while a fade runs (otherwise the gains are constants); a fade that ends inside the speculated block is fine:
    // the pending response (:67-70) is swapped in by the call AFTER the one that ends the fade, with or without this
    if (!g_xf_speculate || out_len == 0 || out_len > c->spec_gains.size() || !c->crossfader.approaching) return;
    Crossfader f = c->crossfader;
    bool all_a = true, all_b = true;
    for (size_t i = 0; i < out_len; i++) {
        const float2 g = f.next_gains();
        c->spec_gains[i] = g;
        all_a = all_a && g.x == 1.f && g.y == 0.f;
        all_b = all_b && g.x == 0.f && g.y == 1.f;
    }
    c->spec_after = f;
    c->spec_len = out_len;
    c->spec_all_a = all_a;
    c->spec_all_b = all_b;
    c->spec_valid = true;
}

extern "C" int fcb_crossfade_process(fcb_crossfade *c, const float *in, size_t in_len, size_t in_stride, float *out,
                                     size_t out_len, size_t out_stride)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    FCB_CUDA(cudaSetDevice(c->device));
    const size_t M = c->max_buffer_size;
    if (out_len > M) return fail(FCB_ERR_PANIC, "index out of bounds: the len is %zu but the index is %zu", M, M);
    // both convolvers consume input[..max_buffer_size] (:72-73): a longer input is never read past that, a shorter
    // one panics inside FFTConvolver::process — either way at most M samples are staged (buffer made in new())
    const size_t n_in = in_len < M ? in_len : M;
    if (g_zero_copy && M && out_len) {
        // the caller's own buffers are page-locked (fcb_host_alloc): the kernels read and write them in place
        float *din = pinned_alias(in), *dout = pinned_alias(out);
        if (din && dout) {
            FCB_TRY(fcb_crossfade_process_dev(c, din, n_in, in_stride, dout, out_len, out_stride));
            crossfade_speculate(c, out_len);
            FCB_CUDA(cudaStreamSynchronize(c->stream));
            c->gains_in_flight = false;
            return FCB_OK;
        }
    }
    if (c->m_in && g_mapped_io && M) { // small batch: the kernels read and write mapped pinned staging themselves
        for (size_t ch = 0; ch < c->C; ch++) memcpy(c->h_in + ch * M, in + ch * in_stride, n_in * sizeof(float));
        FCB_TRY(fcb_crossfade_process_dev(c, c->m_in, n_in, M, c->m_out, out_len, M));
        crossfade_speculate(c, out_len);
        FCB_CUDA(cudaStreamSynchronize(c->stream));
        c->gains_in_flight = false;
        for (size_t ch = 0; ch < c->C; ch++) memcpy(out + ch * out_stride, c->h_out + ch * M, out_len * sizeof(float));
        return FCB_OK;
    }
    if (n_in)
        FCB_CUDA(cudaMemcpy2DAsync(c->d_in, (M ? M : 1) * sizeof(float), in, in_stride * sizeof(float), n_in * sizeof(float),
                                   c->C, cudaMemcpyHostToDevice, c->stream));
    FCB_TRY(fcb_crossfade_process_dev(c, c->d_in, n_in, M ? M : 1, c->d_out, out_len, M ? M : 1));
    if (out_len)
        FCB_CUDA(cudaMemcpy2DAsync(out, out_stride * sizeof(float), c->d_out, (M ? M : 1) * sizeof(float),
                                   out_len * sizeof(float), c->C, cudaMemcpyDeviceToHost, c->stream));
    crossfade_speculate(c, out_len);
    FCB_CUDA(cudaStreamSynchronize(c->stream));
    c->gains_in_flight = false;
    return FCB_OK;
}

// :80-82 is todo!() in the reference.  EXTENSION (semantics written down in DESIGN.md §9 and
// pinned by the CPU restatement the tests check against): forget all audio like FFTConvolver::reset (src/fft_convolver.rs:296-306)
// — both convolvers reset, buffer_a / buffer_b zeroed — and finish a running fade at once: the crossfader lands where
// `mix` leaves it at counter == fading_samples (:261-273), so the response asked for last is the one heard.  A response
// still pending stays pending (:67-70 applies it at the next process()).  fcb_tune("strict_todo", 1): FCB_ERR_TODO.
extern "C" int fcb_crossfade_reset(fcb_crossfade *c)
{
    if (g_strict_todo) return fail(FCB_ERR_TODO, "not yet implemented");
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    FCB_CUDA(cudaSetDevice(c->device));
    FCB_TRY(fftconv_commit_wait(c->a)); // an update whose update() call has returned is part of the state that is kept
    FCB_TRY(fftconv_commit_wait(c->b));
    FCB_TRY(fcb_fftconv_reset(c->a));
    FCB_TRY(fcb_fftconv_reset(c->b));
    const size_t n = c->max_buffer_size ? c->max_buffer_size : 1;
    FCB_CUDA(cudaMemsetAsync(c->buffer_a, 0, c->C * n * sizeof(float), c->stream));
    FCB_CUDA(cudaMemsetAsync(c->buffer_b, 0, c->C * n * sizeof(float), c->stream));
    c->spec_valid = false;
    if (c->crossfader.approaching) {
        c->crossfader.approaching = false;
        c->crossfader.mix_value = c->crossfader.target == 0 ? 0.f : 1.f;
    }
    c->crossfader.counter = 0;
    c->rings_same = c->a->active_seg_count == c->b->active_seg_count; // both rings are all-zero again, current = 0
    return FCB_OK;
}

// #[derive(Clone)] (src/crossfade_convolver.rs:10): deep copy of both convolvers, the crossfader and the buffers
extern "C" int fcb_crossfade_clone(const fcb_crossfade *s, fcb_crossfade **out)
{
    if (!s || !out) return fail(FCB_ERR_ARG, "NULL argument");
    *out = nullptr;
    FCB_CUDA(cudaSetDevice(s->device));
    FCB_TRY(fftconv_commit_wait(s->a)); // updates already accepted are part of the value being cloned
    FCB_TRY(fftconv_commit_wait(s->b));
    FCB_CUDA(cudaStreamSynchronize(s->stream));
    fcb_fftconv *b = nullptr;
    FCB_TRY(fcb_fftconv_clone(s->b, &b)); // a private stream of its own when the source's was private
    fcb_crossfade *c = nullptr;
    int rc = fcb_crossfade_new(&c, b, s->max_response_length, s->max_buffer_size, s->crossfade_samples);
    if (rc != FCB_OK) {
        fcb_fftconv_free(b);
        return rc;
    }
    // new() made A a clone of B; A must be a clone of the source's A
    fcb_fftconv *a = nullptr;
    {
        fcb_fftconv tmp = *s->a;
        tmp.opt.stream = (void *)c->stream;
        rc = fcb_fftconv_clone(&tmp, &a);
    }
    if (rc == FCB_OK && a->eng) rc = fcb_engine_update_reserve(a->eng);
    if (rc != FCB_OK) {
        fcb_fftconv_free(a);
        fcb_crossfade_free(c);
        return rc;
    }
    fcb_fftconv_free(c->a);
    c->a = a;
    c->crossfader = s->crossfader;
    c->spec_valid = false;
    c->response_pending = s->response_pending;
    c->rings_same = s->rings_same;
    const size_t n = s->max_buffer_size ? s->max_buffer_size : 1, rows = s->b->opt.shared_ir ? 1 : s->C;
    auto cp = [&](void *d, const void *src, size_t bytes) {
        if (rc == FCB_OK && bytes && cudaMemcpyAsync(d, src, bytes, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess)
            rc = fail(FCB_ERR_CUDA, "crossfade clone copy failed");
    };
    cp(c->buffer_a, s->buffer_a, s->C * n * sizeof(float));
    cp(c->buffer_b, s->buffer_b, s->C * n * sizeof(float));
    cp(c->stored_response, s->stored_response, rows * (s->stored_len ? s->stored_len : 0) * sizeof(float));
    if (rc == FCB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(FCB_ERR_CUDA, "sync failed");
    if (rc != FCB_OK) {
        fcb_crossfade_free(c);
        return rc;
    }
    *out = c;
    return FCB_OK;
}

extern "C" int fcb_crossfade_sync(fcb_crossfade *c)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    FCB_CUDA(cudaSetDevice(c->device));
    FCB_CUDA(cudaStreamSynchronize(c->stream));
    return FCB_OK;
}

extern "C" int fcb_crossfade_state(const fcb_crossfade *c, int64_t *counter, float *mix_value, int *approaching, int *target)
{
    if (!c) return fail(FCB_ERR_ARG, "NULL convolver");
    if (counter) *counter = c->crossfader.counter;
    if (mix_value) *mix_value = c->crossfader.mix_value;
    if (approaching) *approaching = c->crossfader.approaching ? 1 : 0;
    if (target) *target = c->crossfader.target;
    return FCB_OK;
}
