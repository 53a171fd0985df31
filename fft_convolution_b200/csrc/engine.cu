// engine.cu — layer 1 of the C ABI: device state of C lock-step FFTConvolver channels and the
// stage launches K1 (forward FFT), K2 (delay-line MAC), K3 (inverse FFT + overlap-add +
// epilogues), K5 (IR preparation).  The caller (Rust host or the C++ mirror in host_mirror.cu)
// owns the scheduler scalars `current`, `fill`, `active` (src/fft_convolver.rs:99-101, :91).
#include <cmath>
#include <cstring>
#include <map>
#include <utility>
#include <mutex>
#include <vector>

#include "engine_internal.cuh"
#include "offline_kernels.cuh"

namespace fcb {
thread_local std::string g_last_error = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_resource_calls{0};

// ---- twiddle tables: tw[t] = exp(-2 pi i t / N), t < N, rounded from f64; one per (device, N)
static std::mutex g_tw_mutex;
static std::map<std::pair<int, size_t>, float2 *> g_tw;

int get_twiddles(int device, size_t N, const float2 **out)
{
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    auto key = std::make_pair(device, N);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) {
        *out = it->second;
        return FCB_OK;
    }
    auto root = [N](size_t t) {
        double a = -2.0 * M_PI * (double)t / (double)N;
        return make_float2((float)cos(a), (float)sin(a));
    };
    std::vector<float2> h(N);
    for (size_t t = 0; t < N; t++) h[t] = root(t);
    // pass-ordered copies behind the base table: for Stockham pass p (stride NS, radix R) the factor of
    // (r, k) sits at [off_p + (r-1)*NS + k] — the threads of a warp (consecutive k) then read consecutive
    // entries instead of entries r*k*STEP apart.  Same values as h[r*k*STEP] (pass_tw_offset in fft_kernels.cuh).
    const int logb = ilog2(N / 2);
    size_t ns = 1;
    for (int p = 0; radix_at(logb, p) > 0; p++) {
        const size_t R = (size_t)radix_at(logb, p), step = N / (ns * R);
        if (ns > 1)
            for (size_t r = 1; r < R; r++)
                for (size_t k = 0; k < ns; k++) h.push_back(root(r * k * step));
        ns *= R;
    }
    float2 *d = nullptr;
    FCB_CUDA(cudaMalloc(&d, h.size() * sizeof(float2)));
    FCB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
    g_tw[key] = d;
    *out = d;
    return FCB_OK;
}
} // namespace fcb

using namespace fcb;

struct fcb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    size_t C = 0, B = 0, S = 0, L = 0;
    int logb = 0;
    bool shared_ir = false;
    float2 *ir = nullptr, *ring = nullptr, *premul = nullptr;
    float *overlap = nullptr, *inbuf = nullptr;
    float *scratch = nullptr; // [C][B] output staging for layer-1 callers (fcb_engine_scratch)
    float *stage = nullptr; // IR upload staging
    size_t stage_floats = 0;
    const float2 *tw = nullptr;
    // background IR update (fcb_engine_update_reserve): K5 writes `ir_shadow` on `upd_stream` while the blocks keep
    // reading `ir`; fcb_engine_update_commit swaps the two pointers
    // where the caller's input block lives, remembered per pointer (a driver query per launch would cost the small batches)
    mutable const void *io_probe_ptr = nullptr;
    mutable bool io_probe_host = false;
    // small batches: partial sums and arrival counters of the split whole-block kernels (fused_kernel.cuh SplitArgs)
    float4 *zpart = nullptr;
    unsigned int *zcount = nullptr;
    size_t zslices = 0; // capacity of zpart in (group, slice) entries of 2 x 256 float4
    float2 *ir_shadow = nullptr;
    cudaStream_t upd_stream = nullptr;
    cudaEvent_t upd_done = nullptr, upd_fence = nullptr;
    // multi-block calls (offline_kernels.cuh): workspace for up to mb_cap blocks per pass, allocated on first use
    size_t mb_cap = 0;
    float2 *mb_xnew = nullptr, *mb_premul = nullptr; // [C][mb_cap][B]
    float *mb_y = nullptr;                           // [C][mb_cap][2B]
    float *mb_in = nullptr, *mb_out = nullptr;       // [C][mb_cap*B] staging for host buffers
    // end-to-end pipeline (fcb_engine_process_block_host): channel groups round-robin over streams
    static constexpr int NPIPE = 4;
    cudaStream_t pipe[NPIPE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipe_done[NPIPE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipe_start = nullptr;
    std::vector<cudaEvent_t> pipe_in, pipe_out; // per group: input landed / output computed
    std::vector<size_t> pipe_cut;               // group boundaries, rebuilt only when the group size changes
    size_t pipe_cut_G = 0;

    long long ir_stride() const { return shared_ir ? 0 : (long long)(S * B); }
    long long ring_stride() const { return (long long)(S * B); }
    size_t ir_channels() const { return shared_ir ? 1 : C; }
};

// ---- launch helpers ---------------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember it per (kernel, device)
struct SmemOptIn {
    bool done[64] = {};
    template <typename K>
    int ensure(K kernel, size_t bytes)
    {
        int dev = 0;
        FCB_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !done[dev]) {
            FCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
        return FCB_OK;
    }
};

#define FCB_DISPATCH_LOGB(logb, CALL)                                                       \
    switch (logb) {                                                                         \
    case 0: { constexpr int LB = 0; CALL; } break;                                          \
    case 1: { constexpr int LB = 1; CALL; } break;                                          \
    case 2: { constexpr int LB = 2; CALL; } break;                                          \
    case 3: { constexpr int LB = 3; CALL; } break;                                          \
    case 4: { constexpr int LB = 4; CALL; } break;                                          \
    case 5: { constexpr int LB = 5; CALL; } break;                                          \
    case 6: { constexpr int LB = 6; CALL; } break;                                          \
    case 7: { constexpr int LB = 7; CALL; } break;                                          \
    case 8: { constexpr int LB = 8; CALL; } break;                                          \
    case 9: { constexpr int LB = 9; CALL; } break;                                          \
    case 10: { constexpr int LB = 10; CALL; } break;                                        \
    case 11: { constexpr int LB = 11; CALL; } break;                                        \
    case 12: { constexpr int LB = 12; CALL; } break;                                        \
    case 13: { constexpr int LB = 13; CALL; } break;                                        \
    case 14: { constexpr int LB = 14; CALL; } break;                                        \
    default: return fail(FCB_ERR_UNSUPPORTED, "block size 2^%d not supported (max 16384)", logb); \
    }

template <int LOGB>
static int launch_forward_t(const float2 *tw, cudaStream_t st, const float *src, long long src_stride, int len,
                            float2 *dst, long long dst_stride, int nseg, long long ntransforms)
{
    using P = FftPlan<LOGB>;
    static SmemOptIn optin;
    if (P::SMEM_BYTES > 48 * 1024) FCB_TRY(optin.ensure(k_rfft_forward<LOGB>, P::SMEM_BYTES));
    if (ntransforms <= 0) return FCB_OK;
    long long grid = (ntransforms + P::TPB - 1) / P::TPB;
    k_rfft_forward<LOGB><<<(unsigned)grid, P::CTA, P::SMEM_BYTES, st>>>(src, src_stride, len, dst, dst_stride, nseg,
                                                                        ntransforms, tw);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}

template <int LOGB>
static int launch_inverse_t(const float2 *tw, cudaStream_t st, const IfftArgs &a)
{
    using P = FftPlan<LOGB>;
    static SmemOptIn optin;
    if (P::SMEM_BYTES > 48 * 1024) FCB_TRY(optin.ensure(k_irfft_ola<LOGB>, P::SMEM_BYTES));
    long long grid = (a.nchan + P::TPB - 1) / P::TPB;
    k_irfft_ola<LOGB><<<(unsigned)grid, P::CTA, P::SMEM_BYTES, st>>>(a, tw);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}

// ---- live per-launch timing of K2 (bench.py's roofline.achieved): CUDA events recorded on the
// launching stream around every K2 launch while profiling is on, read back after a sync
struct MacProfile {
    std::mutex mu;
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool, used;
};
static MacProfile g_prof;

static cudaEvent_t prof_before(cudaStream_t s, cudaEvent_t *stop)
{
    std::lock_guard<std::mutex> lock(g_prof.mu);
    if (!g_prof.on) return nullptr;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!g_prof.pool.empty()) {
        ev = g_prof.pool.back();
        g_prof.pool.pop_back();
    } else if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) {
        return nullptr;
    }
    g_prof.used.push_back(ev);
    cudaEventRecord(ev.first, s);
    *stop = ev.second;
    return ev.first;
}

namespace fcb {
cudaEvent_t mac_profile_begin(cudaStream_t s, cudaEvent_t *stop) { return prof_before(s, stop); }
} // namespace fcb

extern "C" int fcb_profile_mac(int enable)
{
    std::lock_guard<std::mutex> lock(g_prof.mu);
    g_prof.on = enable != 0;
    for (auto &ev : g_prof.used) g_prof.pool.push_back(ev);
    g_prof.used.clear();
    return FCB_OK;
}

extern "C" int fcb_profile_mac_read(double *total_ms, uint64_t *launches)
{
    std::lock_guard<std::mutex> lock(g_prof.mu);
    double tot = 0.0;
    for (auto &ev : g_prof.used) {
        FCB_CUDA(cudaEventSynchronize(ev.second));
        float ms = 0.f;
        FCB_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = g_prof.used.size();
    return FCB_OK;
}

// ---- tuning knobs (fcb_tune); the defaults are the measured best (profiles/r01_*sweep*) -----------
static std::atomic<int> g_mac_impl{0};         // K2: 0 = auto (TMA pipeline for B >= 32), 1 = LDG kernel, 2 = TMA
static std::atomic<int> g_mac_stages{3};       // K2 pipeline stages: 2, 3, 4 or 6
static std::atomic<int> g_pipe_group{512};     // channels per group of the copy pipeline (pageable host buffers)
static std::atomic<bool> g_fused_block{true};  // whole blocks with B in 32..512: one fused K1+K2+K3 kernel
static std::atomic<int> g_fused_short{40};     // delay lines up to this many segments use 2-row stages (0 = off)
static std::atomic<int> g_fused_stages{2};     // fused kernel: 2 stages (64 KB) -> 3 CTAs/SM; 3 -> 2 CTAs/SM
static std::atomic<bool> g_tma_io{true};       // fused kernel moves its input/output blocks with bulk copies
static std::atomic<bool> g_shared_reuse{true}; // shared-IR engines: stage each IR tile once per CTA
namespace fcb { std::atomic<bool> g_mimo_tile{true}; } // matrix K2 with in-CTA reuse (0: per-channel K2)
namespace fcb { std::atomic<int> g_mimo_tc{2}; }
namespace fcb { std::atomic<int> g_mimo_tc_min{33}; }  // streams from which mimo_tc == 2 picks the tensor cores
namespace fcb { std::atomic<int> g_mimo_rt_wb{1}; }
namespace fcb { std::atomic<int> g_mimo_rt_r{4}; }
namespace fcb { std::atomic<int> g_mimo_rt_min{1}; }
namespace fcb { std::atomic<int> g_mimo_rt_waves{0}; }
namespace fcb { std::atomic<bool> g_mimo_rt{true}; }   // register-tiled matrix MAC (k_mac_rt) for 2+ streams below that
static std::atomic<bool> g_split{true};        // small batches: cut the delay line of a whole block over several CTAs
static std::atomic<bool> g_multi_block{true}; // calls spanning >= 2 whole blocks run as one time-batched pass       // K4 tensor-core matrix MAC: 0 never, 1 when the shape fits, 2 + NS >= 16

template <int B, int NST>
static int launch_mac_bulk(const MacArgs &a, cudaStream_t st)
{
    using Cfg = MacBulkCfg<B>;
    static SmemOptIn optin;
    FCB_TRY(optin.ensure(k_mac_bulk<B, NST>, Cfg::smem_bytes(NST)));
    long long groups = (a.nchan + Cfg::CPB - 1) / Cfg::CPB;
    k_mac_bulk<B, NST><<<(unsigned)(groups * Cfg::TILES), 256, Cfg::smem_bytes(NST), st>>>(a);
    return FCB_OK;
}

template <int LOGB>
static int launch_mac_t(const MacArgs &a, cudaStream_t st)
{
    constexpr int B = 1 << LOGB;
    if (a.seg_hi <= a.seg_lo) { // no segment to accumulate: pre_multiplied = 0 (src/fft_convolver.rs:245)
        FCB_CUDA(cudaMemsetAsync(a.premul, 0, (size_t)a.nchan * B * sizeof(float2), st));
        return FCB_OK;
    }
    const int impl = g_mac_impl.load();
    cudaEvent_t prof_stop = nullptr;
    const bool profiled = prof_before(st, &prof_stop) != nullptr;
    if constexpr (B == 1) {
        unsigned grid = (unsigned)((a.nchan + 255) / 256);
        k_mac_b1<<<grid, 256, 0, st>>>(a);
    } else {
        bool bulk = impl == 2 || (impl == 0 && B >= 32);
        if constexpr (B < 4) bulk = false;
        if (bulk) {
            if constexpr (B >= 4) {
                switch (g_mac_stages.load()) {
                case 2: FCB_TRY((launch_mac_bulk<B, 2>(a, st))); break;
                case 4: FCB_TRY((launch_mac_bulk<B, 4>(a, st))); break;
                case 6: FCB_TRY((launch_mac_bulk<B, 6>(a, st))); break;
                default: FCB_TRY((launch_mac_bulk<B, 3>(a, st))); break;
                }
            }
        } else {
            constexpr int ROW4 = B / 2;
            constexpr int TX = ROW4 < 256 ? ROW4 : 256;
            constexpr int TILES = ROW4 / TX;
            constexpr int CPB = 256 / TX;
            long long groups = (a.nchan + CPB - 1) / CPB;
            k_mac_v4<B, 8><<<(unsigned)(groups * TILES), 256, 0, st>>>(a);
        }
    }
    if (profiled) cudaEventRecord(prof_stop, st);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}

template <int LOGB>
static int launch_block_fused_shared(const fcb_engine *e, cudaStream_t st, FusedArgs fa, size_t nc);

// Small batches: how many CTAs share one delay line (SplitArgs).  Measured on the B200 (scripts/r02_split_sweep.py,
// 2 s responses, block 512): with fewer channel groups than SMs the split wins by 2x and more (mono: 48 -> 20 us per block);
// at 256 groups it LOSES (0.065 -> 0.076 ms: a full first wave already streams, the partial sums are pure overhead); at 512
// groups — one wave that starts and ends in step, so every CTA is in its FFT phase at the same time — three slices per line
// de-synchronise the phases and win 9 % (0.137 -> 0.125 ms).  Hence: groups <= 148 fill the 4 x 148 CTA slots in whole
// waves; 296 < groups <= 592 take three slices; everything else runs unsplit.  At least two pipeline stages per CTA,
// slices are whole stages.  Returns zsplit = 1 when the split does not apply.
static constexpr size_t kSplitTargetCtas = 4 * 148; // CTA slots of the whole-block kernels (4 per SM at 54 KB each)
static std::atomic<int> g_split_slots{(int)kSplitTargetCtas}; // fcb_tune("split_slots"): CTAs the split aims at (sweeps)
static std::atomic<int> g_split_min_stages{2};                 // fcb_tune("split_min_stages"): pipeline stages per CTA at least
static SplitArgs split_plan(const fcb_engine *e, size_t groups, int seg_lo, int seg_hi, int rows)
{
    SplitArgs sp{};
    sp.zsplit = 1;
    const int nseg = seg_hi - seg_lo;
    const size_t slots = (size_t)g_split_slots.load();
    if (!g_split.load() || !e->zpart || groups == 0 || nseg < 4 * rows) return sp;
    size_t z = 1;
    if (groups <= slots / 4) z = slots / groups; // whole waves only: never a nearly empty second round
    else if (groups > slots / 2 && groups <= slots) z = 3;
    else return sp;
    const size_t zmax = (size_t)nseg / ((size_t)g_split_min_stages.load() * (size_t)rows);
    if (z > zmax) z = zmax;
    if (z < 2) return sp;
    int zlen = (int)((nseg + z - 1) / z);
    zlen = (zlen + rows - 1) / rows * rows;
    z = (size_t)(nseg + zlen - 1) / zlen;
    if (z < 2 || groups * z > e->zslices) return sp;
    sp.zsplit = (int)z;
    sp.zlen = zlen;
    sp.part = e->zpart;
    sp.count = e->zcount;
    return sp;
}

// fcb_tune("k1_late", 1): K1 after the MAC stream when the input block sits in host memory, so that the PCIe pull issued at
// the top of the kernel has the whole stream to land.  Measured on 8 GPUs (profiles/r02_e2e_8gpu_k1_late.jsonl): no
// effect on the end-to-end step (0.98-1.02 ms either way) — the pull is not what the 8-GPU host path pays for.  Off.
static std::atomic<bool> g_k1_late{false};
static bool input_is_host_memory(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// block I/O eligible for one bulk copy per channel: 16-byte aligned rows
static bool tma_io_ok(const float *in, size_t in_stride, const float *out, size_t out_stride, size_t B)
{
    return B >= 4 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) && in_stride % 4 == 0 && out_stride % 4 == 0;
}

// whole block for channels [c0, c0+nc): fused K1+K2+K3 (B in 32..512)
template <int LOGB, int NST, int ROWS = 4>
static int launch_block_fused(const fcb_engine *e, cudaStream_t st, size_t c0, size_t nc, const float *in_dev,
                              size_t in_stride, float *out_dev, size_t out_stride, size_t current, size_t active,
                              const fcb_epilogue *epi)
{
    using Cfg = FusedCfg<LOGB, ROWS>;
    static SmemOptIn optin;
    FCB_TRY(optin.ensure(k_block_fused<LOGB, NST, ROWS>, Cfg::smem_bytes(NST)));
    const size_t B = e->B;
    FusedArgs fa{};
    fa.in = in_dev + c0 * in_stride;
    fa.in_stride = (long long)in_stride;
    fa.mac.ir = e->ir + (e->shared_ir ? 0 : c0) * (long long)(e->S * B);
    fa.mac.ir_stride = e->ir_stride();
    fa.mac.ring = e->ring + c0 * e->ring_stride();
    fa.mac.ring_stride = e->ring_stride();
    fa.mac.current = (int)current;
    fa.mac.active = (int)active;
    fa.mac.nchan = (long long)nc;
    fa.mac.seg_lo = 1;
    fa.mac.seg_hi = (int)active;
    fa.ifft.overlap = e->overlap + c0 * B;
    fa.ifft.out = out_dev + c0 * out_stride;
    fa.ifft.out_stride = (long long)out_stride;
    if (epi) {
        fa.ifft.epi = *epi;
        if (fa.ifft.epi.add0) fa.ifft.epi.add0 += c0 * epi->add_stride;
        if (fa.ifft.epi.add1) fa.ifft.epi.add1 += c0 * epi->add_stride;
        if (fa.ifft.epi.mix_other) fa.ifft.epi.mix_other += c0 * epi->mix_stride;
    }
    const size_t groups = (nc + Cfg::CPB - 1) / Cfg::CPB;
    fa.split.zsplit = 1;
    if (c0 == 0 && nc == e->C) fa.split = split_plan(e, groups, fa.mac.seg_lo, fa.mac.seg_hi, ROWS);
    if (e->shared_ir && g_shared_reuse.load() && ROWS == 4 && fa.split.zsplit == 1) {
        if constexpr (LOGB >= 7) return launch_block_fused_shared<LOGB>(e, st, fa, nc);
    }
    cudaEvent_t prof_stop = nullptr;
    const bool profiled = prof_before(st, &prof_stop) != nullptr;
    const unsigned grid = (unsigned)(groups * (size_t)fa.split.zsplit);
    if (g_tma_io.load() && tma_io_ok(fa.in, in_stride, fa.ifft.out, out_stride, e->B)) {
        static SmemOptIn optin_io;
        if (e->io_probe_ptr != (const void *)in_dev) {
            e->io_probe_ptr = (const void *)in_dev;
            e->io_probe_host = input_is_host_memory(in_dev);
        }
        fa.k1_late = g_k1_late.load() && e->io_probe_host ? 1 : 0;
        FCB_TRY(optin_io.ensure(k_block_fused<LOGB, NST, ROWS, true>, Cfg::smem_bytes(NST, true)));
        k_block_fused<LOGB, NST, ROWS, true><<<grid, 256, Cfg::smem_bytes(NST, true), st>>>(fa, e->tw);
    } else {
        k_block_fused<LOGB, NST, ROWS><<<grid, 256, Cfg::smem_bytes(NST), st>>>(fa, e->tw);
    }
    if (profiled) cudaEventRecord(prof_stop, st);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}

// shared-IR engines: G = 2 channels per thread group, one IR tile per CTA
template <int LOGB>
static int launch_block_fused_shared(const fcb_engine *e, cudaStream_t st, FusedArgs fa, size_t nc)
{
    constexpr int G = 2;
    using Cfg = FusedSharedCfg<LOGB, G>;
    static SmemOptIn optin;
    FCB_TRY(optin.ensure(k_block_fused_shared<LOGB, G>, Cfg::SMEM_BYTES));
    cudaEvent_t prof_stop = nullptr;
    const bool profiled = prof_before(st, &prof_stop) != nullptr;
    const unsigned grid = (unsigned)((nc + Cfg::NSLOT - 1) / Cfg::NSLOT);
    k_block_fused_shared<LOGB, G><<<grid, 256, Cfg::SMEM_BYTES, st>>>(fa, e->tw);
    if (profiled) cudaEventRecord(prof_stop, st);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}

// two convolvers fed the same input (identical rings): one launch, one forward FFT, one ring stream
static std::atomic<bool> g_fused_pair{true};
static int check_sched(const fcb_engine *e, size_t current, size_t active, const char *who);

template <int LOGB, int ROWS>
static int launch_block_fused_pair(const fcb_engine *ea, const fcb_engine *eb, FusedPairArgs fa)
{
    using Cfg = FusedPairCfg<LOGB, ROWS>;
    static SmemOptIn optin;
    FCB_TRY(optin.ensure(k_block_fused_pair<LOGB, ROWS>, Cfg::SMEM_BYTES));
    const size_t groups = (ea->C + Cfg::CPB - 1) / Cfg::CPB;
    fa.split = split_plan(ea, groups, fa.mac.seg_lo, fa.mac.seg_hi, ROWS);
    cudaEvent_t prof_stop = nullptr;
    const bool profiled = prof_before(ea->stream, &prof_stop) != nullptr;
    const unsigned grid = (unsigned)(groups * (size_t)fa.split.zsplit);
    // block I/O as bulk copies whenever every row is 16-byte aligned (device buffers or mapped pinned host staging); a mix
    // that reads another buffer than this launch's out_a keeps the plain-store kernel
    const bool mix_ok = !fa.ifft_b.epi.mix_other || (fa.ifft_b.epi.mix_other == fa.ifft_a.out && (long long)fa.ifft_b.epi.mix_stride == fa.ifft_a.out_stride);
    if (g_tma_io.load() && mix_ok && !fa.ifft_a.epi.mix_other && tma_io_ok(fa.in, (size_t)fa.in_stride, fa.ifft_a.out, (size_t)fa.ifft_a.out_stride, ea->B) &&
        tma_io_ok(fa.in, (size_t)fa.in_stride, fa.ifft_b.out, (size_t)fa.ifft_b.out_stride, ea->B)) {
        static SmemOptIn optin_io;
        FCB_TRY(optin_io.ensure(k_block_fused_pair<LOGB, ROWS, true>, Cfg::SMEM_BYTES));
        fa.mix_from_a = fa.ifft_b.epi.mix_other ? 1 : 0;
        k_block_fused_pair<LOGB, ROWS, true><<<grid, 256, Cfg::SMEM_BYTES, ea->stream>>>(fa, ea->tw);
    } else {
        k_block_fused_pair<LOGB, ROWS><<<grid, 256, Cfg::SMEM_BYTES, ea->stream>>>(fa, ea->tw);
    }
    if (profiled) cudaEventRecord(prof_stop, ea->stream);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    (void)eb;
    return FCB_OK;
}

extern "C" int fcb_engine_pair_ok(const fcb_engine *ea, const fcb_engine *eb, size_t active)
{
    return ea && eb && ea != eb && g_fused_pair.load() && g_fused_block.load() && ea->logb >= 5 && ea->logb <= 9 &&
           ea->device == eb->device && ea->stream == eb->stream && ea->C == eb->C && ea->B == eb->B && ea->S == eb->S &&
           !ea->shared_ir && !eb->shared_ir && active >= 1 && active <= ea->S;
}

extern "C" int fcb_engine_process_block_pair_dev(fcb_engine *ea, fcb_engine *eb, const float *in_dev, size_t in_stride,
                                                 float *out_a, size_t stride_a, const fcb_epilogue *epi_a, float *out_b,
                                                 size_t stride_b, const fcb_epilogue *epi_b, size_t current, size_t active)
{
    return fcb_engine_process_block_pair_copy_dev(ea, eb, in_dev, in_stride, out_a, stride_a, epi_a, out_b, stride_b, epi_b,
                                                  current, active, nullptr, 0);
}

// the same, and the input block is also stored to copy_to (channel stride copy_stride, rows 8-byte aligned): TwoStage's
// append to tail_input (src/fft_convolver.rs:459-461) rides on the launch instead of costing a copy of its own
extern "C" int fcb_engine_process_block_pair_copy_dev(fcb_engine *ea, fcb_engine *eb, const float *in_dev, size_t in_stride,
                                                      float *out_a, size_t stride_a, const fcb_epilogue *epi_a, float *out_b,
                                                      size_t stride_b, const fcb_epilogue *epi_b, size_t current,
                                                      size_t active, float *copy_to, size_t copy_stride)
{
    FCB_TRY(check_sched(ea, current, active, "process_block_pair"));
    if (!fcb_engine_pair_ok(ea, eb, active)) return fail(FCB_ERR_UNSUPPORTED, "process_block_pair: engines do not pair");
    if (!in_dev || !out_a || !out_b) return fail(FCB_ERR_ARG, "process_block_pair: NULL argument");
    FCB_CUDA(cudaSetDevice(ea->device));
    const size_t B = ea->B;
    FusedPairArgs fa{};
    fa.in = in_dev;
    fa.in_stride = (long long)in_stride;
    fa.mac.ir = ea->ir;
    fa.mac.ir_stride = ea->ir_stride();
    fa.mac.ring = ea->ring;
    fa.mac.ring_stride = ea->ring_stride();
    fa.mac.current = (int)current;
    fa.mac.active = (int)active;
    fa.mac.nchan = (long long)ea->C;
    fa.mac.seg_lo = 1;
    fa.mac.seg_hi = (int)active;
    fa.ir_b = eb->ir;
    fa.ir_b_stride = eb->ir_stride();
    fa.ring_b = eb->ring;
    fa.ifft_a.overlap = ea->overlap;
    fa.ifft_a.out = out_a;
    fa.ifft_a.out_stride = (long long)stride_a;
    if (epi_a) fa.ifft_a.epi = *epi_a;
    fa.ifft_b.overlap = eb->overlap;
    fa.ifft_b.out = out_b;
    fa.ifft_b.out_stride = (long long)stride_b;
    if (epi_b) fa.ifft_b.epi = *epi_b;
    if (copy_to && ((uintptr_t)copy_to % 8 || copy_stride % 2)) return fail(FCB_ERR_ARG, "process_block_pair: copy_to rows must be 8-byte aligned");
    fa.copy_in = copy_to;
    fa.copy_stride = (long long)copy_stride;
    (void)B;
    const bool short_line = active <= (size_t)g_fused_short.load();
#define FCB_PAIR_CASE(LB)                                                        \
    case LB:                                                                     \
        return short_line ? launch_block_fused_pair<LB, 2>(ea, eb, fa) : launch_block_fused_pair<LB, 4>(ea, eb, fa);
    switch (ea->logb) {
        FCB_PAIR_CASE(5) FCB_PAIR_CASE(6) FCB_PAIR_CASE(7) FCB_PAIR_CASE(8) FCB_PAIR_CASE(9)
    default: return fail(FCB_ERR_UNSUPPORTED, "process_block_pair: block size not covered");
    }
#undef FCB_PAIR_CASE
}

static bool fused_applicable(const fcb_engine *e, size_t active)
{
    return g_fused_block.load() && e->logb >= 5 && e->logb <= 9 && active >= 1;
}

static int run_block_fused(const fcb_engine *e, cudaStream_t st, size_t c0, size_t nc, const float *in_dev,
                           size_t in_stride, float *out_dev, size_t out_stride, size_t current, size_t active,
                           const fcb_epilogue *epi)
{
    const bool two = g_fused_stages.load() == 2; // 2 stages -> 3 CTAs/SM, 3 stages -> 2 CTAs/SM
    // short delay lines (two-stage head / tail0: 16 segments): the FFT latency is no longer hidden by the
    // stream, so run 2-row stages (32 KB of stages per CTA) and let a fourth CTA per SM cover it
    const bool short_line = active <= (size_t)g_fused_short.load();
#define FCB_FUSED_CASE(LB)                                                                                       \
    case LB:                                                                                                     \
        if (short_line)                                                                                          \
            return launch_block_fused<LB, 2, 2>(e, st, c0, nc, in_dev, in_stride, out_dev, out_stride, current, active, epi); \
        return two ? launch_block_fused<LB, 2>(e, st, c0, nc, in_dev, in_stride, out_dev, out_stride, current, active, epi) \
                   : launch_block_fused<LB, 3>(e, st, c0, nc, in_dev, in_stride, out_dev, out_stride, current, active, epi);
    switch (e->logb) {
        FCB_FUSED_CASE(5) FCB_FUSED_CASE(6) FCB_FUSED_CASE(7) FCB_FUSED_CASE(8) FCB_FUSED_CASE(9)
    default: return fail(FCB_ERR_UNSUPPORTED, "fused block kernel: block size not covered");
    }
#undef FCB_FUSED_CASE
}

template <int LOGB>
static int launch_forward(const fcb_engine *e, const float *src, long long src_stride, int len, float2 *dst,
                          long long dst_stride, int nseg, long long ntransforms, cudaStream_t st = nullptr)
{
    return launch_forward_t<LOGB>(e->tw, st ? st : e->stream, src, src_stride, len, dst, dst_stride, nseg, ntransforms);
}
template <int LOGB>
static int launch_inverse(const fcb_engine *e, const IfftArgs &a, cudaStream_t st = nullptr)
{
    return launch_inverse_t<LOGB>(e->tw, st ? st : e->stream, a);
}
template <int LOGB>
static int launch_mac(const fcb_engine *e, const MacArgs &a, cudaStream_t st = nullptr)
{
    return launch_mac_t<LOGB>(a, st ? st : e->stream);
}

namespace fcb {
int run_forward(int logb, const float2 *tw, cudaStream_t st, const float *src, long long src_stride, int len,
                float2 *dst, long long dst_stride, int nseg, long long ntransforms)
{
    FCB_DISPATCH_LOGB(logb, FCB_TRY(launch_forward_t<LB>(tw, st, src, src_stride, len, dst, dst_stride, nseg, ntransforms)));
    return FCB_OK;
}
int run_mac(int logb, cudaStream_t st, const MacArgs &a)
{
    FCB_DISPATCH_LOGB(logb, FCB_TRY(launch_mac_t<LB>(a, st)));
    return FCB_OK;
}
int run_inverse(int logb, const float2 *tw, cudaStream_t st, const IfftArgs &a)
{
    FCB_DISPATCH_LOGB(logb, FCB_TRY(launch_inverse_t<LB>(tw, st, a)));
    return FCB_OK;
}
} // namespace fcb

template <int B, int OT, int ST>
static int launch_mac_tile(const MacTileArgs &a, cudaStream_t st)
{
    using Cfg = MacTileCfg<B, OT, ST>;
    static SmemOptIn optin;
    FCB_TRY(optin.ensure(k_mac_tile<B, OT, ST>, Cfg::SMEM_BYTES));
    const long long OG = (a.n_out + OT - 1) / OT, SG = (a.n_streams + ST - 1) / ST;
    const long long grid = (long long)Cfg::TILES * a.zchunks * OG * a.n_in * SG;
    cudaEvent_t prof_stop = nullptr;
    const bool profiled = prof_before(st, &prof_stop) != nullptr;
    k_mac_tile<B, OT, ST><<<(unsigned)grid, Cfg::TX, Cfg::SMEM_BYTES, st>>>(a);
    if (profiled) cudaEventRecord(prof_stop, st);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}

namespace fcb {
// one stream: 8 outputs share each ring tile; several streams: 4 outputs x 4 streams per CTA
static inline bool tile_wide_streams(int n_streams) { return n_streams >= 2; }

int mac_tile_plan(int logb, int n_in, int n_out, int n_streams, int nsegs, int *zchunks, int *zlen)
{
    if (logb < 6) return FCB_ERR_UNSUPPORTED; // B < 64: the generic K2 handles it
    const int B = 1 << logb, tiles = B < 512 ? 1 : B / 512;
    const int OT = tile_wide_streams(n_streams) ? 4 : 8, ST = tile_wide_streams(n_streams) ? 4 : 1;
    const long long base = (long long)tiles * ((n_out + OT - 1) / OT) * n_in * ((n_streams + ST - 1) / ST);
    long long z = (4 * 148 + base - 1) / base; // aim at >= 4 CTAs per SM worth of work items
    const long long zmax = nsegs > 8 ? nsegs / 8 : 1;
    if (z > zmax) z = zmax;
    if (z < 1) z = 1;
    // one CTA is resident per SM (its stages fill the shared memory): pick the chunk count near z whose CTA count
    // fills whole waves of 148 best — 19 chunks of the 16 x 16 matrix are 608 CTAs = 4.1 waves (a fifth, nearly
    // empty round), 18 are 576 = 3.9
    {
        double best = 0.0;
        long long best_z = z;
        for (long long c = z > 2 ? z - z / 3 : 1; c <= z + z / 3 && c <= zmax; c++) {
            const double waves = (double)(base * c) / 148.0, eff = waves / (double)(long long)(waves + 0.999999);
            if (eff > best + 1e-9) {
                best = eff;
                best_z = c;
            }
        }
        z = best_z;
    }
    const int len = (int)((nsegs + z - 1) / z);
    *zlen = len > 0 ? len : 1;
    *zchunks = nsegs > 0 ? (nsegs + *zlen - 1) / *zlen : 1;
    return FCB_OK;
}

} // namespace fcb
// test hook (host only): the segment chunking the matrix kernel would use for this problem
extern "C" int fcb_debug_mac_tile_plan(int logb, int n_in, int n_out, int n_streams, int nsegs, int *zchunks, int *zlen)
{
    return fcb::mac_tile_plan(logb, n_in, n_out, n_streams, nsegs, zchunks, zlen);
}
namespace fcb {

int run_mac_tile(int logb, cudaStream_t st, MacTileArgs a, int *zchunks_out)
{
    const int nsegs = a.seg_hi - a.seg_lo;
    FCB_TRY(mac_tile_plan(logb, a.n_in, a.n_out, a.n_streams, nsegs > 0 ? nsegs : 0, &a.zchunks, &a.zlen));
    if (zchunks_out) *zchunks_out = a.zchunks;
    const bool wide = tile_wide_streams(a.n_streams);
#define FCB_TILE_CASE(LB)                                                                  \
    case LB:                                                                               \
        return wide ? launch_mac_tile<(1 << LB), 4, 4>(a, st) : launch_mac_tile<(1 << LB), 8, 1>(a, st);
    switch (logb) {
        FCB_TILE_CASE(6) FCB_TILE_CASE(7) FCB_TILE_CASE(8) FCB_TILE_CASE(9) FCB_TILE_CASE(10) FCB_TILE_CASE(11)
        FCB_TILE_CASE(12) FCB_TILE_CASE(13) FCB_TILE_CASE(14)
    default: return FCB_ERR_UNSUPPORTED;
    }
#undef FCB_TILE_CASE
}
} // namespace fcb

extern "C" void fcb_host_mirror_set_mapped_io(int on);
extern "C" void fcb_host_mirror_set_zero_copy(int on);
extern "C" void fcb_host_mirror_set_strict_todo(int on);
extern "C" void fcb_host_mirror_set_xf_speculate(int on);

extern "C" int fcb_tune(const char *key, int value)
{
    if (!key) return fail(FCB_ERR_ARG, "fcb_tune: NULL key");
    if (!strcmp(key, "mac_impl") && value >= 0 && value <= 2) g_mac_impl = value;
    else if (!strcmp(key, "mac_stages") && (value == 2 || value == 3 || value == 4 || value == 6)) g_mac_stages = value;
    else if (!strcmp(key, "pipe_group") && value >= 1) g_pipe_group = value;
    else if (!strcmp(key, "mimo_tile")) g_mimo_tile = value != 0;
    else if (!strcmp(key, "mimo_tc") && value >= 0 && value <= 2) g_mimo_tc = value;
    else if (!strcmp(key, "mimo_tc_min") && value >= 1) g_mimo_tc_min = value;
    else if (!strcmp(key, "mimo_rt")) g_mimo_rt = value != 0;
    else if (!strcmp(key, "mimo_rt_wb") && (value == 1 || value == 2)) g_mimo_rt_wb = value;
    else if (!strcmp(key, "mimo_rt_min") && value >= 1) g_mimo_rt_min = value;
    else if (!strcmp(key, "mimo_rt_r") && (value == 2 || value == 4)) g_mimo_rt_r = value;
    else if (!strcmp(key, "mimo_rt_waves") && value >= 0 && value <= 16) g_mimo_rt_waves = value;
    else if (!strcmp(key, "multi_block")) g_multi_block = value != 0;
    else if (!strcmp(key, "fused_block")) g_fused_block = value != 0;
    else if (!strcmp(key, "fused_stages") && (value == 2 || value == 3)) g_fused_stages = value;
    else if (!strcmp(key, "fused_short") && value >= 0) g_fused_short = value;
    else if (!strcmp(key, "fused_pair")) g_fused_pair = value != 0;
    else if (!strcmp(key, "split")) g_split = value != 0;
    else if (!strcmp(key, "k1_late")) g_k1_late = value != 0;
    else if (!strcmp(key, "split_slots") && value >= 1) g_split_slots = value;
    else if (!strcmp(key, "split_min_stages") && value >= 1) g_split_min_stages = value;
    else if (!strcmp(key, "shared_reuse")) g_shared_reuse = value != 0;
    else if (!strcmp(key, "tma_io")) g_tma_io = value != 0;
    else if (!strcmp(key, "mapped_io")) fcb_host_mirror_set_mapped_io(value);
    else if (!strcmp(key, "zero_copy")) fcb_host_mirror_set_zero_copy(value);
    else if (!strcmp(key, "strict_todo")) fcb_host_mirror_set_strict_todo(value);
    else if (!strcmp(key, "xf_speculate")) fcb_host_mirror_set_xf_speculate(value);
    else return fail(FCB_ERR_ARG, "fcb_tune: unknown key/value %s=%d", key, value);
    return FCB_OK;
}

// ---- misc exports -----------------------------------------------------------------------------
extern "C" const char *fcb_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *fcb_version(void) { return "fftconv_b200 0.1.0 sm_100a"; }
extern "C" uint64_t fcb_launch_count(void) { return g_launches.load(); }
extern "C" uint64_t fcb_debug_alloc_count(void) { return g_resource_calls.load(); }
extern "C" int fcb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return n;
}
extern "C" void *fcb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        g_last_error = "cudaHostAlloc failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void fcb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ---- lifecycle --------------------------------------------------------------------------------
static int alloc_zero(void **p, size_t bytes, cudaStream_t s)
{
    FCB_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    FCB_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, s));
    return FCB_OK;
}

static int engine_alloc(fcb_engine *e)
{
    const size_t rowsC = e->C * e->S * e->B, rowsI = e->ir_channels() * e->S * e->B;
    FCB_TRY(alloc_zero((void **)&e->ir, rowsI * sizeof(float2), e->stream));
    FCB_TRY(alloc_zero((void **)&e->ring, rowsC * sizeof(float2), e->stream));
    FCB_TRY(alloc_zero((void **)&e->premul, e->C * e->B * sizeof(float2), e->stream));
    FCB_TRY(alloc_zero((void **)&e->overlap, e->C * e->B * sizeof(float), e->stream));
    FCB_TRY(alloc_zero((void **)&e->inbuf, e->C * e->B * sizeof(float), e->stream));
    FCB_TRY(alloc_zero((void **)&e->scratch, e->C * e->B * sizeof(float), e->stream));
    // staging for host IR uploads: whole channels, at most ~64 MB, at least one channel
    size_t per = e->L ? e->L : 1;
    size_t want = e->ir_channels() * per, cap = (size_t)16 << 20;
    e->stage_floats = want < cap ? want : (cap / per ? (cap / per) * per : per);
    FCB_TRY(alloc_zero((void **)&e->stage, e->stage_floats * sizeof(float), e->stream));
    return FCB_OK;
}

static size_t mb_limit(const fcb_engine *e);
static int mb_ensure(fcb_engine *e, size_t need);
static int pipe_streams_ensure(fcb_engine *e);

// everything a process call may need later is created here, so that the audio path never allocates:
// a default multi-block workspace (as many blocks as fit 32 MB, when that is at least 2) and, for batches that take
// the grouped host-buffer pipelines (C >= 1024), their streams and events
static int engine_prepare_process(fcb_engine *e)
{
    if (e->S == 0) return FCB_OK;
    const size_t per_block = e->C * e->B * sizeof(float2) * 4, budget = (size_t)32 << 20;
    size_t nb = budget / per_block;
    if (nb > mb_limit(e)) nb = mb_limit(e);
    if (nb >= 2 && e->logb >= 2) FCB_TRY(mb_ensure(e, nb));
    if (e->logb >= 5 && e->logb <= 9) { // whole-block kernels: split buffers for small batches
        const size_t groups = (e->C + (512 >> e->logb) - 1) / (512 >> e->logb);
        if (groups <= 2 * kSplitTargetCtas) {
            e->zslices = 3 * kSplitTargetCtas + groups;
            FCB_TRY(alloc_zero((void **)&e->zpart, e->zslices * 2 * 256 * sizeof(float4), e->stream));
            FCB_TRY(alloc_zero((void **)&e->zcount, groups * sizeof(unsigned int), e->stream));
        }
    }
    if (e->C >= 1024) {
        FCB_TRY(pipe_streams_ensure(e));
        const size_t G = (size_t)g_pipe_group.load();
        size_t ngroups = e->C / (G ? G : 1) + 4;
        if (ngroups < 8) ngroups = 8;
        while (e->pipe_in.size() < ngroups) {
            cudaEvent_t a = nullptr, b = nullptr;
            FCB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            FCB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            e->pipe_in.push_back(a);
            e->pipe_out.push_back(b);
        }
        e->pipe_cut.reserve(ngroups + 2);
    }
    return FCB_OK;
}

extern "C" int fcb_engine_create(const fcb_engine_desc *d, fcb_engine **out)
{
    if (!d || !out) return fail(FCB_ERR_ARG, "fcb_engine_create: NULL argument");
    *out = nullptr;
    if (d->channels == 0) return fail(FCB_ERR_ARG, "fcb_engine_create: channels must be >= 1");
    size_t B = next_power_of_two(d->block_size); // src/fft_convolver.rs:115
    if (B > 16384) return fail(FCB_ERR_UNSUPPORTED, "block size %zu > 16384 not supported", B);
    FCB_CUDA(cudaSetDevice(d->device));
    fcb_engine *e = new fcb_engine();
    e->device = d->device;
    e->C = d->channels;
    e->B = B;
    e->logb = ilog2(B);
    e->L = d->max_response_length;
    e->S = (size_t)std::ceil((double)e->L / (double)B); // :117
    e->shared_ir = d->shared_ir != 0;
    if (d->stream) {
        e->stream = (cudaStream_t)d->stream;
    } else {
        cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
        if (err != cudaSuccess) {
            delete e;
            return fail(FCB_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(err));
        }
        e->own_stream = true;
    }
    int rc = get_twiddles(e->device, 2 * B, &e->tw);
    if (rc == FCB_OK) rc = engine_alloc(e);
    if (rc == FCB_OK) rc = engine_prepare_process(e);
    if (rc != FCB_OK) {
        fcb_engine_destroy(e);
        return rc;
    }
    *out = e;
    return FCB_OK;
}

extern "C" void fcb_engine_destroy(fcb_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->upd_stream) {
        cudaStreamSynchronize(e->upd_stream);
        cudaStreamDestroy(e->upd_stream);
    }
    if (e->upd_done) cudaEventDestroy(e->upd_done);
    if (e->upd_fence) cudaEventDestroy(e->upd_fence);
    cudaFree(e->ir_shadow);
    cudaFree(e->zpart);
    cudaFree(e->zcount);
    cudaFree(e->ir);
    cudaFree(e->ring);
    cudaFree(e->premul);
    cudaFree(e->overlap);
    cudaFree(e->inbuf);
    cudaFree(e->scratch);
    cudaFree(e->stage);
    cudaFree(e->mb_xnew);
    cudaFree(e->mb_premul);
    cudaFree(e->mb_y);
    cudaFree(e->mb_in);
    cudaFree(e->mb_out);
    for (int i = 0; i < fcb_engine::NPIPE; i++) {
        if (e->pipe[i]) cudaStreamDestroy(e->pipe[i]);
        if (e->pipe_done[i]) cudaEventDestroy(e->pipe_done[i]);
    }
    if (e->pipe_start) cudaEventDestroy(e->pipe_start);
    for (cudaEvent_t ev : e->pipe_in) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->pipe_out) cudaEventDestroy(ev);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" int fcb_engine_clone(const fcb_engine *s, fcb_engine **out)
{
    if (!s || !out) return fail(FCB_ERR_ARG, "fcb_engine_clone: NULL argument");
    fcb_engine_desc d{s->C, s->B, s->L, s->shared_ir ? 1 : 0, s->device, s->own_stream ? nullptr : (void *)s->stream};
    fcb_engine *e = nullptr;
    FCB_TRY(fcb_engine_create(&d, &e));
    // order the copy after everything queued on the source stream
    cudaEvent_t ev;
    FCB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    FCB_CUDA(cudaEventRecord(ev, s->stream));
    FCB_CUDA(cudaStreamWaitEvent(e->stream, ev, 0));
    const size_t rowsC = s->C * s->S * s->B, rowsI = s->ir_channels() * s->S * s->B;
    cudaStream_t st = e->stream;
    FCB_CUDA(cudaMemcpyAsync(e->ir, s->ir, rowsI * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    FCB_CUDA(cudaMemcpyAsync(e->ring, s->ring, rowsC * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    FCB_CUDA(cudaMemcpyAsync(e->premul, s->premul, s->C * s->B * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    FCB_CUDA(cudaMemcpyAsync(e->overlap, s->overlap, s->C * s->B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    FCB_CUDA(cudaMemcpyAsync(e->inbuf, s->inbuf, s->C * s->B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    FCB_CUDA(cudaStreamSynchronize(st));
    cudaEventDestroy(ev);
    if (s->mb_cap > e->mb_cap) FCB_TRY(mb_ensure(e, s->mb_cap));
    if (s->ir_shadow) FCB_TRY(fcb_engine_update_reserve(e));
    *out = e;
    return FCB_OK;
}

extern "C" int fcb_engine_set_stream(fcb_engine *e, void *stream)
{
    if (!e) return fail(FCB_ERR_ARG, "NULL engine");
    FCB_CUDA(cudaStreamSynchronize(e->stream));
    if (e->own_stream) cudaStreamDestroy(e->stream);
    e->own_stream = false;
    if (stream) {
        e->stream = (cudaStream_t)stream;
    } else {
        FCB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        e->own_stream = true;
    }
    return FCB_OK;
}
extern "C" void *fcb_engine_stream(const fcb_engine *e) { return e ? (void *)e->stream : nullptr; }
extern "C" int fcb_engine_sync(fcb_engine *e)
{
    if (!e) return fail(FCB_ERR_ARG, "NULL engine");
    FCB_CUDA(cudaStreamSynchronize(e->stream));
    return FCB_OK;
}
extern "C" float *fcb_engine_scratch(fcb_engine *e) { return e ? e->scratch : nullptr; }
extern "C" float *fcb_engine_input_buffer(fcb_engine *e) { return e ? e->inbuf : nullptr; }
extern "C" size_t fcb_engine_channels(const fcb_engine *e) { return e->C; }
extern "C" size_t fcb_engine_block_size(const fcb_engine *e) { return e->B; }
extern "C" size_t fcb_engine_seg_count(const fcb_engine *e) { return e->S; }

// ---- K5: IR preparation -----------------------------------------------------------------------
// 1 = page-locked host memory (a DMA from it is still in flight when the copy call returns)
static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// K5 of `nchan` responses into `dst` (ir or ir_shadow) on stream `st`.  Host sources are staged through e->stage in
// groups; the groups need no host synchronisation between them — the next group's copy into the staging buffer is
// ordered behind the previous group's K5 by the stream itself.
static int k5_into(fcb_engine *e, float2 *dst, cudaStream_t st, size_t chan0, size_t nchan, const float *irs, size_t len,
                   size_t stride, bool on_device)
{
    const long long rows = (long long)(e->S * e->B);
    if (on_device || len == 0) {
        FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_forward<LB>(e, irs, (long long)stride, (int)len, dst + chan0 * rows, rows,
                                                              (int)e->S, (long long)(nchan * e->S), st)));
        return FCB_OK;
    }
    size_t per_group = e->stage_floats / len;
    if (per_group == 0) return fail(FCB_ERR_CUDA, "IR staging buffer too small");
    for (size_t g0 = 0; g0 < nchan; g0 += per_group) {
        size_t g = nchan - g0 < per_group ? nchan - g0 : per_group;
        FCB_CUDA(cudaMemcpy2DAsync(e->stage, len * sizeof(float), irs + (g0)*stride, stride * sizeof(float),
                                   len * sizeof(float), g, cudaMemcpyHostToDevice, st));
        FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_forward<LB>(e, e->stage, (long long)len, (int)len,
                                                              dst + (chan0 + g0) * rows, rows, (int)e->S,
                                                              (long long)(g * e->S), st)));
    }
    return FCB_OK;
}

static int set_ir_common(fcb_engine *e, size_t chan0, size_t nchan, const float *irs, size_t len, size_t stride,
                         int is_update, bool on_device)
{
    if (!e) return fail(FCB_ERR_ARG, "NULL engine");
    if (len > e->L) return fail(FCB_ERR_PANIC, "New impulse response is longer than initialized length");
    if (chan0 + nchan > e->ir_channels())
        return fail(FCB_ERR_ARG, "set_ir: channel range [%zu,%zu) outside %zu IR channels", chan0, chan0 + nchan,
                    e->ir_channels());
    if (len && !irs) return fail(FCB_ERR_ARG, "set_ir: NULL impulse response");
    if (e->S == 0 || nchan == 0) return FCB_OK; // src/fft_convolver.rs:181-183
    FCB_CUDA(cudaSetDevice(e->device));
    if (is_update) { // :185-188 (fft_buffer and conv are transient on the device)
        size_t c0 = e->shared_ir ? 0 : chan0, nc = e->shared_ir ? e->C : nchan;
        FCB_CUDA(cudaMemsetAsync(e->premul + c0 * e->B, 0, nc * e->B * sizeof(float2), e->stream));
        FCB_CUDA(cudaMemsetAsync(e->overlap + c0 * e->B, 0, nc * e->B * sizeof(float), e->stream));
    }
    if (e->upd_stream) { // the staging buffer is shared with a background update that may still be running
        FCB_CUDA(cudaEventRecord(e->upd_fence, e->upd_stream));
        FCB_CUDA(cudaStreamWaitEvent(e->stream, e->upd_fence, 0));
    }
    FCB_TRY(k5_into(e, e->ir, e->stream, chan0, nchan, irs, len, stride, on_device));
    // the reference's caller may drop or overwrite `response` as soon as update() returns: a DMA out of page-locked
    // memory is still in flight at this point, so wait for it (pageable sources were staged by the driver already)
    if (!on_device && len && is_pinned_host(irs)) FCB_CUDA(cudaStreamSynchronize(e->stream));
    return FCB_OK;
}

extern "C" int fcb_engine_set_ir(fcb_engine *e, size_t chan0, size_t nchan, const float *irs, size_t len,
                                 size_t stride, int is_update)
{
    return set_ir_common(e, chan0, nchan, irs, len, stride, is_update, false);
}
extern "C" int fcb_engine_set_ir_dev(fcb_engine *e, size_t chan0, size_t nchan, const float *irs, size_t len,
                                     size_t stride, int is_update)
{
    return set_ir_common(e, chan0, nchan, irs, len, stride, is_update, true);
}

// ---- background IR update: K5 into a second copy of the spectra while the blocks keep running ----------------
// fcb_engine_update_reserve   once, outside the audio path: the second IR buffer, a side stream, two events
// fcb_engine_update_begin     queue copy + K5 of ALL responses into the shadow buffer on the side stream; returns at
//                             once when `irs` is page-locked or device memory (it must then stay untouched until
//                             fcb_engine_update_ready says 1); pageable memory is staged by the driver during the call
// fcb_engine_update_ready     1 when the shadow buffer is complete (cudaEventQuery, never blocks)
// fcb_engine_update_commit    swap the buffers and zero pre_multiplied / overlap like update() does (:185-188); the
//                             blocks queued after it wait ON THE DEVICE for the K5 if it is still running
extern "C" int fcb_engine_update_reserve(fcb_engine *e)
{
    if (!e) return fail(FCB_ERR_ARG, "NULL engine");
    if (e->ir_shadow || e->S == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_TRY(alloc_zero((void **)&e->ir_shadow, e->ir_channels() * e->S * e->B * sizeof(float2), e->stream));
    int lo = 0, hi = 0;
    FCB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    FCB_CUDA(cudaStreamCreateWithPriority(&e->upd_stream, cudaStreamNonBlocking, lo)); // lowest priority: blocks first
    FCB_CUDA(cudaEventCreateWithFlags(&e->upd_done, cudaEventDisableTiming));
    FCB_CUDA(cudaEventCreateWithFlags(&e->upd_fence, cudaEventDisableTiming));
    FCB_CUDA(cudaEventRecord(e->upd_done, e->upd_stream));
    return FCB_OK;
}
extern "C" int fcb_engine_update_reserved(const fcb_engine *e) { return e && e->ir_shadow ? 1 : 0; }

extern "C" int fcb_engine_update_begin(fcb_engine *e, const float *irs, size_t len, size_t stride, int on_device)
{
    if (!e) return fail(FCB_ERR_ARG, "NULL engine");
    if (!e->ir_shadow) return fail(FCB_ERR_ARG, "update_begin: call fcb_engine_update_reserve first");
    if (len > e->L) return fail(FCB_ERR_PANIC, "New impulse response is longer than initialized length");
    if (len && !irs) return fail(FCB_ERR_ARG, "update_begin: NULL impulse response");
    FCB_CUDA(cudaSetDevice(e->device));
    // the shadow buffer was the active one until the last commit: kernels queued before that commit may still read it
    FCB_CUDA(cudaEventRecord(e->upd_fence, e->stream));
    FCB_CUDA(cudaStreamWaitEvent(e->upd_stream, e->upd_fence, 0));
    FCB_TRY(k5_into(e, e->ir_shadow, e->upd_stream, 0, e->ir_channels(), irs, len, stride, on_device != 0));
    FCB_CUDA(cudaEventRecord(e->upd_done, e->upd_stream));
    return FCB_OK;
}

extern "C" int fcb_engine_update_ready(fcb_engine *e)
{
    if (!e || !e->ir_shadow) return 0;
    cudaError_t q = cudaEventQuery(e->upd_done);
    if (q == cudaErrorNotReady) return 0;
    return q == cudaSuccess ? 1 : -1;
}

// host-blocking wait for the background K5 (a caller that wants its page-locked source buffer back)
extern "C" int fcb_engine_update_wait(fcb_engine *e)
{
    if (!e || !e->upd_stream) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_CUDA(cudaStreamSynchronize(e->upd_stream));
    return FCB_OK;
}

// make `stream` wait (on the device) for the background K5 queued so far, e.g. before its source buffer is rewritten
extern "C" int fcb_engine_update_join(fcb_engine *e, void *stream)
{
    if (!e || !e->upd_done) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, e->upd_done, 0));
    return FCB_OK;
}

extern "C" int fcb_engine_update_commit(fcb_engine *e)
{
    if (!e) return fail(FCB_ERR_ARG, "NULL engine");
    if (!e->ir_shadow) return fail(FCB_ERR_ARG, "update_commit: nothing reserved");
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_CUDA(cudaStreamWaitEvent(e->stream, e->upd_done, 0));
    std::swap(e->ir, e->ir_shadow);
    FCB_CUDA(cudaMemsetAsync(e->premul, 0, e->C * e->B * sizeof(float2), e->stream));
    FCB_CUDA(cudaMemsetAsync(e->overlap, 0, e->C * e->B * sizeof(float), e->stream));
    return FCB_OK;
}

extern "C" int fcb_engine_reset(fcb_engine *e)
{
    if (!e) return fail(FCB_ERR_ARG, "NULL engine");
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_CUDA(cudaMemsetAsync(e->ring, 0, e->C * e->S * e->B * sizeof(float2), e->stream));
    FCB_CUDA(cudaMemsetAsync(e->premul, 0, e->C * e->B * sizeof(float2), e->stream));
    FCB_CUDA(cudaMemsetAsync(e->overlap, 0, e->C * e->B * sizeof(float), e->stream));
    FCB_CUDA(cudaMemsetAsync(e->inbuf, 0, e->C * e->B * sizeof(float), e->stream));
    return FCB_OK;
}

// ---- per-chunk stages -----------------------------------------------------------------------
static int push_common(fcb_engine *e, const float *in, size_t stride, size_t fill, size_t n, cudaMemcpyKind kind)
{
    if (!e || !in) return fail(FCB_ERR_ARG, "push_input: NULL argument");
    if (fill + n > e->B) return fail(FCB_ERR_ARG, "push_input: fill %zu + n %zu exceeds block %zu", fill, n, e->B);
    if (n == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_CUDA(cudaMemcpy2DAsync(e->inbuf + fill, e->B * sizeof(float), in, stride * sizeof(float), n * sizeof(float),
                               e->C, kind, e->stream));
    return FCB_OK;
}
extern "C" int fcb_engine_push_input(fcb_engine *e, const float *in, size_t stride, size_t fill, size_t n)
{
    return push_common(e, in, stride, fill, n, cudaMemcpyHostToDevice);
}
extern "C" int fcb_engine_push_input_dev(fcb_engine *e, const float *in, size_t stride, size_t fill, size_t n)
{
    return push_common(e, in, stride, fill, n, cudaMemcpyDefault); // device or mapped-host pointer
}

static int check_sched(const fcb_engine *e, size_t current, size_t active, const char *who)
{
    if (!e) return fail(FCB_ERR_ARG, "%s: NULL engine", who);
    // `current` may legitimately exceed `active` after update() shrank the IR (quirk of
    // src/fft_convolver.rs:190, :248, :287-291): ring slots are then re-read modulo `active`
    if (active > e->S || (e->S && current >= e->S))
        return fail(FCB_ERR_ARG, "%s: current %zu / active %zu outside seg_count %zu", who, current, active, e->S);
    return FCB_OK;
}

extern "C" int fcb_engine_fft_forward(fcb_engine *e, size_t current, size_t valid)
{
    FCB_TRY(check_sched(e, current, e->S, "fft_forward"));
    if (valid > e->B) return fail(FCB_ERR_ARG, "fft_forward: valid %zu > block %zu", valid, e->B);
    if (e->S == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_forward<LB>(e, e->inbuf, (long long)e->B, (int)valid,
                                                          e->ring + current * e->B, e->ring_stride(), 1,
                                                          (long long)e->C)));
    return FCB_OK;
}

extern "C" int fcb_engine_mac(fcb_engine *e, size_t current, size_t active)
{
    FCB_TRY(check_sched(e, current, active, "mac"));
    if (active == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    MacArgs a{};
    a.ir = e->ir; a.ir_stride = e->ir_stride(); a.ring = e->ring; a.ring_stride = e->ring_stride();
    a.premul = e->premul; a.current = (int)current; a.active = (int)active; a.nchan = (long long)e->C;
    a.seg_lo = 1; a.seg_hi = (int)active; a.ir_seg0 = 0;
    FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_mac<LB>(e, a)));
    return FCB_OK;
}

extern "C" int fcb_engine_ifft_ola(fcb_engine *e, size_t current, size_t fill, size_t n, int block_complete,
                                   float *out_dev, size_t out_stride, const fcb_epilogue *epi)
{
    FCB_TRY(check_sched(e, current, e->S, "ifft_ola"));
    if (!out_dev) return fail(FCB_ERR_ARG, "ifft_ola: NULL output");
    if (fill + n > e->B) return fail(FCB_ERR_ARG, "ifft_ola: fill %zu + n %zu exceeds block %zu", fill, n, e->B);
    if (e->S == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    IfftArgs a{};
    a.ring_cur = e->ring + current * e->B;
    a.ring_stride = e->ring_stride();
    a.ir0 = e->ir;
    a.ir_stride = e->ir_stride();
    a.premul = e->premul;
    a.overlap = e->overlap;
    a.out = out_dev;
    a.out_stride = (long long)out_stride;
    a.fill = (int)fill;
    a.n = (int)n;
    a.block_complete = block_complete;
    a.nchan = (long long)e->C;
    if (epi) a.epi = *epi;
    FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_inverse<LB>(e, a)));
    return FCB_OK;
}

extern "C" int fcb_engine_fetch(fcb_engine *e, float *out_host, size_t host_stride, const float *src_dev,
                                size_t dev_stride, size_t n)
{
    if (!e || !out_host || !src_dev) return fail(FCB_ERR_ARG, "fetch: NULL argument");
    FCB_CUDA(cudaSetDevice(e->device));
    if (n)
        FCB_CUDA(cudaMemcpy2DAsync(out_host, host_stride * sizeof(float), src_dev, dev_stride * sizeof(float),
                                   n * sizeof(float), e->C, cudaMemcpyDeviceToHost, e->stream));
    FCB_CUDA(cudaStreamSynchronize(e->stream));
    return FCB_OK;
}

extern "C" int fcb_engine_process_block_dev(fcb_engine *e, const float *in_dev, size_t in_stride, float *out_dev,
                                            size_t out_stride, size_t current, size_t active,
                                            const fcb_epilogue *epi)
{
    FCB_TRY(check_sched(e, current, active, "process_block"));
    if (!in_dev || !out_dev) return fail(FCB_ERR_ARG, "process_block: NULL argument");
    if (active == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    if (fused_applicable(e, active))
        return run_block_fused(e, e->stream, 0, e->C, in_dev, in_stride, out_dev, out_stride, current, active, epi);
    // K1 straight from the caller's block: a full block leaves the input buffer empty again (:280-281).
    // (Running K1 on a side stream underneath K2 was measured in round 1: no gain — K1 then competes
    // with the HBM-bound K2 for the same bandwidth.)
    FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_forward<LB>(e, in_dev, (long long)in_stride, (int)e->B,
                                                          e->ring + current * e->B, e->ring_stride(), 1,
                                                          (long long)e->C)));
    FCB_TRY(fcb_engine_mac(e, current, active));
    return fcb_engine_ifft_ola(e, current, 0, e->B, 1, out_dev, out_stride, epi);
}

// ---- multi-block calls: nblocks whole blocks of every channel in one time-batched pass -------------
template <int LOGB, int T, int P = (T < 4 ? T : 4)>
static int launch_mac_time_t(const MacTimeArgs &a, cudaStream_t st)
{
    constexpr int B = 1 << LOGB, ROW4 = B / 2, TX = ROW4 < 256 ? ROW4 : 256, TILES = ROW4 / TX, CPB = 256 / TX;
    const long long ngq = (a.nblocks + T - 1) / T, cgroups = (a.nchan + CPB - 1) / CPB;
    k_mac_time<B, T, P><<<(unsigned)(TILES * ngq * cgroups), 256, 0, st>>>(a);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}

extern "C" int fcb_engine_multi_block_ok(const fcb_engine *e, size_t current, size_t active)
{
    return e && g_multi_block.load() && e->logb >= 2 && active >= 1 && current < active;
}

// most blocks one pass may hold: 256 MB per workspace buffer, at least 4, at most 1024
static size_t mb_limit(const fcb_engine *e)
{
    const size_t per_block = e->C * e->B * sizeof(float2);
    size_t cap = ((size_t)256 << 20) / per_block;
    return cap < 4 ? 4 : cap > 1024 ? 1024 : cap;
}

static void mb_release(fcb_engine *e)
{
    cudaFree(e->mb_xnew);
    cudaFree(e->mb_premul);
    cudaFree(e->mb_y);
    cudaFree(e->mb_in);
    cudaFree(e->mb_out);
    e->mb_xnew = e->mb_premul = nullptr;
    e->mb_y = e->mb_in = e->mb_out = nullptr;
    e->mb_cap = 0;
}

// workspace for `need` blocks per pass.  Called from fcb_engine_create (a small default) and from
// fcb_engine_multi_block_reserve only — never from a process call (src/lib.rs:8: no allocation on the audio path).
static int mb_ensure(fcb_engine *e, size_t need)
{
    if (need > mb_limit(e)) need = mb_limit(e);
    if (need <= e->mb_cap) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    if (e->mb_cap) {
        FCB_CUDA(cudaStreamSynchronize(e->stream));
        mb_release(e);
    }
    const size_t per_block = e->C * e->B * sizeof(float2);
    if (cudaMalloc((void **)&e->mb_xnew, need * per_block) != cudaSuccess ||
        cudaMalloc((void **)&e->mb_premul, need * per_block) != cudaSuccess ||
        cudaMalloc((void **)&e->mb_y, need * per_block) != cudaSuccess ||
        cudaMalloc((void **)&e->mb_in, need * per_block / 2) != cudaSuccess ||
        cudaMalloc((void **)&e->mb_out, need * per_block / 2) != cudaSuccess) {
        cudaGetLastError();
        mb_release(e); // nothing half-allocated survives a failure
        return fail(FCB_ERR_CUDA, "multi-block workspace for %zu blocks (%zu MB) does not fit", need, need * per_block * 4 >> 20);
    }
    e->mb_cap = need;
    return FCB_OK;
}

extern "C" size_t fcb_engine_multi_block_capacity(fcb_engine *e) { return e ? mb_limit(e) : 0; }
extern "C" size_t fcb_engine_multi_block_reserved(const fcb_engine *e) { return e ? e->mb_cap : 0; }

// size the multi-block workspace (create reserves what fits FCB_MB_DEFAULT_BYTES; callers whose buffers span more
// blocks call this once, outside the audio path)
extern "C" int fcb_engine_multi_block_reserve(fcb_engine *e, size_t nblocks)
{
    if (!e) return fail(FCB_ERR_ARG, "multi_block_reserve: NULL engine");
    return mb_ensure(e, nblocks);
}

// in / out: device pointers, or host pointers when host_io (staged through the workspace).  Caller rotates
// `current` nblocks times afterwards (src/fft_convolver.rs:287-291).  Output is bit-identical to nblocks calls of
// fcb_engine_process_block_dev.
// the time-batched pass for channels [c0, c0 + nc) on stream st; din / dout point at channel c0's first sample (device)
static int process_blocks_range(fcb_engine *e, cudaStream_t st, size_t c0, size_t nc, const float *din, size_t dstride_in,
                                float *dout, size_t dstride_out, size_t current, size_t active, size_t NB,
                                const fcb_epilogue *epi)
{
    const size_t B = e->B;
    const long long ir_off = e->shared_ir ? 0 : (long long)(c0 * e->S * B);
    float2 *xnew = e->mb_xnew + c0 * NB * B, *premul = e->mb_premul + c0 * NB * B;
    float *y = e->mb_y + c0 * NB * 2 * B;
    // K1 for every (channel, block) -> xnew[c][d]
    FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_forward<LB>(e, din, (long long)dstride_in, (int)(NB * B), xnew, (long long)(NB * B),
                                                          (int)NB, (long long)(nc * NB), st)));
    MacTimeArgs m{};
    m.ir = e->ir + ir_off;
    m.ir_stride = e->ir_stride();
    m.ring = e->ring + c0 * e->ring_stride();
    m.ring_stride = e->ring_stride();
    m.xnew = xnew;
    m.premul = premul;
    m.current = (int)current;
    m.active = (int)active;
    m.nblocks = (int)NB;
    m.nchan = (long long)nc;
    cudaEvent_t prof_stop = nullptr;
    const bool profiled = prof_before(st, &prof_stop) != nullptr;
#define FCB_MAC_TIME_CASE(LB)                                                      \
    case LB:                                                                       \
        if (NB >= 3) FCB_TRY((launch_mac_time_t<LB, 4>(m, st)));                   \
        else FCB_TRY((launch_mac_time_t<LB, 2>(m, st)));                           \
        break;
    switch (e->logb) {
        FCB_MAC_TIME_CASE(2) FCB_MAC_TIME_CASE(3) FCB_MAC_TIME_CASE(4) FCB_MAC_TIME_CASE(5) FCB_MAC_TIME_CASE(6)
        FCB_MAC_TIME_CASE(7) FCB_MAC_TIME_CASE(8) FCB_MAC_TIME_CASE(9) FCB_MAC_TIME_CASE(10) FCB_MAC_TIME_CASE(11)
        FCB_MAC_TIME_CASE(12) FCB_MAC_TIME_CASE(13) FCB_MAC_TIME_CASE(14)
    default: return fail(FCB_ERR_UNSUPPORTED, "process_blocks: block size");
    }
#undef FCB_MAC_TIME_CASE
    if (profiled) cudaEventRecord(prof_stop, st);
    // K3 in raw mode: conv_d = pre_multiplied_d + X_d * H_0, inverse FFT, /N -> y[c][d][2B]
    IfftArgs a{};
    a.ring_cur = xnew;
    a.ring_stride = (long long)B;
    a.ir0 = e->ir + ir_off;
    a.ir_stride = e->ir_stride();
    a.ir_div = (long long)NB;
    a.premul = premul;
    a.raw_out = y;
    a.nchan = (long long)(nc * NB);
    FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_inverse<LB>(e, a, st)));
    // overlap-add across the call's blocks (+ epilogue), then the new overlap and the ring
    fcb_epilogue ep;
    memset(&ep, 0, sizeof ep);
    if (epi) {
        ep = *epi;
        if (ep.add0) ep.add0 += c0 * epi->add_stride;
        if (ep.add1) ep.add1 += c0 * epi->add_stride;
        if (ep.mix_other) ep.mix_other += c0 * epi->mix_stride;
    }
    const long long nout = (long long)(nc * NB * B);
    k_ola_time<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(y, e->overlap + c0 * B, dout, (long long)dstride_out, (int)B, (int)NB, nout, ep);
    FCB_CUDA(cudaMemcpy2DAsync(e->overlap + c0 * B, B * sizeof(float), y + (NB - 1) * 2 * B + B, NB * 2 * B * sizeof(float),
                               B * sizeof(float), nc, cudaMemcpyDeviceToDevice, st));
    const size_t first = NB > active ? NB - active : 0;
    const long long nring = (long long)(nc * (NB - first) * (B / 2));
    k_ring_update<<<(unsigned)((nring + 255) / 256), 256, 0, st>>>(xnew, e->ring + c0 * e->ring_stride(), e->ring_stride(), (int)B, (int)NB,
                                                                   (int)current, (int)active, (int)first, nring);
    g_launches += 2;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}

static int pipe_streams_ensure(fcb_engine *e)
{
    if (e->pipe_start) return FCB_OK;
    FCB_CUDA(cudaEventCreateWithFlags(&e->pipe_start, cudaEventDisableTiming));
    for (int i = 0; i < fcb_engine::NPIPE; i++) {
        FCB_CUDA(cudaStreamCreateWithFlags(&e->pipe[i], cudaStreamNonBlocking));
        FCB_CUDA(cudaEventCreateWithFlags(&e->pipe_done[i], cudaEventDisableTiming));
    }
    return FCB_OK;
}

// in / out: device pointers, or host pointers when host_io (staged through the workspace; with 1024 channels or more
// the channels are cut into groups whose H2D copy, pass and D2H copy overlap on separate streams).  Caller rotates
// `current` nblocks times afterwards (src/fft_convolver.rs:287-291).  Output is bit-identical to nblocks calls of
// fcb_engine_process_block_dev.
extern "C" int fcb_engine_process_blocks(fcb_engine *e, const float *in, size_t in_stride, float *out, size_t out_stride,
                                         size_t current, size_t active, size_t nblocks, const fcb_epilogue *epi,
                                         int host_io)
{
    FCB_TRY(check_sched(e, current, active, "process_blocks"));
    if (!in || !out) return fail(FCB_ERR_ARG, "process_blocks: NULL argument");
    if (!fcb_engine_multi_block_ok(e, current, active)) return fail(FCB_ERR_UNSUPPORTED, "process_blocks: not applicable here");
    FCB_CUDA(cudaSetDevice(e->device));
    if (nblocks == 0) return FCB_OK;
    if (nblocks > e->mb_cap)
        return fail(FCB_ERR_ARG, "process_blocks: %zu blocks exceed the reserved workspace (%zu; fcb_engine_multi_block_reserve)",
                    nblocks, e->mb_cap);
    const size_t B = e->B, C = e->C, NB = nblocks, row = NB * B;
    cudaStream_t st = e->stream;
    if (!host_io) return process_blocks_range(e, st, 0, C, in, in_stride, out, out_stride, current, active, NB, epi);
    if (C < 1024) {
        FCB_CUDA(cudaMemcpy2DAsync(e->mb_in, row * sizeof(float), in, in_stride * sizeof(float), row * sizeof(float), C,
                                   cudaMemcpyHostToDevice, st));
        FCB_TRY(process_blocks_range(e, st, 0, C, e->mb_in, row, e->mb_out, row, current, active, NB, epi));
        FCB_CUDA(cudaMemcpy2DAsync(out, out_stride * sizeof(float), e->mb_out, row * sizeof(float), row * sizeof(float), C,
                                   cudaMemcpyDeviceToHost, st));
        FCB_CUDA(cudaStreamSynchronize(st));
        return FCB_OK;
    }
    // many channels: groups of channels, all H2D copies on one stream running ahead, the passes on two compute streams
    // as their input lands, the D2H copies on a fourth stream — the PCIe traffic hides under the HBM-bound passes
    FCB_TRY(pipe_streams_ensure(e)); // created with the engine (C >= 1024)
    const size_t ngroups = 8, per = (C + ngroups - 1) / ngroups;
    while (e->pipe_in.size() < ngroups) { // sized in fcb_engine_create; grows only after fcb_tune("pipe_group")
        cudaEvent_t a = nullptr, b = nullptr;
        FCB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        FCB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        e->pipe_in.push_back(a);
        e->pipe_out.push_back(b);
    }
    cudaStream_t s_in = e->pipe[0], s_out = e->pipe[1];
    FCB_CUDA(cudaEventRecord(e->pipe_start, st));
    for (int i = 0; i < fcb_engine::NPIPE; i++) FCB_CUDA(cudaStreamWaitEvent(e->pipe[i], e->pipe_start, 0));
    for (size_t g = 0; g < ngroups; g++) {
        const size_t c0 = g * per;
        if (c0 >= C) break;
        const size_t nc = C - c0 < per ? C - c0 : per;
        cudaStream_t sc = e->pipe[2 + g % (fcb_engine::NPIPE - 2)];
        FCB_CUDA(cudaMemcpy2DAsync(e->mb_in + c0 * row, row * sizeof(float), in + c0 * in_stride, in_stride * sizeof(float),
                                   row * sizeof(float), nc, cudaMemcpyHostToDevice, s_in));
        FCB_CUDA(cudaEventRecord(e->pipe_in[g], s_in));
        FCB_CUDA(cudaStreamWaitEvent(sc, e->pipe_in[g], 0));
        FCB_TRY(process_blocks_range(e, sc, c0, nc, e->mb_in + c0 * row, row, e->mb_out + c0 * row, row, current, active, NB, epi));
        FCB_CUDA(cudaEventRecord(e->pipe_out[g], sc));
        FCB_CUDA(cudaStreamWaitEvent(s_out, e->pipe_out[g], 0));
        FCB_CUDA(cudaMemcpy2DAsync(out + c0 * out_stride, out_stride * sizeof(float), e->mb_out + c0 * row, row * sizeof(float),
                                   row * sizeof(float), nc, cudaMemcpyDeviceToHost, s_out));
    }
    FCB_CUDA(cudaEventRecord(e->pipe_done[1], s_out));
    FCB_CUDA(cudaStreamWaitEvent(st, e->pipe_done[1], 0));
    for (int i = 2; i < fcb_engine::NPIPE; i++) { // the engine stream also orders after the compute streams
        FCB_CUDA(cudaEventRecord(e->pipe_done[i], e->pipe[i]));
        FCB_CUDA(cudaStreamWaitEvent(st, e->pipe_done[i], 0));
    }
    FCB_CUDA(cudaStreamSynchronize(st));
    return FCB_OK;
}

// Full block, host buffers, pipelined: the channels are cut into groups; each group's pinned H2D
// copy, K1, K2, K3 and D2H copy are queued on one of NPIPE streams, so the PCIe copies of one
// group overlap the HBM-bound K2 of another.  Groups are independent channels, so the result is
// the one the unpipelined sequence gives.  Returns after every group's output is in host memory.
extern "C" int fcb_engine_process_block_host(fcb_engine *e, const float *in, size_t in_stride, float *out,
                                             size_t out_stride, size_t current, size_t active, size_t group_channels)
{
    FCB_TRY(check_sched(e, current, active, "process_block_host"));
    if (!in || !out) return fail(FCB_ERR_ARG, "process_block_host: NULL argument");
    if (active == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_TRY(pipe_streams_ensure(e)); // created with the engine (C >= 1024)
    const size_t B = e->B, C = e->C;
    size_t G = group_channels ? group_channels : (size_t)g_pipe_group.load();
    if (G > C) G = C;
    // group boundaries: full groups of G with a short first and last group (G/4), so the pipeline
    // fills and drains on a quarter-size copy instead of a full one
    std::vector<size_t> &cut = e->pipe_cut;
    if (e->pipe_cut_G != G) { // first call with this grouping (not in the steady state)
    cut.clear();
    cut.push_back(0);
    if (C >= 4 * G && G >= 4) {
        const size_t edge = G / 4;
        cut.push_back(edge);
        size_t c = edge;
        while (c + G + edge <= C) {
            c += G;
            cut.push_back(c);
        }
        if (C - c > edge) cut.push_back(C - edge);
        cut.push_back(C);
    } else {
        for (size_t c = G; c < C; c += G) cut.push_back(c);
        cut.push_back(C);
    }
    e->pipe_cut_G = G;
    }
    const size_t ngroups = cut.size() - 1;
    while (e->pipe_in.size() < ngroups) { // grows on the first call with this grouping only
        cudaEvent_t a = nullptr, b = nullptr;
        FCB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        FCB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        e->pipe_in.push_back(a);
        e->pipe_out.push_back(b);
    }
    // stream roles: pipe[0] = all H2D copies (run ahead at PCIe speed), pipe[1] = all D2H copies,
    // pipe[2..] = compute, groups alternating so one group's K2 tail overlaps the next group's head
    cudaStream_t s_in = e->pipe[0], s_out = e->pipe[1];
    FCB_CUDA(cudaEventRecord(e->pipe_start, e->stream));
    for (int i = 0; i < fcb_engine::NPIPE; i++) FCB_CUDA(cudaStreamWaitEvent(e->pipe[i], e->pipe_start, 0));
    // enqueue order: copy-in of group g+1 is queued right after group g's kernel, so the first kernel
    // is in the GPU's queue after two driver calls instead of after every copy has been queued
    auto queue_copy_in = [&](size_t g) -> int {
        const size_t c0 = cut[g], nc = cut[g + 1] - cut[g];
        FCB_CUDA(cudaMemcpy2DAsync(e->inbuf + c0 * B, B * sizeof(float), in + c0 * in_stride, in_stride * sizeof(float),
                                   B * sizeof(float), nc, cudaMemcpyHostToDevice, s_in));
        FCB_CUDA(cudaEventRecord(e->pipe_in[g], s_in));
        return FCB_OK;
    };
    FCB_TRY(queue_copy_in(0));
    if (ngroups > 1) FCB_TRY(queue_copy_in(1));
    for (size_t g = 0; g < ngroups; g++) {
        cudaStream_t st = e->pipe[2 + g % (fcb_engine::NPIPE - 2)];
        const size_t c0 = cut[g], nc = cut[g + 1] - cut[g];
        float *d_in = e->inbuf + c0 * B, *d_out = e->scratch + c0 * B;
        if (g + 2 < ngroups) FCB_TRY(queue_copy_in(g + 2));
        FCB_CUDA(cudaStreamWaitEvent(st, e->pipe_in[g], 0));
        if (fused_applicable(e, active)) {
            FCB_TRY(run_block_fused(e, st, c0, nc, e->inbuf, B, e->scratch, B, current, active, nullptr));
            FCB_CUDA(cudaEventRecord(e->pipe_out[g], st));
            FCB_CUDA(cudaStreamWaitEvent(s_out, e->pipe_out[g], 0));
            FCB_CUDA(cudaMemcpy2DAsync(out + c0 * out_stride, out_stride * sizeof(float), d_out, B * sizeof(float),
                                       B * sizeof(float), nc, cudaMemcpyDeviceToHost, s_out));
            continue;
        }
        FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_forward<LB>(e, d_in, (long long)B, (int)B,
                                                              e->ring + c0 * e->ring_stride() + current * B,
                                                              e->ring_stride(), 1, (long long)nc, st)));
        MacArgs m{};
        m.ir = e->ir + (e->shared_ir ? 0 : c0) * (long long)(e->S * B); m.ir_stride = e->ir_stride();
        m.ring = e->ring + c0 * e->ring_stride(); m.ring_stride = e->ring_stride();
        m.premul = e->premul + c0 * B; m.current = (int)current; m.active = (int)active; m.nchan = (long long)nc;
        m.seg_lo = 1; m.seg_hi = (int)active; m.ir_seg0 = 0;
        FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_mac<LB>(e, m, st)));
        IfftArgs a{};
        a.ring_cur = e->ring + c0 * e->ring_stride() + current * B;
        a.ring_stride = e->ring_stride();
        a.ir0 = e->ir + (e->shared_ir ? 0 : c0) * (long long)(e->S * B);
        a.ir_stride = e->ir_stride();
        a.premul = e->premul + c0 * B;
        a.overlap = e->overlap + c0 * B;
        a.out = d_out;
        a.out_stride = (long long)B;
        a.fill = 0;
        a.n = (int)B;
        a.block_complete = 1;
        a.nchan = (long long)nc;
        FCB_DISPATCH_LOGB(e->logb, FCB_TRY(launch_inverse<LB>(e, a, st)));
        FCB_CUDA(cudaEventRecord(e->pipe_out[g], st));
        FCB_CUDA(cudaStreamWaitEvent(s_out, e->pipe_out[g], 0));
        FCB_CUDA(cudaMemcpy2DAsync(out + c0 * out_stride, out_stride * sizeof(float), d_out, B * sizeof(float),
                                   B * sizeof(float), nc, cudaMemcpyDeviceToHost, s_out));
    }
    // the last D2H completes after every compute stream it waited on
    FCB_CUDA(cudaEventRecord(e->pipe_done[1], s_out));
    FCB_CUDA(cudaStreamWaitEvent(e->stream, e->pipe_done[1], 0));
    FCB_CUDA(cudaStreamSynchronize(e->stream));
    return FCB_OK;
}

// ---- debug readback in the reference layout (K = B+1 interleaved complex) ------------------------
static int read_row(fcb_engine *e, const float2 *row, float *out_2k)
{
    std::vector<float2> h(e->B);
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_CUDA(cudaStreamSynchronize(e->stream));
    FCB_CUDA(cudaMemcpy(h.data(), row, e->B * sizeof(float2), cudaMemcpyDeviceToHost));
    out_2k[0] = h[0].x;
    out_2k[1] = 0.f;
    for (size_t k = 1; k < e->B; k++) {
        out_2k[2 * k] = h[k].x;
        out_2k[2 * k + 1] = h[k].y;
    }
    out_2k[2 * e->B] = h[0].y;
    out_2k[2 * e->B + 1] = 0.f;
    return FCB_OK;
}
static int write_row(fcb_engine *e, float2 *row, const float *in_2k)
{
    std::vector<float2> h(e->B);
    h[0] = make_float2(in_2k[0], in_2k[2 * e->B]);
    for (size_t k = 1; k < e->B; k++) h[k] = make_float2(in_2k[2 * k], in_2k[2 * k + 1]);
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_CUDA(cudaStreamSynchronize(e->stream));
    FCB_CUDA(cudaMemcpy(row, h.data(), e->B * sizeof(float2), cudaMemcpyHostToDevice));
    return FCB_OK;
}
#define FCB_CHECK_ROW(e, chan, seg, nch)                                                   \
    if (!(e) || (chan) >= (nch) || (seg) >= (e)->S) return fail(FCB_ERR_ARG, "row index out of range")

extern "C" int fcb_engine_read_ir_segment(fcb_engine *e, size_t chan, size_t seg, float *out)
{
    FCB_CHECK_ROW(e, chan, seg, e->ir_channels());
    return read_row(e, e->ir + (chan * e->S + seg) * e->B, out);
}
extern "C" int fcb_engine_read_ring_segment(fcb_engine *e, size_t chan, size_t seg, float *out)
{
    FCB_CHECK_ROW(e, chan, seg, e->C);
    return read_row(e, e->ring + (chan * e->S + seg) * e->B, out);
}
extern "C" int fcb_engine_write_ir_segment(fcb_engine *e, size_t chan, size_t seg, const float *in)
{
    FCB_CHECK_ROW(e, chan, seg, e->ir_channels());
    return write_row(e, e->ir + (chan * e->S + seg) * e->B, in);
}
extern "C" int fcb_engine_write_ring_segment(fcb_engine *e, size_t chan, size_t seg, const float *in)
{
    FCB_CHECK_ROW(e, chan, seg, e->C);
    return write_row(e, e->ring + (chan * e->S + seg) * e->B, in);
}
extern "C" int fcb_engine_read_premul(fcb_engine *e, size_t chan, float *out)
{
    if (!e || chan >= e->C) return fail(FCB_ERR_ARG, "channel out of range");
    return read_row(e, e->premul + chan * e->B, out);
}
extern "C" int fcb_engine_read_overlap(fcb_engine *e, size_t chan, float *out)
{
    if (!e || chan >= e->C) return fail(FCB_ERR_ARG, "channel out of range");
    FCB_CUDA(cudaSetDevice(e->device));
    FCB_CUDA(cudaStreamSynchronize(e->stream));
    FCB_CUDA(cudaMemcpy(out, e->overlap + chan * e->B, e->B * sizeof(float), cudaMemcpyDeviceToHost));
    return FCB_OK;
}
