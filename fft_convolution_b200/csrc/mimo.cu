// mimo.cu — convolution matrix (BASELINE configs[4]): OUT x IN FFTConvolver-equivalents sharing
// the IN input-spectrum rings, optionally NS independent streams sharing the one IR matrix, and
// optionally one IR-partition shard of a multi-GPU job.
//
// Semantics: y_out = sum_in FFTConvolver(h[out][in]).process(x_in), i.e. OUT*IN
// reference convolvers (src/fft_convolver.rs:86-307) with outputs summed over `in`.  Because the
// inverse FFT and overlap-add are linear, the sum over `in` is taken in the frequency domain and
// each output gets ONE inverse FFT and ONE overlap buffer; the delay-line sum over segments
// (:244-255) is associative, so a shard may own only a contiguous range of IR segments and the
// partial spectra of all shards are summed (NCCL all-reduce, done by the caller) before K3.
// Full blocks only (n == B per call).
#include <cmath>
#include <cstring>

#include "engine_internal.cuh"
#include "mimo_tc.cuh"
#include "mimo_rt.cuh"

using namespace fcb;

namespace fcb {
extern std::atomic<bool> g_mimo_tile;
extern std::atomic<int> g_mimo_tc; // 0 never, 1 whenever the shape fits, 2 (default) when it fits and NS >= g_mimo_tc_min
extern std::atomic<int> g_mimo_tc_min; // fewer streams than this run on the FP32 pipes (k_mac_rt)
extern std::atomic<bool> g_mimo_rt;    // register-tiled matrix MAC for 2+ streams (0: the shared-memory tile kernel)
extern std::atomic<int> g_mimo_rt_wb;  // warps side by side along the bins in k_mac_rt (1 or 2)
extern std::atomic<int> g_mimo_rt_min; // streams from which k_mac_rt replaces the tile kernel
extern std::atomic<int> g_mimo_rt_r;   // segments per pipeline stage (4: 3 stages, 2: 6 stages)
extern std::atomic<int> g_mimo_rt_waves; // waves of resident CTAs the segment chunking aims at

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// f32 tensor [d3][d2][d1][d0] (d0 contiguous, pitches in bytes; rank 3 when d3 == 0), box [..1][b1][32] with
// the 128-byte swizzle the K-major UMMA descriptors expect; out-of-bounds elements read as zero
static int tc_encode_map(CUtensorMap *tm, void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t pitch1,
                         uint64_t pitch2, uint64_t pitch3, uint32_t b1)
{
    static TensorMapEncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FCB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p) return fail(FCB_ERR_CUDA, "cuTensorMapEncodeTiled is not exported by this driver");
        fn = (TensorMapEncodeTiledFn)p;
    }
    cuuint64_t dims[4] = {d0, d1, d2, d3};
    cuuint64_t strides[3] = {pitch1, pitch2, pitch3};
    cuuint32_t box[4] = {2 * TC_KSEG, b1, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d3 ? 4 : 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FCB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return FCB_OK;
}

// f32 tensor [d3][d2][d1][d0] (d0 contiguous, pitches in bytes), box [b3][1][b1][b0], no swizzle; out-of-bounds reads as zero
static int rt_encode_map(CUtensorMap *tm, void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t pitch1,
                         uint64_t pitch2, uint64_t pitch3, uint32_t b0, uint32_t b1, uint32_t b3)
{
    static TensorMapEncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FCB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p) return fail(FCB_ERR_CUDA, "cuTensorMapEncodeTiled is not exported by this driver");
        fn = (TensorMapEncodeTiledFn)p;
    }
    cuuint64_t dims[4] = {d0, d1, d2, d3};
    cuuint64_t strides[3] = {pitch1, pitch2, pitch3};
    cuuint32_t box[4] = {b0, b1, 1, b3};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FCB_ERR_CUDA, "cuTensorMapEncodeTiled (matrix operands) failed (%d)", (int)r);
    return FCB_OK;
}

// k_mac_rt: CTA shape by problem (8 or 16 outputs x 4, 8 or 16 streams) and the segment chunking that fills ONE wave of
// resident CTAs (the partial spectra have no input dimension, so a chunk is cheap: Z * NS * OUT rows)
struct RtPlan {
    int wo = 1, ws = 1, wb = 1, r = 4, zchunks = 1, out_groups = 1, stream_groups = 1;
    size_t smem = 0;
    int threads = 32, per_sm = 1;
};
template <int WO, int WS, int WB, int R>
static void rt_shape(RtPlan &p)
{
    using Cfg = RtCfg<WO, WS, WB, R>;
    p.r = R;
    p.wo = WO;
    p.ws = WS;
    p.wb = WB;
    p.smem = Cfg::SMEM;
    p.threads = Cfg::THREADS;
    const int by_smem = (int)((227 * 1024) / (Cfg::SMEM + 1024));
    p.per_sm = Cfg::MIN_CTAS < by_smem ? Cfg::MIN_CTAS : by_smem;
}
static RtPlan rt_plan(int B, int n_out, int n_streams, int nsegs)
{
    RtPlan p;
    const bool o16 = n_out > 8;
    // streams per CTA: 16, 8 or 4 — the widest shape whose padding (stream rows that exist only as zeros) stays under a
    // quarter of the work: 24 streams run as 3 groups of 8, not as 2 groups of 16
    int sw = 1;
    for (int c = 4; c >= 1; c /= 2) {
        const int padded = (n_streams + 4 * c - 1) / (4 * c) * (4 * c);
        if (4 * (padded - n_streams) <= padded || c == 1) {
            sw = c;
            break;
        }
    }
    // two warps side by side along the bins (512-byte runs per TMA row instead of 256) when the block has the bins
    const bool wide = B >= 2 * RT_BINS && g_mimo_rt_wb.load() >= 2;
    const int wo = o16 ? 2 : 1, wb = wide ? 2 : 1, r = g_mimo_rt_r.load();
#define FCB_RT_SHAPE(WO, WS, WB)                                                      \
    if (wo == WO && sw == WS && wb == WB) {                                           \
        if (r == 2) rt_shape<WO, WS, WB, 2>(p);                                       \
        else rt_shape<WO, WS, WB, 4>(p);                                              \
    }
    FCB_RT_SHAPE(2, 4, 1) FCB_RT_SHAPE(2, 2, 1) FCB_RT_SHAPE(2, 1, 1) FCB_RT_SHAPE(1, 4, 1) FCB_RT_SHAPE(1, 2, 1) FCB_RT_SHAPE(1, 1, 1)
    FCB_RT_SHAPE(2, 4, 2) FCB_RT_SHAPE(2, 2, 2) FCB_RT_SHAPE(2, 1, 2) FCB_RT_SHAPE(1, 4, 2) FCB_RT_SHAPE(1, 2, 2) FCB_RT_SHAPE(1, 1, 2)
#undef FCB_RT_SHAPE
    p.out_groups = (n_out + 8 * p.wo - 1) / (8 * p.wo);
    p.stream_groups = (n_streams + 4 * p.ws - 1) / (4 * p.ws);
    const int bins = RT_BINS * p.wb;
    const long long base = (long long)((B + bins - 1) / bins) * p.out_groups * p.stream_groups;
    const long long slots = 148LL * p.per_sm;
    // one wave of CTAs while the operands' HBM time dominates (up to 8 streams: fewer partial rows), two from there on
    // (the FP32 pipes dominate and the second wave evens out the CTAs' finishing times); measured, profiles/r02_k_mac_rt_notes.txt
    const int waves = g_mimo_rt_waves.load() > 0 ? g_mimo_rt_waves.load() : n_streams <= 8 ? 1 : 2;
    long long z = slots * waves / base;
    const long long zmax = nsegs / 8 > 0 ? nsegs / 8 : 1; // at least eight segments per chunk
    z = z > zmax ? zmax : z < 1 ? 1 : z;
    p.zchunks = nsegs > 0 ? (int)z : 0;
    return p;
}
template <int WO, int WS, int WB, int R>
static int rt_launch_t(const RtArgs &a, const CUtensorMap &tm_ir, const CUtensorMap &tm_ring, cudaStream_t st)
{
    using Cfg = RtCfg<WO, WS, WB, R>;
    static bool opted = false;
    if (!opted) {
        if (cudaFuncSetAttribute(k_mac_rt<WO, WS, WB, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM) != cudaSuccess)
            return fail(FCB_ERR_CUDA, "k_mac_rt: cannot opt in to %zu bytes of shared memory", (size_t)Cfg::SMEM);
        opted = true;
    }
    const long long grid = (long long)((a.B + Cfg::BINS - 1) / Cfg::BINS) * a.zchunks * a.out_groups * a.stream_groups;
    cudaEvent_t prof_stop = nullptr;
    const bool profiled = mac_profile_begin(st, &prof_stop) != nullptr;
    k_mac_rt<WO, WS, WB, R><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, st>>>(a, tm_ir, tm_ring);
    if (profiled) cudaEventRecord(prof_stop, st);
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return FCB_OK;
}
static int rt_launch(const RtPlan &p, const RtArgs &a, const CUtensorMap &tm_ir, const CUtensorMap &tm_ring, cudaStream_t st)
{
#define FCB_RT_CASE(WO, WS, WB)                                                                                              \
    if (p.wo == WO && p.ws == WS && p.wb == WB)                                                                              \
        return p.r == 2 ? rt_launch_t<WO, WS, WB, 2>(a, tm_ir, tm_ring, st) : rt_launch_t<WO, WS, WB, 4>(a, tm_ir, tm_ring, st);
    FCB_RT_CASE(2, 4, 1) FCB_RT_CASE(2, 2, 1) FCB_RT_CASE(2, 1, 1) FCB_RT_CASE(1, 4, 1) FCB_RT_CASE(1, 2, 1) FCB_RT_CASE(1, 1, 1)
    FCB_RT_CASE(2, 4, 2) FCB_RT_CASE(2, 2, 2) FCB_RT_CASE(2, 1, 2) FCB_RT_CASE(1, 4, 2) FCB_RT_CASE(1, 2, 2) FCB_RT_CASE(1, 1, 2)
#undef FCB_RT_CASE
    return fail(FCB_ERR_ARG, "k_mac_rt: no such shape");
}

// ---- peer exchange of the partial spectra (replaces the NCCL all-reduce between the MAC and K3) ----
// Every shard's reduce kernel stores its partial conv spectra straight into EVERY shard's inbox slot
// [parity][me] over NVLink (plain peer stores), then the last CTA publishes a release/system-scope flag
// per peer; K3 on each shard acquires the G flags and sums the G slots in rank order — every rank ends
// with bit-identical spectra, no collective kernel, no host round trip.  Inboxes are double-buffered by
// block parity: a peer can run at most one block ahead (it needs my flag to finish its block).
struct PeerPub {
    float2 *inbox[FCB_MAX_PEERS];       // peer g's inbox base ([2][G][n_conv] float2)
    unsigned int *flags[FCB_MAX_PEERS]; // peer g's flags ([2][G])
    unsigned int *done;                 // local CTA counter for the last-CTA pattern
    long long n_conv;
    unsigned int seq;                   // block counter, 1-based
    int G, me;
    // reduce-scatter form (fcb_mimo_peer_set_scatter): shard g finishes rows [rows*g/G, rows*(g+1)/G) only, so a partial
    // row travels to its owner alone — 1/G of the all-gather form's NVLink bytes — and lands in the owner's slot
    // [parity][source][row - first owned row]
    int scatter, B;
    long long rows, slot_stride;        // rows = NS*OUT; slot_stride = float2 per (parity, source) slot
};

__device__ __forceinline__ void peer_store(const PeerPub &p, long long idx, float2 v)
{
    if (p.scatter) {
        const long long row = idx / p.B;
        const int g = (int)(((row + 1) * p.G - 1) / p.rows);          // owner of `row`
        const long long local = (row - p.rows * g / p.G) * p.B + idx % p.B;
        p.inbox[g][((long long)(p.seq & 1) * p.G + p.me) * p.slot_stride + local] = v;
        return;
    }
    const long long off = ((long long)(p.seq & 1) * p.G + p.me) * p.n_conv + idx;
    for (int g = 0; g < p.G; g++) p.inbox[g][off] = v;
}
// after every thread of the grid has stored: the last CTA to arrive raises my flag on every peer
__device__ __forceinline__ void peer_signal(const PeerPub &p, unsigned int ctas)
{
    __syncthreads(); // the CTA's stores happen-before thread 0's fences below (fences are cumulative)
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        if (atomicAdd(p.done, 1u) == ctas - 1) {
            *p.done = 0;
            __threadfence_system();
            for (int g = 0; g < p.G; g++) st_release_sys(p.flags[g] + (p.seq & 1) * p.G + p.me, p.seq);
        }
    }
}

// conv[s][o][k] = sum_in sum_z part[z][s][o][in][k]  +  sum_in X[s][in][cur][k] * H[o][in][seg 0][k]
// (the segment-0 product, src/fft_convolver.rs:256-261, only on the shard that owns segment 0).
// part_in == 0: the partial rows are already summed over `in` (k_mac_rt): part[z][s][o][k], added by lane 0.
// CTA = 32 bins x 8 input lanes of one (stream, out): lane y sums its inputs y, y+8, ... over all
// z chunks, then the 8 lane sums are added in fixed order — deterministic, and parallel enough
// that the Z*IN partial rows (19 MB at 16x16, Z = 19) stream at memory speed.
__global__ void __launch_bounds__(256)
k_mimo_reduce(const float2 *__restrict__ premul, const float2 *__restrict__ ring_cur, long long ring_stride,
              const float2 *__restrict__ ir0, long long ir_stride, float2 *__restrict__ conv, int B, int n_in,
              int n_out, long long n_so, int zchunks, int part_in, PeerPub pub)
{
    __shared__ float2 lane_sum[8][32];
    const int kt = (B + 31) / 32;
    const long long so = blockIdx.x / kt; // stream*OUT + out
    const int k = (blockIdx.x % kt) * 32 + threadIdx.x;
    const int y = threadIdx.y;
    const long long s = so / n_out, o = so % n_out;
    float ar = 0.f, ai = 0.f;
    if (k < B) {
        for (int in = y; in < n_in; in += 8) {
            float2 p = make_float2(0.f, 0.f);
            if (part_in) {
                p = premul[(so * n_in + in) * B + k];
                for (int z = 1; z < zchunks; z++) {
                    float2 q = premul[(((long long)z * n_so + so) * n_in + in) * B + k];
                    p.x += q.x;
                    p.y += q.y;
                }
            }
            if (ir0) {
                float2 x = ring_cur[(s * n_in + in) * ring_stride + k];
                float2 h = __ldg(&ir0[(o * n_in + in) * ir_stride + k]);
                float pr, pi;
                if (k == 0) {
                    pr = __fmul_rn(x.x, h.x);
                    pi = __fmul_rn(x.y, h.y);
                } else {
                    pr = __fsub_rn(__fmul_rn(x.x, h.x), __fmul_rn(x.y, h.y));
                    pi = __fadd_rn(__fmul_rn(x.x, h.y), __fmul_rn(x.y, h.x));
                }
                p.x = __fadd_rn(p.x, pr);
                p.y = __fadd_rn(p.y, pi);
            }
            ar = __fadd_rn(ar, p.x);
            ai = __fadd_rn(ai, p.y);
        }
        if (!part_in && y == 0)
            for (int z = 0; z < zchunks; z++) {
                float2 q = premul[((long long)z * n_so + so) * B + k];
                ar = __fadd_rn(ar, q.x);
                ai = __fadd_rn(ai, q.y);
            }
    }
    lane_sum[y][threadIdx.x] = make_float2(ar, ai);
    __syncthreads();
    if (y == 0 && k < B) {
        float2 t = lane_sum[0][threadIdx.x];
#pragma unroll
        for (int l = 1; l < 8; l++) {
            t.x = __fadd_rn(t.x, lane_sum[l][threadIdx.x].x);
            t.y = __fadd_rn(t.y, lane_sum[l][threadIdx.x].y);
        }
        if (pub.G > 0) peer_store(pub, so * B + k, t);
        else conv[so * B + k] = t;
    }
    if (pub.G > 0) peer_signal(pub, gridDim.x);
}

// tensor-core path: conv[so][k] = sum over input groups of part[g][so][k], ascending g (k_tc_reduce), peer variant
__global__ void __launch_bounds__(256)
k_tc_reduce_peer(const float2 *__restrict__ part, long long n, int groups, PeerPub pub)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx < n) {
        float2 t = part[idx];
        for (int g = 1; g < groups; g++) {
            float2 q = part[(long long)g * n + idx];
            t.x += q.x;
            t.y += q.y;
        }
        peer_store(pub, idx, t);
    }
    peer_signal(pub, gridDim.x);
}

} // namespace fcb

struct fcb_mimo {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    size_t n_in = 0, n_out = 0, n_streams = 1, B = 0, L = 0, S = 0;
    int logb = 0;
    size_t seg_lo = 0, seg_hi = 0; // IR segments owned by this shard
    size_t current = 0;            // ring slot of the next block (src/fft_convolver.rs:99)
    float2 *ir = nullptr;          // [OUT*IN][seg_hi-seg_lo][B]
    float2 *ring = nullptr;        // [NS*IN][S][B]
    float2 *premul = nullptr;      // [Z*NS*OUT*IN][B]  (Z segment chunks of the tile kernel, >= 1)
    int zmax = 1;
    // register-tiled matrix MAC (k_mac_rt, mimo_rt.cuh): 2+ streams below the tensor-core threshold
    bool rt = false;
    RtPlan rt_plan_;
    CUtensorMap tm_rt_ir, tm_rt_ring;
    float2 *conv = nullptr;        // [NS*OUT][B]  (this shard's partial until all-reduced)
    float *overlap = nullptr;      // [NS*OUT][B]
    float *io_in = nullptr, *io_out = nullptr; // staging for the host-pointer call
    float *stage = nullptr;
    size_t stage_floats = 0;
    const float2 *tw = nullptr;
    // tensor-core path (K4, mimo_tc.cuh): transposed operands instead of ir / ring / premul
    bool tc = false;
    float2 *ring_t = nullptr;  // [B][IN][nblk][128][16], slot = 16 * block + column
    float *ir_t = nullptr;     // [2 copies][B][IN][2*OUT][2*rowsP]; segment row r at position TC_LEAD + copy + r
    float2 *xcur = nullptr;    // [NS*IN][B] spectra of the current block before the scatter
    float2 *part_tc = nullptr; // [groups][NS*OUT][B]
    float2 *ir_tmp = nullptr;  // [tmp_pairs][rows][B] K5 output before the transposition
    size_t nblk = 0, rowsP = 0, tmp_pairs = 0, nsp = 0; // nsp: stream rows per ring tile
    size_t out_groups = 1, stream_groups = 1;           // groups of 16 outputs / 128 streams
    size_t ring_t_elems() const { return B * n_in * stream_groups * nblk * nsp * TC_KSEG; }
    int tc_groups = 1;
    CUtensorMap tm_ring, tm_ir[2];
    // K1 runs beside the MAC: the MAC only reads ring slots older than the current block (segments >= 1)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_k1 = nullptr;
    // peer exchange (see PeerPub)
    size_t shard_index = 0, shard_count = 1;
    unsigned char *inbox = nullptr;      // [2][G][n_conv] float2 then [2][G] flags, one allocation (IPC-exported)
    unsigned int *peer_done = nullptr;   // CTA counter of the last-CTA pattern
    int *peer_err_h = nullptr, *peer_err_d = nullptr; // error word in mapped pinned memory (host view / device alias):
                                                      // K3 sets it when a flag never arrives, every later call reads it
    void *peer_base[FCB_MAX_PEERS] = {}; // every shard's inbox in my address space
    bool peer_opened[FCB_MAX_PEERS] = {};
    bool peer_on = false;
    bool peer_scatter = false;           // reduce-scatter form: this shard finishes (and outputs) its own rows only
    unsigned int peer_seq = 0;
    // overlapped finish (fcb_mimo_set_overlap, peer exchange only): K3 of block n — the kernel that waits for the peers —
    // runs on its own stream beside K1 / the MAC of block n + 1; only the reduce of block n + 1 waits for it
    bool ovl_fin = false, fin_pending = false;
    cudaStream_t fin = nullptr;
    cudaEvent_t ev_red = nullptr, ev_fin = nullptr;
    size_t rows_total() const { return n_streams * n_out; }
    size_t row_lo() const { return rows_total() * shard_index / shard_count; }
    size_t row_hi() const { return rows_total() * (shard_index + 1) / shard_count; }
    size_t slot_stride() const { return peer_scatter ? ((rows_total() + shard_count - 1) / shard_count) * B : n_conv(); }
    size_t n_conv() const { return n_streams * n_out * B; }
    size_t inbox_data_bytes() const { return 2 * shard_count * n_conv() * sizeof(float2); }
    PeerPub pub() const
    {
        PeerPub p{};
        if (!peer_on) return p;
        for (size_t g = 0; g < shard_count; g++) {
            p.inbox[g] = reinterpret_cast<float2 *>(peer_base[g]);
            p.flags[g] = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned char *>(peer_base[g]) + inbox_data_bytes());
        }
        p.done = peer_done;
        p.n_conv = (long long)n_conv();
        p.seq = peer_seq;
        p.G = (int)shard_count;
        p.me = (int)shard_index;
        p.scatter = peer_scatter ? 1 : 0;
        p.B = (int)B;
        p.rows = (long long)rows_total();
        p.slot_stride = (long long)slot_stride();
        return p;
    }
    size_t ir_copy_floats() const { return B * n_in * 32 * out_groups * 2 * rowsP; }

    size_t rows() const { return seg_hi - seg_lo; }
};

// Overlapped finish.  The reduce of block n + 1 publishes into the inbox slots (parity n + 1) and, two blocks on, a
// peer's reduce overwrites the slots K3 of block n reads: the protocol's "a peer runs at most one block ahead" rests on
// every shard's reduce coming after its own previous K3.  With K3 on its own stream that order is kept by two events:
// reduce(n + 1) waits for K3(n), K3(n) waits for reduce(n) — and nothing else does, so K1 and the MAC of block n + 1 run
// while K3 of block n waits for the peers' flags.
static int overlap_before_reduce(fcb_mimo *m)
{
    if (m->ovl_fin && m->fin_pending) FCB_CUDA(cudaStreamWaitEvent(m->stream, m->ev_fin, 0));
    return FCB_OK;
}
static int overlap_after_reduce(fcb_mimo *m)
{
    if (m->ovl_fin) FCB_CUDA(cudaEventRecord(m->ev_red, m->stream));
    return FCB_OK;
}
// everything queued on the convolver's stream from here on comes after every K3 launched so far
static int overlap_join(fcb_mimo *m)
{
    if (m->fin_pending) FCB_CUDA(cudaStreamWaitEvent(m->stream, m->ev_fin, 0));
    return FCB_OK;
}

extern "C" void fcb_mimo_destroy(fcb_mimo *m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->fin) {
        cudaStreamSynchronize(m->fin);
        cudaStreamDestroy(m->fin);
    }
    if (m->ev_red) cudaEventDestroy(m->ev_red);
    if (m->ev_fin) cudaEventDestroy(m->ev_fin);
    cudaFree(m->ir);
    cudaFree(m->ring);
    cudaFree(m->premul);
    cudaFree(m->conv);
    cudaFree(m->overlap);
    cudaFree(m->io_in);
    cudaFree(m->io_out);
    cudaFree(m->stage);
    cudaFree(m->ring_t);
    cudaFree(m->ir_t);
    cudaFree(m->xcur);
    cudaFree(m->part_tc);
    cudaFree(m->ir_tmp);
    for (size_t g = 0; g < FCB_MAX_PEERS; g++)
        if (m->peer_opened[g]) cudaIpcCloseMemHandle(m->peer_base[g]);
    cudaFree(m->inbox);
    cudaFree(m->peer_done);
    if (m->peer_err_h) cudaFreeHost(m->peer_err_h);
    if (m->side) cudaStreamDestroy(m->side);
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_k1) cudaEventDestroy(m->ev_k1);
    if (m->own_stream && m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

static int mimo_alloc(void **p, size_t bytes, cudaStream_t s)
{
    FCB_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    FCB_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, s));
    return FCB_OK;
}

extern "C" int fcb_mimo_create(const fcb_mimo_desc *d, fcb_mimo **out)
{
    if (!d || !out) return fail(FCB_ERR_ARG, "fcb_mimo_create: NULL argument");
    *out = nullptr;
    if (!d->n_in || !d->n_out) return fail(FCB_ERR_ARG, "fcb_mimo_create: empty matrix");
    const size_t ns = d->n_streams ? d->n_streams : 1, shards = d->shard_count ? d->shard_count : 1;
    if (d->shard_index >= shards) return fail(FCB_ERR_ARG, "fcb_mimo_create: shard %zu of %zu", d->shard_index, shards);
    const size_t B = next_power_of_two(d->block_size);
    if (B > 16384) return fail(FCB_ERR_UNSUPPORTED, "block size %zu > 16384 not supported", B);
    FCB_CUDA(cudaSetDevice(d->device));
    fcb_mimo *m = new fcb_mimo();
    m->device = d->device;
    m->n_in = d->n_in;
    m->n_out = d->n_out;
    m->n_streams = ns;
    m->B = B;
    m->logb = ilog2(B);
    m->L = d->max_response_length;
    m->S = (size_t)std::ceil((double)m->L / (double)B);
    m->shard_index = d->shard_index;
    m->shard_count = shards;
    m->seg_lo = m->S * d->shard_index / shards;
    m->seg_hi = m->S * (d->shard_index + 1) / shards;
    if (d->stream) m->stream = (cudaStream_t)d->stream;
    else {
        cudaError_t err = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
        if (err != cudaSuccess) {
            delete m;
            return fail(FCB_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(err));
        }
        m->own_stream = true;
    }
    const size_t pairs = m->n_out * m->n_in;
    int rc = get_twiddles(m->device, 2 * B, &m->tw);
    const int tc_mode = g_mimo_tc.load();
    m->tc = tc_mode != 0 && (tc_mode == 1 || ns >= (size_t)g_mimo_tc_min.load()) && m->rows() > 0 && B >= 2;
    if (m->tc) {
        m->nblk = (m->S + TC_KSEG - 1) / TC_KSEG;
        m->stream_groups = (ns + TC_M - 1) / TC_M;
        m->out_groups = (m->n_out + 15) / 16;
        m->nsp = m->stream_groups > 1 ? (size_t)TC_M : (ns + 7) & ~(size_t)7;
        m->rowsP = (TC_LEAD + 1 + m->rows() + 1) & ~(size_t)1; // positions per IR row; pitch a multiple of 16 bytes
        const size_t per_bin = m->out_groups * m->stream_groups;
        size_t groups = (6 * 148 + B * per_bin - 1) / (B * per_bin); // ~6 waves of CTAs
        m->tc_groups = (int)(groups < 1 ? 1 : groups > m->n_in ? m->n_in : groups);
        const size_t per_pair = m->rows() * B * sizeof(float2);
        m->tmp_pairs = ((size_t)256 << 20) / per_pair;
        if (m->tmp_pairs < 1) m->tmp_pairs = 1;
        if (m->tmp_pairs > pairs) m->tmp_pairs = pairs;
        if (!rc) rc = mimo_alloc((void **)&m->ring_t, m->ring_t_elems() * sizeof(float2), m->stream);
        if (!rc) rc = mimo_alloc((void **)&m->ir_t, 2 * m->ir_copy_floats() * sizeof(float), m->stream);
        if (!rc) rc = mimo_alloc((void **)&m->xcur, ns * m->n_in * B * sizeof(float2), m->stream);
        if (!rc) rc = mimo_alloc((void **)&m->part_tc, (size_t)m->tc_groups * ns * m->n_out * B * sizeof(float2), m->stream);
        if (!rc) rc = mimo_alloc((void **)&m->ir_tmp, m->tmp_pairs * per_pair, m->stream);
        const size_t tile = m->nsp * TC_KSEG * sizeof(float2);
        if (!rc)
            rc = tc_encode_map(&m->tm_ring, m->ring_t, 2 * TC_KSEG, m->nsp, m->stream_groups * m->nblk, B * m->n_in,
                               TC_KSEG * sizeof(float2), tile, m->stream_groups * m->nblk * tile, (uint32_t)m->nsp);
        for (size_t sh = 0; sh < 2 && !rc; sh++)
            rc = tc_encode_map(&m->tm_ir[sh], m->ir_t + sh * m->ir_copy_floats(), 2 * (TC_LEAD + sh + m->rows()), 32 * m->out_groups,
                               B * m->n_in, 0, 2 * m->rowsP * sizeof(float), 32 * m->out_groups * 2 * m->rowsP * sizeof(float), 0, 32);
        if (!rc && cudaFuncSetAttribute(k_mimo_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<16>::SMEM) != cudaSuccess)
            rc = fail(FCB_ERR_CUDA, "k_mimo_tc: cannot opt in to %zu bytes of shared memory", TcCfg<16>::SMEM);
    } else {
        if (!rc) rc = mimo_alloc((void **)&m->ir, pairs * m->rows() * B * sizeof(float2), m->stream);
        if (!rc) rc = mimo_alloc((void **)&m->ring, ns * m->n_in * m->S * B * sizeof(float2), m->stream);
        int z = 1, zl = 1;
        if (mac_tile_plan(m->logb, (int)m->n_in, (int)m->n_out, (int)ns, (int)m->rows(), &z, &zl) == FCB_OK) m->zmax = z;
        size_t part_rows = (size_t)m->zmax * ns * pairs;
        const size_t first = m->seg_lo > 1 ? m->seg_lo : 1; // segment 0 belongs to the reduce kernel
        m->rt = g_mimo_rt.load() && ns >= (size_t)g_mimo_rt_min.load() && B >= (size_t)RT_BINS && m->seg_hi > first;
        if (m->rt) {
            m->rt_plan_ = rt_plan((int)B, (int)m->n_out, (int)ns, (int)(m->seg_hi - first));
            const size_t need = (size_t)m->rt_plan_.zchunks * ns * m->n_out;
            if (need > part_rows) part_rows = need;
        }
        if (!rc) rc = mimo_alloc((void **)&m->premul, part_rows * B * sizeof(float2), m->stream);
        if (!rc && m->rt) {
            const uint64_t row = B * sizeof(float2), per_pair = m->rows() * row, per_ring = m->S * row;
            rc = rt_encode_map(&m->tm_rt_ir, m->ir + (first - m->seg_lo) * B, 2 * B, m->seg_hi - first, m->n_in, m->n_out, row,
                               per_pair, m->n_in * per_pair, 2 * RT_BINS * m->rt_plan_.wb, m->rt_plan_.r, 8 * m->rt_plan_.wo);
            if (!rc)
                rc = rt_encode_map(&m->tm_rt_ring, m->ring, 2 * B, m->S, m->n_in, ns, row, per_ring, m->n_in * per_ring,
                                   2 * RT_BINS * m->rt_plan_.wb, m->rt_plan_.r, 4 * m->rt_plan_.ws);
        }
    }
    if (!rc) rc = mimo_alloc((void **)&m->conv, ns * m->n_out * B * sizeof(float2), m->stream);
    if (!rc) rc = mimo_alloc((void **)&m->overlap, ns * m->n_out * B * sizeof(float), m->stream);
    if (!rc) rc = mimo_alloc((void **)&m->io_in, ns * m->n_in * B * sizeof(float), m->stream);
    if (!rc) rc = mimo_alloc((void **)&m->io_out, ns * m->n_out * B * sizeof(float), m->stream);
    const size_t per = m->L ? m->L : 1, cap = (size_t)16 << 20;
    m->stage_floats = pairs * per < cap ? pairs * per : (cap / per ? (cap / per) * per : per);
    if (!rc) rc = mimo_alloc((void **)&m->stage, m->stage_floats * sizeof(float), m->stream);
    if (!rc && (cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&m->ev_k1, cudaEventDisableTiming) != cudaSuccess))
        rc = fail(FCB_ERR_CUDA, "mimo: side stream / events");
    if (!rc && cudaStreamSynchronize(m->stream) != cudaSuccess) rc = fail(FCB_ERR_CUDA, "sync failed");
    if (rc) {
        fcb_mimo_destroy(m);
        return rc;
    }
    *out = m;
    return FCB_OK;
}

// K5 over this shard's segment range: irs = [OUT][IN][len] host, stride `len`
extern "C" int fcb_mimo_set_ir(fcb_mimo *m, const float *irs, size_t len)
{
    if (!m) return fail(FCB_ERR_ARG, "NULL mimo");
    if (len > m->L) return fail(FCB_ERR_PANIC, "max_response_length must be at least the length of the initial impulse response");
    if (len && !irs) return fail(FCB_ERR_ARG, "NULL impulse responses");
    if (m->rows() == 0) return FCB_OK;
    FCB_CUDA(cudaSetDevice(m->device));
    const size_t pairs = m->n_out * m->n_in, B = m->B;
    const long long dst_stride = (long long)(m->rows() * B);
    const size_t off = m->seg_lo * B;                 // first sample of this shard's first segment
    const int rem = len > off ? (int)(len - off) : 0; // samples of the IR at or past it
    if (len == 0) {
        if (m->tc) {
            FCB_CUDA(cudaMemsetAsync(m->ir_t, 0, 2 * m->ir_copy_floats() * sizeof(float), m->stream));
            return FCB_OK;
        }
        return run_forward(m->logb, m->tw, m->stream, m->stage, 0, 0, m->ir, dst_stride, (int)m->rows(),
                           (long long)(pairs * m->rows()));
    }
    size_t per_group = m->stage_floats / len;
    if (per_group == 0) return fail(FCB_ERR_CUDA, "IR staging buffer too small");
    if (m->tc && per_group > m->tmp_pairs) per_group = m->tmp_pairs;
    for (size_t g0 = 0; g0 < pairs; g0 += per_group) {
        const size_t g = pairs - g0 < per_group ? pairs - g0 : per_group;
        FCB_CUDA(cudaMemcpyAsync(m->stage, irs + g0 * len, g * len * sizeof(float), cudaMemcpyHostToDevice, m->stream));
        float2 *dst = m->tc ? m->ir_tmp : m->ir + g0 * dst_stride;
        FCB_TRY(run_forward(m->logb, m->tw, m->stream, m->stage + off, (long long)len, rem, dst, dst_stride, (int)m->rows(),
                            (long long)(g * m->rows())));
        if (m->tc) {
            const long long total = (long long)(g * m->rows() * B);
            k_tc_build_ir<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(m->ir_tmp, m->ir_t, (int)B, (int)m->n_in,
                                                                                   (int)m->n_out, (int)m->rows(), (long long)m->rowsP,
                                                                                   (long long)g0, total, (long long)m->ir_copy_floats());
            g_launches++;
            FCB_CUDA(cudaGetLastError());
        }
        if (g0 + per_group < pairs) FCB_CUDA(cudaStreamSynchronize(m->stream));
    }
    return FCB_OK;
}

extern "C" int fcb_mimo_reset(fcb_mimo *m)
{
    if (!m) return fail(FCB_ERR_ARG, "NULL mimo");
    FCB_CUDA(cudaSetDevice(m->device));
    FCB_TRY(overlap_join(m)); // K3 of the last block may still be adding into `overlap` on its own stream
    const size_t ns = m->n_streams, B = m->B;
    if (m->tc) FCB_CUDA(cudaMemsetAsync(m->ring_t, 0, m->ring_t_elems() * sizeof(float2), m->stream));
    else FCB_CUDA(cudaMemsetAsync(m->ring, 0, ns * m->n_in * m->S * B * sizeof(float2), m->stream));
    FCB_CUDA(cudaMemsetAsync(m->overlap, 0, ns * m->n_out * B * sizeof(float), m->stream));
    m->current = 0;
    return FCB_OK;
}

// Peer exchange health: K3 raises the mapped error word when a shard's flag never arrives (it then emits silence for
// that block instead of summing a stale inbox).  Reading it is one host load, so every partial / finish / sync call
// checks it and fails loudly from the first call after the fault on — no synchronisation needed.
static int peer_check(const fcb_mimo *m)
{
    if (m->peer_on && m->peer_err_h && *reinterpret_cast<volatile int *>(m->peer_err_h))
        return fail(FCB_ERR_CUDA, "peer exchange: a shard's partial spectra never arrived (flag wait timed out); "
                                  "output blocks since then are silence");
    return FCB_OK;
}

// K1 on the NS*IN input blocks, K2 over this shard's segments, reduction over `in` -> conv
extern "C" int fcb_mimo_partial_dev(fcb_mimo *m, const float *in_dev, size_t in_stride)
{
    if (!m || !in_dev) return fail(FCB_ERR_ARG, "fcb_mimo_partial_dev: NULL argument");
    if (m->S == 0) return FCB_OK;
    FCB_TRY(peer_check(m));
    FCB_CUDA(cudaSetDevice(m->device));
    const size_t B = m->B, ns = m->n_streams, pairs = m->n_out * m->n_in;
    if (m->peer_on) m->peer_seq++;
    if (m->tc) {
        // K1 -> scatter into column `current` of the K-major ring -> K4 (every owned segment,
        // segment 0 included) -> sum over input groups
        FCB_TRY(run_forward(m->logb, m->tw, m->stream, in_dev, (long long)in_stride, (int)B, m->xcur, (long long)B, 1,
                            (long long)(ns * m->n_in)));
        const long long nx = (long long)(ns * m->n_in * B);
        k_tc_scatter_ring<<<(unsigned)((nx + 255) / 256), 256, 0, m->stream>>>(m->xcur, m->ring_t, (int)B, (int)m->n_in, nx,
                                                                                (long long)m->nblk, (int)m->current, (int)m->nsp,
                                                                                (int)m->stream_groups);
        TcArgs t{};
        t.part = m->part_tc;
        t.B = (int)B;
        t.n_in = (int)m->n_in;
        t.n_streams = (int)ns;
        t.n_out = (int)m->n_out;
        t.out_groups = (int)m->out_groups;
        t.stream_groups = (int)m->stream_groups;
        t.nblk = (int)m->nblk;
        t.rows_pad = (int)m->nsp;
        t.S = (int)m->S;
        t.current = (int)m->current;
        t.seg_lo = (int)m->seg_lo;
        t.seg_hi = (int)m->seg_hi;
        t.groups = m->tc_groups;
        cudaEvent_t prof_stop = nullptr;
        const bool profiled = mac_profile_begin(m->stream, &prof_stop) != nullptr;
        k_mimo_tc<16><<<(unsigned)(B * m->tc_groups * m->out_groups * m->stream_groups), TC_THREADS, TcCfg<16>::SMEM, m->stream>>>(t, m->tm_ring, m->tm_ir[0], m->tm_ir[1]);
        if (profiled) cudaEventRecord(prof_stop, m->stream);
        const long long nc = (long long)(ns * m->n_out * B);
        FCB_TRY(overlap_before_reduce(m));
        if (m->peer_on) k_tc_reduce_peer<<<(unsigned)((nc + 255) / 256), 256, 0, m->stream>>>(m->part_tc, nc, m->tc_groups, m->pub());
        else k_tc_reduce<<<(unsigned)((nc + 255) / 256), 256, 0, m->stream>>>(m->part_tc, m->conv, nc, m->tc_groups);
        g_launches += 3;
        FCB_CUDA(cudaGetLastError());
        return overlap_after_reduce(m);
    }
    const long long ring_stride = (long long)(m->S * B);
    // K1 writes ring slot `current`; the MAC below reads only the older slots (segments >= 1) and the reduce kernel is
    // the first to need the new spectra (segment 0): K1 runs on a side stream beside the MAC and joins before the reduce
    FCB_CUDA(cudaEventRecord(m->ev_fork, m->stream)); // after the previous block's last reader of that slot
    FCB_CUDA(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
    FCB_TRY(run_forward(m->logb, m->tw, m->side, in_dev, (long long)in_stride, (int)B, m->ring + m->current * B,
                        ring_stride, 1, (long long)(ns * m->n_in)));
    FCB_CUDA(cudaEventRecord(m->ev_k1, m->side));
    const int seg_lo = (int)(m->seg_lo > 1 ? m->seg_lo : 1), seg_hi = (int)m->seg_hi;
    const long long ir_stride = (long long)(m->rows() * B);
    int zchunks = 1, part_in = 1;
    bool done = false;
    if (m->rt) {
        RtArgs t{};
        t.part = m->premul;
        t.B = (int)B;
        t.n_in = (int)m->n_in;
        t.n_out = (int)m->n_out;
        t.n_streams = (int)ns;
        t.S = (int)m->S;
        t.current = (int)m->current;
        t.seg_lo = seg_lo;
        t.seg_hi = seg_hi;
        t.seg_base = seg_lo;
        t.zchunks = m->rt_plan_.zchunks;
        t.out_groups = m->rt_plan_.out_groups;
        t.stream_groups = m->rt_plan_.stream_groups;
        FCB_TRY(rt_launch(m->rt_plan_, t, m->tm_rt_ir, m->tm_rt_ring, m->stream));
        zchunks = t.zchunks;
        part_in = 0;
        done = true;
    }
    if (!done && g_mimo_tile.load() && m->zmax >= 1 && seg_hi > seg_lo) {
        MacTileArgs t{};
        t.ir = m->ir;
        t.ir_stride = ir_stride;
        t.ring = m->ring;
        t.ring_stride = ring_stride;
        t.part = m->premul;
        t.current = (int)m->current;
        t.active = (int)m->S;
        t.seg_lo = seg_lo;
        t.seg_hi = seg_hi;
        t.ir_seg0 = (int)m->seg_lo;
        t.n_in = (int)m->n_in;
        t.n_out = (int)m->n_out;
        t.n_streams = (int)ns;
        int rc = run_mac_tile(m->logb, m->stream, t, &zchunks);
        if (rc == FCB_OK) done = true;
        else if (rc != FCB_ERR_UNSUPPORTED) return rc;
        if (done && zchunks > m->zmax) return fail(FCB_ERR_CUDA, "mimo: tile plan grew past its workspace");
    }
    if (!done) {
        zchunks = 1;
        MacArgs a{};
        a.ir = m->ir;
        a.ir_stride = ir_stride;
        a.ring = m->ring;
        a.ring_stride = ring_stride;
        a.premul = m->premul;
        a.current = (int)m->current;
        a.active = (int)m->S;
        a.nchan = (long long)(ns * pairs);
        a.seg_lo = seg_lo;
        a.seg_hi = seg_hi;
        a.ir_seg0 = (int)m->seg_lo;
        a.ir_mod = (long long)pairs;
        a.ring_div = (long long)pairs;
        a.ring_mul = (long long)m->n_in;
        a.ring_mod = (long long)m->n_in;
        FCB_TRY(run_mac(m->logb, m->stream, a));
    }
    FCB_CUDA(cudaStreamWaitEvent(m->stream, m->ev_k1, 0));
    FCB_TRY(overlap_before_reduce(m));
    const bool owns0 = m->seg_lo == 0 && m->seg_hi > 0;
    const long long n_so = (long long)(ns * m->n_out);
    k_mimo_reduce<<<(unsigned)(n_so * ((B + 31) / 32)), dim3(32, 8), 0, m->stream>>>(
        m->premul, m->ring + m->current * B, ring_stride, owns0 ? m->ir : nullptr, ir_stride, m->conv, (int)B,
        (int)m->n_in, (int)m->n_out, n_so, zchunks, part_in, m->pub());
    g_launches++;
    FCB_CUDA(cudaGetLastError());
    return overlap_after_reduce(m);
}

extern "C" float *fcb_mimo_conv_buffer(fcb_mimo *m, size_t *n_floats)
{
    if (!m) return nullptr;
    if (n_floats) *n_floats = 2 * m->n_streams * m->n_out * m->B;
    return reinterpret_cast<float *>(m->conv);
}

// K3 on rows [row_lo, row_hi) of the (reduced) conv: inverse FFT, /N, overlap-add, overlap save; rotates `current`.
// out_dev is the base of the FULL [NS*OUT][B] output (row r lands at out_dev + r*out_stride).
static int mimo_finish_rows(fcb_mimo *m, float *out_dev, size_t out_stride, size_t row_lo, size_t row_hi)
{
    FCB_TRY(peer_check(m));
    FCB_CUDA(cudaSetDevice(m->device));
    const size_t B = m->B;
    if (m->S == 0) {
        if (row_hi > row_lo)
            FCB_CUDA(cudaMemset2DAsync(out_dev + row_lo * out_stride, out_stride * sizeof(float), 0, B * sizeof(float), row_hi - row_lo, m->stream));
        return FCB_OK;
    }
    if (row_hi > row_lo) {
        IfftArgs a{};
        a.ir0 = nullptr; // conv is complete
        a.premul = m->conv + row_lo * B;
        a.overlap = m->overlap + row_lo * B;
        a.out = out_dev + row_lo * out_stride;
        a.out_stride = (long long)out_stride;
        a.fill = 0;
        a.n = (int)B;
        a.block_complete = 1;
        a.nchan = (long long)(row_hi - row_lo);
        if (m->peer_on) { // sum the G shards' partial spectra out of my inbox once their flags have arrived
            const size_t G = m->shard_count, par = m->peer_seq & 1, stride = m->slot_stride();
            a.gather = reinterpret_cast<const float2 *>(m->inbox) + par * G * stride + (m->peer_scatter ? 0 : row_lo * B);
            a.gather_stride = (long long)stride;
            a.gather_flags = reinterpret_cast<const unsigned int *>(m->inbox + m->inbox_data_bytes()) + par * G;
            a.gather_seq = m->peer_seq;
            a.gather_n = (int)G;
            a.gather_err = m->peer_err_d;
        }
        if (m->ovl_fin && m->peer_on) {
            FCB_CUDA(cudaStreamWaitEvent(m->fin, m->ev_red, 0));
            FCB_TRY(run_inverse(m->logb, m->tw, m->fin, a));
            FCB_CUDA(cudaEventRecord(m->ev_fin, m->fin));
            m->fin_pending = true;
        } else {
            FCB_TRY(run_inverse(m->logb, m->tw, m->stream, a));
        }
    }
    m->current = m->current > 0 ? m->current - 1 : m->S - 1; // src/fft_convolver.rs:287-291
    return FCB_OK;
}

// Overlapped finish (see overlap_before_reduce): set on EVERY shard before the first block, peer exchange only.
// While on, the output of a block is complete — in the order of the convolver's stream — after fcb_mimo_join, and for
// the host after fcb_mimo_sync (inputs are taken in stream order as before: K1 and the MAC stay on that stream).
extern "C" int fcb_mimo_set_overlap(fcb_mimo *m, int on)
{
    if (!m) return fail(FCB_ERR_ARG, "NULL mimo");
    if (m->peer_seq != 0) return fail(FCB_ERR_ARG, "fcb_mimo_set_overlap: set before the first block");
    FCB_CUDA(cudaSetDevice(m->device));
    if (on && !m->fin) {
        FCB_CUDA(cudaStreamCreateWithFlags(&m->fin, cudaStreamNonBlocking));
        FCB_CUDA(cudaEventCreateWithFlags(&m->ev_red, cudaEventDisableTiming));
        FCB_CUDA(cudaEventCreateWithFlags(&m->ev_fin, cudaEventDisableTiming));
    }
    m->ovl_fin = on != 0;
    return FCB_OK;
}
// the convolver's stream waits (on the device) for every K3 launched so far
extern "C" int fcb_mimo_join(fcb_mimo *m)
{
    if (!m) return fail(FCB_ERR_ARG, "NULL mimo");
    FCB_CUDA(cudaSetDevice(m->device));
    return overlap_join(m);
}

// every row — or, in the reduce-scatter form of the peer exchange, the rows this shard owns
extern "C" int fcb_mimo_finish_dev(fcb_mimo *m, float *out_dev, size_t out_stride)
{
    if (!m || !out_dev) return fail(FCB_ERR_ARG, "fcb_mimo_finish_dev: NULL argument");
    if (m->peer_on && m->peer_scatter) return mimo_finish_rows(m, out_dev, out_stride, m->row_lo(), m->row_hi());
    return mimo_finish_rows(m, out_dev, out_stride, 0, m->rows_total());
}

// K3 for rows [row_lo, row_hi) only (the caller reduce-scattered the conv buffer: NCCL reduce_scatter in place of all-reduce)
extern "C" int fcb_mimo_finish_rows_dev(fcb_mimo *m, float *out_dev, size_t out_stride, size_t row_lo, size_t row_hi)
{
    if (!m || !out_dev) return fail(FCB_ERR_ARG, "fcb_mimo_finish_rows_dev: NULL argument");
    if (row_lo > row_hi || row_hi > m->rows_total()) return fail(FCB_ERR_ARG, "fcb_mimo_finish_rows_dev: rows [%zu, %zu) of %zu", row_lo, row_hi, m->rows_total());
    if (m->peer_on) return fail(FCB_ERR_ARG, "fcb_mimo_finish_rows_dev: the peer exchange picks the rows itself (fcb_mimo_peer_set_scatter)");
    return mimo_finish_rows(m, out_dev, out_stride, row_lo, row_hi);
}

// reduce-scatter form of the peer exchange: set on EVERY shard before the first block
extern "C" int fcb_mimo_peer_set_scatter(fcb_mimo *m, int on)
{
    if (!m) return fail(FCB_ERR_ARG, "NULL mimo");
    if (m->peer_seq != 0) return fail(FCB_ERR_ARG, "fcb_mimo_peer_set_scatter: set before the first block");
    m->peer_scatter = on != 0;
    return FCB_OK;
}
extern "C" int fcb_mimo_owned_rows(const fcb_mimo *m, size_t *lo, size_t *hi)
{
    if (!m) return fail(FCB_ERR_ARG, "NULL mimo");
    if (lo) *lo = m->row_lo();
    if (hi) *hi = m->row_hi();
    return FCB_OK;
}

// single-shard convenience with host buffers: in [NS*IN][B], out [NS*OUT][B]
extern "C" int fcb_mimo_process(fcb_mimo *m, const float *in, float *out)
{
    if (!m || !in || !out) return fail(FCB_ERR_ARG, "fcb_mimo_process: NULL argument");
    if (m->seg_lo != 0 || m->seg_hi != m->S)
        return fail(FCB_ERR_ARG, "fcb_mimo_process: this object is one shard of several; use partial/all-reduce/finish");
    FCB_CUDA(cudaSetDevice(m->device));
    const size_t B = m->B, ns = m->n_streams;
    FCB_CUDA(cudaMemcpyAsync(m->io_in, in, ns * m->n_in * B * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    FCB_TRY(fcb_mimo_partial_dev(m, m->io_in, B));
    FCB_TRY(fcb_mimo_finish_dev(m, m->io_out, B));
    FCB_CUDA(cudaMemcpyAsync(out, m->io_out, ns * m->n_out * B * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    FCB_CUDA(cudaStreamSynchronize(m->stream));
    return FCB_OK;
}

extern "C" int fcb_mimo_sync(fcb_mimo *m)
{
    if (!m) return fail(FCB_ERR_ARG, "NULL mimo");
    FCB_CUDA(cudaSetDevice(m->device));
    FCB_CUDA(cudaStreamSynchronize(m->stream));
    if (m->fin) FCB_CUDA(cudaStreamSynchronize(m->fin));
    return peer_check(m);
}

// ---- peer exchange set-up -------------------------------------------------------------------
static int peer_alloc(fcb_mimo *m)
{
    if (m->inbox) return FCB_OK;
    if (m->shard_count > FCB_MAX_PEERS) return fail(FCB_ERR_UNSUPPORTED, "peer exchange supports up to %d shards", FCB_MAX_PEERS);
    FCB_CUDA(cudaSetDevice(m->device));
    const size_t bytes = m->inbox_data_bytes() + 256;
    FCB_CUDA(cudaMalloc((void **)&m->inbox, bytes));
    FCB_CUDA(cudaMemset(m->inbox, 0, bytes));
    FCB_CUDA(cudaMalloc((void **)&m->peer_done, 2 * sizeof(unsigned int)));
    FCB_CUDA(cudaMemset(m->peer_done, 0, 2 * sizeof(unsigned int)));
    FCB_CUDA(cudaHostAlloc((void **)&m->peer_err_h, sizeof(int), cudaHostAllocMapped));
    *m->peer_err_h = 0;
    FCB_CUDA(cudaHostGetDevicePointer((void **)&m->peer_err_d, m->peer_err_h, 0));
    FCB_CUDA(cudaDeviceSynchronize());
    return FCB_OK;
}

extern "C" int fcb_mimo_peer_export(fcb_mimo *m, unsigned char *handle_out)
{
    if (!m || !handle_out) return fail(FCB_ERR_ARG, "fcb_mimo_peer_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == FCB_PEER_HANDLE_BYTES, "IPC handle size");
    FCB_TRY(peer_alloc(m));
    cudaIpcMemHandle_t h;
    FCB_CUDA(cudaIpcGetMemHandle(&h, m->inbox));
    memcpy(handle_out, &h, sizeof h);
    return FCB_OK;
}

extern "C" int fcb_mimo_peer_attach(fcb_mimo *m, const unsigned char *handles)
{
    if (!m || !handles) return fail(FCB_ERR_ARG, "fcb_mimo_peer_attach: NULL argument");
    if (!m->inbox) return fail(FCB_ERR_ARG, "fcb_mimo_peer_attach: call fcb_mimo_peer_export first");
    FCB_CUDA(cudaSetDevice(m->device));
    for (size_t g = 0; g < m->shard_count; g++) {
        if (g == m->shard_index) {
            m->peer_base[g] = m->inbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + g * FCB_PEER_HANDLE_BYTES, sizeof h);
        FCB_CUDA(cudaIpcOpenMemHandle(&m->peer_base[g], h, cudaIpcMemLazyEnablePeerAccess));
        m->peer_opened[g] = true;
    }
    m->peer_on = true;
    return FCB_OK;
}

extern "C" void *fcb_mimo_peer_inbox(fcb_mimo *m)
{
    if (!m || peer_alloc(m) != FCB_OK) return nullptr;
    return m->inbox;
}

// same-process variant (tests: several shards of one job on one GPU): inboxes as plain device pointers
extern "C" int fcb_mimo_peer_attach_ptrs(fcb_mimo *m, void *const *inboxes)
{
    if (!m || !inboxes) return fail(FCB_ERR_ARG, "fcb_mimo_peer_attach_ptrs: NULL argument");
    FCB_TRY(peer_alloc(m));
    for (size_t g = 0; g < m->shard_count; g++) m->peer_base[g] = g == m->shard_index ? (void *)m->inbox : inboxes[g];
    m->peer_on = true;
    return FCB_OK;
}

extern "C" void *fcb_mimo_stream(fcb_mimo *m) { return m ? (void *)m->stream : nullptr; }
// K4's stage enumeration for one input (host copy of what the TMA producer computes): for stage c the ring slot
// block, the IR copy and the IR position paired with the block's first slot.  No device needed (tests).
extern "C" int fcb_debug_tc_stages(int S, int current, int seg_lo, int seg_hi, int max_stages, int *blk, int *copy, int *pos0)
{
    TcArgs a{};
    a.S = S;
    a.current = current;
    a.seg_lo = seg_lo;
    a.seg_hi = seg_hi;
    const TcSpan span(a);
    const int n = span.per_input();
    for (int c = 0; c < n && c < max_stages; c++) span.chunk(c, blk[c], copy[c], pos0[c]);
    return n;
}

extern "C" int fcb_mimo_uses_tensor_cores(const fcb_mimo *m) { return !m ? 0 : m->tc ? 1 : m->rt ? 2 : 0; }
extern "C" size_t fcb_mimo_block_size(const fcb_mimo *m) { return m->B; }
extern "C" size_t fcb_mimo_seg_count(const fcb_mimo *m) { return m->S; }
extern "C" int fcb_mimo_segment_range(const fcb_mimo *m, size_t *lo, size_t *hi)
{
    if (!m) return fail(FCB_ERR_ARG, "NULL mimo");
    if (lo) *lo = m->seg_lo;
    if (hi) *hi = m->seg_hi;
    return FCB_OK;
}
