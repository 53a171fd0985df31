// engine_internal.cuh — launch entry points shared by engine.cu and mimo.cu (not part of the C ABI)
#pragma once

#include "common.cuh"
#include "fft_kernels.cuh"
#include "mac_kernels.cuh"
#include "fused_kernel.cuh"

namespace fcb {

// twiddle table tw[t] = exp(-2 pi i t / N), t < N (device memory, cached per (device, N))
int get_twiddles(int device, size_t N, const float2 **out);

// K1/K5, K2, K3 for block size 2^logb on stream st
int run_forward(int logb, const float2 *tw, cudaStream_t st, const float *src, long long src_stride, int len,
                float2 *dst, long long dst_stride, int nseg, long long ntransforms);
int run_mac(int logb, cudaStream_t st, const MacArgs &a);
int run_inverse(int logb, const float2 *tw, cudaStream_t st, const IfftArgs &a);
// matrix K2 with in-CTA reuse; returns FCB_ERR_UNSUPPORTED for block sizes it is not built for
int run_mac_tile(int logb, cudaStream_t st, MacTileArgs a, int *zchunks_out);
// part rows needed by run_mac_tile for this problem (upper bound on zchunks * NS * OUT * IN)
int mac_tile_plan(int logb, int n_in, int n_out, int n_streams, int nsegs, int *zchunks, int *zlen);

// fcb_profile_mac hook for MAC kernels launched outside engine.cu: returns nullptr when profiling is
// off, else records the start event on `s` and hands back the stop event to record after the launch
cudaEvent_t mac_profile_begin(cudaStream_t s, cudaEvent_t *stop);

} // namespace fcb
