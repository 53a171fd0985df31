// read_peak.cu — what a pure read stream can reach on this GPU (context for roofline.frac > 1:
// MEASURED_PEAKS.json's denominator is a read+write copy).  nvcc -arch=sm_100a -O3 read_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_read(const float4 *__restrict__ p, size_t n4, float *sink)
{
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) {
        float4 v = __ldcs(p + i);
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) *sink = acc;
}

int main()
{
    const size_t bytes = (size_t)8 << 30, n4 = bytes / 16;
    float4 *d;
    float *sink;
    cudaMalloc(&d, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(d, 0, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int grid : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        float best = 1e9f;
        for (int r = 0; r < 12; r++) {
            cudaEventRecord(a);
            k_read<<<grid, 256>>>(d, n4, sink);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (r >= 2 && ms < best) best = ms;
        }
        printf("read-only stream, grid %d x 256: %.1f GB/s (8 GiB, best of 10)\n", grid, bytes / (best * 1e-3) / 1e9);
    }
    // copy for comparison (read + write bytes)
    float4 *e;
    cudaMalloc(&e, bytes / 2);
    float best = 1e9f;
    for (int r = 0; r < 12; r++) {
        cudaEventRecord(a);
        cudaMemcpyAsync(e, d, bytes / 2, cudaMemcpyDeviceToDevice);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r >= 2 && ms < best) best = ms;
    }
    printf("cudaMemcpy D2D 4 GiB (read+write bytes): %.1f GB/s\n", bytes / (best * 1e-3) / 1e9);
    return 0;
}
