// api_cost_probe.cu — host-side cost of the runtime calls a synchronous small-batch block is made of (B200 box, one thread):
//   nvcc -O2 -arch=sm_100a scripts/api_cost_probe.cu -o /tmp/api_cost && /tmp/api_cost
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>

__global__ void k_empty(int *p) { if (p) *p = 1; }
__global__ void k_spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }

static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F> static double per_call(int n, F f)
{
    for (int i = 0; i < 200; i++) f();
    cudaDeviceSynchronize();
    const double t0 = now();
    for (int i = 0; i < n; i++) f();
    const double t = (now() - t0) / n;
    cudaDeviceSynchronize();
    return t;
}

int main()
{
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    float *h, *d, *hm, *dm;
    cudaHostAlloc(&h, 1 << 16, cudaHostAllocDefault);
    cudaHostAlloc(&hm, 1 << 16, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&dm, hm, 0);
    cudaMalloc(&d, 1 << 16);
    cudaEvent_t ev;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventRecord(ev, st);
    cudaStreamSynchronize(st);
    cudaPointerAttributes a;
    const int n = 20000;
    printf("{\"cudaSetDevice\": %.3f", per_call(n, [&] { cudaSetDevice(0); }));
    printf(", \"cudaPointerGetAttributes_pinned\": %.3f", per_call(n, [&] { cudaPointerGetAttributes(&a, h); }));
    int stackvar;
    printf(", \"cudaPointerGetAttributes_pageable\": %.3f", per_call(n, [&] { cudaPointerGetAttributes(&a, &stackvar); cudaGetLastError(); }));
    printf(", \"cudaEventSynchronize_done\": %.3f", per_call(n, [&] { cudaEventSynchronize(ev); }));
    printf(", \"cudaEventQuery_done\": %.3f", per_call(n, [&] { cudaEventQuery(ev); }));
    printf(", \"cudaStreamSynchronize_idle\": %.3f", per_call(n, [&] { cudaStreamSynchronize(st); }));
    printf(", \"launch_empty_async\": %.3f", per_call(n, [&] { k_empty<<<1, 32, 0, st>>>(nullptr); }));
    printf(", \"launch_empty_and_sync\": %.3f", per_call(n, [&] { k_empty<<<1, 32, 0, st>>>(nullptr); cudaStreamSynchronize(st); }));
    printf(", \"memcpy4k_h2d_and_launch_and_sync\": %.3f", per_call(n, [&] { cudaMemcpyAsync(d, h, 4096, cudaMemcpyHostToDevice, st); k_empty<<<1, 32, 0, st>>>(nullptr); cudaStreamSynchronize(st); }));
    printf(", \"memcpy4k_h2d_event_launch_and_sync\": %.3f", per_call(n, [&] { cudaMemcpyAsync(d, h, 4096, cudaMemcpyHostToDevice, st); cudaEventRecord(ev, st); k_empty<<<1, 32, 0, st>>>(nullptr); cudaStreamSynchronize(st); }));
    printf(", \"memcpy4k_h2d_async_call\": %.3f", per_call(2000, [&] { cudaMemcpyAsync(d, h, 4096, cudaMemcpyHostToDevice, st); }));
    printf(", \"cudaEventRecord_call\": %.3f", per_call(2000, [&] { cudaEventRecord(ev, st); }));
    // kernel of ~100 us, then sync: what the sync adds after the kernel ends
    const double spin100 = per_call(500, [&] { k_spin<<<1, 32, 0, st>>>(190000); cudaStreamSynchronize(st); });
    printf(", \"spin_kernel_190k_cycles_and_sync\": %.3f", spin100);
    // completion seen through a flag in mapped memory written by the kernel itself
    volatile int *flag = (volatile int *)hm;
    printf(", \"launch_empty_flag_in_mapped_memory\": %.3f", per_call(n, [&] { *flag = 0; k_empty<<<1, 32, 0, st>>>((int *)dm); while (*flag == 0) {} }));
    printf("}\n");
    return 0;
}
