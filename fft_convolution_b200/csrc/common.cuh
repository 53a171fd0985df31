// common.cuh — error plumbing shared by the engine and the host mirror.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/fftconv_b200.h"

namespace fcb {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;
// every device / pinned allocation, stream or event creation and their releases made by this library
// (fcb_debug_alloc_count): update() and process() must not move it in the steady state (src/lib.rs:8)
extern std::atomic<uint64_t> g_resource_calls;

inline int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// count the resource calls where they are made (a function-like macro is not re-expanded inside itself)
#define cudaMalloc(...) (++::fcb::g_resource_calls, cudaMalloc(__VA_ARGS__))
#define cudaFree(...) (++::fcb::g_resource_calls, cudaFree(__VA_ARGS__))
#define cudaHostAlloc(...) (++::fcb::g_resource_calls, cudaHostAlloc(__VA_ARGS__))
#define cudaFreeHost(...) (++::fcb::g_resource_calls, cudaFreeHost(__VA_ARGS__))
#define cudaStreamCreateWithFlags(...) (++::fcb::g_resource_calls, cudaStreamCreateWithFlags(__VA_ARGS__))
#define cudaStreamCreateWithPriority(...) (++::fcb::g_resource_calls, cudaStreamCreateWithPriority(__VA_ARGS__))
#define cudaStreamDestroy(...) (++::fcb::g_resource_calls, cudaStreamDestroy(__VA_ARGS__))
#define cudaEventCreateWithFlags(...) (++::fcb::g_resource_calls, cudaEventCreateWithFlags(__VA_ARGS__))
#define cudaEventCreate(...) (++::fcb::g_resource_calls, cudaEventCreate(__VA_ARGS__))
#define cudaEventDestroy(...) (++::fcb::g_resource_calls, cudaEventDestroy(__VA_ARGS__))

#define FCB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t err__ = (expr);                                                                 \
        if (err__ != cudaSuccess)                                                                   \
            return ::fcb::fail(FCB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                               __FILE__, __LINE__);                                                 \
    } while (0)

#define FCB_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != FCB_OK) return rc__; \
    } while (0)

inline size_t next_power_of_two(size_t v)
{
    size_t p = 1; // usize::next_power_of_two(0) == 1
    while (p < v) p <<= 1;
    return p;
}

inline int ilog2(size_t v)
{
    int l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

} // namespace fcb
