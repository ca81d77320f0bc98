// Shared helpers for the acvd_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace acvd {

constexpr int kThreads = 256;
constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr int kMaxRing = 64;          // connexity test handles up to 64 same-cluster ring neighbours

struct CudaError {
    cudaError_t code;
    const char* what;
    const char* file;
    int line;
};

#define ACVD_CUDA(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) throw ::acvd::CudaError{_e, #expr, __FILE__, __LINE__}; \
    } while (0)

#define ACVD_LAUNCH_CHECK() ACVD_CUDA(cudaGetLastError())

// grid sized as a multiple of the SM count for grid-stride kernels
inline int grid_for(int64_t n, int threads = kThreads, int blocks_per_sm = 8) {
    int64_t need = (n + threads - 1) / threads;
    int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync) on the stream of the context
// whose API call is running: temporaries of one call are recycled by the next instead of going back to the driver
// (cudaMalloc / cudaFree of multi-GB scratch cost more than the kernels that use it).  Outside an API call
// (context teardown) the plain synchronous calls are used.
inline thread_local cudaStream_t g_alloc_stream = nullptr;
inline thread_local bool g_alloc_async = false;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        if (count == 0) return;
        if (g_alloc_async) ACVD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), g_alloc_stream));
        else ACVD_CUDA(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void release() {
        if (p) {
            if (g_alloc_async) cudaFreeAsync(p, g_alloc_stream);
            else cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// order-preserving map float -> uint32 (smaller float => smaller key)
__device__ __forceinline__ uint32_t ordered_float_bits(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// warp-aggregated counter add: one atomic per warp
__device__ __forceinline__ void warp_count_add(unsigned long long* ctr, unsigned v) {
    unsigned s = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(ctr, (unsigned long long)s);
}

}  // namespace acvd
