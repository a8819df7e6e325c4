// Shared helpers for the funsor_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>
#include <math.h>

#include "../../include/funsor_b200.h"

namespace fb2 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_device();  // FB2_OK if the current device is sm_100; caches the answer

#define FB2_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            ::fb2::set_error(__VA_ARGS__);  \
            return FB2_ERR_INVALID_ARGUMENT; \
        }                                   \
    } while (0)

#define FB2_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            ::fb2::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,         \
                             cudaGetErrorString(e_));                                     \
            return FB2_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define FB2_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                              \
        if (e_ != cudaSuccess) {                                                          \
            ::fb2::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,     \
                             cudaGetErrorString(e_));                                     \
            return FB2_ERR_CUDA;                                                          \
        }                                                                                 \
        ::fb2::count_launch();                                                            \
    } while (0)

constexpr float kNegInf = -INFINITY;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Arena {
    char* base;
    size_t size;
    size_t used = 0;
    Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        if (base == nullptr || used + bytes > size) {
            used = size + 1;
            return nullptr;
        }
        T* p = reinterpret_cast<T*>(base + used);
        used += bytes;
        return p;
    }
    bool ok() const { return used <= size; }
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace fb2
