// Shared host-side helpers for libvitae_b200.so: error reporting, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "../../include/vitae_b200.h"

namespace vitae {

int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launch_count;  // kernels enqueued by this library (vitae_launch_count)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define VITAE_REQUIRE(cond, ...)                          \
    do {                                                  \
        if (!(cond)) return ::vitae::set_error(-2, __VA_ARGS__); \
    } while (0)

#define VITAE_CHECK_LAUNCH(name)                                                                    \
    do {                                                                                            \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess) return ::vitae::set_error(-3, "%s launch: %s", name, cudaGetErrorString(e__)); \
        ::vitae::g_launch_count.fetch_add(1, std::memory_order_relaxed);                            \
    } while (0)

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}

}  // namespace vitae
