// Shared host-side helpers for libvitae_b200.so: error reporting, launch checks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
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

// Programmatic dependent launch (PDL): every kernel of this library is enqueued with
// cudaLaunchAttributeProgrammaticStreamSerialization, so that its launch latency and prologue (barrier init, TMEM
// allocation, descriptor prefetch) overlap the tail of its stream predecessor; inside a CUDA-graph capture the
// attribute becomes a programmatic edge.  Contract for every kernel: call pdl_trigger() once its CTA holds all the
// on-chip resources it will ever acquire, and pdl_wait() before its first global-memory access (ptx.cuh).
// VITAE_PDL=0 in the environment launches with plain stream serialisation instead (A/B measurements).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Tensor map over a row-major matrix, or (d2 > 0) a stack of d2 matrices of d1 rows (gemm_tcgen05.cu); cached per key.
int make_tmap(CUtensorMap* out, const void* ptr, uint32_t esize, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld,
              uint32_t b0, uint32_t b1);

}  // namespace vitae
