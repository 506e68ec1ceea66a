// Micro-benchmarks of the NVLS primitives (tools/symm_probe.py): how the achieved NVLink bandwidth of multimem.ld_reduce /
// multimem.st depends on the grid, the loads kept in flight per thread and the memory-ordering qualifier.  Debug entry
// points only: nothing on the training path calls them.
#include "common.h"
#include "ptx.cuh"

namespace vitae {

template <int DEPTH, bool WEAK>
__global__ void __launch_bounds__(256)
mc_reduce_probe_kernel(const float* __restrict__ mc, float* __restrict__ out, long long n4) {
    const long long stride = gridDim.x * 256ll;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += stride * DEPTH) {
        float4 x[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const long long j = i + d * stride;
            if (j < n4) {
                if constexpr (WEAK)
                    asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(x[d].x), "=f"(x[d].y), "=f"(x[d].z), "=f"(x[d].w) : "l"(mc + (j << 2)));
                else
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(x[d].x), "=f"(x[d].y), "=f"(x[d].z), "=f"(x[d].w) : "l"(mc + (j << 2)));
            }
        }
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const long long j = i + d * stride;
            if (j < n4) *reinterpret_cast<float4*>(out + (j << 2)) = x[d];
        }
    }
}

template <bool WEAK>
__global__ void __launch_bounds__(256)
mc_copy_probe_kernel(const float* __restrict__ src, float* __restrict__ mc, long long n4) {
    const long long stride = gridDim.x * 256ll;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += stride) {
        const float4 v = *reinterpret_cast<const float4*>(src + (i << 2));
        if constexpr (WEAK)
            asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + (i << 2)), "f"(v.x), "f"(v.y),
                         "f"(v.z), "f"(v.w));
        else
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + (i << 2)), "f"(v.x),
                         "f"(v.y), "f"(v.z), "f"(v.w));
    }
}

// plain peer-to-peer copy (dst is a peer's buffer mapped into this process): the unicast NVLink write rate
__global__ void __launch_bounds__(256)
p2p_copy_probe_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4) {
    const long long stride = gridDim.x * 256ll;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += stride)
        *reinterpret_cast<float4*>(dst + (i << 2)) = *reinterpret_cast<const float4*>(src + (i << 2));
}

}  // namespace vitae

using namespace vitae;

// mode: depth (1, 4, 8) + 100 for .weak
extern "C" int vitae_debug_mc_reduce(const float* mc, float* out, long long n, int mode, int blocks, void* stream) {
    VITAE_REQUIRE(mc && out && n > 0 && n % 4 == 0 && blocks > 0, "debug_mc_reduce: bad arguments");
    const long long n4 = n >> 2;
    cudaStream_t st = as_stream(stream);
    switch (mode) {
        case 1: mc_reduce_probe_kernel<1, false><<<blocks, 256, 0, st>>>(mc, out, n4); break;
        case 4: mc_reduce_probe_kernel<4, false><<<blocks, 256, 0, st>>>(mc, out, n4); break;
        case 8: mc_reduce_probe_kernel<8, false><<<blocks, 256, 0, st>>>(mc, out, n4); break;
        case 101: mc_reduce_probe_kernel<1, true><<<blocks, 256, 0, st>>>(mc, out, n4); break;
        case 104: mc_reduce_probe_kernel<4, true><<<blocks, 256, 0, st>>>(mc, out, n4); break;
        case 108: mc_reduce_probe_kernel<8, true><<<blocks, 256, 0, st>>>(mc, out, n4); break;
        default: return set_error(-1, "debug_mc_reduce: mode %d", mode);
    }
    VITAE_CHECK_LAUNCH("debug_mc_reduce");
    return 0;
}

// mode 0: multimem.st.relaxed.sys, 1: multimem.st.weak, 2: plain stores (dst = a unicast peer pointer)
extern "C" int vitae_debug_mc_copy(const float* src, float* dst, long long n, int mode, int blocks, void* stream) {
    VITAE_REQUIRE(src && dst && n > 0 && n % 4 == 0 && blocks > 0, "debug_mc_copy: bad arguments");
    const long long n4 = n >> 2;
    cudaStream_t st = as_stream(stream);
    if (mode == 0) mc_copy_probe_kernel<false><<<blocks, 256, 0, st>>>(src, dst, n4);
    else if (mode == 1) mc_copy_probe_kernel<true><<<blocks, 256, 0, st>>>(src, dst, n4);
    else if (mode == 2) p2p_copy_probe_kernel<<<blocks, 256, 0, st>>>(src, dst, n4);
    else return set_error(-1, "debug_mc_copy: mode %d", mode);
    VITAE_CHECK_LAUNCH("debug_mc_copy");
    return 0;
}
