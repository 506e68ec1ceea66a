// On-device intensity normalisation of incoming volumes: the reference normalises every volume on the
// host inside Dataset.__getitem__ (dataset/egd_dataset/egd.py:44-50, dataset/brats_dataset/brats.py:26-32) and ships fp32
// over PCIe -- 134 MB per 4 x 4 x 128^3 batch, more PCIe time than half a B200 training step.  Here the host ships the
// RAW volume in its storage type (uint16 / int16 / uint8 scanner intensities, fp16 / bf16, or fp32) and two kernels
// produce the fp32 volume the step reads:
//   pass 1  per group (a channel of a sample, or a whole sample) partial sums / extrema, fixed block order -> deterministic
//   pass 2  every block re-reduces its group's partials (a few dozen values) and writes (x - a) * s
// Modes (ref lines): 0 z-score per channel, unbiased variance (egd.py:45-47); 1 z-score per sample (brats.py:27-29);
//                    2 min-max of the sample to [-1, 1] (egd.py:48-50, brats.py:30-32).
// HBM-bound: esize + 4 bytes written + esize re-read per voxel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"
#include "ptx.cuh"

namespace vitae {

constexpr int ING_THREADS = 256;
constexpr int ING_VEC = 8;                     // elements per thread per iteration
constexpr int ING_MAX_PARTS = 64;              // partial blocks per group

enum RawType : int { RAW_F32 = 0, RAW_F16 = 1, RAW_BF16 = 2, RAW_U16 = 3, RAW_I16 = 4, RAW_U8 = 5 };

__host__ __device__ inline int raw_esize(int t) { return t == RAW_F32 ? 4 : (t == RAW_U8 ? 1 : 2); }

// 8 consecutive raw elements starting at element index i (i % 8 == 0) as floats
template <int TYPE>
__device__ __forceinline__ void load8(const void* __restrict__ src, long long i, float (&x)[8]) {
    if (TYPE == RAW_F32) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + (i >> 2));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + (i >> 2) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else if (TYPE == RAW_U8) {
        const uint2 a = __ldcs(reinterpret_cast<const uint2*>(src) + (i >> 3));
        const unsigned w[2] = {a.x, a.y};
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = static_cast<float>((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
    } else {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(src) + (i >> 3));
        const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned lo = w[k] & 0xffffu, hi = w[k] >> 16;
            if (TYPE == RAW_F16) {
                x[2 * k] = __half2float(__ushort_as_half(static_cast<unsigned short>(lo)));
                x[2 * k + 1] = __half2float(__ushort_as_half(static_cast<unsigned short>(hi)));
            } else if (TYPE == RAW_BF16) {
                x[2 * k] = __uint_as_float(lo << 16);
                x[2 * k + 1] = __uint_as_float(hi << 16);
            } else if (TYPE == RAW_U16) {
                x[2 * k] = static_cast<float>(lo);
                x[2 * k + 1] = static_cast<float>(hi);
            } else {
                x[2 * k] = static_cast<float>(static_cast<short>(lo));
                x[2 * k + 1] = static_cast<float>(static_cast<short>(hi));
            }
        }
    }
}

struct IngPart {
    double sum, sumsq;
    float mn, mx;
};

// pass 1: block (part, group) reduces elements [part * per_part, ...) of its group
template <int TYPE>
__global__ void __launch_bounds__(ING_THREADS)
ingest_stats_kernel(const void* __restrict__ raw, long long group_elems, long long per_part, IngPart* __restrict__ parts) {
    pdl_trigger();
    pdl_wait();
    const int part = blockIdx.x, nparts = gridDim.x, group = blockIdx.y;
    const long long g0 = static_cast<long long>(group) * group_elems;
    const long long lo = part * per_part, hi = min(group_elems, lo + per_part);
    double s = 0.0, ss = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    for (long long i = lo + threadIdx.x * ING_VEC; i < hi; i += ING_THREADS * ING_VEC) {
        float x[8];
        load8<TYPE>(raw, g0 + i, x);
        float fs = 0.f, fss = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            fs += x[k];
            fss = fmaf(x[k], x[k], fss);
            mn = fminf(mn, x[k]);
            mx = fmaxf(mx, x[k]);
        }
        s += static_cast<double>(fs);          // 8-element fp32 partials, fp64 across them
        ss += static_cast<double>(fss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ IngPart sh[ING_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = IngPart{s, ss, mn, mx};
    __syncthreads();
    if (threadIdx.x == 0) {
        IngPart t = sh[0];
        for (int w = 1; w < ING_THREADS / 32; ++w) {
            t.sum += sh[w].sum; t.sumsq += sh[w].sumsq;
            t.mn = fminf(t.mn, sh[w].mn); t.mx = fmaxf(t.mx, sh[w].mx);
        }
        parts[static_cast<long long>(group) * nparts + part] = t;
    }
}

// pass 2: out = (x - a) * s with (a, s) of the element's group; stats[group] = {a, s} is written by part 0 for the caller
template <int TYPE>
__global__ void __launch_bounds__(ING_THREADS)
ingest_apply_kernel(const void* __restrict__ raw, float* __restrict__ out, long long group_elems, long long per_part,
                    const IngPart* __restrict__ parts, int nparts, int mode, float* __restrict__ stats) {
    pdl_trigger();
    pdl_wait();
    const int part = blockIdx.x, group = blockIdx.y;
    __shared__ float sh_a, sh_s;
    if (threadIdx.x == 0) {
        double s = 0.0, ss = 0.0;
        float mn = INFINITY, mx = -INFINITY;
        for (int k = 0; k < nparts; ++k) {
            const IngPart t = parts[static_cast<long long>(group) * nparts + k];
            s += t.sum; ss += t.sumsq;
            mn = fminf(mn, t.mn); mx = fmaxf(mx, t.mx);
        }
        float a, sc;
        if (mode == 2) {                       // 2 * (x - min) / (max - min) - 1  ==  (x - (min + max) / 2) * 2 / (max - min)
            a = 0.5f * (mn + mx);
            sc = 2.0f / (mx - mn);
        } else {
            const double n = static_cast<double>(group_elems);
            const double mean = s / n;
            const double var = (ss - n * mean * mean) / (n - 1.0);          // torch.var default: unbiased
            a = static_cast<float>(mean);
            sc = static_cast<float>(1.0 / sqrt(var));
        }
        sh_a = a;
        sh_s = sc;
        if (part == 0 && stats) {
            stats[2 * group] = a;
            stats[2 * group + 1] = sc;
        }
    }
    __syncthreads();
    const float a = sh_a, sc = sh_s;
    const long long g0 = static_cast<long long>(group) * group_elems;
    const long long lo = part * per_part, hi = min(group_elems, lo + per_part);
    for (long long i = lo + threadIdx.x * ING_VEC; i < hi; i += ING_THREADS * ING_VEC) {
        float x[8];
        load8<TYPE>(raw, g0 + i, x);
        float4 o0, o1;
        o0.x = (x[0] - a) * sc; o0.y = (x[1] - a) * sc; o0.z = (x[2] - a) * sc; o0.w = (x[3] - a) * sc;
        o1.x = (x[4] - a) * sc; o1.y = (x[5] - a) * sc; o1.z = (x[6] - a) * sc; o1.w = (x[7] - a) * sc;
        float4* dst = reinterpret_cast<float4*>(out + g0 + i);
        dst[0] = o0;
        dst[1] = o1;
    }
}

template <int TYPE>
static int run_ingest(const void* raw, float* out, int groups, long long group_elems, int mode, void* workspace, float* stats,
                      cudaStream_t st) {
    long long per_part = ceil_div<long long>(group_elems, ING_MAX_PARTS);
    const long long quantum = ING_THREADS * ING_VEC;
    per_part = ceil_div<long long>(per_part, quantum) * quantum;
    const int nparts = static_cast<int>(ceil_div<long long>(group_elems, per_part));
    launch_kernel(ingest_stats_kernel<TYPE>, dim3(nparts, groups), dim3(ING_THREADS), 0, st, raw, group_elems, per_part,
                  static_cast<IngPart*>(workspace));
    VITAE_CHECK_LAUNCH("ingest_stats");
    launch_kernel(ingest_apply_kernel<TYPE>, dim3(nparts, groups), dim3(ING_THREADS), 0, st, raw, out, group_elems, per_part,
                  static_cast<const IngPart*>(workspace), nparts, mode, stats);
    VITAE_CHECK_LAUNCH("ingest_apply");
    return 0;
}

}  // namespace vitae

using namespace vitae;

extern "C" size_t vitae_ingest_workspace_bytes(int B, int C) {
    return static_cast<size_t>(B) * C * ING_MAX_PARTS * sizeof(IngPart);
}

extern "C" int vitae_ingest_normalize(const void* raw, int raw_type, float* out, int B, int C, long long voxels, int mode,
                                      void* workspace, float* stats, void* stream) {
    VITAE_REQUIRE(raw && out && workspace, "ingest_normalize: null pointer");
    VITAE_REQUIRE(raw_type >= RAW_F32 && raw_type <= RAW_U8, "ingest_normalize: raw_type %d (0 f32, 1 f16, 2 bf16, 3 u16, 4 i16, 5 u8)", raw_type);
    VITAE_REQUIRE(mode >= 0 && mode <= 2, "ingest_normalize: mode %d (0 z-score per channel, 1 z-score per sample, 2 min-max per sample)", mode);
    VITAE_REQUIRE(B > 0 && C > 0 && voxels > 0 && voxels % 8 == 0, "ingest_normalize: B=%d C=%d voxels=%lld (voxels per channel must be a multiple of 8)", B, C, voxels);
    VITAE_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "ingest_normalize: buffers must be 16-byte aligned");
    const int groups = mode == 0 ? B * C : B;
    const long long group_elems = mode == 0 ? voxels : voxels * C;
    cudaStream_t st = as_stream(stream);
    switch (raw_type) {
        case RAW_F32: return run_ingest<RAW_F32>(raw, out, groups, group_elems, mode, workspace, stats, st);
        case RAW_F16: return run_ingest<RAW_F16>(raw, out, groups, group_elems, mode, workspace, stats, st);
        case RAW_BF16: return run_ingest<RAW_BF16>(raw, out, groups, group_elems, mode, workspace, stats, st);
        case RAW_U16: return run_ingest<RAW_U16>(raw, out, groups, group_elems, mode, workspace, stats, st);
        case RAW_I16: return run_ingest<RAW_I16>(raw, out, groups, group_elems, mode, workspace, stats, st);
        default: return run_ingest<RAW_U8>(raw, out, groups, group_elems, mode, workspace, stats, st);
    }
}
