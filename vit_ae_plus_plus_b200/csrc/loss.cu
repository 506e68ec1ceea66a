// Masked patch-reconstruction loss (forward + backward) that reads the raw NCDHW volume directly: the reference's
// patchify permute-copy (model/vit_autoenc.py:100-113, 134 MB at ViT-B/16 128^3x4 B=4) never materialises.
// Only removed patches (mask == 1) are read.  HBM-bound; one CTA per patch, one thread per voxel (C channels:
// contiguous in pred, strided by V^3 in the volume and coalesced across the warp along px).
#include "common.h"
#include "ptx.cuh"

namespace vitae {

template <typename T>
__device__ __forceinline__ float ld_pred(const T* p);
template <>
__device__ __forceinline__ float ld_pred<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_pred<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T, int C>
__device__ __forceinline__ void load_pred_vec(const T* p, float (&v)[C]) {
    if constexpr (C == 4 && sizeof(T) == 2) {
        const uint2 raw = *reinterpret_cast<const uint2*>(p);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else if constexpr (C == 4 && sizeof(T) == 4) {
        const float4 raw = *reinterpret_cast<const float4*>(p);
        v[0] = raw.x; v[1] = raw.y; v[2] = raw.z; v[3] = raw.w;
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = ld_pred<T>(p + c);
    }
}

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;  // valid in thread 0
}

// CT = compile-time channel count (0 = runtime C, scalar path)
template <typename T, int CT>
__global__ void __launch_bounds__(256)
masked_mse_fwd_kernel(const T* __restrict__ pred, const float* __restrict__ vol, const float* __restrict__ mask,
                      float* __restrict__ patch_sums, int Crt, int V, int p, int g) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sh[8];
    const int C = CT ? CT : Crt;
    const int L = g * g * g;
    const int b = blockIdx.x / L, l = blockIdx.x % L;
    if (mask[blockIdx.x] == 0.f) {
        if (threadIdx.x == 0) patch_sums[blockIdx.x] = 0.f;
        return;
    }
    const int gz = l / (g * g), gy = (l / g) % g, gx = l % g;
    const size_t V3 = static_cast<size_t>(V) * V * V;
    const size_t P = static_cast<size_t>(p) * p * p * C;
    const T* prow = pred + (static_cast<size_t>(b) * (L + 1) + 1 + l) * P;
    const float* vb = vol + static_cast<size_t>(b) * C * V3;
    float acc = 0.f;
    const int nvox = p * p * p;
    for (int v = threadIdx.x; v < nvox; v += blockDim.x) {
        const int px = v % p, py = (v / p) % p, pz = v / (p * p);
        const size_t voff = (static_cast<size_t>(gz * p + pz) * V + (gy * p + py)) * V + gx * p + px;
        if constexpr (CT != 0) {
            float pv[CT ? CT : 1];
            load_pred_vec<T, CT ? CT : 1>(prow + static_cast<size_t>(v) * CT, pv);
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float d = pv[c] - __ldg(vb + c * V3 + voff);
                acc += d * d;
            }
        } else {
            for (int c = 0; c < C; ++c) {
                const float d = ld_pred<T>(prow + static_cast<size_t>(v) * C + c) - __ldg(vb + c * V3 + voff);
                acc += d * d;
            }
        }
    }
    const float tot = block_sum_256(acc, sh);
    if (threadIdx.x == 0) patch_sums[blockIdx.x] = tot;
}

// loss_out[0] = sum(patch_sums) / (P * sum(mask)); loss_out[1] = sum(mask).  One block, fixed order.
__global__ void __launch_bounds__(256)
masked_mse_finalize_kernel(const float* __restrict__ patch_sums, const float* __restrict__ mask, int n, float P,
                           float* __restrict__ loss_out) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sh[8];
    float s = 0.f, m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s += patch_sums[i];
        m += mask[i];
    }
    const float st = block_sum_256(s, sh);
    __syncthreads();
    const float mt = block_sum_256(m, sh);
    if (threadIdx.x == 0) {
        loss_out[0] = st / (P * mt);
        loss_out[1] = mt;
    }
}

template <typename T, int CT>
__global__ void __launch_bounds__(256)
masked_mse_bwd_kernel(const T* __restrict__ pred, const float* __restrict__ vol, const float* __restrict__ mask,
                      const float* __restrict__ mask_sum, const float* __restrict__ dloss,
                      __nv_bfloat16* __restrict__ dpred, int Crt, int V, int p, int g) {
    pdl_trigger();
    pdl_wait();
    const int C = CT ? CT : Crt;
    const int L = g * g * g;
    const int b = blockIdx.x / (L + 1), t = blockIdx.x % (L + 1);  // token row incl. cls
    const size_t P = static_cast<size_t>(p) * p * p * C;
    __nv_bfloat16* drow = dpred + (static_cast<size_t>(b) * (L + 1) + t) * P;
    const bool live = t > 0 && mask[static_cast<size_t>(b) * L + (t - 1)] != 0.f;
    if (!live) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (size_t i = threadIdx.x * 8ull; i < P; i += blockDim.x * 8ull) {
            if (i + 8 <= P) *reinterpret_cast<uint4*>(drow + i) = z;
            else for (size_t k = i; k < P; ++k) drow[k] = __float2bfloat16(0.f);
        }
        return;
    }
    const int l = t - 1;
    const float coef = 2.0f * (*dloss) / (static_cast<float>(P) * (*mask_sum));
    const int gz = l / (g * g), gy = (l / g) % g, gx = l % g;
    const size_t V3 = static_cast<size_t>(V) * V * V;
    const T* prow = pred + (static_cast<size_t>(b) * (L + 1) + t) * P;
    const float* vb = vol + static_cast<size_t>(b) * C * V3;
    const int nvox = p * p * p;
    for (int v = threadIdx.x; v < nvox; v += blockDim.x) {
        const int px = v % p, py = (v / p) % p, pz = v / (p * p);
        const size_t voff = (static_cast<size_t>(gz * p + pz) * V + (gy * p + py)) * V + gx * p + px;
        if constexpr (CT == 4) {
            float pv[4];
            load_pred_vec<T, 4>(prow + static_cast<size_t>(v) * 4, pv);
            float d[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) d[c] = coef * (pv[c] - __ldg(vb + c * V3 + voff));
            uint2 pk;
            pk.x = pack_bf16(d[0], d[1]);
            pk.y = pack_bf16(d[2], d[3]);
            *reinterpret_cast<uint2*>(drow + static_cast<size_t>(v) * 4) = pk;
        } else {
            for (int c = 0; c < C; ++c) {
                const float d = coef * (ld_pred<T>(prow + static_cast<size_t>(v) * C + c) - __ldg(vb + c * V3 + voff));
                drow[static_cast<size_t>(v) * C + c] = __float2bfloat16(d);
            }
        }
    }
}

template <typename T>
static int launch_fwd(const T* pred, const float* vol, const float* mask, float* patch_sums, float* loss_out, int B, int C,
                      int V, int p, cudaStream_t st) {
    const int g = V / p, L = g * g * g;
    const int n = B * L;
    if (C == 4) launch_kernel(masked_mse_fwd_kernel<T, 4>, dim3(n), dim3(256), 0, st, pred, vol, mask, patch_sums, C, V, p, g);
    else if (C == 1) launch_kernel(masked_mse_fwd_kernel<T, 1>, dim3(n), dim3(256), 0, st, pred, vol, mask, patch_sums, C, V, p, g);
    else if (C == 2) launch_kernel(masked_mse_fwd_kernel<T, 2>, dim3(n), dim3(256), 0, st, pred, vol, mask, patch_sums, C, V, p, g);
    else launch_kernel(masked_mse_fwd_kernel<T, 0>, dim3(n), dim3(256), 0, st, pred, vol, mask, patch_sums, C, V, p, g);
    VITAE_CHECK_LAUNCH("masked_mse_fwd");
    launch_kernel(masked_mse_finalize_kernel, dim3(1), dim3(256), 0, st, patch_sums, mask, n, static_cast<float>(p) * p * p * C, loss_out);
    VITAE_CHECK_LAUNCH("masked_mse_finalize");
    return 0;
}

template <typename T>
static int launch_bwd(const T* pred, const float* vol, const float* mask, const float* mask_sum, const float* dloss,
                      __nv_bfloat16* dpred, int B, int C, int V, int p, cudaStream_t st) {
    const int g = V / p, L = g * g * g;
    const int n = B * (L + 1);
    if (C == 4) launch_kernel(masked_mse_bwd_kernel<T, 4>, dim3(n), dim3(256), 0, st, pred, vol, mask, mask_sum, dloss, dpred, C, V, p, g);
    else launch_kernel(masked_mse_bwd_kernel<T, 0>, dim3(n), dim3(256), 0, st, pred, vol, mask, mask_sum, dloss, dpred, C, V, p, g);
    VITAE_CHECK_LAUNCH("masked_mse_bwd");
    return 0;
}

}  // namespace vitae

using namespace vitae;

extern "C" int vitae_masked_mse_fwd(const void* pred, int pred_is_bf16, const float* vol, const float* mask,
                                    float* patch_sums, float* loss_out, int B, int C, int V, int p, void* stream) {
    VITAE_REQUIRE(pred && vol && mask && patch_sums && loss_out, "masked_mse_fwd: null pointer");
    VITAE_REQUIRE(B > 0 && C > 0 && p > 0 && V % p == 0, "masked_mse_fwd: bad geometry V=%d p=%d", V, p);
    if (pred_is_bf16)
        return launch_fwd(static_cast<const __nv_bfloat16*>(pred), vol, mask, patch_sums, loss_out, B, C, V, p, as_stream(stream));
    return launch_fwd(static_cast<const float*>(pred), vol, mask, patch_sums, loss_out, B, C, V, p, as_stream(stream));
}

extern "C" int vitae_masked_mse_bwd(const void* pred, int pred_is_bf16, const float* vol, const float* mask,
                                    const float* mask_sum, const float* dloss, void* dpred_bf16, int B, int C, int V,
                                    int p, void* stream) {
    VITAE_REQUIRE(pred && vol && mask && mask_sum && dloss && dpred_bf16, "masked_mse_bwd: null pointer");
    VITAE_REQUIRE(B > 0 && C > 0 && p > 0 && V % p == 0, "masked_mse_bwd: bad geometry V=%d p=%d", V, p);
    VITAE_REQUIRE((static_cast<long long>(p) * p * p * C) % 8 == 0, "masked_mse_bwd: P must be a multiple of 8");
    if (pred_is_bf16)
        return launch_bwd(static_cast<const __nv_bfloat16*>(pred), vol, mask, mask_sum, dloss, static_cast<__nv_bfloat16*>(dpred_bf16), B, C, V, p, as_stream(stream));
    return launch_bwd(static_cast<const float*>(pred), vol, mask, mask_sum, dloss, static_cast<__nv_bfloat16*>(dpred_bf16), B, C, V, p, as_stream(stream));
}
