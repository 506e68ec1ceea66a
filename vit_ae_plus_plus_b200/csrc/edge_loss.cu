// Auxiliary edge-map loss of the MAE (model/vit_autoenc.py:221-224, shipped default use_edge_map = yes):
//     raw_edge = mse( sobel(unpatchify(pred)),  sobel(gaussian_blur(target, sigma = 2)) )
// Sobel: model/model_utils/sobel_filter.py:10-45 -- three 3x3x3 directional kernels per channel (smoothing [1,2,1] x
// [1,2,1] x derivative), zero padding, sqrt(gx^2+gy^2+gz^2) summed over the channels.  Blur: gaussian_filter.py:5-26 --
// dense ks^3 kernel = outer product of 11 normalised taps (linspace(-6, 6, 11) quirk: taps 1.2 apart), zero padding; run
// here as three 1-D passes (identical under zero padding).
//
// All kernels are stencils over the [B*C, V, V, V] fp32 volumes.  Their HBM traffic is close to the algorithmic minimum
// already with neighbours served by L1/L2 (ncu: 134 MB read per blur pass of a 134 MB volume), so what bounds them is the
// ISSUE rate: the kernels below are written for few instructions per voxel -- compile-time tap counts, separable
// evaluation (the three Sobel kernels share one [1,2,1] / (-1,0,1) reduction over z before the x and y combinations),
// x-neighbours exchanged by warp shuffles instead of scalar loads, several outputs per thread along the blur axis, 16-byte
// vector accesses.  Forward for pred keeps F_{i,c} = D * g_{i,c} / |g_c| (3 bf16 per channel and voxel; D = E_pred -
// E_target) so that the backward is one transposed-stencil pass:
//     d raw_edge / d P_c(u) = 2 / (B V^3) * sum_i sum_d K_i[d] * F_{i,c}(u + 1 - d)
// (|g| = 0 -> contribution 0: the reference's sqrt'(0) * 0 would be NaN there).  The gradient is added to
// the bf16 dpred buffer in the patch layout (pz, py, px, c) of model/vit_autoenc.py:100-113.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace vitae {

struct Taps {
    float t[16];
    int n;
};

// Thread geometry of every stencil kernel: a thread owns 4 consecutive x (one float4), a warp 128 consecutive x of ONE
// row (block = (32, 4): warp index = threadIdx.y, so row validity is warp-uniform and shuffles sit in uniform branches),
// grid = (ceil(V/128), rows / 4, planes * V): no 64-bit div/mod per thread, all row accesses are coalesced vectors.
constexpr int XT = 4;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ size_t vox(int V, int plane, int z, int y, int x) {
    return ((static_cast<size_t>(plane) * V + z) * V + y) * V + x;
}

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ float4 ld4(const float* __restrict__ p, bool ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) v = *reinterpret_cast<const float4*>(p);
    return v;
}
// 4 bf16 -> fp32
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* __restrict__ p, bool ok) {
    uint2 m = make_uint2(0u, 0u);
    if (ok) m = *reinterpret_cast<const uint2*>(p);
    return make_float4(bf_lo(m.x), bf_hi(m.x), bf_lo(m.y), bf_hi(m.y));
}
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// r[0..5] = values x0-1 .. x0+4 of a row quantity whose own 4 values are v: the neighbours come from the adjacent lanes
// (all 32 lanes must call; lanes / rows outside the volume carry zeros), the warp's outer edges are zero unless the row
// spans several warps (V > 128), in which case the caller patches r[0] of lane 0 and r[5] of lane 31.
__device__ __forceinline__ void widen(const float4 v, float (&r)[6]) {
    float left = __shfl_up_sync(FULL, v.w, 1), right = __shfl_down_sync(FULL, v.x, 1);
    if (threadIdx.x == 0) left = 0.f;
    if (threadIdx.x == 31) right = 0.f;
    r[0] = left; r[1] = v.x; r[2] = v.y; r[3] = v.z; r[4] = v.w; r[5] = right;
}

// out(u) = sum_k taps[k] * in(u + (k - NT/2) e_axis), zero outside; AXIS 0 = x (fastest), 1 = y, 2 = z.  Along y / z a
// thread produces MT = 4 consecutive outputs from one sliding set of MT + NT - 1 row loads (3.5 loads per output for 11 taps).
constexpr int BLUR_MT = 4;
template <int NT, int AXIS>
__global__ void __launch_bounds__(128)
blur_axis_kernel(const float* __restrict__ in, float* __restrict__ out, int V, const Taps taps) {
    pdl_trigger();
    pdl_wait();
    constexpr int H = NT / 2;
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * XT;
    if (x0 >= V) return;
    if constexpr (AXIS == 0) {
        // BLUR_MT independent rows per thread (longer-lived blocks, BLUR_MT x the loads in flight)
        const int yb = (blockIdx.y * 4 + threadIdx.y) * BLUR_MT, z = blockIdx.z % V, plane = blockIdx.z / V;
        constexpr int LQ = (H + 3) / 4;                        // aligned float4 chunks on each side of the own one
#pragma unroll
        for (int j = 0; j < BLUR_MT; ++j) {
            const int y = yb + j;
            if (y >= V) break;
            float w[(2 * LQ + 1) * 4];
            const float* row = in + vox(V, plane, z, y, 0);
#pragma unroll
            for (int q = 0; q < 2 * LQ + 1; ++q) {
                const int xs = x0 + 4 * (q - LQ);
                const float4 v = (xs >= 0 && xs < V) ? *reinterpret_cast<const float4*>(row + xs) : make_float4(0.f, 0.f, 0.f, 0.f);
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const float t = taps.t[k];
                acc.x += t * w[4 * LQ + 0 + k - H]; acc.y += t * w[4 * LQ + 1 + k - H];
                acc.z += t * w[4 * LQ + 2 + k - H]; acc.w += t * w[4 * LQ + 3 + k - H];
            }
            *reinterpret_cast<float4*>(out + vox(V, plane, z, y, x0)) = acc;
        }
    } else {
        // first output coordinate along the axis (a0) and the fixed other one
        int y, z, plane;
        if (AXIS == 1) {
            y = (blockIdx.y * 4 + threadIdx.y) * BLUR_MT; z = blockIdx.z % V; plane = blockIdx.z / V;
            if (y >= V) return;
        } else {
            const int zg = V / BLUR_MT;                        // V % 4 == 0
            y = blockIdx.y * 4 + threadIdx.y; z = (blockIdx.z % zg) * BLUR_MT; plane = blockIdx.z / zg;
            if (y >= V) return;
        }
        const int a0 = AXIS == 1 ? y : z;
        const size_t stride = AXIS == 1 ? static_cast<size_t>(V) : static_cast<size_t>(V) * V;
        const float* base = in + vox(V, plane, z, y, x0);       // element (a0) of the column walked along the axis
        float4 acc[BLUR_MT];
#pragma unroll
        for (int j = 0; j < BLUR_MT; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < BLUR_MT + NT - 1; ++r) {
            const int a = a0 + r - H;
            if (a >= 0 && a < V) {
                const float4 v = *reinterpret_cast<const float4*>(base + (static_cast<long long>(r) - H) * static_cast<long long>(stride));
#pragma unroll
                for (int j = 0; j < BLUR_MT; ++j) {
                    const int k = r - j;
                    if (k >= 0 && k < NT) {
                        const float t = taps.t[k];
                        acc[j].x += t * v.x; acc[j].y += t * v.y; acc[j].z += t * v.z; acc[j].w += t * v.w;
                    }
                }
            }
        }
        float* o = out + vox(V, plane, z, y, x0);
#pragma unroll
        for (int j = 0; j < BLUR_MT; ++j)
            if (a0 + j < V) *reinterpret_cast<float4*>(o + j * stride) = acc[j];
    }
}

// pred bf16 [B, Nd = L+1, P] (row 0 of each sample = cls) -> predvol fp32 [B*C, V, V, V]; plane = b here.  The 4 voxels
// of a thread are 4C consecutive bf16 of one patch row (p % 4 == 0): one or two vector loads.
template <int C>
__global__ void __launch_bounds__(128)
unpatchify_kernel(const __nv_bfloat16* __restrict__ pred, float* __restrict__ vol, int V, int p) {
    pdl_trigger();
    pdl_wait();
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * XT;
    const int y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z % V, plane = blockIdx.z / V;
    if (x0 >= V || y >= V) return;
    const int g = V / p;
    const int l = ((z / p) * g + (y / p)) * g + (x0 / p);          // 4 consecutive x stay inside one patch (p % 4 == 0)
    const int within = (((z % p) * p + (y % p)) * p + (x0 % p)) * C;
    const size_t L1 = static_cast<size_t>(g) * g * g + 1;
    const __nv_bfloat16* src = pred + (static_cast<size_t>(plane) * L1 + 1 + l) * (static_cast<size_t>(p) * p * p * C) + within;
    uint32_t wds[2 * C];                                            // XT * C bf16 = 2C words, 8C-byte aligned
    if constexpr (C == 4) {
        const uint4 a = *reinterpret_cast<const uint4*>(src), b = *reinterpret_cast<const uint4*>(src + 8);
        wds[0] = a.x; wds[1] = a.y; wds[2] = a.z; wds[3] = a.w; wds[4] = b.x; wds[5] = b.y; wds[6] = b.z; wds[7] = b.w;
    } else if constexpr (C == 2) {
        const uint4 a = *reinterpret_cast<const uint4*>(src);
        wds[0] = a.x; wds[1] = a.y; wds[2] = a.z; wds[3] = a.w;
    } else {
        const uint2 a = *reinterpret_cast<const uint2*>(src);
        wds[0] = a.x; wds[1] = a.y;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float v[XT];
#pragma unroll
        for (int i = 0; i < XT; ++i) {
            const int e = i * C + c;                                // element index inside the 4C run
            v[i] = (e & 1) ? bf_hi(wds[e >> 1]) : bf_lo(wds[e >> 1]);
        }
        *reinterpret_cast<float4*>(vol + vox(V, plane * C + c, z, y, x0)) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// ---- Sobel, evaluated separably while walking along y --------------------------------------------------------------
// s = [1,2,1]; g0: d/dx (left - right), g1: d/dy (y+1 minus y-1), g2: d/dz (z+1 minus z-1)   (weight[0..2] of
// model/model_utils/sobel_filter.py:10-35).  A thread owns 4 x of YT consecutive rows.  Per input row it reduces the
// three z planes to Rs = [1,2,1]_z and Rd = (-1,0,1)_z (6 wide) and takes the x combinations once
//     A = Rs(x-1) - Rs(x+1),  Sm = [1,2,1]_x Rs,  Cz = [1,2,1]_x Rd
// and an output row is  g0 = [1,2,1]_y A,  g1 = Sm(y+1) - Sm(y-1),  g2 = [1,2,1]_y Cz  of the last three input rows:
// 4.5 row loads per output row instead of 9 and a third of the additions.

struct SobelRow {
    float A[XT], Sm[XT], Cz[XT];
};

// the three z planes of one input row (own 4 x), zero outside the volume
struct Rows3 {
    float4 m, c, p;
};
// px: element (z, yr, x0) of the channel volume; ok: lane and row inside the volume; zm / zp: planes z-1 / z+1 exist
template <bool CENTRE = true, typename T>
__device__ __forceinline__ Rows3 load_rows3(const T* __restrict__ px, bool ok, bool zm, bool zp, int VV) {
    Rows3 r;
    r.m = ld4(px - VV, ok && zm);
    r.c = CENTRE ? ld4(px, ok) : make_float4(0.f, 0.f, 0.f, 0.f);
    r.p = ld4(px + VV, ok && zp);
    return r;
}

// [1,2,1]_z and (-1,0,1)_z of a row's three planes, widened to x0-1 .. x0+4 (the reductions are taken on the own 4 values
// BEFORE the neighbour exchange: 4 shuffles per row)
template <bool WANT_S = true, bool WANT_D = true, typename T>
__device__ __forceinline__ void z_reduce(const Rows3& w, const T* __restrict__ px, bool ok, bool zm, bool zp, bool multi, int V,
                                         int VV, int x0, float (&Rs)[6], float (&Rd)[6]) {
    if (WANT_S)
        widen(make_float4((w.m.x + w.p.x) + 2.f * w.c.x, (w.m.y + w.p.y) + 2.f * w.c.y, (w.m.z + w.p.z) + 2.f * w.c.z,
                          (w.m.w + w.p.w) + 2.f * w.c.w), Rs);
    if (WANT_D) widen(make_float4(w.p.x - w.m.x, w.p.y - w.m.y, w.p.z - w.m.z, w.p.w - w.m.w), Rd);
    if (multi && ok) {                                              // V > 128: the warp's outer neighbours are real voxels
        if (threadIdx.x == 0 && x0 > 0) {
            const float a = zm ? ld1(px - VV - 1) : 0.f, c = WANT_S ? ld1(px - 1) : 0.f, d = zp ? ld1(px + VV - 1) : 0.f;
            if (WANT_S) Rs[0] = (a + d) + 2.f * c;
            if (WANT_D) Rd[0] = d - a;
        }
        if (threadIdx.x == 31 && x0 + 4 < V) {
            const float a = zm ? ld1(px - VV + 4) : 0.f, c = WANT_S ? ld1(px + 4) : 0.f, d = zp ? ld1(px + VV + 4) : 0.f;
            if (WANT_S) Rs[5] = (a + d) + 2.f * c;
            if (WANT_D) Rd[5] = d - a;
        }
    }
}

__device__ __forceinline__ void sobel_row(const Rows3& w, const float* __restrict__ px, bool ok, bool zm, bool zp, bool multi,
                                          int V, int VV, int x0, SobelRow& o) {
    float Rs[6], Rd[6];
    z_reduce(w, px, ok, zm, zp, multi, V, VV, x0, Rs, Rd);
#pragma unroll
    for (int i = 0; i < XT; ++i) {
        o.A[i] = Rs[i] - Rs[i + 2];
        o.Sm[i] = (Rs[i] + Rs[i + 2]) + 2.f * Rs[i + 1];
        o.Cz[i] = (Rd[i] + Rd[i + 2]) + 2.f * Rd[i + 1];
    }
}

__device__ __forceinline__ void sobel_combine(const SobelRow& p, const SobelRow& q, const SobelRow& n, float (&g0)[XT],
                                              float (&g1)[XT], float (&g2)[XT]) {
#pragma unroll
    for (int i = 0; i < XT; ++i) {
        g0[i] = (p.A[i] + n.A[i]) + 2.f * q.A[i];
        g1[i] = n.Sm[i] - p.Sm[i];
        g2[i] = (p.Cz[i] + n.Cz[i]) + 2.f * q.Cz[i];
    }
}

// |g| and 1/|g| (0 for g = 0) from one MUFU.RSQ: 2 ulp, far inside the bf16 F planes and the loss tolerance
__device__ __forceinline__ float grad_mag(float g0, float g1, float g2, float& inv) {
    const float m2 = g0 * g0 + g1 * g1 + g2 * g2;
    inv = m2 > 0.f ? rsqrtf(m2) : 0.f;
    return m2 * inv;
}

#define SOBEL_COORDS                                                                      \
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * XT;                                  \
    const int y0 = (blockIdx.y * 4 + threadIdx.y) * YT;                                   \
    const int z = blockIdx.z % V, b = blockIdx.z / V;                                     \
    const bool xvalid = x0 < V, multi = gridDim.x > 1;                                    \
    const bool zm = z > 0, zp = z + 1 < V;                                                \
    const int VV = V * V;

// E(b, v) = sum_c |sobel(vol_c)(v)|   (target branch, no gradient).  plane = b.
template <int YT>
__global__ void __launch_bounds__(128, YT == 4 ? 6 : 8)
sobel_edge_kernel(const float* __restrict__ vol, float* __restrict__ E, int C, int V) {
    pdl_trigger();
    pdl_wait();
    SOBEL_COORDS
    if (y0 >= V) return;                                            // warp-uniform
    float e[YT][XT];
#pragma unroll
    for (int j = 0; j < YT; ++j)
#pragma unroll
        for (int i = 0; i < XT; ++i) e[j][i] = 0.f;
    for (int c = 0; c < C; ++c) {
        const float* px = vol + vox(V, b * C + c, z, y0, x0);
        SobelRow w[3];
        Rows3 nxt = load_rows3(px - V, xvalid && y0 > 0, zm, zp, VV);
#pragma unroll
        for (int r = 0; r < YT + 2; ++r) {
            const bool ok = xvalid && static_cast<unsigned>(y0 - 1 + r) < static_cast<unsigned>(V);
            const Rows3 cur = nxt;                                  // the next row's loads are in flight during this row's math
            if (r + 1 < YT + 2) nxt = load_rows3(px + r * V, xvalid && y0 + r < V, zm, zp, VV);
            sobel_row(cur, px + (r - 1) * V, ok, zm, zp, multi, V, VV, x0, w[r % 3]);
            if (r >= 2) {
                float g0[XT], g1[XT], g2[XT];
                sobel_combine(w[(r - 2) % 3], w[(r - 1) % 3], w[r % 3], g0, g1, g2);
#pragma unroll
                for (int i = 0; i < XT; ++i) {
                    float inv;
                    e[r - 2][i] += grad_mag(g0[i], g1[i], g2[i], inv);
                }
            }
        }
    }
    if (xvalid) {
#pragma unroll
        for (int j = 0; j < YT; ++j)
            if (y0 + j < V) *reinterpret_cast<float4*>(E + vox(V, b, z, y0 + j, x0)) = make_float4(e[j][0], e[j][1], e[j][2], e[j][3]);
    }
}

// Pred branch: D = sum_c |sobel(vol_c)| - E_tgt (stored in `resid`), F_{i,c} = D g_{i,c} / |g_c| (bf16, 3 planes per channel)
// and block sums of D^2.  The normalised gradients are stored while the channels are walked and scaled by D (known only
// after the last channel) in a second touch of the thread's own stores.  plane = b.
template <int C, int YT>
__global__ void __launch_bounds__(128, YT == 4 ? 6 : 8)
sobel_pred_kernel(const float* __restrict__ vol, float* __restrict__ resid, const float* __restrict__ E_tgt,
                  __nv_bfloat16* __restrict__ F, float* __restrict__ partials, int V) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[4];
    SOBEL_COORDS
    float sq = 0.f;
    if (y0 < V) {                                                   // warp-uniform
        float e[YT][XT];
#pragma unroll
        for (int j = 0; j < YT; ++j)
#pragma unroll
            for (int i = 0; i < XT; ++i) e[j][i] = 0.f;
        const size_t plane_stride = static_cast<size_t>(V) * VV;
#pragma unroll 1
        for (int c = 0; c < C; ++c) {                               // not unrolled: registers (occupancy) over code size
            const float* px = vol + vox(V, b * C + c, z, y0, x0);
            __nv_bfloat16* Fc = F + vox(V, (b * C + c) * 3, z, y0, x0);
            SobelRow w[3];
            Rows3 nxt = load_rows3(px - V, xvalid && y0 > 0, zm, zp, VV);
#pragma unroll
            for (int r = 0; r < YT + 2; ++r) {
                const bool ok = xvalid && static_cast<unsigned>(y0 - 1 + r) < static_cast<unsigned>(V);
                const Rows3 cur = nxt;                              // the next row's loads are in flight during this row's math
                if (r + 1 < YT + 2) nxt = load_rows3(px + r * V, xvalid && y0 + r < V, zm, zp, VV);
                sobel_row(cur, px + (r - 1) * V, ok, zm, zp, multi, V, VV, x0, w[r % 3]);
                if (r >= 2) {
                    float g0[XT], g1[XT], g2[XT];
                    sobel_combine(w[(r - 2) % 3], w[(r - 1) % 3], w[r % 3], g0, g1, g2);
#pragma unroll
                    for (int i = 0; i < XT; ++i) {
                        float inv;
                        e[r - 2][i] += grad_mag(g0[i], g1[i], g2[i], inv);
                        g0[i] *= inv; g1[i] *= inv; g2[i] *= inv;
                    }
                    if (xvalid) {
                        __nv_bfloat16* f = Fc + (r - 2) * V;
                        *reinterpret_cast<uint2*>(f) = make_uint2(pack_bf16(g0[0], g0[1]), pack_bf16(g0[2], g0[3]));
                        *reinterpret_cast<uint2*>(f + plane_stride) = make_uint2(pack_bf16(g1[0], g1[1]), pack_bf16(g1[2], g1[3]));
                        *reinterpret_cast<uint2*>(f + 2 * plane_stride) = make_uint2(pack_bf16(g2[0], g2[1]), pack_bf16(g2[2], g2[3]));
                    }
                }
            }
        }
        if (xvalid) {                                               // rows y0 .. y0+3 all exist (V % 4 == 0)
#pragma unroll
            for (int j = 0; j < YT; ++j) {
                const size_t o = vox(V, b, z, y0 + j, x0);
                const float4 t = *reinterpret_cast<const float4*>(E_tgt + o);
                e[j][0] -= t.x; e[j][1] -= t.y; e[j][2] -= t.z; e[j][3] -= t.w;      // e becomes the residual D
                *reinterpret_cast<float4*>(resid + o) = make_float4(e[j][0], e[j][1], e[j][2], e[j][3]);
                sq += e[j][0] * e[j][0] + e[j][1] * e[j][1] + e[j][2] * e[j][2] + e[j][3] * e[j][3];
            }
#pragma unroll 1
            for (int ck = 0; ck < 3 * C; ++ck) {                    // the thread's own stores of above: n -> F = D n
                __nv_bfloat16* f = F + vox(V, b * C * 3 + ck, z, y0, x0);
                uint2 n[YT];
#pragma unroll
                for (int j = 0; j < YT; ++j) n[j] = *reinterpret_cast<const uint2*>(f + j * V);
#pragma unroll
                for (int j = 0; j < YT; ++j)
                    *reinterpret_cast<uint2*>(f + j * V) = make_uint2(pack_bf16(e[j][0] * bf_lo(n[j].x), e[j][1] * bf_hi(n[j].x)),
                                                                      pack_bf16(e[j][2] * bf_lo(n[j].y), e[j][3] * bf_hi(n[j].y)));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(FULL, sq, o);
    if (threadIdx.x == 0) red[threadIdx.y] = sq;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0)
        partials[(static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] =
            (red[0] + red[1]) + (red[2] + red[3]);
}

__global__ void __launch_bounds__(256)
edge_loss_finalize_kernel(const float* __restrict__ partials, long long n, float inv_count, float* __restrict__ loss_out) {
    pdl_trigger();
    pdl_wait();
    __shared__ double sh[256];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) s += static_cast<double>(partials[i]);   // fixed order -> deterministic
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_out[0] = static_cast<float>(sh[0] * inv_count);
}

// dpred[b, 1 + l, (pz,py,px,c)] += (*upstream) * 2 / (B V^3) * sum_i sum_d K_i[d] F_{i,c}(u + 1 - d); plane = b.
// With S = [1,2,1] and the offsets taken on the contributing voxel:
//   out = Dx( Sz Sy F0 ) + Sx( Sz (F1(y-1) - F1(y+1)) + Sy (F2(z-1) - F2(z+1)) ),   Dx(f)(x) = f(x+1) - f(x-1)
// evaluated like the forward: per input row the z reductions U0 = Sz F0, U1 = Sz F1, U2 = F2(z-1) - F2(z+1) (6 wide), then
// G = Dx U0 + Sx U2 and Bq = Sx U1 once per row, and an output row is Sy G + Bq(y-1) - Bq(y+1) of the last three rows.
template <int C, int YT>
__global__ void __launch_bounds__(128, YT == 4 ? 6 : 8)
edge_loss_bwd_kernel(const __nv_bfloat16* __restrict__ F, const float* __restrict__ upstream,
                     __nv_bfloat16* __restrict__ dpred, int V, int p, float two_over_count) {
    pdl_trigger();
    pdl_wait();
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * XT;
    const int y0 = (blockIdx.y * 4 + threadIdx.y) * YT;
    const int z = blockIdx.z % V, b = blockIdx.z / V;
    if (y0 >= V) return;                                            // warp-uniform
    const bool xvalid = x0 < V, multi = gridDim.x > 1;
    const int VV = V * V;
    const size_t plane_stride = static_cast<size_t>(V) * VV;
    // per-channel results wait in shared memory (thread-private slots, no barrier) for the channel-interleaved store:
    // 16 C values per row would otherwise cost 64 registers and with them half of the resident warps
    __shared__ float4 sacc[C * YT * 128];
    const int tid = threadIdx.y * 32 + threadIdx.x;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
        const __nv_bfloat16* F0 = F + vox(V, (b * C + c) * 3, z, y0, x0);   // own voxel run in plane 0; rows by 32-bit offsets
        const __nv_bfloat16* F1 = F0 + plane_stride;
        const __nv_bfloat16* F2 = F1 + plane_stride;
        float G[3][XT], Bq[3][XT];
        const bool zm = z > 0, zp = z + 1 < V;
        Rows3 n0 = load_rows3(F0 - V, xvalid && y0 > 0, zm, zp, VV), n1 = load_rows3(F1 - V, xvalid && y0 > 0, zm, zp, VV),
              n2 = load_rows3<false>(F2 - V, xvalid && y0 > 0, zm, zp, VV);
#pragma unroll
        for (int r = 0; r < YT + 2; ++r) {
            const bool ok = xvalid && static_cast<unsigned>(y0 - 1 + r) < static_cast<unsigned>(V);
            const Rows3 c0 = n0, c1 = n1, c2 = n2;                  // the next row's loads are in flight during this row's math
            if (r + 1 < YT + 2) {
                const bool okn = xvalid && y0 + r < V;
                n0 = load_rows3(F0 + r * V, okn, zm, zp, VV);
                n1 = load_rows3(F1 + r * V, okn, zm, zp, VV);
                n2 = load_rows3<false>(F2 + r * V, okn, zm, zp, VV);
            }
            const int ro = (r - 1) * V;
            float U0[6], U1[6], U2[6], unused[6];
            z_reduce<true, false>(c0, F0 + ro, ok, zm, zp, multi, V, VV, x0, U0, unused);      // U0 = Sz F0
            z_reduce<true, false>(c1, F1 + ro, ok, zm, zp, multi, V, VV, x0, U1, unused);      // U1 = Sz F1
            z_reduce<false, true>(c2, F2 + ro, ok, zm, zp, multi, V, VV, x0, unused, U2);      // U2 = F2(z+1) - F2(z-1)
#pragma unroll
            for (int i = 0; i < XT; ++i) {
                G[r % 3][i] = (U0[i + 2] - U0[i]) - ((U2[i] + U2[i + 2]) + 2.f * U2[i + 1]);
                Bq[r % 3][i] = (U1[i] + U1[i + 2]) + 2.f * U1[i + 1];
            }
            if (r >= 2) {
                const int pp = (r - 2) % 3, q = (r - 1) % 3, n = r % 3;
                float o[XT];
#pragma unroll
                for (int i = 0; i < XT; ++i) o[i] = ((G[pp][i] + G[n][i]) + 2.f * G[q][i]) + (Bq[pp][i] - Bq[n][i]);
                sacc[(c * YT + (r - 2)) * 128 + tid] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    if (!xvalid) return;
    const float scale = (*upstream) * two_over_count;
    const int g = V / p;
    const size_t L1 = static_cast<size_t>(g) * g * g + 1, P = static_cast<size_t>(p) * p * p * C;
#pragma unroll
    for (int j = 0; j < YT; ++j) {
        const int y = y0 + j;
        if (y >= V) continue;
        const int l = ((z / p) * g + (y / p)) * g + (x0 / p);
        const int within = (((z % p) * p + (y % p)) * p + (x0 % p)) * C;
        __nv_bfloat16* dst = dpred + (static_cast<size_t>(b) * L1 + 1 + l) * P + within;
        float acc[XT][C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float4 v = sacc[(c * YT + j) * 128 + tid];
            acc[0][c] = v.x; acc[1][c] = v.y; acc[2][c] = v.z; acc[3][c] = v.w;
        }
        // the thread's XT * C values are one 8C-byte aligned run of the patch row: vector read-modify-write
        uint32_t wds[2 * C];
        if constexpr (C == 4) {
            const uint4 a = *reinterpret_cast<const uint4*>(dst), bq = *reinterpret_cast<const uint4*>(dst + 8);
            wds[0] = a.x; wds[1] = a.y; wds[2] = a.z; wds[3] = a.w; wds[4] = bq.x; wds[5] = bq.y; wds[6] = bq.z; wds[7] = bq.w;
        } else if constexpr (C == 2) {
            const uint4 a = *reinterpret_cast<const uint4*>(dst);
            wds[0] = a.x; wds[1] = a.y; wds[2] = a.z; wds[3] = a.w;
        } else {
            const uint2 a = *reinterpret_cast<const uint2*>(dst);
            wds[0] = a.x; wds[1] = a.y;
        }
#pragma unroll
        for (int w = 0; w < 2 * C; ++w) {
            const int e0 = 2 * w, e1 = 2 * w + 1;                   // element index e = i * C + c
            const float lo = bf_lo(wds[w]) + scale * acc[e0 / C][e0 % C];
            const float hi = bf_hi(wds[w]) + scale * acc[e1 / C][e1 % C];
            wds[w] = pack_bf16(lo, hi);
        }
        if constexpr (C == 4) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            *reinterpret_cast<uint4*>(dst + 8) = make_uint4(wds[4], wds[5], wds[6], wds[7]);
        } else if constexpr (C == 2) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        } else {
            *reinterpret_cast<uint2*>(dst) = make_uint2(wds[0], wds[1]);
        }
    }
}

static inline dim3 edge_grid(int V, int planes) { return dim3(ceil_div(V, 128), ceil_div(V, 4), planes * V); }
// kernels walking YT rows per thread
static inline dim3 walk_grid(int V, int planes, int yt) { return dim3(ceil_div(V, 128), ceil_div(V, 4 * yt), planes * V); }

template <int NT>
static void launch_blur(const float* in, float* out, int V, int planes, int axis, const Taps& t, cudaStream_t st) {
    const dim3 blk(32, 4);
    if (axis == 0)
        launch_kernel(blur_axis_kernel<NT, 0>, dim3(ceil_div(V, 128), ceil_div(V, 4 * BLUR_MT), planes * V), blk, 0, st, in, out, V, t);
    else if (axis == 1)
        launch_kernel(blur_axis_kernel<NT, 1>, dim3(ceil_div(V, 128), ceil_div(V, 4 * BLUR_MT), planes * V), blk, 0, st, in, out, V, t);
    else
        launch_kernel(blur_axis_kernel<NT, 2>, dim3(ceil_div(V, 128), ceil_div(V, 4), planes * (V / BLUR_MT)), blk, 0, st, in, out, V, t);
}

static void blur_pass(const float* in, float* out, int V, int planes, int axis, const Taps& t, cudaStream_t st) {
    switch (t.n) {
        case 1: launch_blur<1>(in, out, V, planes, axis, t, st); break;
        case 3: launch_blur<3>(in, out, V, planes, axis, t, st); break;
        case 5: launch_blur<5>(in, out, V, planes, axis, t, st); break;
        case 7: launch_blur<7>(in, out, V, planes, axis, t, st); break;
        case 9: launch_blur<9>(in, out, V, planes, axis, t, st); break;
        case 11: launch_blur<11>(in, out, V, planes, axis, t, st); break;      // sigma = 2, the reference's call site
        case 13: launch_blur<13>(in, out, V, planes, axis, t, st); break;
        default: launch_blur<15>(in, out, V, planes, axis, t, st); break;
    }
}

}  // namespace vitae

using namespace vitae;

// rows a thread of the Sobel kernels walks: 4 (default) or 2 (VITAE_EDGE_YT=2, experiment: more, lighter threads)
static int edge_yt() {
    static const int v = [] {
        const char* e = std::getenv("VITAE_EDGE_YT");
        return (e && std::atoi(e) == 2) ? 2 : 4;
    }();
    return v;
}

// scratch layout (floats): [0, bc) volume A / predvol, [bc, 2bc) volume B, [2bc, 2bc + 1.5bc) F planes (bf16, 3 per
// channel voxel), then the block partials
static inline size_t edge_partials_count(int B, int V) {
    return static_cast<size_t>(ceil_div(V, 128)) * ceil_div(V, 4 * edge_yt()) * B * V;     // blocks of sobel_pred_kernel
}

extern "C" size_t vitae_edge_scratch_floats(int B, int C, int V) {
    const size_t bc = static_cast<size_t>(B) * C * V * V * V;
    return 2 * bc + (3 * bc + 1) / 2 + edge_partials_count(B, V) + 64;
}

// E_tgt[b, v] = sum_c |sobel(blur(vol_c))|(v)          (target branch of model/vit_autoenc.py:221-224; no gradient)
extern "C" int vitae_edge_target(const float* vol, const float* taps, int ntaps, float* scratch, float* E_tgt, int B, int C,
                                 int V, void* stream) {
    VITAE_REQUIRE(vol && taps && scratch && E_tgt, "edge_target: null pointer");
    VITAE_REQUIRE(B > 0 && C > 0 && V >= 4 && V % 4 == 0 && ntaps > 0 && ntaps <= 15 && (ntaps & 1),
                  "edge_target: bad sizes (V %% 4 == 0, ntaps odd <= 15)");
    VITAE_REQUIRE(static_cast<long long>(B) * C * V <= 65535, "edge_target: B*C*V exceeds the grid limit");
    Taps t;
    t.n = ntaps;
    for (int i = 0; i < 16; ++i) t.t[i] = i < ntaps ? taps[i] : 0.f;
    const size_t bc = static_cast<size_t>(B) * C * V * V * V;
    float* sa = scratch;
    float* sb = scratch + bc;
    cudaStream_t st = as_stream(stream);
    blur_pass(vol, sa, V, B * C, 2, t, st);
    VITAE_CHECK_LAUNCH("edge blur z");
    blur_pass(sa, sb, V, B * C, 1, t, st);
    VITAE_CHECK_LAUNCH("edge blur y");
    blur_pass(sb, sa, V, B * C, 0, t, st);
    VITAE_CHECK_LAUNCH("edge blur x");
    if (edge_yt() == 4)
        launch_kernel(sobel_edge_kernel<4>, walk_grid(V, B, 4), dim3(32, 4), 0, st, static_cast<const float*>(sa), E_tgt, C, V);
    else
        launch_kernel(sobel_edge_kernel<2>, walk_grid(V, B, 2), dim3(32, 4), 0, st, static_cast<const float*>(sa), E_tgt, C, V);
    VITAE_CHECK_LAUNCH("edge sobel target");
    return 0;
}

// loss_out[0] = mean_{b,v} (E_pred - E_tgt)^2; writes D (`resid`, [B, V^3]) and keeps F = D g / |g| (in scratch) for
// vitae_edge_loss_bwd.  pred bf16 [B, L+1, P] (cls row first).
extern "C" int vitae_edge_loss_fwd(const void* pred_bf16, const float* E_tgt, float* scratch, float* resid, float* loss_out,
                                   int B, int C, int V, int p, void* stream) {
    VITAE_REQUIRE(pred_bf16 && E_tgt && scratch && resid && loss_out, "edge_loss_fwd: null pointer");
    VITAE_REQUIRE(B > 0 && V >= 4 && V % 4 == 0 && p % 4 == 0 && V % p == 0, "edge_loss_fwd: bad sizes (V, p multiples of 4)");
    VITAE_REQUIRE(C == 1 || C == 2 || C == 4, "edge_loss_fwd: in_chans must be 1, 2 or 4 (got %d)", C);
    VITAE_REQUIRE(static_cast<long long>(B) * C * V <= 65535, "edge_loss_fwd: B*C*V exceeds the grid limit");
    const size_t bc = static_cast<size_t>(B) * C * V * V * V, bv = static_cast<size_t>(B) * V * V * V;
    float* predvol = scratch;
    auto* F = reinterpret_cast<__nv_bfloat16*>(scratch + 2 * bc);
    float* partials = scratch + 2 * bc + (3 * bc + 1) / 2;
    cudaStream_t st = as_stream(stream);
    const auto* pr = static_cast<const __nv_bfloat16*>(pred_bf16);
    const float* pv = predvol;
    const dim3 blk(32, 4);
    const int yt = edge_yt();
    if (C == 4) {
        launch_kernel(unpatchify_kernel<4>, edge_grid(V, B), blk, 0, st, pr, predvol, V, p);
        VITAE_CHECK_LAUNCH("edge unpatchify");
        if (yt == 4) launch_kernel(sobel_pred_kernel<4, 4>, walk_grid(V, B, 4), blk, 0, st, pv, resid, E_tgt, F, partials, V);
        else launch_kernel(sobel_pred_kernel<4, 2>, walk_grid(V, B, 2), blk, 0, st, pv, resid, E_tgt, F, partials, V);
    } else if (C == 2) {
        launch_kernel(unpatchify_kernel<2>, edge_grid(V, B), blk, 0, st, pr, predvol, V, p);
        VITAE_CHECK_LAUNCH("edge unpatchify");
        if (yt == 4) launch_kernel(sobel_pred_kernel<2, 4>, walk_grid(V, B, 4), blk, 0, st, pv, resid, E_tgt, F, partials, V);
        else launch_kernel(sobel_pred_kernel<2, 2>, walk_grid(V, B, 2), blk, 0, st, pv, resid, E_tgt, F, partials, V);
    } else {
        launch_kernel(unpatchify_kernel<1>, edge_grid(V, B), blk, 0, st, pr, predvol, V, p);
        VITAE_CHECK_LAUNCH("edge unpatchify");
        if (yt == 4) launch_kernel(sobel_pred_kernel<1, 4>, walk_grid(V, B, 4), blk, 0, st, pv, resid, E_tgt, F, partials, V);
        else launch_kernel(sobel_pred_kernel<1, 2>, walk_grid(V, B, 2), blk, 0, st, pv, resid, E_tgt, F, partials, V);
    }
    VITAE_CHECK_LAUNCH("edge sobel pred");
    launch_kernel(edge_loss_finalize_kernel, dim3(1), dim3(256), 0, st, static_cast<const float*>(partials),
                  static_cast<long long>(edge_partials_count(B, V)), static_cast<float>(1.0 / static_cast<double>(bv)), loss_out);
    VITAE_CHECK_LAUNCH("edge finalize");
    return 0;
}

// dpred (bf16 [B, L+1, P]) += (*upstream) * d raw_edge / d pred.  Reads F from scratch; `resid` is not read (the residual
// is folded into F by the forward) and only checked.
extern "C" int vitae_edge_loss_bwd(const float* resid, const float* scratch, const float* upstream, void* dpred_bf16, int B,
                                   int C, int V, int p, void* stream) {
    VITAE_REQUIRE(resid && scratch && upstream && dpred_bf16, "edge_loss_bwd: null pointer");
    VITAE_REQUIRE(B > 0 && V >= 4 && V % 4 == 0 && p % 4 == 0 && V % p == 0, "edge_loss_bwd: bad sizes (V, p multiples of 4)");
    VITAE_REQUIRE(C == 1 || C == 2 || C == 4, "edge_loss_bwd: in_chans must be 1, 2 or 4 (got %d)", C);
    const size_t bc = static_cast<size_t>(B) * C * V * V * V, bv = static_cast<size_t>(B) * V * V * V;
    const auto* F = reinterpret_cast<const __nv_bfloat16*>(scratch + 2 * bc);
    const float k = static_cast<float>(2.0 / static_cast<double>(bv));
    auto* dp = static_cast<__nv_bfloat16*>(dpred_bf16);
    cudaStream_t st = as_stream(stream);
    const dim3 blk(32, 4);
#define VITAE_BWD(CC)                                                                                                    \
    if (edge_yt() == 4) launch_kernel(edge_loss_bwd_kernel<CC, 4>, walk_grid(V, B, 4), blk, 0, st, F, upstream, dp, V, p, k); \
    else launch_kernel(edge_loss_bwd_kernel<CC, 2>, walk_grid(V, B, 2), blk, 0, st, F, upstream, dp, V, p, k)
    if (C == 4) { VITAE_BWD(4); }
    else if (C == 2) { VITAE_BWD(2); }
    else { VITAE_BWD(1); }
#undef VITAE_BWD
    VITAE_CHECK_LAUNCH("edge_loss_bwd");
    return 0;
}
