// Auxiliary edge-map loss of the MAE (model/vit_autoenc.py:221-224, shipped default use_edge_map = yes):
//     raw_edge = mse( sobel(unpatchify(pred)),  sobel(gaussian_blur(target, sigma = 2)) )
// Sobel: model/model_utils/sobel_filter.py:10-45 -- three 3x3x3 directional kernels per channel (smoothing [1,2,1] x
// [1,2,1] x derivative), zero padding, sqrt(gx^2+gy^2+gz^2) summed over the channels.  Blur: gaussian_filter.py:5-26 --
// dense ks^3 kernel = outer product of 11 normalised taps (linspace(-6, 6, 11) quirk: taps 1.2 apart), zero padding; run
// here as three 1-D passes (identical under zero padding).
//
// All kernels are stencils over the [B*C, V, V, V] fp32 volumes: HBM / L2 bound, one thread per voxel, neighbours served
// by L1.  Forward for pred keeps the normalised gradients n_i = g_i / |g| (3 per channel and voxel) and the residual
// D = E_pred - E_target, so that the backward is one transposed-stencil pass:
//     d raw_edge / d P_c(u) = 2 / (B V^3) * sum_i sum_d K_i[d] * D(u + 1 - d) * n_{i,c}(u + 1 - d)
// (|g| = 0 -> contribution 0: the reference's sqrt'(0) * 0 would be NaN there).  The gradient is added to
// the bf16 dpred buffer in the patch layout (pz, py, px, c) of model/vit_autoenc.py:100-113.
#include "common.h"
#include "ptx.cuh"

namespace vitae {

struct Taps {
    float t[16];
    int n;
};

// Thread geometry of every stencil kernel: a thread owns 4 consecutive x (one float4), a warp 128 consecutive x of one
// row, block = (32, 4) -> 4 rows; grid = (ceil(V/128), ceil(V/4), BC*V): no 64-bit div/mod per thread, neighbours along
// x are reused from registers, all row accesses are coalesced 16-byte vectors.
constexpr int XT = 4;
#define EDGE_COORDS(nplanes)                                                      \
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * XT;                           \
    const int y = blockIdx.y * 4 + threadIdx.y;                                    \
    const int z = blockIdx.z % V;                                                  \
    const int plane = blockIdx.z / V; /* b*C + c, or b */                          \
    (void)(nplanes);                                                               \
    if (x0 >= V || y >= V) return;

__device__ __forceinline__ size_t vox(int V, int plane, int z, int y, int x) {
    return ((static_cast<size_t>(plane) * V + z) * V + y) * V + x;
}

// 6 consecutive values x0-1 .. x0+4 of row (plane, z, y) with zero padding (the row itself may be out of range)
__device__ __forceinline__ void load_row6(const float* __restrict__ P, int V, int plane, int z, int y, int x0, float (&r)[6]) {
    if (z < 0 || z >= V || y < 0 || y >= V) {
#pragma unroll
        for (int i = 0; i < 6; ++i) r[i] = 0.f;
        return;
    }
    const float* row = P + vox(V, plane, z, y, 0);
    const float4 m = *reinterpret_cast<const float4*>(row + x0);
    r[0] = x0 > 0 ? row[x0 - 1] : 0.f;
    r[1] = m.x; r[2] = m.y; r[3] = m.z; r[4] = m.w;
    r[5] = x0 + 4 < V ? row[x0 + 4] : 0.f;
}

// out(u) = sum_k taps[k] * in(u + (k - n/2) e_axis), zero outside; axis 0 = x (fastest), 1 = y, 2 = z
__global__ void __launch_bounds__(128)
blur_axis_kernel(const float* __restrict__ in, float* __restrict__ out, int V, int axis, const Taps taps) {
    pdl_trigger();
    pdl_wait();
    EDGE_COORDS(0)
    const int h = taps.n / 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (axis == 0) {
        // values x0-8 .. x0+11 cover every tap of the 4 outputs for n <= 17 taps: five aligned float4
        float w[20];
        const float* row = in + vox(V, plane, z, y, 0);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int xs = x0 - 8 + 4 * q;
            const float4 v = (xs >= 0 && xs < V) ? *reinterpret_cast<const float4*>(row + xs) : make_float4(0.f, 0.f, 0.f, 0.f);
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < taps.n) {
                const float t = taps.t[k];
                acc.x += t * w[8 + 0 + k - h]; acc.y += t * w[8 + 1 + k - h];
                acc.z += t * w[8 + 2 + k - h]; acc.w += t * w[8 + 3 + k - h];
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < taps.n) {
                const int yy = axis == 1 ? y + k - h : y, zz = axis == 2 ? z + k - h : z;
                if (yy >= 0 && yy < V && zz >= 0 && zz < V) {
                    const float4 v = *reinterpret_cast<const float4*>(in + vox(V, plane, zz, yy, x0));
                    const float t = taps.t[k];
                    acc.x += t * v.x; acc.y += t * v.y; acc.z += t * v.z; acc.w += t * v.w;
                }
            }
        }
    }
    *reinterpret_cast<float4*>(out + vox(V, plane, z, y, x0)) = acc;
}

// pred bf16 [B, Nd = L+1, P] (row 0 of each sample = cls) -> predvol fp32 [B*C, V, V, V]; plane = b here
template <int C>
__global__ void __launch_bounds__(128)
unpatchify_kernel(const __nv_bfloat16* __restrict__ pred, float* __restrict__ vol, int V, int p) {
    pdl_trigger();
    pdl_wait();
    EDGE_COORDS(0)
    const int g = V / p;
    const int l = ((z / p) * g + (y / p)) * g + (x0 / p);          // 4 consecutive x stay inside one patch (p % 4 == 0)
    const int within = (((z % p) * p + (y % p)) * p + (x0 % p)) * C;
    const size_t L1 = static_cast<size_t>(g) * g * g + 1;
    const __nv_bfloat16* src = pred + (static_cast<size_t>(plane) * L1 + 1 + l) * (static_cast<size_t>(p) * p * p * C) + within;
    float v[XT][C];
#pragma unroll
    for (int i = 0; i < XT; ++i)
#pragma unroll
        for (int c = 0; c < C; ++c) v[i][c] = __bfloat162float(src[i * C + c]);
#pragma unroll
    for (int c = 0; c < C; ++c)
        *reinterpret_cast<float4*>(vol + vox(V, plane * C + c, z, y, x0)) = make_float4(v[0][c], v[1][c], v[2][c], v[3][c]);
}

// Sobel responses of the 4 voxels x0..x0+3 of row (plane, z, y): s = [1,2,1]; g0: d/dx (+,0,-), g1: d/dy (-,0,+),
// g2: d/dz (-,0,+)   (weight[0..2] of model/model_utils/sobel_filter.py:10-35)
__device__ __forceinline__ void sobel4(const float* __restrict__ P, int V, int plane, int z, int y, int x0, float (&g0)[XT],
                                       float (&g1)[XT], float (&g2)[XT]) {
    const float s[3] = {1.f, 2.f, 1.f};
#pragma unroll
    for (int i = 0; i < XT; ++i) g0[i] = g1[i] = g2[i] = 0.f;
#pragma unroll
    for (int dz = 0; dz < 3; ++dz)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            float r[6];
            load_row6(P, V, plane, z + dz - 1, y + dy - 1, x0, r);
            const float wy = dy == 0 ? -1.f : (dy == 2 ? 1.f : 0.f);   // derivative taps along y / z: (-,0,+)
            const float wz = dz == 0 ? -1.f : (dz == 2 ? 1.f : 0.f);
#pragma unroll
            for (int i = 0; i < XT; ++i) {
                const float sm = r[i] + 2.f * r[i + 1] + r[i + 2];     // [1,2,1] along x
                g0[i] += s[dz] * s[dy] * (r[i] - r[i + 2]);
                g1[i] += s[dz] * wy * sm;
                g2[i] += wz * s[dy] * sm;
            }
        }
}

// E(b, v) = sum_c |sobel(vol_c)(v)|.  WITH_GRAD: also n = g / |g| per channel (bf16), D = E - E_tgt (stored in E), and
// block sums of D^2.  plane = b.
template <bool WITH_GRAD>
__global__ void __launch_bounds__(128)
sobel_edge_kernel(const float* __restrict__ vol, float* __restrict__ E, const float* __restrict__ E_tgt,
                  __nv_bfloat16* __restrict__ nrm, float* __restrict__ partials, int C, int V) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[4];
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * XT;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int z = blockIdx.z % V;
    const int b = blockIdx.z / V;
    float sq = 0.f;
    if (x0 < V && y < V) {
        float e[XT] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < C; ++c) {
            float g0[XT], g1[XT], g2[XT];
            sobel4(vol, V, b * C + c, z, y, x0, g0, g1, g2);
            uint2 p0, p1, p2;
            float n0[XT], n1[XT], n2[XT];
#pragma unroll
            for (int i = 0; i < XT; ++i) {
                const float mag = sqrtf(g0[i] * g0[i] + g1[i] * g1[i] + g2[i] * g2[i]);
                e[i] += mag;
                const float inv = mag > 0.f ? 1.f / mag : 0.f;
                n0[i] = g0[i] * inv; n1[i] = g1[i] * inv; n2[i] = g2[i] * inv;
            }
            if (WITH_GRAD) {
                p0.x = pack_bf16(n0[0], n0[1]); p0.y = pack_bf16(n0[2], n0[3]);
                p1.x = pack_bf16(n1[0], n1[1]); p1.y = pack_bf16(n1[2], n1[3]);
                p2.x = pack_bf16(n2[0], n2[1]); p2.y = pack_bf16(n2[2], n2[3]);
                *reinterpret_cast<uint2*>(nrm + vox(V, (b * C + c) * 3 + 0, z, y, x0)) = p0;
                *reinterpret_cast<uint2*>(nrm + vox(V, (b * C + c) * 3 + 1, z, y, x0)) = p1;
                *reinterpret_cast<uint2*>(nrm + vox(V, (b * C + c) * 3 + 2, z, y, x0)) = p2;
            }
        }
        float* eo = E + vox(V, b, z, y, x0);
        if (WITH_GRAD) {
            const float4 t = *reinterpret_cast<const float4*>(E_tgt + vox(V, b, z, y, x0));
            const float4 d = make_float4(e[0] - t.x, e[1] - t.y, e[2] - t.z, e[3] - t.w);
            *reinterpret_cast<float4*>(eo) = d;          // the residual D replaces E (only D is needed afterwards)
            sq = d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        } else {
            *reinterpret_cast<float4*>(eo) = make_float4(e[0], e[1], e[2], e[3]);
        }
    }
    if (WITH_GRAD) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (threadIdx.x == 0) red[threadIdx.y] = sq;
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0)
            partials[(static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] =
                (red[0] + red[1]) + (red[2] + red[3]);
    }
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

// 6 consecutive bf16 values x0-1 .. x0+4 of a row of the normalised-gradient planes, zero padded
__device__ __forceinline__ void load_row6_bf16(const __nv_bfloat16* __restrict__ row, int V, int x0, float (&r)[6]) {
    const uint2 m = *reinterpret_cast<const uint2*>(row + x0);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&m.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&m.y));
    r[0] = x0 > 0 ? __bfloat162float(row[x0 - 1]) : 0.f;
    r[1] = a.x; r[2] = a.y; r[3] = b.x; r[4] = b.y;
    r[5] = x0 + 4 < V ? __bfloat162float(row[x0 + 4]) : 0.f;
}

// dpred[b, 1 + l, (pz,py,px,c)] += (*upstream) * 2 / (B V^3) * sum_i sum_d K_i[d] D(u+1-d) n_{i,c}(u+1-d); plane = b
template <int C>
__global__ void __launch_bounds__(128)
edge_loss_bwd_kernel(const float* __restrict__ D, const __nv_bfloat16* __restrict__ nrm, const float* __restrict__ upstream,
                     __nv_bfloat16* __restrict__ dpred, int V, int p, float two_over_count) {
    pdl_trigger();
    pdl_wait();
    EDGE_COORDS(0)
    const int b = plane;
    const float s[3] = {1.f, 2.f, 1.f};
    float acc[XT][C];
#pragma unroll
    for (int i = 0; i < XT; ++i)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[i][c] = 0.f;
#pragma unroll
    for (int dz = 0; dz < 3; ++dz)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int zz = z + 1 - dz, yy = y + 1 - dy;      // v = u + 1 - d
            if (zz < 0 || zz >= V || yy < 0 || yy >= V) continue;
            float dr[6];
            load_row6(D, V, b, zz, yy, x0, dr);
            // kernel taps at (dz, dy): K0 = s s dk[dx] (dk = +,0,-), K1 = s (-dk[dy]) s[dx], K2 = (-dk[dz]) s s[dx]
            const float k1y = dy == 0 ? -1.f : (dy == 2 ? 1.f : 0.f);
            const float k2z = dz == 0 ? -1.f : (dz == 2 ? 1.f : 0.f);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float n0[6], n1[6], n2[6];
                load_row6_bf16(nrm + vox(V, (b * C + c) * 3 + 0, zz, yy, 0), V, x0, n0);
                if (dy != 1) load_row6_bf16(nrm + vox(V, (b * C + c) * 3 + 1, zz, yy, 0), V, x0, n1);
                if (dz != 1) load_row6_bf16(nrm + vox(V, (b * C + c) * 3 + 2, zz, yy, 0), V, x0, n2);
#pragma unroll
                for (int i = 0; i < XT; ++i) {
                    // neighbours along x: v_x = u_x + 1 - dx -> row index i + 2 - dx; dx = 0: +1 tap of K0, dx = 2: -1 tap
                    float t = s[dz] * s[dy] * (dr[i + 2] * n0[i + 2] - dr[i] * n0[i]);
                    if (dy != 1)
                        t += s[dz] * k1y * (dr[i + 2] * n1[i + 2] + 2.f * dr[i + 1] * n1[i + 1] + dr[i] * n1[i]);
                    if (dz != 1)
                        t += k2z * s[dy] * (dr[i + 2] * n2[i + 2] + 2.f * dr[i + 1] * n2[i + 1] + dr[i] * n2[i]);
                    acc[i][c] += t;
                }
            }
        }
    const float scale = (*upstream) * two_over_count;
    const int g = V / p;
    const int l = ((z / p) * g + (y / p)) * g + (x0 / p);
    const int within = (((z % p) * p + (y % p)) * p + (x0 % p)) * C;
    const size_t L1 = static_cast<size_t>(g) * g * g + 1;
    __nv_bfloat16* dst = dpred + (static_cast<size_t>(b) * L1 + 1 + l) * (static_cast<size_t>(p) * p * p * C) + within;
#pragma unroll
    for (int i = 0; i < XT; ++i)
#pragma unroll
        for (int c = 0; c < C; ++c)
            dst[i * C + c] = __float2bfloat16(__bfloat162float(dst[i * C + c]) + scale * acc[i][c]);
}

static inline dim3 edge_grid(int V, int planes) { return dim3(ceil_div(V, 128), ceil_div(V, 4), planes * V); }

}  // namespace vitae

using namespace vitae;

// scratch layout (floats): [0, bc) volume A / predvol, [bc, 2bc) volume B, [2bc, 2bc + 1.5bc) normalised gradients (bf16,
// 3 per channel voxel), then the block partials
static inline size_t edge_partials_count(int B, int V) {
    return static_cast<size_t>(ceil_div(V, 128)) * ceil_div(V, 4) * B * V;
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
    const dim3 blk(32, 4);
    launch_kernel(blur_axis_kernel, edge_grid(V, B * C), blk, 0, st, vol, sa, V, 2, t);
    VITAE_CHECK_LAUNCH("edge blur z");
    launch_kernel(blur_axis_kernel, edge_grid(V, B * C), blk, 0, st, static_cast<const float*>(sa), sb, V, 1, t);
    VITAE_CHECK_LAUNCH("edge blur y");
    launch_kernel(blur_axis_kernel, edge_grid(V, B * C), blk, 0, st, static_cast<const float*>(sb), sa, V, 0, t);
    VITAE_CHECK_LAUNCH("edge blur x");
    launch_kernel(sobel_edge_kernel<false>, edge_grid(V, B), blk, 0, st, static_cast<const float*>(sa), E_tgt,
                  static_cast<const float*>(nullptr), static_cast<__nv_bfloat16*>(nullptr), static_cast<float*>(nullptr), C, V);
    VITAE_CHECK_LAUNCH("edge sobel target");
    return 0;
}

// loss_out[0] = mean_{b,v} (E_pred - E_tgt)^2; keeps D (in `resid`, [B, V^3]) and the normalised gradients (in scratch) for
// vitae_edge_loss_bwd.  pred bf16 [B, L+1, P] (cls row first).
extern "C" int vitae_edge_loss_fwd(const void* pred_bf16, const float* E_tgt, float* scratch, float* resid, float* loss_out,
                                   int B, int C, int V, int p, void* stream) {
    VITAE_REQUIRE(pred_bf16 && E_tgt && scratch && resid && loss_out, "edge_loss_fwd: null pointer");
    VITAE_REQUIRE(B > 0 && V >= 4 && V % 4 == 0 && p % 4 == 0 && V % p == 0, "edge_loss_fwd: bad sizes (V, p multiples of 4)");
    VITAE_REQUIRE(C == 1 || C == 2 || C == 4, "edge_loss_fwd: in_chans must be 1, 2 or 4 (got %d)", C);
    VITAE_REQUIRE(static_cast<long long>(B) * C * V <= 65535, "edge_loss_fwd: B*C*V exceeds the grid limit");
    const size_t bc = static_cast<size_t>(B) * C * V * V * V, bv = static_cast<size_t>(B) * V * V * V;
    float* predvol = scratch;
    auto* nrm = reinterpret_cast<__nv_bfloat16*>(scratch + 2 * bc);
    float* partials = scratch + 2 * bc + (3 * bc + 1) / 2;
    cudaStream_t st = as_stream(stream);
    const auto* pr = static_cast<const __nv_bfloat16*>(pred_bf16);
    const dim3 blk(32, 4);
    if (C == 4) launch_kernel(unpatchify_kernel<4>, edge_grid(V, B), blk, 0, st, pr, predvol, V, p);
    else if (C == 2) launch_kernel(unpatchify_kernel<2>, edge_grid(V, B), blk, 0, st, pr, predvol, V, p);
    else launch_kernel(unpatchify_kernel<1>, edge_grid(V, B), blk, 0, st, pr, predvol, V, p);
    VITAE_CHECK_LAUNCH("edge unpatchify");
    launch_kernel(sobel_edge_kernel<true>, edge_grid(V, B), blk, 0, st, static_cast<const float*>(predvol), resid, E_tgt, nrm,
                  partials, C, V);
    VITAE_CHECK_LAUNCH("edge sobel pred");
    launch_kernel(edge_loss_finalize_kernel, dim3(1), dim3(256), 0, st, static_cast<const float*>(partials),
                  static_cast<long long>(edge_partials_count(B, V)), static_cast<float>(1.0 / static_cast<double>(bv)), loss_out);
    VITAE_CHECK_LAUNCH("edge finalize");
    return 0;
}

// dpred (bf16 [B, L+1, P]) += (*upstream) * d raw_edge / d pred
extern "C" int vitae_edge_loss_bwd(const float* resid, const float* scratch, const float* upstream, void* dpred_bf16, int B,
                                   int C, int V, int p, void* stream) {
    VITAE_REQUIRE(resid && scratch && upstream && dpred_bf16, "edge_loss_bwd: null pointer");
    VITAE_REQUIRE(B > 0 && V >= 4 && V % 4 == 0 && p % 4 == 0 && V % p == 0, "edge_loss_bwd: bad sizes (V, p multiples of 4)");
    VITAE_REQUIRE(C == 1 || C == 2 || C == 4, "edge_loss_bwd: in_chans must be 1, 2 or 4 (got %d)", C);
    const size_t bc = static_cast<size_t>(B) * C * V * V * V, bv = static_cast<size_t>(B) * V * V * V;
    const auto* nrm = reinterpret_cast<const __nv_bfloat16*>(scratch + 2 * bc);
    const float k = static_cast<float>(2.0 / static_cast<double>(bv));
    auto* dp = static_cast<__nv_bfloat16*>(dpred_bf16);
    cudaStream_t st = as_stream(stream);
    const dim3 blk(32, 4);
    if (C == 4) launch_kernel(edge_loss_bwd_kernel<4>, edge_grid(V, B), blk, 0, st, resid, nrm, upstream, dp, V, p, k);
    else if (C == 2) launch_kernel(edge_loss_bwd_kernel<2>, edge_grid(V, B), blk, 0, st, resid, nrm, upstream, dp, V, p, k);
    else launch_kernel(edge_loss_bwd_kernel<1>, edge_grid(V, B), blk, 0, st, resid, nrm, upstream, dp, V, p, k);
    VITAE_CHECK_LAUNCH("edge_loss_bwd");
    return 0;
}
