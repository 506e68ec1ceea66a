// LayerNorm forward / backward (one warp per row, warp-shuffle reductions, float4 I/O) and the column-sum
// reduction used for bias / LayerNorm-affine gradients.  HBM-bound kernels.
#include "common.h"
#include "ptx.cuh"

namespace vitae {

constexpr int LN_WARPS = 8;        // rows per CTA iteration
constexpr int LN_MAX_VEC = 8;      // float4 per lane -> D <= 1024
constexpr int LN_BWD_MAX_BLOCKS = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// x fp32 [rows, D] -> y (bf16 and/or fp32), mean, rstd.  D % 4 == 0, D <= 1024.
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y16, float* __restrict__ y32, float* __restrict__ mean,
                     float* __restrict__ rstd, int rows, int D, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = D >> 2;  // float4 per row
    for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
        float4 v[LN_MAX_VEC];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                v[i] = xr[c];
                s += v[i].x + v[i].y + v[i].z + v[i].w;
            }
        }
        const float mu = warp_sum(s) / D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                const float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu;
                q += a * a + b * b + cc * cc + d * d;
            }
        }
        const float rs = rsqrtf(warp_sum(q) / D + eps);
        if (lane == 0) {
            if (mean) mean[row] = mu;
            if (rstd) rstd[row] = rs;
        }
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                const float4 g = reinterpret_cast<const float4*>(gamma)[c];
                const float4 b = reinterpret_cast<const float4*>(beta)[c];
                float4 o;
                o.x = (v[i].x - mu) * rs * g.x + b.x;
                o.y = (v[i].y - mu) * rs * g.y + b.y;
                o.z = (v[i].z - mu) * rs * g.z + b.z;
                o.w = (v[i].w - mu) * rs * g.w + b.w;
                if (y16) {
                    uint2 pk;
                    pk.x = pack_bf16(o.x, o.y);
                    pk.y = pack_bf16(o.z, o.w);
                    reinterpret_cast<uint2*>(y16 + static_cast<size_t>(row) * D)[c] = pk;
                }
                if (y32) reinterpret_cast<float4*>(y32 + static_cast<size_t>(row) * D)[c] = o;
            }
        }
    }
}

// dx_out = dx_in + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// partials[0][blk][:] += dy * xhat (dgamma), partials[1][blk][:] += dy (dbeta)
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32, const float* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ dx_in, float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx16,
                     float* __restrict__ partials, int rows, int D) {
    __shared__ float red[LN_WARPS][32 * 4 + 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    float4 dg[LN_MAX_VEC], db[LN_MAX_VEC];
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
        const float mu = mean[row], rs = rstd[row];
        const size_t off = static_cast<size_t>(row) * D;
        float4 xh[LN_MAX_VEC], g[LN_MAX_VEC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                const float4 xv = reinterpret_cast<const float4*>(x + off)[c];
                float4 d;
                if (dy16) {
                    const uint2 raw = reinterpret_cast<const uint2*>(dy16 + off)[c];
                    const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
                    const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
                    d = make_float4(lo.x, lo.y, hi.x, hi.y);
                } else {
                    d = reinterpret_cast<const float4*>(dy32 + off)[c];
                }
                const float4 gm = reinterpret_cast<const float4*>(gamma)[c];
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
                s1 += g[i].x + g[i].y + g[i].z + g[i].w;
                s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
                dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
                db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
            }
        }
        const float m1 = warp_sum(s1) / D, m2 = warp_sum(s2) / D;
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                float4 o;
                o.x = rs * (g[i].x - m1 - xh[i].x * m2);
                o.y = rs * (g[i].y - m1 - xh[i].y * m2);
                o.z = rs * (g[i].z - m1 - xh[i].z * m2);
                o.w = rs * (g[i].w - m1 - xh[i].w * m2);
                if (dx_in) {
                    const float4 p = reinterpret_cast<const float4*>(dx_in + off)[c];
                    o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                }
                reinterpret_cast<float4*>(dx_out + off)[c] = o;
                if (dx16) {
                    uint2 pk;
                    pk.x = pack_bf16(o.x, o.y);
                    pk.y = pack_bf16(o.z, o.w);
                    reinterpret_cast<uint2*>(dx16 + off)[c] = pk;
                }
            }
        }
    }
    // cross-warp reduction of the affine-gradient partials, one 128-column slab at a time
    float* pg = partials + static_cast<size_t>(blockIdx.x) * D;
    float* pb = partials + static_cast<size_t>(gridDim.x + blockIdx.x) * D;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
        if (32 * i >= nvec) break;
        for (int pass = 0; pass < 2; ++pass) {
            const float4 val = pass == 0 ? dg[i] : db[i];
            __syncthreads();
            red[warp][lane * 4 + 0] = val.x;
            red[warp][lane * 4 + 1] = val.y;
            red[warp][lane * 4 + 2] = val.z;
            red[warp][lane * 4 + 3] = val.w;
            __syncthreads();
            if (threadIdx.x < 128) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < LN_WARPS; ++w) s += red[w][threadIdx.x];
                const int col = 128 * i + threadIdx.x;
                if (col < D) (pass == 0 ? pg : pb)[col] = s;
            }
        }
    }
}

constexpr int CS_ROWS_PER_BLOCK = 64;

// stage 1: block (bx, by) sums rows [by*64, by*64+64) for 256 columns starting at bx*256 -> ws[by][col]
template <typename T>
__global__ void __launch_bounds__(256) colsum_stage1_kernel(const T* __restrict__ in, int rows, int cols, int ld,
                                                            float* __restrict__ ws) {
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= cols) return;
    const int r0 = blockIdx.y * CS_ROWS_PER_BLOCK;
    const int r1 = min(rows, r0 + CS_ROWS_PER_BLOCK);
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += static_cast<float>(in[static_cast<size_t>(r) * ld + col]);
    ws[static_cast<size_t>(blockIdx.y) * cols + col] = s;
}
__global__ void __launch_bounds__(256) colsum_stage2_kernel(const float* __restrict__ ws, int nblk, int cols,
                                                            float* __restrict__ out, int accumulate) {
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= cols) return;
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += ws[static_cast<size_t>(b) * cols + col];
    out[col] = accumulate ? out[col] + s : s;
}

}  // namespace vitae

using namespace vitae;

extern "C" int vitae_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                                   float* mean, float* rstd, int rows, int D, float eps, void* stream) {
    VITAE_REQUIRE(x && gamma && beta && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
    VITAE_REQUIRE(rows > 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAX_VEC, "layernorm_fwd: unsupported D=%d rows=%d", D, rows);
    const int blocks = std::min(ceil_div(rows, LN_WARPS), 148 * 4);
    layernorm_fwd_kernel<<<blocks, LN_WARPS * 32, 0, as_stream(stream)>>>(
        x, gamma, beta, static_cast<__nv_bfloat16*>(y_bf16), y_f32, mean, rstd, rows, D, eps);
    VITAE_CHECK_LAUNCH("layernorm_fwd");
    return 0;
}

extern "C" int vitae_layernorm_bwd_blocks(int rows) { return std::max(1, std::min(ceil_div(rows, LN_WARPS), LN_BWD_MAX_BLOCKS)); }

extern "C" int vitae_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma,
                                   const float* mean, const float* rstd, const float* dx_in, float* dx_out,
                                   void* dx_out_bf16, float* partials, int rows, int D, void* stream) {
    VITAE_REQUIRE((dy_bf16 != nullptr) != (dy_f32 != nullptr), "layernorm_bwd: exactly one of dy_bf16/dy_f32");
    VITAE_REQUIRE(x && gamma && mean && rstd && dx_out && partials, "layernorm_bwd: null pointer");
    VITAE_REQUIRE(rows > 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAX_VEC, "layernorm_bwd: unsupported D=%d rows=%d", D, rows);
    const int blocks = vitae_layernorm_bwd_blocks(rows);
    layernorm_bwd_kernel<<<blocks, LN_WARPS * 32, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, gamma, mean, rstd, dx_in, dx_out,
        static_cast<__nv_bfloat16*>(dx_out_bf16), partials, rows, D);
    VITAE_CHECK_LAUNCH("layernorm_bwd");
    return 0;
}

extern "C" int vitae_colsum_blocks(int rows) { return ceil_div(rows, CS_ROWS_PER_BLOCK); }

extern "C" int vitae_colsum(const void* in_bf16, const float* in_f32, int rows, int cols, int ld, float* out,
                            int accumulate, float* workspace, void* stream) {
    VITAE_REQUIRE((in_bf16 != nullptr) != (in_f32 != nullptr), "colsum: exactly one input");
    VITAE_REQUIRE(out && workspace && rows > 0 && cols > 0 && ld >= cols, "colsum: bad arguments");
    const int nblk = vitae_colsum_blocks(rows);
    dim3 grid(ceil_div(cols, 256), nblk);
    if (in_bf16)
        colsum_stage1_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(in_bf16), rows, cols, ld, workspace);
    else
        colsum_stage1_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(in_f32, rows, cols, ld, workspace);
    VITAE_CHECK_LAUNCH("colsum_stage1");
    colsum_stage2_kernel<<<ceil_div(cols, 256), 256, 0, as_stream(stream)>>>(workspace, nblk, cols, out, accumulate);
    VITAE_CHECK_LAUNCH("colsum_stage2");
    return 0;
}
