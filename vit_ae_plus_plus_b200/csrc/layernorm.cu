// LayerNorm forward / backward (one warp per row, warp-shuffle reductions, float4 I/O) and the column-sum
// reduction used for bias / LayerNorm-affine gradients.  HBM-bound kernels.
#include "common.h"
#include <string.h>

#include "ptx.cuh"

namespace vitae {

constexpr int LN_WARPS = 8;        // rows per CTA iteration
// kernels are templated on LN_MAX_VEC = float4 per lane (D <= 128 * LN_MAX_VEC): registers follow the actual width

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// x fp32 [rows, D] -> y (bf16 and/or fp32), mean, rstd.  D % 4 == 0, D <= 128 * LN_MAX_VEC.
template <int LN_MAX_VEC>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y16, float* __restrict__ y32, float* __restrict__ mean,
                     float* __restrict__ rstd, int rows, int D, float eps) {
    pdl_trigger();
    pdl_wait();
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

// dx_out = dx_in + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma.   Row-wise only: this kernel sits on
// the backward critical path (dgrad chain), so the column reductions that give the affine / bias gradients are a
// separate kernel (layernorm_param_grads_kernel) that runs off that path.
template <int LN_MAX_VEC>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32, const float* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ dx_in, float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx16,
                     int rows, int D) {
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = D >> 2;
    for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
        const float mu = mean[row], rs = rstd[row];
        const size_t off = static_cast<size_t>(row) * D;
        float4 xh[LN_MAX_VEC], g[LN_MAX_VEC], din[LN_MAX_VEC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                const float4 xv = reinterpret_cast<const float4*>(x + off)[c];
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (dy16) {
                    const uint2 raw = reinterpret_cast<const uint2*>(dy16 + off)[c];
                    const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
                    const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
                    d = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
                if (dy32) {   // both given: the two upstream gradients are summed
                    const float4 e = reinterpret_cast<const float4*>(dy32 + off)[c];
                    d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
                }
                if (dx_in) din[i] = reinterpret_cast<const float4*>(dx_in + off)[c];
                const float4 gm = reinterpret_cast<const float4*>(gamma)[c];
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
                s1 += g[i].x + g[i].y + g[i].z + g[i].w;
                s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
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
                    o.x += din[i].x; o.y += din[i].y; o.z += din[i].z; o.w += din[i].w;
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
}

// Column reductions of the LayerNorm backward in ONE launch: out0 = dgamma = sum dy * xhat, out1 = dbeta = sum dy,
// out2 = column sums of dx_out (= gradient of the bias that was added to this residual stream).
// Block = 32 columns x 8 row lanes (coalesced 128-byte row segments); grid = (D / 32 strips, row slices).  Every block
// writes its slice's partial sums to partials[3][slices][D]; the last block of a strip to arrive (ticket counter, reset
// for the next call) adds the slices in fixed order -> deterministic.
__global__ void __launch_bounds__(256)
layernorm_param_grads_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32,
                             const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                             const float* __restrict__ dx_out, float* __restrict__ partials,
                             unsigned int* __restrict__ counters, float* __restrict__ out0, float* __restrict__ out1,
                             float* __restrict__ out2, int accumulate, int rows, int D, int rows_per_slice) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[3][8][33];
    __shared__ unsigned int ticket_sh;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * rows_per_slice, r1 = min(rows, r0 + rows_per_slice);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (col < D) {
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += 8) {
            const size_t o = static_cast<size_t>(r) * D + col;
            const float d = (dy16 ? __bfloat162float(dy16[o]) : 0.f) + (dy32 ? dy32[o] : 0.f);
            const float xh = (x[o] - mean[r]) * rstd[r];
            a0 += d * xh;
            a1 += d;
            if (dx_out) a2 += dx_out[o];
        }
    }
    red[0][ty][tx] = a0; red[1][ty][tx] = a1; red[2][ty][tx] = a2;
    __syncthreads();
    const int nsl = gridDim.y;
    if (ty < 3 && col < D) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[ty][k][tx];
        partials[(static_cast<size_t>(ty) * nsl + blockIdx.y) * D + col] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket_sh = atomicAdd(&counters[blockIdx.x], 1u);
    __syncthreads();
    if (ticket_sh != static_cast<unsigned int>(nsl - 1)) return;
    __threadfence();
    // last block of this strip: the 8 row lanes share the slices (fixed assignment and order -> deterministic)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float t = 0.f;
        if (col < D)
            for (int y = ty; y < nsl; y += 8) t += __ldcg(partials + (static_cast<size_t>(k) * nsl + y) * D + col);
        red[k][ty][tx] = t;
    }
    __syncthreads();
    if (ty < 3 && col < D) {
        float* out = ty == 0 ? out0 : (ty == 1 ? out1 : out2);
        if (out) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[ty][k][tx];
            out[col] = accumulate ? out[col] + t : t;
        }
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
}

// out[part][col] (+)= sum_blk partials[part][blk][col]: finishes LayerNorm-backward partials (up to 3 outputs, 1 launch).
// Block = 32 columns x 8 row lanes; lane ty adds blocks ty, ty+8, ...; the 8 lane sums are added in fixed order.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partials, int nblk, int D, float* __restrict__ out0,
                       float* __restrict__ out1, float* __restrict__ out2, int accumulate) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    float* out = blockIdx.y == 0 ? out0 : (blockIdx.y == 1 ? out1 : out2);
    if (out == nullptr) return;
    float s0 = 0.f, s1 = 0.f;
    if (col < D) {
        const float* p = partials + static_cast<size_t>(blockIdx.y) * nblk * D + col;
        int b = ty;
        for (; b + 8 < nblk; b += 16) {
            s0 += p[static_cast<size_t>(b) * D];
            s1 += p[static_cast<size_t>(b + 8) * D];
        }
        if (b < nblk) s0 += p[static_cast<size_t>(b) * D];
    }
    red[ty][tx] = s0 + s1;
    __syncthreads();
    if (ty == 0 && col < D) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][tx];
        out[col] = accumulate ? out[col] + s : s;
    }
}

// Column sums in ONE launch.  Block (bx, by): 256 columns starting at bx*256 (lane = 8 consecutive columns, a warp
// = one 256-column row segment), rows of slice by (warps stride the rows, 4 rows in flight per warp).  With more
// than one row slice each block writes its partial row to the workspace and the last block to arrive (ticket
// counter) adds the slices in fixed order -> deterministic; the counter is reset for the next call.
constexpr int CS_COLS = 256;

template <typename T>
__device__ __forceinline__ void cs_load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void cs_load8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void cs_load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ in, int rows, int cols, int ld, float* __restrict__ out, int accumulate,
              float* __restrict__ ws_partials, unsigned int* __restrict__ counters, int rows_per_slice,
              const float* __restrict__ scale_ptr) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[8][CS_COLS + 8];
    __shared__ unsigned int ticket_sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * CS_COLS + lane * 8;
    const int r0 = blockIdx.y * rows_per_slice;
    const int r1 = min(rows, r0 + rows_per_slice);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < cols) {
        int r = r0 + warp;
        for (; r + 24 < r1; r += 32) {
            float v0[8], v1[8], v2[8], v3[8];
            cs_load8<T>(in + static_cast<size_t>(r) * ld + c, v0);
            cs_load8<T>(in + static_cast<size_t>(r + 8) * ld + c, v1);
            cs_load8<T>(in + static_cast<size_t>(r + 16) * ld + c, v2);
            cs_load8<T>(in + static_cast<size_t>(r + 24) * ld + c, v3);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
        }
        for (; r < r1; r += 8) {
            float v0[8];
            cs_load8<T>(in + static_cast<size_t>(r) * ld + c, v0);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v0[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
    __syncthreads();
    const int col = blockIdx.x * CS_COLS + threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    const float scale = scale_ptr ? *scale_ptr : 1.0f;
    if (gridDim.y == 1) {
        if (col < cols) out[col] = accumulate ? out[col] + s * scale : s * scale;
        return;
    }
    if (col < cols) ws_partials[static_cast<size_t>(blockIdx.y) * cols + col] = s;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket_sh = atomicAdd(&counters[blockIdx.x], 1u);
    __syncthreads();
    if (ticket_sh != gridDim.y - 1) return;
    __threadfence();
    if (col < cols) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;   // four independent loads in flight, fixed association
        const int ns = static_cast<int>(gridDim.y);
        int y = 0;
        for (; y + 4 <= ns; y += 4) {
            t0 += __ldcg(ws_partials + static_cast<size_t>(y) * cols + col);
            t1 += __ldcg(ws_partials + static_cast<size_t>(y + 1) * cols + col);
            t2 += __ldcg(ws_partials + static_cast<size_t>(y + 2) * cols + col);
            t3 += __ldcg(ws_partials + static_cast<size_t>(y + 3) * cols + col);
        }
        for (; y < ns; ++y) t0 += __ldcg(ws_partials + static_cast<size_t>(y) * cols + col);
        const float t = ((t0 + t1) + (t2 + t3)) * scale;
        out[col] = accumulate ? out[col] + t : t;
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
}

// ---------------------------------------------------------------------------------------------------------------
// All column reductions of one transformer block's backward in ONE launch (they used to be six: two bias column sums,
// two LayerNorm affine-gradient reductions, two residual-gradient column sums; each a few microseconds of pure latency).
// A job is either a plain column sum of a bf16 / fp32 matrix (out0[c] = sum_r a[r, c]) or a LayerNorm reduction
// (out0 = dgamma = sum_r dy * xhat, out1 = dbeta = sum_r dy with dy = a (+ a2)).  Grid = (256-column strips of all jobs,
// row slices); lane = 8 consecutive columns, warp = one row, the 8 warps stride the rows of the slice; slices are added in
// fixed order by the last block of a strip to arrive (ticket) -> deterministic.
// ---------------------------------------------------------------------------------------------------------------
struct ColJob {
    const void* a;        // bf16 or fp32 [rows, ld]
    const float* a2;      // optional second upstream gradient (fp32, LayerNorm jobs)
    const float* x;       // LayerNorm job: forward input fp32 [rows, cols]; nullptr: plain column sum
    const float* mean;
    const float* rstd;
    float* out0;
    float* out1;
    int cols, ld, a_is_bf16, strip0;
};
struct ColJobs {
    ColJob j[6];
    int n, rows, rows_per_slice, accumulate, total_cols;   // total_cols = 256 * total strips
};

__device__ __forceinline__ void col_load8(const ColJob& jb, size_t off, float (&v)[8]) {
    if (jb.a_is_bf16) cs_load8<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(jb.a) + off, v);
    else cs_load8<float>(static_cast<const float*>(jb.a) + off, v);
}

__global__ void __launch_bounds__(256)
block_colreduce_kernel(const ColJobs jobs, float* __restrict__ partials, unsigned int* __restrict__ counters) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[2][8][CS_COLS + 8];
    __shared__ unsigned int ticket_sh;
    int ji = 0;
#pragma unroll
    for (int k = 1; k < 6; ++k)
        if (k < jobs.n && static_cast<int>(blockIdx.x) >= jobs.j[k].strip0) ji = k;
    const ColJob& jb = jobs.j[ji];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = (blockIdx.x - jb.strip0) * CS_COLS + lane * 8;
    const int r0 = blockIdx.y * jobs.rows_per_slice, r1 = min(jobs.rows, r0 + jobs.rows_per_slice);
    const bool ln = jb.x != nullptr;
    float acc0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, acc1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < jb.cols) {
        if (!ln) {
            int r = r0 + warp;
            for (; r + 24 < r1; r += 32) {
                float v0[8], v1[8], v2[8], v3[8];
                col_load8(jb, static_cast<size_t>(r) * jb.ld + c, v0);
                col_load8(jb, static_cast<size_t>(r + 8) * jb.ld + c, v1);
                col_load8(jb, static_cast<size_t>(r + 16) * jb.ld + c, v2);
                col_load8(jb, static_cast<size_t>(r + 24) * jb.ld + c, v3);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc0[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
            }
            for (; r < r1; r += 8) {
                float v0[8];
                col_load8(jb, static_cast<size_t>(r) * jb.ld + c, v0);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc0[i] += v0[i];
            }
        } else {
#pragma unroll 4
            for (int r = r0 + warp; r < r1; r += 8) {
                float d[8], xv[8];
                col_load8(jb, static_cast<size_t>(r) * jb.ld + c, d);
                cs_load8<float>(jb.x + static_cast<size_t>(r) * jb.cols + c, xv);
                if (jb.a2) {
                    float e[8];
                    cs_load8<float>(jb.a2 + static_cast<size_t>(r) * jb.cols + c, e);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d[i] += e[i];
                }
                const float mu = jb.mean[r], rs = jb.rstd[r];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc0[i] += d[i] * ((xv[i] - mu) * rs);
                    acc1[i] += d[i];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        red[0][warp][lane * 8 + i] = acc0[i];
        red[1][warp][lane * 8 + i] = acc1[i];
    }
    __syncthreads();
    const int col = (blockIdx.x - jb.strip0) * CS_COLS + threadIdx.x;     // one column per thread from here on
    const size_t gcol = static_cast<size_t>(blockIdx.x) * CS_COLS + threadIdx.x;
    const int nsl = gridDim.y;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        s0 += red[0][w][threadIdx.x];
        s1 += red[1][w][threadIdx.x];
    }
    if (nsl == 1) {
        if (col < jb.cols) {
            jb.out0[col] = jobs.accumulate ? jb.out0[col] + s0 : s0;
            if (ln && jb.out1) jb.out1[col] = jobs.accumulate ? jb.out1[col] + s1 : s1;
        }
        return;
    }
    partials[(static_cast<size_t>(0) * nsl + blockIdx.y) * jobs.total_cols + gcol] = s0;
    if (ln) partials[(static_cast<size_t>(1) * nsl + blockIdx.y) * jobs.total_cols + gcol] = s1;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket_sh = atomicAdd(&counters[blockIdx.x], 1u);
    __syncthreads();
    if (ticket_sh != static_cast<unsigned int>(nsl - 1)) return;
    __threadfence();
    if (col < jb.cols) {
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            float* out = o == 0 ? jb.out0 : jb.out1;
            if ((o == 1 && !ln) || out == nullptr) continue;
            float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // eight loads in flight, fixed association
            const float* pp = partials + static_cast<size_t>(o) * nsl * jobs.total_cols + gcol;
            int y = 0;
            for (; y + 8 <= nsl; y += 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) t[q] += __ldcg(pp + static_cast<size_t>(y + q) * jobs.total_cols);
            }
            for (; y < nsl; ++y) t[0] += __ldcg(pp + static_cast<size_t>(y) * jobs.total_cols);
            const float tt = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
            out[col] = jobs.accumulate ? out[col] + tt : tt;
        }
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
}

}  // namespace vitae

using namespace vitae;

// One launch for up to 6 column-reduction jobs over matrices with the same number of rows (see block_colreduce_kernel).
// jobs: HOST array of vitae_col_job, read during the call.  workspace: vitae_block_colreduce_workspace_bytes(...) bytes,
// zero-filled before first use (ticket counters, self-resetting), not shared by concurrent calls.
constexpr int BCR_MAX_SLICES = 32;
static inline int bcr_slices(int rows, int strips) {
    int s = std::max(1, (6 * 148) / std::max(1, strips));
    s = std::min(s, std::max(1, rows / 32));
    return std::min(s, BCR_MAX_SLICES);
}

// Sized for the largest number of row slices any launch over `rows` rows can use: a launch with FEWER jobs than the set
// the workspace was sized for has fewer strips and therefore more slices, so the bound must not depend on the strips.
extern "C" size_t vitae_block_colreduce_workspace_bytes(int rows, int total_cols_padded) {
    const int slices = std::min(BCR_MAX_SLICES, std::max(1, rows / 32));
    return 4096 + static_cast<size_t>(2) * slices * total_cols_padded * sizeof(float);
}

extern "C" int vitae_block_colreduce(const vitae_col_job* jobs, int njobs, int rows, int accumulate, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    VITAE_REQUIRE(jobs && njobs > 0 && njobs <= 6 && rows > 0 && workspace, "block_colreduce: bad arguments");
    ColJobs js;
    memset(&js, 0, sizeof(js));
    int strips = 0;
    for (int i = 0; i < njobs; ++i) {
        const vitae_col_job& j = jobs[i];
        VITAE_REQUIRE(j.a && j.out0 && j.cols > 0 && j.cols % 8 == 0 && j.ld % 8 == 0 && j.ld >= j.cols,
                      "block_colreduce: job %d: cols / ld must be multiples of 8", i);
        VITAE_REQUIRE(!j.x || (j.mean && j.rstd), "block_colreduce: job %d: LayerNorm job needs mean / rstd", i);
        ColJob& d = js.j[i];
        d.a = j.a; d.a2 = j.a2; d.x = j.x; d.mean = j.mean; d.rstd = j.rstd; d.out0 = j.out0; d.out1 = j.out1;
        d.cols = j.cols; d.ld = j.ld; d.a_is_bf16 = j.a_is_bf16; d.strip0 = strips;
        strips += ceil_div(j.cols, CS_COLS);
    }
    VITAE_REQUIRE(strips <= 1000, "block_colreduce: too many columns");
    js.n = njobs; js.rows = rows; js.accumulate = accumulate; js.total_cols = strips * CS_COLS;
    const int slices = bcr_slices(rows, strips);
    js.rows_per_slice = ceil_div(rows, slices);
    const int eff = ceil_div(rows, js.rows_per_slice);
    VITAE_REQUIRE(workspace_bytes >= 4096 + static_cast<size_t>(2) * eff * js.total_cols * sizeof(float),
                  "block_colreduce: workspace too small");
    auto* counters = static_cast<unsigned int*>(workspace);
    auto* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 4096);
    launch_kernel(block_colreduce_kernel, dim3(strips, eff), dim3(256), 0, as_stream(stream), js, partials, counters);
    VITAE_CHECK_LAUNCH("block_colreduce");
    return 0;
}

static inline int ln_vec_class(int D) {
    const int v = ceil_div(D, 128);
    return v <= 1 ? 1 : (v <= 2 ? 2 : (v <= 4 ? 4 : (v <= 6 ? 6 : 8)));
}

extern "C" int vitae_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                                   float* mean, float* rstd, int rows, int D, float eps, void* stream) {
    VITAE_REQUIRE(x && gamma && beta && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
    VITAE_REQUIRE(rows > 0 && D > 0 && D % 4 == 0 && D <= 1024, "layernorm_fwd: unsupported D=%d rows=%d", D, rows);
    const int blocks = std::min(ceil_div(rows, LN_WARPS), 148 * 4);
    auto* y16 = static_cast<__nv_bfloat16*>(y_bf16);
    cudaStream_t st = as_stream(stream);
#define VITAE_LN_FWD(V) launch_kernel(layernorm_fwd_kernel<V>, dim3(blocks), dim3(LN_WARPS * 32), 0, st, x, gamma, beta, y16, y_f32, mean, rstd, rows, D, eps)
    switch (ln_vec_class(D)) {
        case 1: VITAE_LN_FWD(1); break;
        case 2: VITAE_LN_FWD(2); break;
        case 4: VITAE_LN_FWD(4); break;
        case 6: VITAE_LN_FWD(6); break;
        default: VITAE_LN_FWD(8); break;
    }
#undef VITAE_LN_FWD
    VITAE_CHECK_LAUNCH("layernorm_fwd");
    return 0;
}

extern "C" int vitae_layernorm_bwd_blocks(int rows) { return std::max(1, std::min(ceil_div(rows, 32), 64)); }

extern "C" int vitae_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma,
                                   const float* mean, const float* rstd, const float* dx_in, float* dx_out,
                                   void* dx_out_bf16, int rows, int D, void* stream) {
    VITAE_REQUIRE(dy_bf16 != nullptr || dy_f32 != nullptr, "layernorm_bwd: need dy_bf16 and/or dy_f32");
    VITAE_REQUIRE(x && gamma && mean && rstd && dx_out, "layernorm_bwd: null pointer");
    VITAE_REQUIRE(rows > 0 && D > 0 && D % 4 == 0 && D <= 1024, "layernorm_bwd: unsupported D=%d rows=%d", D, rows);
    const int blocks = std::min(ceil_div(rows, LN_WARPS), 148 * 4);
    const auto* dy16 = static_cast<const __nv_bfloat16*>(dy_bf16);
    auto* dx16 = static_cast<__nv_bfloat16*>(dx_out_bf16);
    cudaStream_t st = as_stream(stream);
#define VITAE_LN_BWD(V) launch_kernel(layernorm_bwd_kernel<V>, dim3(blocks), dim3(LN_WARPS * 32), 0, st, dy16, dy_f32, x, gamma, mean, rstd, dx_in, dx_out, dx16, rows, D)
    switch (ln_vec_class(D)) {
        case 1: VITAE_LN_BWD(1); break;
        case 2: VITAE_LN_BWD(2); break;
        case 4: VITAE_LN_BWD(4); break;
        case 6: VITAE_LN_BWD(6); break;
        default: VITAE_LN_BWD(8); break;
    }
#undef VITAE_LN_BWD
    VITAE_CHECK_LAUNCH("layernorm_bwd");
    return 0;
}

extern "C" size_t vitae_layernorm_param_grads_workspace_bytes(int rows, int D) {
    return 1024 + static_cast<size_t>(3) * vitae_layernorm_bwd_blocks(rows) * D * sizeof(float);
}

extern "C" int vitae_layernorm_param_grads(const void* dy_bf16, const float* dy_f32, const float* x, const float* mean,
                                           const float* rstd, const float* dx_out, void* workspace, float* dgamma,
                                           float* dbeta, float* dbias, int accumulate, int rows, int D, void* stream) {
    VITAE_REQUIRE(dy_bf16 != nullptr || dy_f32 != nullptr, "layernorm_param_grads: need dy_bf16 and/or dy_f32");
    VITAE_REQUIRE(x && mean && rstd && workspace && rows > 0 && D > 0 && D <= 32 * 256, "layernorm_param_grads: bad arguments");
    VITAE_REQUIRE(!dbias || dx_out, "layernorm_param_grads: dbias needs dx_out");
    const int nblk = vitae_layernorm_bwd_blocks(rows);
    const int rows_per_slice = ceil_div(rows, nblk);
    auto* counters = static_cast<unsigned int*>(workspace);                       // [<= 256] zero at first use, self-resetting
    auto* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 1024);
    dim3 grid(ceil_div(D, 32), nblk);
    launch_kernel(layernorm_param_grads_kernel, grid, dim3(256), 0, as_stream(stream),
                  static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, mean, rstd, dx_out, partials, counters, dgamma, dbeta,
                  dbias, accumulate, rows, D, rows_per_slice);
    VITAE_CHECK_LAUNCH("layernorm_param_grads");
    return 0;
}

extern "C" int vitae_reduce_partials(const float* partials, int nblk, int D, float* out0, float* out1, float* out2,
                                     int accumulate, void* stream) {
    VITAE_REQUIRE(partials && nblk > 0 && D > 0 && (out0 || out1 || out2), "reduce_partials: bad arguments");
    dim3 grid(ceil_div(D, 32), 3);
    launch_kernel(reduce_partials_kernel, dim3(grid), dim3(256), 0, as_stream(stream), partials, nblk, D, out0, out1, out2, accumulate);
    VITAE_CHECK_LAUNCH("reduce_partials");
    return 0;
}

static inline int colsum_slices(int rows, int cols) {
    const int strips = ceil_div(cols, CS_COLS);
    int s = std::max(1, (2 * 148) / strips);
    s = std::min(s, std::max(1, rows / 32));
    return std::min(s, 16);
}

extern "C" size_t vitae_colsum_workspace_bytes(int rows, int cols) {
    const int strips = ceil_div(cols, CS_COLS);
    return static_cast<size_t>(strips) * sizeof(unsigned int) + 256 + static_cast<size_t>(colsum_slices(rows, cols)) * cols * sizeof(float);
}

extern "C" int vitae_colsum(const void* in_bf16, const float* in_f32, int rows, int cols, int ld, float* out,
                            int accumulate, void* workspace, const float* scale_ptr, void* stream) {
    VITAE_REQUIRE((in_bf16 != nullptr) != (in_f32 != nullptr), "colsum: exactly one input");
    VITAE_REQUIRE(out && workspace && rows > 0 && cols > 0 && ld >= cols, "colsum: bad arguments");
    VITAE_REQUIRE(cols % 8 == 0 && ld % 8 == 0, "colsum: cols and ld must be multiples of 8 (cols=%d ld=%d)", cols, ld);
    const int strips = ceil_div(cols, CS_COLS);
    const int slices = colsum_slices(rows, cols);
    const int rows_per_slice = ceil_div(rows, slices);
    const int eff_slices = ceil_div(rows, rows_per_slice);
    auto* counters = static_cast<unsigned int*>(workspace);
    auto* parts = reinterpret_cast<float*>(static_cast<char*>(workspace) + ((static_cast<size_t>(strips) * sizeof(unsigned int) + 255) / 256) * 256);
    dim3 grid(strips, eff_slices);
    if (in_bf16)
        launch_kernel(colsum_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(in_bf16), rows, cols, ld, out, accumulate, parts, counters, rows_per_slice, scale_ptr);
    else
        launch_kernel(colsum_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), in_f32, rows, cols, ld, out, accumulate, parts, counters, rows_per_slice, scale_ptr);
    VITAE_CHECK_LAUNCH("colsum");
    return 0;
}
