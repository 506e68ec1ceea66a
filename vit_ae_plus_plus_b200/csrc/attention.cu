// Fused multi-head self-attention, flash-style: softmax(q k^T * scale) v with scores kept on-chip
// (model/vit.py:112-121 materialises a [B,H,N,N] fp32 score tensor; 67 MB per decoder block at ViT-B/16 128^3).
//
// Layouts: qkv bf16 [B, N, 3, H, hd] (output of the qkv Linear as-is), out/dout bf16 [B, N, H*hd],
// lse/delta fp32 [B, H, N].  hd in {32, 64}; N ragged (tail masked).
//
// Round-1 implementation: warp-level mma.sync (m16n8k16 bf16, fp32 accumulate), ldmatrix from padded shared
// memory, cp.async staging, online softmax in the exp2 domain with quad-shuffle row reductions.  Attention is
// 7-14 % of the path's FLOPs; the dense contractions run on tcgen05 (gemm_tcgen05.cu).
// Backward = two kernels: dQ (one CTA per query tile, loops over kv tiles; also produces delta = rowsum(dO*O) for its
// rows) and dK/dV (one CTA per kv tile, loops over query tiles).  No atomics -> deterministic.
#include "common.h"
#include "ptx.cuh"

namespace vitae {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int ATT_THREADS = 128;
constexpr int TILE = 64;

// 64 x HD bf16 tile: global rows [row0, row0+64) (row pitch `pitch` elements) -> padded smem [64][HD+8]; rows past
// nrows are zero-filled.
template <int HD>
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* s, const __nv_bfloat16* g, size_t pitch, int row0, int nrows) {
    constexpr int LD = HD + 8;
    constexpr int CPR = HD / 8;
    for (int i = threadIdx.x; i < TILE * CPR; i += ATT_THREADS) {
        const int r = i / CPR, c = i % CPR;
        const bool valid = (row0 + r) < nrows;
        const __nv_bfloat16* src = g + static_cast<size_t>(valid ? (row0 + r) : 0) * pitch + c * 8;
        cp_async_16(smem_u32(s + r * LD + c * 8), src, valid);
    }
}

// A fragments (16 rows x HD) of this warp's rows from a padded smem tile
template <int HD>
__device__ __forceinline__ void load_a_frags(const __nv_bfloat16* s, int warp, int lane, uint32_t (&f)[HD / 16][4]) {
    constexpr int LD = HD + 8;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t addr = smem_u32(s + (warp * 16 + (lane & 15)) * LD + ks * 16 + (lane >> 4) * 8);
        ldmatrix_x4(addr, f[ks][0], f[ks][1], f[ks][2], f[ks][3]);
    }
}

// acc[8][4] (16 x 64) = A(16 x HD, frags) * T^T where T is a padded smem tile [64][HD] (rows = output columns)
// jpn: how many 16-column pairs of T hold valid rows (4 = all; VITAE_ATTN_TAIL passes fewer for a ragged tail tile: the
// skipped accumulators stay 0)
template <int HD>
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[HD / 16][4], const __nv_bfloat16* t, int lane,
                                            const int jpn = 4) {
    constexpr int LD = HD + 8;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            if (jp >= jpn) continue;
            uint32_t b0, b1, b2, b3;
            const uint32_t addr = smem_u32(t + (jp * 16 + (lane & 7) + 8 * (lane >> 4)) * LD + ks * 16 + 8 * ((lane >> 3) & 1));
            ldmatrix_x4(addr, b0, b1, b2, b3);
            mma_bf16_16816(acc[2 * jp], a[ks], b0, b1);
            mma_bf16_16816(acc[2 * jp + 1], a[ks], b2, b3);
        }
    }
}

// acc[HD/8][4] (16 x HD) += P(16 x 64, fp32 in accumulator layout, rounded to bf16) * T, T = padded smem tile [64][HD]
// ksn: how many 16-row steps of T (= 16-column pairs of P) are non-zero (4 = all)
template <int HD>
__device__ __forceinline__ void mma_p_tile(float (&acc)[HD / 8][4], const float (&p)[8][4], const __nv_bfloat16* t, int lane,
                                           const int ksn = 4) {
    constexpr int LD = HD + 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        if (ks >= ksn) continue;
        uint32_t a[4];
        a[0] = pack_bf16(p[2 * ks][0], p[2 * ks][1]);
        a[1] = pack_bf16(p[2 * ks][2], p[2 * ks][3]);
        a[2] = pack_bf16(p[2 * ks + 1][0], p[2 * ks + 1][1]);
        a[3] = pack_bf16(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
        for (int dp = 0; dp < HD / 16; ++dp) {
            uint32_t b0, b1, b2, b3;
            const uint32_t addr = smem_u32(t + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + dp * 16 + 8 * (lane >> 4));
            ldmatrix_x4_trans(addr, b0, b1, b2, b3);
            mma_bf16_16816(acc[2 * dp], a, b0, b1);
            mma_bf16_16816(acc[2 * dp + 1], a, b2, b3);
        }
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Ragged tails (opt-in build -DVITAE_ATTN_TAIL, not yet measured): the decoder's N = 513 = 8 * 64 + 1 leaves ONE valid
// row / column in the 9th tile of every head, i.e. 81 tile pairs are computed where 64.3 are needed.  With the flag the
// loops of a tail tile are bounded by its valid 16-column pairs and warps without a valid row skip the math (they still
// take part in the cooperative loads and barriers).  Without it TAIL_PAIRS is the constant 4 and WARP_HAS_ROWS true: the
// generated code is the unbounded one.
#ifdef VITAE_ATTN_TAIL
#define TAIL_PAIRS(first, n) min(4, (min(TILE, (n) - (first)) + 15) >> 4)
#define WARP_HAS_ROWS(first, n) ((first) < (n))
#else
#define TAIL_PAIRS(first, n) 4
#define WARP_HAS_ROWS(first, n) true
#endif

// ------------------------------------------------------------------------------------------------------ forward
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int H,
                float scale_log2) {
    PDL_TRIGGER_EARLY();
    pdl_wait();
    constexpr int LD = HD + 8;
    __shared__ __align__(16) __nv_bfloat16 sQ[TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sK[2][TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sV[2][TILE * LD];
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * TILE;
    const int D = H * HD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t pitch = static_cast<size_t>(3) * D;
    const __nv_bfloat16* gq = qkv + static_cast<size_t>(b) * N * pitch + h * HD;
    const __nv_bfloat16* gk = gq + D;
    const __nv_bfloat16* gv = gq + 2 * D;

    load_tile_async<HD>(sQ, gq, pitch, q0, N);
    load_tile_async<HD>(sK[0], gk, pitch, 0, N);
    load_tile_async<HD>(sV[0], gv, pitch, 0, N);
    cp_async_commit();

    const int ntiles = ceil_div(N, TILE);
    float o[HD / 8][4];
#pragma unroll
    for (int j = 0; j < HD / 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    uint32_t qf[HD / 16][4];

    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) {
            load_tile_async<HD>(sK[buf ^ 1], gk, pitch, (t + 1) * TILE, N);
            load_tile_async<HD>(sV[buf ^ 1], gv, pitch, (t + 1) * TILE, N);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (t == 0) load_a_frags<HD>(sQ, warp, lane, qf);
        if (!WARP_HAS_ROWS(q0 + warp * 16, N)) {       // no valid query row in this warp: barriers and loads only
            __syncthreads();
            continue;
        }

        float s[8][4];
        const int jpn = TAIL_PAIRS(t * TILE, N);
        mma_a_tileT<HD>(s, qf, sK[buf], lane, jpn);
        const int kv0 = t * TILE;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j >= 2 * jpn) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = kv0 + 8 * j + 2 * (lane & 3) + (e & 1);
                const float v = col < N ? s[j][e] * scale_log2 : -INFINITY;
                s[j][e] = v;
                if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
            }
        }
        const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j >= 2 * jpn) {
                s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
                continue;
            }
            s[j][0] = exp2f(s[j][0] - mn0); s[j][1] = exp2f(s[j][1] - mn0);
            s[j][2] = exp2f(s[j][2] - mn1); s[j][3] = exp2f(s[j][3] - mn1);
            rs0 += s[j][0] + s[j][1];
            rs1 += s[j][2] + s[j][3];
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int j = 0; j < HD / 8; ++j) {
            o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
        }
        mma_p_tile<HD>(o, s, sV[buf], lane, jpn);
        __syncthreads();
    }
    PDL_TRIGGER_LATE();
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) {
        const int d = 8 * j + 2 * (lane & 3);
        if (r0 < N) *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * N + r0) * D + h * HD + d) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
        if (r1 < N) *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * N + r1) * D + h * HD + d) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
    }
    if ((lane & 3) == 0) {
        float* lrow = lse + (static_cast<size_t>(b) * H + h) * N;
        if (r0 < N) lrow[r0] = (m0 + log2f(l0)) * LN2;
        if (r1 < N) lrow[r1] = (m1 + log2f(l1)) * LN2;
    }
}

// ------------------------------------------------------------------------------------------------------ backward
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                   const float* __restrict__ lse, float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int N, int H,
                   float scale, float scale_log2) {
    PDL_TRIGGER_EARLY();
    pdl_wait();
    constexpr int LD = HD + 8;
    __shared__ __align__(16) __nv_bfloat16 sQ[TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sdO[TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sK[TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sV[TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sO[TILE * LD];   // forward output rows of this query tile (for delta)
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * TILE;
    const int D = H * HD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t pitch = static_cast<size_t>(3) * D;
    const __nv_bfloat16* gq = qkv + static_cast<size_t>(b) * N * pitch + h * HD;
    const __nv_bfloat16* gk = gq + D;
    const __nv_bfloat16* gv = gq + 2 * D;
    const __nv_bfloat16* gdo = dout + static_cast<size_t>(b) * N * D + h * HD;
    const __nv_bfloat16* go = out + static_cast<size_t>(b) * N * D + h * HD;

    load_tile_async<HD>(sQ, gq, pitch, q0, N);
    load_tile_async<HD>(sdO, gdo, D, q0, N);
    load_tile_async<HD>(sO, go, D, q0, N);
    load_tile_async<HD>(sK, gk, pitch, 0, N);
    load_tile_async<HD>(sV, gv, pitch, 0, N);
    cp_async_commit();

    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    const float* lrow = lse + (static_cast<size_t>(b) * H + h) * N;
    float* drow = delta + (static_cast<size_t>(b) * H + h) * N;
    const float lse0 = r0 < N ? lrow[r0] * LOG2E : 0.f, lse1 = r1 < N ? lrow[r1] * LOG2E : 0.f;
    float dl0 = 0.f, dl1 = 0.f;   // delta = rowsum(dO * O), computed below from the staged tiles

    float dq[HD / 8][4];
#pragma unroll
    for (int j = 0; j < HD / 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[j][e] = 0.f;
    uint32_t qf[HD / 16][4], dof[HD / 16][4];
    const int ntiles = ceil_div(N, TILE);
    for (int t = 0; t < ntiles; ++t) {
        cp_async_wait<0>();
        __syncthreads();
        if (t == 0) {
            load_a_frags<HD>(sQ, warp, lane, qf);
            load_a_frags<HD>(sdO, warp, lane, dof);
            // delta for this thread's two rows: the quad shares a row, each lane covers 2 of every 8 columns
            const int lr0 = warp * 16 + (lane >> 2), lr1 = lr0 + 8;
#pragma unroll
            for (int j = 0; j < HD / 8; ++j) {
                const int d = 8 * j + 2 * (lane & 3);
                const float2 a0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sdO + lr0 * LD + d));
                const float2 o0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sO + lr0 * LD + d));
                const float2 a1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sdO + lr1 * LD + d));
                const float2 o1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sO + lr1 * LD + d));
                dl0 += a0.x * o0.x + a0.y * o0.y;
                dl1 += a1.x * o1.x + a1.y * o1.y;
            }
            dl0 = quad_sum(dl0);
            dl1 = quad_sum(dl1);
            if ((lane & 3) == 0) {   // the dK/dV kernel (launched next) reads delta from global memory
                if (r0 < N) drow[r0] = dl0;
                if (r1 < N) drow[r1] = dl1;
            }
        }
        if (WARP_HAS_ROWS(q0 + warp * 16, N)) {        // (always true without VITAE_ATTN_TAIL)
            float s[8][4], dp[8][4];
            const int jpn = TAIL_PAIRS(t * TILE, N);
            mma_a_tileT<HD>(s, qf, sK, lane, jpn);
            mma_a_tileT<HD>(dp, dof, sV, lane, jpn);
            const int kv0 = t * TILE;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j >= 2 * jpn) {
                    s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
                    continue;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = kv0 + 8 * j + 2 * (lane & 3) + (e & 1);
                    const float p = col < N ? exp2f(s[j][e] * scale_log2 - (e < 2 ? lse0 : lse1)) : 0.f;
                    s[j][e] = p * (dp[j][e] - (e < 2 ? dl0 : dl1));
                }
            }
            mma_p_tile<HD>(dq, s, sK, lane, jpn);
        }
        __syncthreads();
        if (t + 1 < ntiles) {
            load_tile_async<HD>(sK, gk, pitch, (t + 1) * TILE, N);
            load_tile_async<HD>(sV, gv, pitch, (t + 1) * TILE, N);
            cp_async_commit();
        }
    }
    PDL_TRIGGER_LATE();
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) {
        const int d = 8 * j + 2 * (lane & 3);
        if (r0 < N) *reinterpret_cast<uint32_t*>(dqkv + (static_cast<size_t>(b) * N + r0) * pitch + h * HD + d) = pack_bf16(dq[j][0] * scale, dq[j][1] * scale);
        if (r1 < N) *reinterpret_cast<uint32_t*>(dqkv + (static_cast<size_t>(b) * N + r1) * pitch + h * HD + d) = pack_bf16(dq[j][2] * scale, dq[j][3] * scale);
    }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int N, int H, float scale, float scale_log2) {
    PDL_TRIGGER_EARLY();
    pdl_wait();
    constexpr int LD = HD + 8;
    __shared__ __align__(16) __nv_bfloat16 sQ[TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sdO[TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sK[TILE * LD];
    __shared__ __align__(16) __nv_bfloat16 sV[TILE * LD];
    __shared__ float sLse[TILE], sDelta[TILE];
    const int h = blockIdx.y, b = blockIdx.z;
    const int kv0 = blockIdx.x * TILE;
    const int D = H * HD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t pitch = static_cast<size_t>(3) * D;
    const __nv_bfloat16* gq = qkv + static_cast<size_t>(b) * N * pitch + h * HD;
    const __nv_bfloat16* gk = gq + D;
    const __nv_bfloat16* gv = gq + 2 * D;
    const __nv_bfloat16* gdo = dout + static_cast<size_t>(b) * N * D + h * HD;
    const float* lrow = lse + (static_cast<size_t>(b) * H + h) * N;
    const float* drow = delta + (static_cast<size_t>(b) * H + h) * N;

    load_tile_async<HD>(sK, gk, pitch, kv0, N);
    load_tile_async<HD>(sV, gv, pitch, kv0, N);
    load_tile_async<HD>(sQ, gq, pitch, 0, N);
    load_tile_async<HD>(sdO, gdo, D, 0, N);
    cp_async_commit();

    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int j = 0; j < HD / 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dk[j][e] = 0.f; dv[j][e] = 0.f; }
    uint32_t kf[HD / 16][4], vf[HD / 16][4];
    const int ntiles = ceil_div(N, TILE);
    for (int t = 0; t < ntiles; ++t) {
        if (threadIdx.x < TILE) {
            const int r = t * TILE + threadIdx.x;
            sLse[threadIdx.x] = r < N ? lrow[r] * LOG2E : INFINITY;  // padded query rows -> p = 0
            sDelta[threadIdx.x] = r < N ? drow[r] : 0.f;
        }
        cp_async_wait<0>();
        __syncthreads();
        if (t == 0) {
            load_a_frags<HD>(sK, warp, lane, kf);
            load_a_frags<HD>(sV, warp, lane, vf);
        }
        if (WARP_HAS_ROWS(kv0 + warp * 16, N)) {       // (always true without VITAE_ATTN_TAIL)
            float pt[8][4], dpt[8][4];
            const int jpn = TAIL_PAIRS(t * TILE, N);   // valid query columns of this Q tile
            mma_a_tileT<HD>(pt, kf, sQ, lane, jpn);     // S^T  (kv rows x q cols)
            mma_a_tileT<HD>(dpt, vf, sdO, lane, jpn);   // dP^T
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j >= 2 * jpn) {
                    pt[j][0] = pt[j][1] = pt[j][2] = pt[j][3] = 0.f;
                    dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
                    continue;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int qc = 8 * j + 2 * (lane & 3) + (e & 1);
                    const float p = exp2f(pt[j][e] * scale_log2 - sLse[qc]);
                    pt[j][e] = p;
                    dpt[j][e] = p * (dpt[j][e] - sDelta[qc]);
                }
            }
            mma_p_tile<HD>(dv, pt, sdO, lane, jpn);
            mma_p_tile<HD>(dk, dpt, sQ, lane, jpn);
        }
        __syncthreads();
        if (t + 1 < ntiles) {
            load_tile_async<HD>(sQ, gq, pitch, (t + 1) * TILE, N);
            load_tile_async<HD>(sdO, gdo, D, (t + 1) * TILE, N);
            cp_async_commit();
        }
    }
    PDL_TRIGGER_LATE();
    const int r0 = kv0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) {
        const int d = 8 * j + 2 * (lane & 3);
        if (r0 < N) {
            __nv_bfloat16* base = dqkv + (static_cast<size_t>(b) * N + r0) * pitch + h * HD + d;
            *reinterpret_cast<uint32_t*>(base + D) = pack_bf16(dk[j][0] * scale, dk[j][1] * scale);
            *reinterpret_cast<uint32_t*>(base + 2 * D) = pack_bf16(dv[j][0], dv[j][1]);
        }
        if (r1 < N) {
            __nv_bfloat16* base = dqkv + (static_cast<size_t>(b) * N + r1) * pitch + h * HD + d;
            *reinterpret_cast<uint32_t*>(base + D) = pack_bf16(dk[j][2] * scale, dk[j][3] * scale);
            *reinterpret_cast<uint32_t*>(base + 2 * D) = pack_bf16(dv[j][2], dv[j][3]);
        }
    }
}

}  // namespace vitae

namespace vitae {

// Round-1 kernels, reachable through VITAE_ATTN_LEGACY=1 only (A/B timing against attention_tc.cu); not part of the C ABI.
int attention_fwd_legacy(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, void* stream) {
    VITAE_REQUIRE(qkv && out && lse, "attention_fwd: null pointer");
    VITAE_REQUIRE(B > 0 && N > 0 && H > 0 && (hd == 16 || hd == 32 || hd == 64), "attention_fwd: unsupported shape B=%d N=%d H=%d hd=%d", B, N, H, hd);
    dim3 grid(ceil_div(N, TILE), H, B);
    const float sl2 = scale * LOG2E;
    if (hd == 64)
        launch_kernel(attn_fwd_kernel<64>, dim3(grid), dim3(ATT_THREADS), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, N, H, sl2);
    else if (hd == 32)
        launch_kernel(attn_fwd_kernel<32>, dim3(grid), dim3(ATT_THREADS), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, N, H, sl2);
    else
        launch_kernel(attn_fwd_kernel<16>, dim3(grid), dim3(ATT_THREADS), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, N, H, sl2);
    VITAE_CHECK_LAUNCH("attention_fwd");
    return 0;
}

int attention_bwd_legacy(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                         void* dqkv, int B, int N, int H, int hd, float scale, void* stream) {
    VITAE_REQUIRE(qkv && out && dout && lse && delta && dqkv, "attention_bwd: null pointer");
    VITAE_REQUIRE(B > 0 && N > 0 && H > 0 && (hd == 16 || hd == 32 || hd == 64), "attention_bwd: unsupported shape B=%d N=%d H=%d hd=%d", B, N, H, hd);
    cudaStream_t st = as_stream(stream);
    const auto* q = static_cast<const __nv_bfloat16*>(qkv);
    const auto* o = static_cast<const __nv_bfloat16*>(out);
    const auto* g = static_cast<const __nv_bfloat16*>(dout);
    auto* dq = static_cast<__nv_bfloat16*>(dqkv);
    dim3 grid(ceil_div(N, TILE), H, B);
    const float sl2 = scale * LOG2E;
    if (hd == 64) {
        launch_kernel(attn_bwd_dq_kernel<64>, dim3(grid), dim3(ATT_THREADS), 0, st, q, o, g, lse, delta, dq, N, H, scale, sl2);
        VITAE_CHECK_LAUNCH("attention_bwd_dq");
        launch_kernel(attn_bwd_dkv_kernel<64>, dim3(grid), dim3(ATT_THREADS), 0, st, q, g, lse, delta, dq, N, H, scale, sl2);
    } else if (hd == 32) {
        launch_kernel(attn_bwd_dq_kernel<32>, dim3(grid), dim3(ATT_THREADS), 0, st, q, o, g, lse, delta, dq, N, H, scale, sl2);
        VITAE_CHECK_LAUNCH("attention_bwd_dq");
        launch_kernel(attn_bwd_dkv_kernel<32>, dim3(grid), dim3(ATT_THREADS), 0, st, q, g, lse, delta, dq, N, H, scale, sl2);
    } else {
        launch_kernel(attn_bwd_dq_kernel<16>, dim3(grid), dim3(ATT_THREADS), 0, st, q, o, g, lse, delta, dq, N, H, scale, sl2);
        VITAE_CHECK_LAUNCH("attention_bwd_dq");
        launch_kernel(attn_bwd_dkv_kernel<16>, dim3(grid), dim3(ATT_THREADS), 0, st, q, g, lse, delta, dq, N, H, scale, sl2);
    }
    VITAE_CHECK_LAUNCH("attention_bwd_dkv");
    return 0;
}

}  // namespace vitae
