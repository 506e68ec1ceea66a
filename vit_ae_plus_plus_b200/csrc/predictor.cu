// Contrastive head of ContrastiveMAEViT: the pieces of
//   predictor = Linear(D, D, bias=False) -> BatchNorm1d(D) -> ReLU -> Linear(D, D)      (model/vit_autoenc.py:263-268)
// that are not GEMMs, and the loop's cosine loss -(cos(p1, z2).mean() + cos(p2, z1).mean()) / 2 * contr_weight
// (utils/train_one_epoch.py:32,113-114).  The two Linear layers run on the tcgen05 GEMM (gemm_tcgen05.cu).
//
// BatchNorm1d in training mode normalises every feature column with the batch statistics of the M = B * (keep + 1) token
// rows (biased variance) and updates running_mean / running_var (momentum, unbiased variance).  A block owns 32 columns
// (lane = column: 128-byte coalesced rows), its 8 warps stride the rows; the column's rows are read twice from L2
// (mean, then centred sum of squares: no E[x^2] - mean^2 cancellation), in fixed order -> deterministic.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace vitae {

constexpr int BN_WARPS = 8;

__device__ __forceinline__ float bn_block_colsum(float v, float (*sh)[32], int warp, int lane) {
    sh[warp][lane] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < BN_WARPS; ++w) t += sh[w][lane];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(BN_WARPS * 32)
bn_relu_fwd_kernel(const float* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   __nv_bfloat16* __restrict__ act, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, int M, int D) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sh[BN_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + lane;
    const bool ok = c < D;
    float s = 0.f;
    if (ok)
        for (int r = warp; r < M; r += BN_WARPS) s += h[static_cast<size_t>(r) * D + c];
    const float mean = bn_block_colsum(s, sh, warp, lane) / M;
    float q = 0.f;
    if (ok)
        for (int r = warp; r < M; r += BN_WARPS) {
            const float d = h[static_cast<size_t>(r) * D + c] - mean;
            q = fmaf(d, d, q);
        }
    const float ssq = bn_block_colsum(q, sh, warp, lane);
    const float var = ssq / M;
    const float rstd = rsqrtf(var + eps);
    if (ok) {
        const float g = gamma[c], b = beta[c];
        for (int r = warp; r < M; r += BN_WARPS) {
            const float y = (h[static_cast<size_t>(r) * D + c] - mean) * rstd * g + b;
            act[static_cast<size_t>(r) * D + c] = __float2bfloat16(fmaxf(y, 0.f));
        }
        if (warp == 0) {
            mean_out[c] = mean;
            rstd_out[c] = rstd;
            if (running_mean) {       // torch: running = (1 - momentum) * running + momentum * stat, unbiased variance
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (M > 1 ? ssq / (M - 1) : var);
            }
        }
    }
}

// dact: gradient w.r.t. the ReLU output (bf16); dh = gradient w.r.t. the BatchNorm input (bf16, GEMM operand);
// dgamma / dbeta (+)= column sums.
__global__ void __launch_bounds__(BN_WARPS * 32)
bn_relu_bwd_kernel(const __nv_bfloat16* __restrict__ dact, const float* __restrict__ h, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                   __nv_bfloat16* __restrict__ dh, float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate, int M,
                   int D) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sh[BN_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + lane;
    const bool ok = c < D;
    const float mu = ok ? mean[c] : 0.f, rs = ok ? rstd[c] : 0.f, g = ok ? gamma[c] : 0.f, b = ok ? beta[c] : 0.f;
    float sdy = 0.f, sdyx = 0.f;
    if (ok)
        for (int r = warp; r < M; r += BN_WARPS) {
            const float xh = (h[static_cast<size_t>(r) * D + c] - mu) * rs;
            const float dy = xh * g + b > 0.f ? __bfloat162float(dact[static_cast<size_t>(r) * D + c]) : 0.f;
            sdy += dy;
            sdyx = fmaf(dy, xh, sdyx);
        }
    const float db = bn_block_colsum(sdy, sh, warp, lane);
    const float dg = bn_block_colsum(sdyx, sh, warp, lane);
    if (ok) {
        const float k = g * rs, inv_m = 1.f / M;
        for (int r = warp; r < M; r += BN_WARPS) {
            const float xh = (h[static_cast<size_t>(r) * D + c] - mu) * rs;
            const float dy = xh * g + b > 0.f ? __bfloat162float(dact[static_cast<size_t>(r) * D + c]) : 0.f;
            dh[static_cast<size_t>(r) * D + c] = __float2bfloat16(k * (dy - inv_m * (db + xh * dg)));
        }
        if (warp == 0) {
            dgamma[c] = accumulate ? dgamma[c] + dg : dg;
            dbeta[c] = accumulate ? dbeta[c] + db : db;
        }
    }
}

// ---- cosine loss: rows a_i, b_i -> cos_i = <a, b> / (max(|a|, eps) * max(|b|, eps))   (torch.nn.CosineSimilarity, eps 1e-8)
// stats[pair][row] = {dot, |a|, |b|}; partial[pair][block] = sum of cos over the block's rows (fixed order)
constexpr int COS_ROWS = 8;   // warps (= rows) per block
__global__ void __launch_bounds__(COS_ROWS * 32)
cosine_rows_kernel(const float* __restrict__ a0, const float* __restrict__ b0, const float* __restrict__ a1,
                   const float* __restrict__ b1, int M, int D, float* __restrict__ stats, float* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sh[COS_ROWS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.y;
    const float* a = pair ? a1 : a0;
    const float* b = pair ? b1 : b0;
    const int r = blockIdx.x * COS_ROWS + warp;
    float cosv = 0.f;
    if (r < M) {
        float dot = 0.f, na = 0.f, nb = 0.f;
        for (int c = lane; c < D; c += 32) {
            const float x = a[static_cast<size_t>(r) * D + c], y = b[static_cast<size_t>(r) * D + c];
            dot = fmaf(x, y, dot);
            na = fmaf(x, x, na);
            nb = fmaf(y, y, nb);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            na += __shfl_xor_sync(0xffffffffu, na, o);
            nb += __shfl_xor_sync(0xffffffffu, nb, o);
        }
        na = sqrtf(na);
        nb = sqrtf(nb);
        cosv = dot / (fmaxf(na, 1e-8f) * fmaxf(nb, 1e-8f));
        if (lane == 0) {
            float* st = stats + (static_cast<size_t>(pair) * M + r) * 3;
            st[0] = dot; st[1] = na; st[2] = nb;
        }
    }
    if (lane == 0) sh[warp] = cosv;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < COS_ROWS; ++w) t += sh[w];
        partial[pair * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
cosine_finalize_kernel(const float* __restrict__ partial, int nblk, int M, float weight, float* __restrict__ loss) {
    pdl_trigger();
    pdl_wait();
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < 2 * nblk; i += 256) s += static_cast<double>(partial[i]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = static_cast<float>(-0.5 * static_cast<double>(weight) * sh[0] / M);
}

// d loss / d a_i = upstream * (-weight / (2 M)) * (b / (|a||b|) - cos * a / |a|^2); b is detached (no gradient)
__global__ void __launch_bounds__(COS_ROWS * 32)
cosine_bwd_kernel(const float* __restrict__ a0, const float* __restrict__ b0, const float* __restrict__ a1,
                  const float* __restrict__ b1, int M, int D, const float* __restrict__ stats, float weight,
                  const float* __restrict__ upstream, float* __restrict__ da0, float* __restrict__ da1) {
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.y;
    const float* a = pair ? a1 : a0;
    const float* b = pair ? b1 : b0;
    float* da = pair ? da1 : da0;
    const int r = blockIdx.x * COS_ROWS + warp;
    if (r >= M) return;
    const float* st = stats + (static_cast<size_t>(pair) * M + r) * 3;
    const float dot = st[0], na = fmaxf(st[1], 1e-8f), nb = fmaxf(st[2], 1e-8f);
    const float k = upstream[0] * (-0.5f * weight / M);
    const float inv = 1.f / (na * nb);
    const float cosv = dot * inv;
    const float self = st[1] > 1e-8f ? cosv / (na * na) : 0.f;     // the clamp has zero derivative below eps
    for (int c = lane; c < D; c += 32) {
        const size_t i = static_cast<size_t>(r) * D + c;
        da[i] = k * (b[i] * inv - self * a[i]);
    }
}

}  // namespace vitae

using namespace vitae;

extern "C" int vitae_bn_relu_fwd(const float* h, const float* gamma, const float* beta, float eps, void* act_bf16,
                                 float* mean, float* rstd, float* running_mean, float* running_var, float momentum, int M,
                                 int D, void* stream) {
    VITAE_REQUIRE(h && gamma && beta && act_bf16 && mean && rstd, "bn_relu_fwd: null pointer");
    VITAE_REQUIRE(M > 0 && D > 0 && (!running_mean) == (!running_var), "bn_relu_fwd: bad arguments M=%d D=%d", M, D);
    launch_kernel(bn_relu_fwd_kernel, dim3(ceil_div(D, 32)), dim3(BN_WARPS * 32), 0, as_stream(stream), h, gamma, beta, eps,
                  static_cast<__nv_bfloat16*>(act_bf16), mean, rstd, running_mean, running_var, momentum, M, D);
    VITAE_CHECK_LAUNCH("bn_relu_fwd");
    return 0;
}

extern "C" int vitae_bn_relu_bwd(const void* dact_bf16, const float* h, const float* gamma, const float* beta,
                                 const float* mean, const float* rstd, void* dh_bf16, float* dgamma, float* dbeta,
                                 int accumulate, int M, int D, void* stream) {
    VITAE_REQUIRE(dact_bf16 && h && gamma && beta && mean && rstd && dh_bf16 && dgamma && dbeta, "bn_relu_bwd: null pointer");
    VITAE_REQUIRE(M > 0 && D > 0, "bn_relu_bwd: bad arguments M=%d D=%d", M, D);
    launch_kernel(bn_relu_bwd_kernel, dim3(ceil_div(D, 32)), dim3(BN_WARPS * 32), 0, as_stream(stream),
                  static_cast<const __nv_bfloat16*>(dact_bf16), h, gamma, beta, mean, rstd, static_cast<__nv_bfloat16*>(dh_bf16),
                  dgamma, dbeta, accumulate, M, D);
    VITAE_CHECK_LAUNCH("bn_relu_bwd");
    return 0;
}

extern "C" size_t vitae_cosine_loss_workspace_floats(int M) {
    return static_cast<size_t>(2) * M * 3 + 2 * static_cast<size_t>(ceil_div(M, COS_ROWS));
}

extern "C" int vitae_cosine_loss_fwd(const float* p1, const float* z2, const float* p2, const float* z1, int M, int D,
                                     float weight, float* workspace, float* loss, void* stream) {
    VITAE_REQUIRE(p1 && z2 && p2 && z1 && workspace && loss && M > 0 && D > 0, "cosine_loss_fwd: bad arguments");
    const int nblk = ceil_div(M, COS_ROWS);
    float* stats = workspace;
    float* partial = workspace + static_cast<size_t>(2) * M * 3;
    launch_kernel(cosine_rows_kernel, dim3(nblk, 2), dim3(COS_ROWS * 32), 0, as_stream(stream), p1, z2, p2, z1, M, D, stats, partial);
    VITAE_CHECK_LAUNCH("cosine_rows");
    launch_kernel(cosine_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), static_cast<const float*>(partial), nblk, M, weight, loss);
    VITAE_CHECK_LAUNCH("cosine_finalize");
    return 0;
}

extern "C" int vitae_cosine_loss_bwd(const float* p1, const float* z2, const float* p2, const float* z1, int M, int D,
                                     float weight, const float* workspace, const float* upstream, float* dp1, float* dp2,
                                     void* stream) {
    VITAE_REQUIRE(p1 && z2 && p2 && z1 && workspace && upstream && dp1 && dp2 && M > 0 && D > 0, "cosine_loss_bwd: bad arguments");
    launch_kernel(cosine_bwd_kernel, dim3(ceil_div(M, COS_ROWS), 2), dim3(COS_ROWS * 32), 0, as_stream(stream), p1, z2, p2, z1, M, D,
                  workspace, weight, upstream, dp1, dp2);
    VITAE_CHECK_LAUNCH("cosine_loss_bwd");
    return 0;
}
