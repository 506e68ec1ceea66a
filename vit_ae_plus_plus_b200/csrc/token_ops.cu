// Index / gather kernels of the MAE path: per-sample random masking (argsort of noise), kept-patch im2col for the
// patch embed, token assembly (cls / mask-token rows), gradient row gathers, multi-tensor bf16 cast, fused AdamW.
// All HBM- or latency-bound; coalesced, vectorised where the layout allows.
#include "common.h"
#include <string.h>

#include "ptx.cuh"

namespace vitae {

// ---------------------------------------------------------------------------------------------------------------
// random masking: one CTA per sample, bitonic sort of (noise bits, index) 64-bit keys in shared memory.
// noise in [0,1) => the fp32 bit pattern orders like the value; the index in the low word makes the sort stable
// (== torch.argsort(stable=True); torch's default argsort is only unspecified on exact ties).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
random_masking_kernel(const float* __restrict__ noise, int* __restrict__ ids_shuffle, int* __restrict__ ids_restore,
                      float* __restrict__ mask, int L, int Lpow2, int len_keep) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ unsigned long long keys[];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < Lpow2; i += blockDim.x) {
        unsigned long long k = ~0ull;  // padding sorts last
        if (i < L) {
            float v = noise[static_cast<size_t>(b) * L + i];
            unsigned int u = __float_as_uint(v);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // total order for any finite float
            k = (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned int>(i);
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= Lpow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < Lpow2 / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const unsigned long long a = keys[lo], c = keys[hi];
                if ((a > c) == asc) {
                    keys[lo] = c;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int rank = threadIdx.x; rank < L; rank += blockDim.x) {
        const int idx = static_cast<int>(keys[rank] & 0xffffffffu);
        ids_shuffle[static_cast<size_t>(b) * L + rank] = idx;
        ids_restore[static_cast<size_t>(b) * L + idx] = rank;
        mask[static_cast<size_t>(b) * L + idx] = rank < len_keep ? 0.f : 1.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// im2col of the kept patches: cols[(b*keep + j), (c, pz, py, px)] = vol[b, c, gz*p+pz, gy*p+py, gx*p+px]
// one thread moves 4 consecutive px (float4 load, 8-byte bf16 store); p % 4 == 0.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_patches_kernel(const float* __restrict__ vol, const int* __restrict__ ids_shuffle, __nv_bfloat16* __restrict__ cols,
                      int C, int V, int p, int g, int L, int keep) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x;  // b*keep + j
    const int b = row / keep, j = row % keep;
    const int patch = ids_shuffle[static_cast<size_t>(b) * L + j];
    const int gz = patch / (g * g), gy = (patch / g) % g, gx = patch % g;
    const int p4 = p >> 2;
    const int per_patch = C * p * p * p4;  // float4 chunks
    const size_t vol_b = static_cast<size_t>(b) * C * V * V * V;
    __nv_bfloat16* dst = cols + static_cast<size_t>(row) * C * p * p * p;
    for (int i = threadIdx.x; i < per_patch; i += blockDim.x) {
        const int x4 = i % p4;
        const int py = (i / p4) % p;
        const int pz = (i / (p4 * p)) % p;
        const int c = i / (p4 * p * p);
        const size_t src = vol_b + ((static_cast<size_t>(c) * V + (gz * p + pz)) * V + (gy * p + py)) * V + gx * p + x4 * 4;
        const float4 v = *reinterpret_cast<const float4*>(vol + src);
        uint2 pk;
        pk.x = pack_bf16(v.x, v.y);
        pk.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + static_cast<size_t>(i) * 4) = pk;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Row maps that turn the reference's cat / gather / repeat token shuffling (model/vit_autoenc.py:147,168-170,184-190)
// into GEMM-epilogue row scatters: everything is a function of ids_shuffle.  Ne = keep+1, Nd = L+1.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_row_maps_kernel(const int* __restrict__ ids_shuffle, int B, int L, int keep, int* __restrict__ enc_tok_rows,
                      int* __restrict__ enc_cls_rows, int* __restrict__ pe_pos_rows, int* __restrict__ dec_rows_of_enc,
                      int* __restrict__ dec_pos_rows_of_enc, int* __restrict__ masked_dec_rows,
                      int* __restrict__ masked_pos_rows) {
    pdl_trigger();
    pdl_wait();
    const int Ne = keep + 1, Nd = L + 1, nmask = L - keep;
    const int total = B * L;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int b = i / L, r = i % L;  // r = rank in the shuffled order
        const int patch = ids_shuffle[i];
        if (r < keep) {
            enc_tok_rows[b * keep + r] = b * Ne + 1 + r;
            pe_pos_rows[b * keep + r] = 1 + patch;
            dec_rows_of_enc[b * Ne + 1 + r] = b * Nd + 1 + patch;
            dec_pos_rows_of_enc[b * Ne + 1 + r] = 1 + patch;
        } else {
            masked_dec_rows[b * nmask + (r - keep)] = b * Nd + 1 + patch;
            masked_pos_rows[b * nmask + (r - keep)] = 1 + patch;
        }
        if (r == 0) {
            enc_cls_rows[b] = b * Ne;
            dec_rows_of_enc[b * Ne] = b * Nd;
            dec_pos_rows_of_enc[b * Ne] = 0;
        }
    }
}

// dst[row_idx[i] or i, :] = src0[src0_rows ? src0_rows[i] : 0, :] + src1[src1_rows ? src1_rows[i] : 0, :]
__global__ void __launch_bounds__(128)
fill_rows_kernel(float* __restrict__ dst, const int* __restrict__ row_idx, int D, const float* __restrict__ src0,
                 const int* __restrict__ src0_rows, const float* __restrict__ src1, const int* __restrict__ src1_rows) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x;
    const int r = row_idx ? row_idx[i] : i;
    const float* a = src0 + static_cast<size_t>(src0_rows ? src0_rows[i] : 0) * D;
    const float* c = src1 ? src1 + static_cast<size_t>(src1_rows ? src1_rows[i] : 0) * D : nullptr;
    for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
        float4 v = *reinterpret_cast<const float4*>(a + d);
        if (c) {
            const float4 w = *reinterpret_cast<const float4*>(c + d);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        *reinterpret_cast<float4*>(dst + static_cast<size_t>(r) * D + d) = v;
    }
}

__global__ void __launch_bounds__(128)
gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ row_idx, int D, __nv_bfloat16* __restrict__ dst16,
                   float* __restrict__ dst32) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x;
    const float* a = src + static_cast<size_t>(row_idx ? row_idx[i] : i) * D;
    for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(a + d);
        if (dst16) {
            uint2 pk;
            pk.x = pack_bf16(v.x, v.y);
            pk.y = pack_bf16(v.z, v.w);
            *reinterpret_cast<uint2*>(dst16 + static_cast<size_t>(i) * D + d) = pk;
        }
        if (dst32) *reinterpret_cast<float4*>(dst32 + static_cast<size_t>(i) * D + d) = v;
    }
}

// out[d] (+)= sum_i src[row_idx[i], d]; block = 32 columns x 32 row lanes: lane ty sums rows ty, ty+32, ... (4 loads in
// flight), then the 32 lane sums are added in fixed order (deterministic)
__global__ void __launch_bounds__(1024)
sum_rows_kernel(const float* __restrict__ src, const int* __restrict__ row_idx, int nrows, int D, float* __restrict__ out,
                int accumulate) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + tx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (d < D) {
        int i = ty;
        for (; i + 96 < nrows; i += 128) {
            const int r0 = row_idx ? row_idx[i] : i, r1 = row_idx ? row_idx[i + 32] : i + 32;
            const int r2 = row_idx ? row_idx[i + 64] : i + 64, r3 = row_idx ? row_idx[i + 96] : i + 96;
            s0 += src[static_cast<size_t>(r0) * D + d];
            s1 += src[static_cast<size_t>(r1) * D + d];
            s2 += src[static_cast<size_t>(r2) * D + d];
            s3 += src[static_cast<size_t>(r3) * D + d];
        }
        for (; i < nrows; i += 32) s0 += src[static_cast<size_t>(row_idx ? row_idx[i] : i) * D + d];
    }
    red[ty][tx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ty == 0 && d < D) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) s += red[k][tx];
        out[d] = accumulate ? out[d] + s : s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// L2 prefetch: the step touches ~1 GB of weights and saved activations once per pass, each GEMM / attention kernel is a
// few microseconds long and starts with a cold first load (~1 us from HBM instead of ~0.3 us from L2).  The 126 MB L2
// holds a transformer block's weights and activations several times over, so a tiny kernel on the side lane pulls the
// NEXT block's regions into L2 (cp.async.bulk.prefetch.L2, no registers / shared memory involved) while the current block
// computes.
// ---------------------------------------------------------------------------------------------------------------
struct PrefetchTable {
    unsigned long long ptr[12];
    unsigned long long bytes[12];
    int n;
};
constexpr unsigned int PREFETCH_CHUNK = 16384;

__global__ void __launch_bounds__(32)
prefetch_l2_kernel(const PrefetchTable t) {
    pdl_trigger();
    // no pdl_wait: only brings lines into L2 (the point of coherence), it neither reads values nor writes anything
    if (threadIdx.x != 0) return;
    unsigned int c = blockIdx.x;
    for (int r = 0; r < t.n; ++r) {
        const unsigned long long nchunks = (t.bytes[r] + PREFETCH_CHUNK - 1) / PREFETCH_CHUNK;
        for (unsigned long long i = c; i < nchunks; i += gridDim.x) {
            const unsigned long long off = i * PREFETCH_CHUNK;
            const unsigned long long rem = t.bytes[r] - off;
            const unsigned int sz = static_cast<unsigned int>(rem < PREFETCH_CHUNK ? (rem & ~15ull) : PREFETCH_CHUNK);
            if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(t.ptr[r] + off), "r"(sz) : "memory");
        }
        c = (c + static_cast<unsigned int>(nchunks % gridDim.x)) % gridDim.x;   // rotate so short regions spread over CTAs
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 master parameters -> flat bf16 shadow, all tensors in one launch
// ---------------------------------------------------------------------------------------------------------------
struct CastRecord {
    unsigned long long src;
    unsigned long long dst_off;
    unsigned long long numel;
};
constexpr int CAST_CHUNK = 4096;  // elements per work item

__global__ void __launch_bounds__(256)
cast_params_kernel(const CastRecord* __restrict__ table, int ntensors, __nv_bfloat16* __restrict__ dst) {
    pdl_trigger();
    pdl_wait();
    // blockIdx.y = tensor, blockIdx.x strides over that tensor's chunks
    const CastRecord rec = table[blockIdx.y];
    const float* src = reinterpret_cast<const float*>(rec.src);
    __nv_bfloat16* out = dst + rec.dst_off;
    const unsigned long long n = rec.numel;
    const bool vec_ok = ((rec.src & 15ull) == 0) && ((rec.dst_off & 7ull) == 0);
    for (unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * CAST_CHUNK; base < n;
         base += static_cast<unsigned long long>(gridDim.x) * CAST_CHUNK) {
        const unsigned long long end = min(n, base + CAST_CHUNK);
        if (vec_ok) {
            for (unsigned long long i = base + threadIdx.x * 8ull; i < end; i += 256ull * 8ull) {
                if (i + 8 <= end) {
                    const float4 a = *reinterpret_cast<const float4*>(src + i);
                    const float4 b = *reinterpret_cast<const float4*>(src + i + 4);
                    uint4 pk;
                    pk.x = pack_bf16(a.x, a.y); pk.y = pack_bf16(a.z, a.w);
                    pk.z = pack_bf16(b.x, b.y); pk.w = pack_bf16(b.z, b.w);
                    *reinterpret_cast<uint4*>(out + i) = pk;
                } else {
                    for (unsigned long long k = i; k < end; ++k) out[k] = __float2bfloat16(src[k]);
                }
            }
        } else {
            for (unsigned long long i = base + threadIdx.x; i < end; i += 256) out[i] = __float2bfloat16(src[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused AdamW (decoupled weight decay, torch.optim.AdamW update order) over a flat buffer
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ p16, long long n, float lr, float b1, float b2, float eps, float wd, float bc1,
             float bc2, const float* __restrict__ inv_scale, const float* __restrict__ found_inf) {
    pdl_trigger();
    pdl_wait();
    if (found_inf && *found_inf != 0.f) return;
    const float is = inv_scale ? *inv_scale : 1.f;
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = (blockIdx.x * 256ll + threadIdx.x) * 4; i < n; i += gridDim.x * 256ll * 4) {
        if (i + 4 <= n) {
            float4 pp = *reinterpret_cast<float4*>(p + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gr = ga[k] * is;
                pa[k] *= (1.f - lr * wd);
                ma[k] = b1 * ma[k] + (1.f - b1) * gr;
                va[k] = b2 * va[k] + (1.f - b2) * gr * gr;
                pa[k] -= step_size * ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps);
            }
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
            if (p16) {
                uint2 pk;
                pk.x = pack_bf16(pp.x, pp.y);
                pk.y = pack_bf16(pp.z, pp.w);
                *reinterpret_cast<uint2*>(p16 + i) = pk;
            }
        } else {
            for (long long k = i; k < n; ++k) {
                const float gr = g[k] * is;
                float pv = p[k] * (1.f - lr * wd);
                const float mv = b1 * m[k] + (1.f - b1) * gr;
                const float vv2 = b2 * v[k] + (1.f - b2) * gr * gr;
                pv -= step_size * mv / (sqrtf(vv2) * inv_sqrt_bc2 + eps);
                p[k] = pv; m[k] = mv; v[k] = vv2;
                if (p16) p16[k] = __float2bfloat16(pv);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Optimizer step over the flat buffers: GradScaler.unscale_ + found_inf + get_grad_norm_
// (utils/misc.py:259-292) as one reduction pass, then GradScaler.update + AdamW (both param groups) in one pass.
// Control block ctl (device fp32[8]): [0] loss scale, [1] growth tracker, [2] found_inf, [3] 1/scale used by this step,
// [4] unscaled global gradient L2 norm, [5] number of optimizer steps taken (skipped steps do not count).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sh[8];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
        const float4 v = __ldcs(g4 + i);
        a0 += v.x * v.x; a1 += v.y * v.y; a2 += v.z * v.z; a3 += v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long k = n4 << 2; k < n; ++k) a0 += g[k] * g[k];
    float v = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w];
        partials[blockIdx.x] = t;
    }
}

// sum of n fp32 partials in double -> out[0] (fp32): the sharded step publishes ONE value per rank to its peers
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ in, int n, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += static_cast<double>(in[i]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = static_cast<float>(sh[0]);
}

// nsrc > 1 (data parallel, sharded step): nblk partials from each of the ranks' buffers (peer memory), summed in a fixed
// order so that every rank arrives at the same control block.
struct PartialSources {
    const float* src[8];
};
__global__ void __launch_bounds__(256)
optim_finalize_kernel(const PartialSources ps, int nsrc, int nblk, float* __restrict__ ctl, float growth_factor,
                      float backoff_factor, int growth_interval, int use_scaler) {
    pdl_trigger();
    pdl_wait();
    __shared__ double sh[256];
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (r < nsrc)
            for (int i = threadIdx.x; i < nblk; i += 256) s += static_cast<double>(__ldcv(ps.src[r] + i));
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double tot = sh[0];
        const float scale = use_scaler ? ctl[0] : 1.f;
        const float inv = 1.f / scale;
        const bool bad = !isfinite(tot) || !isfinite(static_cast<float>(tot));
        ctl[2] = (use_scaler && bad) ? 1.f : 0.f;
        ctl[3] = inv;
        ctl[4] = static_cast<float>(sqrt(tot)) * inv;
        if (use_scaler) {                        // torch.amp.GradScaler.update semantics
            if (bad) {
                ctl[0] = scale * backoff_factor;
                ctl[1] = 0.f;
            } else {
                const float ok = ctl[1] + 1.f;
                if (ok >= static_cast<float>(growth_interval)) {
                    const float grown = scale * growth_factor;
                    if (isfinite(grown)) ctl[0] = grown;
                    ctl[1] = 0.f;
                } else {
                    ctl[1] = ok;
                }
            }
        }
        if (!(use_scaler && bad)) ctl[5] += 1.f;
    }
}

// hyper (by value): [ngroups][8] = {lr, beta1, beta2, eps, weight_decay, -, -, -}; group_of_chunk[i >> 6] selects the
// parameter group of element i (tensors start on 64-element boundaries).  torch.optim.AdamW update order.
struct AdamHyper {
    float h[8][8];
};
struct AdamCoef {
    float decay, b1, b2, eps, step_size, isb2;
};
__device__ __forceinline__ void adam_prologue(float (*hs)[8], const AdamHyper& hyper, int ngroups, const float* ctl) {
    if (threadIdx.x < ngroups) {
        const float* h = hyper.h[threadIdx.x];
        const double step = static_cast<double>(ctl[5]);
        const double bc1 = 1.0 - pow(static_cast<double>(h[1]), step);
        const double bc2 = 1.0 - pow(static_cast<double>(h[2]), step);
        hs[threadIdx.x][0] = 1.f - h[0] * h[4];
        hs[threadIdx.x][1] = h[1];
        hs[threadIdx.x][2] = h[2];
        hs[threadIdx.x][3] = h[3];
        hs[threadIdx.x][4] = static_cast<float>(static_cast<double>(h[0]) / bc1);
        hs[threadIdx.x][5] = static_cast<float>(1.0 / sqrt(bc2));
    }
    __syncthreads();
}
// torch.optim.AdamW update order, four elements
__device__ __forceinline__ void adam_update4(float4& pp, const float4& gg, float4& mm, float4& vv, const float* c, float is) {
    const float decay = c[0], b1 = c[1], b2 = c[2], eps = c[3], step_size = c[4], isb2 = c[5];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float gr = ga[k] * is;
        pa[k] *= decay;
        ma[k] = b1 * ma[k] + (1.f - b1) * gr;
        va[k] = b2 * va[k] + (1.f - b2) * gr * gr;
        pa[k] -= step_size * ma[k] / (sqrtf(va[k]) * isb2 + eps);
    }
}

__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  __nv_bfloat16* __restrict__ p16, long long n, const unsigned char* __restrict__ group_of_chunk,
                  const AdamHyper hyper, int ngroups, const float* __restrict__ ctl) {
    pdl_trigger();
    pdl_wait();
    __shared__ float hs[8][8];   // per group: lr*wd factor, b1, b2, eps, step_size, inv_sqrt_bc2
    if (ctl[2] != 0.f) return;   // inf / nan gradients: GradScaler skips the step
    adam_prologue(hs, hyper, ngroups, ctl);
    const float is = ctl[3];
    const long long n4 = n >> 2;   // n is a multiple of 64
    for (long long i4 = blockIdx.x * 256ll + threadIdx.x; i4 < n4; i4 += gridDim.x * 256ll) {
        const long long i = i4 << 2;
        const int grp = group_of_chunk[i >> 6];
        if (grp >= ngroups) continue;           // padding / frozen chunk
        float4 pp = *reinterpret_cast<float4*>(p + i);
        const float4 gg = __ldcs(reinterpret_cast<const float4*>(g + i));
        float4 mm = *reinterpret_cast<float4*>(m + i);
        float4 vv = *reinterpret_cast<float4*>(v + i);
        adam_update4(pp, gg, mm, vv, hs[grp], is);
        *reinterpret_cast<float4*>(p + i) = pp;
        *reinterpret_cast<float4*>(m + i) = mm;
        *reinterpret_cast<float4*>(v + i) = vv;
        if (p16) {
            uint2 pk;
            pk.x = pack_bf16(pp.x, pp.y);
            pk.y = pack_bf16(pp.z, pp.w);
            *reinterpret_cast<uint2*>(p16 + i) = pk;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Data parallel, sharded optimizer step over NVLink peer memory (dp.ShardedStep).  Every rank's flat gradient / master /
// shadow buffers are symmetric allocations mapped into every other rank's address space; DpPeers carries the ranks' base
// addresses of one such buffer (index = rank).  Ownership is block-cyclic and fixed for the life of the buffers: granule q
// (2^gshift elements) belongs to rank q % world, so any slice [lo, hi) of the flat index space -- the backward's stage
// slices -- splits evenly and an element never changes its owner.
// ---------------------------------------------------------------------------------------------------------------
struct DpPeers {
    void* base[8];
};
struct DpRange {          // this rank's part of the slice [lo, hi): granules q0, q0 + world, ... (nq of them)
    long long lo, hi, q0, nq;
    int gshift, world;
};
// element index of vector `t` (VEC elements each) of this rank's part, or -1 outside [lo, hi)
template <int VEC>
__device__ __forceinline__ long long dp_elem(const DpRange& rg, long long t) {
    const int vshift = rg.gshift - (VEC == 8 ? 3 : 2);
    const long long q = rg.q0 + (t >> vshift) * rg.world;
    const long long e = (q << rg.gshift) + ((t & ((1ll << vshift) - 1)) * VEC);
    return (e >= rg.lo && e < rg.hi) ? e : -1;
}

// this rank's part of the MEAN gradient: own copy + the peers' copies, pulled over NVLink with plain loads in rank order
// (every element is reduced by exactly one rank, so the replicas cannot diverge); the result replaces this rank's copy; its
// per-block sums of squares go to partials (vitae_optim_finalize_peers' input).
template <int W>
__global__ void __launch_bounds__(256)
dp_reduce_shard_kernel(const DpPeers peers, float* __restrict__ own, const DpRange rg, float inv_world,
                       float* __restrict__ partials) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sh[8];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const long long nt = rg.nq << (rg.gshift - 2);
    constexpr int U = W <= 2 ? 4 : 2;      // elements in flight per thread: U * W 16-byte loads
    const long long stride = gridDim.x * 256ll;
    for (long long t0 = blockIdx.x * 256ll + threadIdx.x; t0 < nt; t0 += stride * U) {
        float4 x[U][W];
        long long e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long t = t0 + u * stride;
            e[u] = t < nt ? dp_elem<4>(rg, t) : -1;
            if (e[u] >= 0) {
#pragma unroll
                for (int r = 0; r < W; ++r)
                    x[u][r] = __ldcs(reinterpret_cast<const float4*>(static_cast<const float*>(peers.base[r]) + e[u]));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e[u] >= 0) {
                float4 s = x[u][0];
#pragma unroll
                for (int r = 1; r < W; ++r) { s.x += x[u][r].x; s.y += x[u][r].y; s.z += x[u][r].z; s.w += x[u][r].w; }
                s.x *= inv_world; s.y *= inv_world; s.z *= inv_world; s.w *= inv_world;
                *reinterpret_cast<float4*>(own + e[u]) = s;
                a0 += s.x * s.x; a1 += s.y * s.y; a2 += s.z * s.z; a3 += s.w * s.w;
            }
        }
    }
    float t = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += sh[w];
        partials[blockIdx.x] = r;
    }
}

// AdamW over this rank's part of [lo, hi) (moments local to the owner), then the all-gather in the same kernel: the new
// bf16 shadow goes to every rank; the fp32 master goes to the peers only for the chunks flagged in f32_chunk (tensors the
// kernels read in fp32 on every rank: biases, LayerNorm affine, tokens) -- the peers' masters of the big GEMM weights go
// stale and are pulled from their owner on demand (dp.ShardedStep.sync_master).  Eight elements per thread: 16-byte
// NVLink stores of the shadow.
template <int W>
__global__ void __launch_bounds__(256)
adamw_shard_kernel(const DpPeers p32s, const DpPeers p16s, int rank, float* __restrict__ p, const DpRange rg,
                   const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   const unsigned char* __restrict__ group_of_chunk, const unsigned char* __restrict__ f32_chunk,
                   const AdamHyper hyper, int ngroups, const float* __restrict__ ctl) {
    pdl_trigger();
    pdl_wait();
    __shared__ float hs[8][8];
    if (ctl[2] != 0.f) return;   // skipped step: nothing changes on any rank
    adam_prologue(hs, hyper, ngroups, ctl);
    const float is = ctl[3];
    const long long nt = rg.nq << (rg.gshift - 3);
    for (long long t = blockIdx.x * 256ll + threadIdx.x; t < nt; t += gridDim.x * 256ll) {
        const long long i = dp_elem<8>(rg, t);
        if (i < 0) continue;
        const long long chunk = i >> 6;
        const int grp = group_of_chunk[chunk];
        if (grp >= ngroups) continue;
        float4 pa = *reinterpret_cast<float4*>(p + i), pb = *reinterpret_cast<float4*>(p + i + 4);
        const float4 ga = __ldcs(reinterpret_cast<const float4*>(g + i)), gb = __ldcs(reinterpret_cast<const float4*>(g + i + 4));
        float4 ma = *reinterpret_cast<float4*>(m + i), mb = *reinterpret_cast<float4*>(m + i + 4);
        float4 va = *reinterpret_cast<float4*>(v + i), vb = *reinterpret_cast<float4*>(v + i + 4);
        adam_update4(pa, ga, ma, va, hs[grp], is);
        adam_update4(pb, gb, mb, vb, hs[grp], is);
        *reinterpret_cast<float4*>(m + i) = ma; *reinterpret_cast<float4*>(m + i + 4) = mb;
        *reinterpret_cast<float4*>(v + i) = va; *reinterpret_cast<float4*>(v + i + 4) = vb;
        uint4 pk;
        pk.x = pack_bf16(pa.x, pa.y); pk.y = pack_bf16(pa.z, pa.w);
        pk.z = pack_bf16(pb.x, pb.y); pk.w = pack_bf16(pb.z, pb.w);
        const bool wide = f32_chunk[chunk] != 0;
#pragma unroll
        for (int r = 0; r < W; ++r) {
            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p16s.base[r]) + i) = pk;
            if (r == rank || wide) {
                float* q = static_cast<float*>(p32s.base[r]) + i;
                *reinterpret_cast<float4*>(q) = pa;
                *reinterpret_cast<float4*>(q + 4) = pb;
            }
        }
    }
}

}  // namespace vitae

using namespace vitae;

extern "C" int vitae_random_masking(const float* noise, int32_t* ids_shuffle, int32_t* ids_restore, float* mask, int B,
                                    int L, int len_keep, void* stream) {
    VITAE_REQUIRE(noise && ids_shuffle && ids_restore && mask, "random_masking: null pointer");
    VITAE_REQUIRE(B > 0 && L > 0 && L <= 8192 && len_keep >= 0 && len_keep <= L, "random_masking: bad sizes B=%d L=%d keep=%d", B, L, len_keep);
    int lp = 2;
    while (lp < L) lp <<= 1;
    const int threads = std::min(1024, std::max(32, lp / 2));
    const size_t key_bytes = lp * sizeof(unsigned long long);
    if (key_bytes > 48 * 1024) {      // L > 4096 (keys padded to 8192): past the default 48 KB dynamic shared-memory limit
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(random_masking_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e != cudaSuccess) return set_error(-3, "random_masking: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            attr_set = true;
        }
    }
    launch_kernel(random_masking_kernel, dim3(B), dim3(threads), key_bytes, as_stream(stream), noise, ids_shuffle, ids_restore, mask, L, lp, len_keep);
    VITAE_CHECK_LAUNCH("random_masking");
    return 0;
}

extern "C" int vitae_build_row_maps(const int32_t* ids_shuffle, int B, int L, int keep, int32_t* enc_tok_rows,
                                    int32_t* enc_cls_rows, int32_t* pe_pos_rows, int32_t* dec_rows_of_enc,
                                    int32_t* dec_pos_rows_of_enc, int32_t* masked_dec_rows, int32_t* masked_pos_rows,
                                    void* stream) {
    VITAE_REQUIRE(ids_shuffle && enc_tok_rows && enc_cls_rows && pe_pos_rows && dec_rows_of_enc && dec_pos_rows_of_enc &&
                      masked_dec_rows && masked_pos_rows, "build_row_maps: null pointer");
    VITAE_REQUIRE(B > 0 && L > 0 && keep > 0 && keep <= L, "build_row_maps: bad sizes B=%d L=%d keep=%d", B, L, keep);
    const int blocks = std::min(ceil_div(B * L, 256), 148);
    launch_kernel(build_row_maps_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), ids_shuffle, B, L, keep, enc_tok_rows, enc_cls_rows, pe_pos_rows,
                                                                 dec_rows_of_enc, dec_pos_rows_of_enc, masked_dec_rows, masked_pos_rows);
    VITAE_CHECK_LAUNCH("build_row_maps");
    return 0;
}

extern "C" int vitae_im2col_patches(const float* vol, const int32_t* ids_shuffle, void* cols_bf16, int B, int C, int V,
                                    int p, int L, int keep, void* stream) {
    VITAE_REQUIRE(vol && ids_shuffle && cols_bf16, "im2col: null pointer");
    VITAE_REQUIRE(p % 4 == 0 && V % p == 0 && keep > 0, "im2col: need p %% 4 == 0 and V %% p == 0 (V=%d p=%d)", V, p);
    const int g = V / p;
    VITAE_REQUIRE(g * g * g == L, "im2col: L=%d does not match (V/p)^3", L);
    launch_kernel(im2col_patches_kernel, dim3(B * keep), dim3(256), 0, as_stream(stream), vol, ids_shuffle, static_cast<__nv_bfloat16*>(cols_bf16), C, V, p, g, L, keep);
    VITAE_CHECK_LAUNCH("im2col_patches");
    return 0;
}

extern "C" int vitae_fill_rows(float* dst, const int32_t* row_idx, int nrows, int D, const float* src0,
                               const int32_t* src0_rows, const float* src1, const int32_t* src1_rows, void* stream) {
    VITAE_REQUIRE(dst && src0 && nrows >= 0 && D % 4 == 0, "fill_rows: bad arguments");
    if (nrows == 0) return 0;
    launch_kernel(fill_rows_kernel, dim3(nrows), dim3(128), 0, as_stream(stream), dst, row_idx, D, src0, src0_rows, src1, src1_rows);
    VITAE_CHECK_LAUNCH("fill_rows");
    return 0;
}

extern "C" int vitae_gather_rows(const float* src, const int32_t* row_idx, int nrows, int D, void* dst_bf16,
                                 float* dst_f32, void* stream) {
    VITAE_REQUIRE(src && (dst_bf16 || dst_f32) && nrows >= 0 && D % 4 == 0, "gather_rows: bad arguments");
    if (nrows == 0) return 0;
    launch_kernel(gather_rows_kernel, dim3(nrows), dim3(128), 0, as_stream(stream), src, row_idx, D, static_cast<__nv_bfloat16*>(dst_bf16), dst_f32);
    VITAE_CHECK_LAUNCH("gather_rows");
    return 0;
}

extern "C" int vitae_sum_rows(const float* src, const int32_t* row_idx, int nrows, int D, float* out, int accumulate,
                              void* stream) {
    VITAE_REQUIRE(src && out && nrows >= 0 && D > 0, "sum_rows: bad arguments");
    launch_kernel(sum_rows_kernel, dim3(ceil_div(D, 32)), dim3(1024), 0, as_stream(stream), src, row_idx, nrows, D, out, accumulate);
    VITAE_CHECK_LAUNCH("sum_rows");
    return 0;
}

extern "C" int vitae_prefetch_l2(const void* const* ptrs, const size_t* bytes, int n, void* stream) {
    VITAE_REQUIRE(ptrs && bytes && n > 0 && n <= 12, "prefetch_l2: 1..12 regions");
    PrefetchTable t;
    memset(&t, 0, sizeof(t));
    unsigned long long total = 0;
    for (int i = 0; i < n; ++i) {
        // 16-byte granularity: shrink the region to its aligned interior
        unsigned long long a = reinterpret_cast<unsigned long long>(ptrs[i]);
        unsigned long long e = a + bytes[i];
        a = (a + 15ull) & ~15ull;
        e &= ~15ull;
        t.ptr[i] = a;
        t.bytes[i] = e > a ? e - a : 0;
        total += t.bytes[i];
    }
    t.n = n;
    if (total == 0) return 0;
    const int blocks = static_cast<int>(std::min<unsigned long long>(ceil_div<unsigned long long>(total, PREFETCH_CHUNK), 148));
    launch_kernel(prefetch_l2_kernel, dim3(blocks), dim3(32), 0, as_stream(stream), t);
    VITAE_CHECK_LAUNCH("prefetch_l2");
    return 0;
}

extern "C" int vitae_cast_params_bf16(const void* table, int ntensors, void* dst_bf16, long long total_elems, void* stream) {
    VITAE_REQUIRE(table && dst_bf16 && ntensors > 0, "cast_params: bad arguments");
    // enough x-blocks that the largest tensors are spread over the machine; small tensors exit after one chunk
    const long long avg = total_elems / ntensors + 1;
    int bx = static_cast<int>(std::min<long long>(64, std::max<long long>(1, (4 * avg) / CAST_CHUNK)));
    dim3 grid(bx, ntensors);
    launch_kernel(cast_params_kernel, dim3(grid), dim3(256), 0, as_stream(stream), static_cast<const CastRecord*>(table), ntensors, static_cast<__nv_bfloat16*>(dst_bf16));
    VITAE_CHECK_LAUNCH("cast_params");
    return 0;
}

extern "C" int vitae_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                                long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                float bias_corr1, float bias_corr2, const float* inv_scale, const float* found_inf,
                                void* stream) {
    VITAE_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0, "adamw: bad arguments");
    const int blocks = static_cast<int>(std::min<long long>(ceil_div<long long>(n, 1024), 148 * 8));
    launch_kernel(adamw_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(param_bf16), n, lr,
                                                        beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, inv_scale, found_inf);
    VITAE_CHECK_LAUNCH("adamw");
    return 0;
}

constexpr int OPT_NORM_BLOCKS = 148 * 4;

// fp32 <-> bf16 copies of a flat gradient slice (the data-parallel gradient exchange moves bf16, dp.py)
template <bool TO_BF16>
__global__ void __launch_bounds__(256) cast_flat_kernel(const void* __restrict__ src, void* __restrict__ dst, long long n) {
    pdl_trigger();
    pdl_wait();
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
        if (TO_BF16) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + i);
            uint2 pk;
            pk.x = pack_bf16(v.x, v.y);
            pk.y = pack_bf16(v.z, v.w);
            reinterpret_cast<uint2*>(dst)[i] = pk;
        } else {
            const uint2 pk = __ldcs(reinterpret_cast<const uint2*>(src) + i);
            float4 v;
            v.x = __uint_as_float(pk.x << 16); v.y = __uint_as_float(pk.x & 0xffff0000u);
            v.z = __uint_as_float(pk.y << 16); v.w = __uint_as_float(pk.y & 0xffff0000u);
            reinterpret_cast<float4*>(dst)[i] = v;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long k = n4 << 2; k < n; ++k) {
            if (TO_BF16) static_cast<__nv_bfloat16*>(dst)[k] = __float2bfloat16(static_cast<const float*>(src)[k]);
            else static_cast<float*>(dst)[k] = __bfloat162float(static_cast<const __nv_bfloat16*>(src)[k]);
        }
}

extern "C" int vitae_cast_f32_to_bf16(const float* src, void* dst_bf16, long long n, int max_blocks, void* stream) {
    VITAE_REQUIRE(src && dst_bf16 && n > 0, "cast_f32_to_bf16: bad arguments");
    VITAE_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_bf16) & 7) == 0, "cast_f32_to_bf16: alignment");
    const int cap = max_blocks > 0 ? max_blocks : 148 * 4;
    const int blocks = static_cast<int>(std::min<long long>(ceil_div<long long>(n, 1024), cap));
    launch_kernel(cast_flat_kernel<true>, dim3(blocks), dim3(256), 0, as_stream(stream), static_cast<const void*>(src), dst_bf16, n);
    VITAE_CHECK_LAUNCH("cast_f32_to_bf16");
    return 0;
}

extern "C" int vitae_cast_bf16_to_f32(const void* src_bf16, float* dst, long long n, int max_blocks, void* stream) {
    VITAE_REQUIRE(src_bf16 && dst && n > 0, "cast_bf16_to_f32: bad arguments");
    VITAE_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src_bf16) & 7) == 0, "cast_bf16_to_f32: alignment");
    const int cap = max_blocks > 0 ? max_blocks : 148 * 4;
    const int blocks = static_cast<int>(std::min<long long>(ceil_div<long long>(n, 1024), cap));
    launch_kernel(cast_flat_kernel<false>, dim3(blocks), dim3(256), 0, as_stream(stream), src_bf16, static_cast<void*>(dst), n);
    VITAE_CHECK_LAUNCH("cast_bf16_to_f32");
    return 0;
}

extern "C" size_t vitae_optim_workspace_bytes(void) { return 2 * OPT_NORM_BLOCKS * sizeof(float); }

extern "C" int vitae_optim_prepare(const float* grad, long long n, const float* grad2, long long n2, float* ctl,
                                   float* workspace, float growth_factor, float backoff_factor, int growth_interval,
                                   int use_scaler, void* stream) {
    VITAE_REQUIRE(grad && ctl && workspace && n > 0, "optim_prepare: bad arguments");
    VITAE_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad2) & 15) == 0,
                  "optim_prepare: grad buffers must be 16-byte aligned");
    int blocks = static_cast<int>(std::min<long long>(ceil_div<long long>(n, 1024), OPT_NORM_BLOCKS));
    launch_kernel(grad_sqnorm_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), grad, n, workspace);
    VITAE_CHECK_LAUNCH("grad_sqnorm");
    if (grad2 && n2 > 0) {   // a second gradient region (parameters outside the main flat buffer) joins the same norm
        const int blocks2 = static_cast<int>(std::min<long long>(ceil_div<long long>(n2, 1024), OPT_NORM_BLOCKS));
        launch_kernel(grad_sqnorm_kernel, dim3(blocks2), dim3(256), 0, as_stream(stream), grad2, n2, workspace + blocks);
        VITAE_CHECK_LAUNCH("grad_sqnorm");
        blocks += blocks2;
    }
    PartialSources ps{};
    ps.src[0] = workspace;
    launch_kernel(optim_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), ps, 1, blocks, ctl, growth_factor, backoff_factor, growth_interval, use_scaler);
    VITAE_CHECK_LAUNCH("optim_finalize");
    return 0;
}

// The two halves of vitae_optim_prepare on their own: squared-norm partials of a gradient SLICE (the backward computes them
// per stage on a side lane, underneath the later stages) and the control-block update over any number of partials.
extern "C" int vitae_grad_sqnorm_blocks(long long n, int max_blocks) {
    const int cap = max_blocks > 0 ? std::min(max_blocks, OPT_NORM_BLOCKS) : OPT_NORM_BLOCKS;
    return static_cast<int>(std::min<long long>(ceil_div<long long>(n, 1024), cap));
}

extern "C" int vitae_grad_sqnorm(const float* grad, long long n, float* partials, int max_blocks, void* stream) {
    VITAE_REQUIRE(grad && partials && n > 0, "grad_sqnorm: bad arguments");
    VITAE_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad_sqnorm: grad must be 16-byte aligned");
    const int blocks = vitae_grad_sqnorm_blocks(n, max_blocks);
    launch_kernel(grad_sqnorm_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), grad, n, partials);
    VITAE_CHECK_LAUNCH("grad_sqnorm");
    return 0;
}

extern "C" int vitae_optim_finalize(const float* partials, int npartials, float* ctl, float growth_factor, float backoff_factor,
                                    int growth_interval, int use_scaler, void* stream) {
    VITAE_REQUIRE(partials && ctl && npartials > 0, "optim_finalize: bad arguments");
    PartialSources ps{};
    ps.src[0] = partials;
    launch_kernel(optim_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), ps, 1, npartials, ctl, growth_factor,
                  backoff_factor, growth_interval, use_scaler);
    VITAE_CHECK_LAUNCH("optim_finalize");
    return 0;
}

extern "C" int vitae_sum_partials(const float* partials, int n, float* out, void* stream) {
    VITAE_REQUIRE(partials && out && n > 0, "sum_partials: bad arguments");
    launch_kernel(sum_partials_kernel, dim3(1), dim3(256), 0, as_stream(stream), partials, n, out);
    VITAE_CHECK_LAUNCH("sum_partials");
    return 0;
}

extern "C" int vitae_optim_finalize_peers(void* const* partial_ptrs, int world, int npartials, float* ctl, float growth_factor,
                                          float backoff_factor, int growth_interval, int use_scaler, void* stream) {
    VITAE_REQUIRE(partial_ptrs && ctl && npartials > 0 && world >= 1 && world <= 8, "optim_finalize_peers: bad arguments");
    PartialSources ps{};
    for (int r = 0; r < world; ++r) {
        VITAE_REQUIRE(partial_ptrs[r], "optim_finalize_peers: null pointer for rank %d", r);
        ps.src[r] = static_cast<const float*>(partial_ptrs[r]);
    }
    launch_kernel(optim_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), ps, world, npartials, ctl, growth_factor,
                  backoff_factor, growth_interval, use_scaler);
    VITAE_CHECK_LAUNCH("optim_finalize_peers");
    return 0;
}

extern "C" int vitae_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                                long long n, const unsigned char* group_of_chunk, const float* hyper, int ngroups,
                                const float* ctl, int max_blocks, void* stream) {
    VITAE_REQUIRE(param && grad && exp_avg && exp_avg_sq && group_of_chunk && hyper && ctl, "adamw_flat: null pointer");
    VITAE_REQUIRE(n > 0 && n % 64 == 0 && ngroups > 0 && ngroups <= 8, "adamw_flat: n=%lld must be a multiple of 64, 1..8 groups", n);
    const int cap = max_blocks > 0 ? max_blocks : 148 * 8;
    const int blocks = static_cast<int>(std::min<long long>(ceil_div<long long>(n, 1024), cap));
    AdamHyper hy;
    memset(&hy, 0, sizeof(hy));
    memcpy(hy.h, hyper, sizeof(float) * 8 * ngroups);   // host pointer, read during this call
    launch_kernel(adamw_flat_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(param_bf16), n,
                                                             group_of_chunk, hy, ngroups, ctl);
    VITAE_CHECK_LAUNCH("adamw_flat");
    return 0;
}

template <int W>
static void launch_dp_reduce(int blocks, cudaStream_t st, const DpPeers& peers, int rank, const DpRange& rg, float inv_world,
                             float* partials) {
    launch_kernel(dp_reduce_shard_kernel<W>, dim3(blocks), dim3(256), 0, st, peers, static_cast<float*>(peers.base[rank]), rg,
                  inv_world, partials);
}
template <int W>
static void launch_adamw_shard(int blocks, cudaStream_t st, const DpPeers& p32s, const DpPeers& p16s, int rank, const DpRange& rg,
                               const float* g, float* m, float* v, const unsigned char* goc, const unsigned char* f32c,
                               const AdamHyper& hy, int ngroups, const float* ctl) {
    launch_kernel(adamw_shard_kernel<W>, dim3(blocks), dim3(256), 0, st, p32s, p16s, rank, static_cast<float*>(p32s.base[rank]), rg,
                  g, m, v, goc, f32c, hy, ngroups, ctl);
}
#define DP_DISPATCH_WORLD(world, CALL)                                                            \
    switch (world) {                                                                                \
        case 2: CALL(2); break;                                                                     \
        case 3: CALL(3); break;                                                                     \
        case 4: CALL(4); break;                                                                     \
        case 5: CALL(5); break;                                                                     \
        case 6: CALL(6); break;                                                                     \
        case 7: CALL(7); break;                                                                     \
        case 8: CALL(8); break;                                                                     \
        default: return set_error(-1, "data-parallel shard kernels: world size %d not in 2..8", world); \
    }

static int fill_peers(DpPeers& d, void* const* ptrs, int world, size_t align, const char* what) {
    memset(&d, 0, sizeof(d));
    for (int r = 0; r < world; ++r) {
        if (!ptrs[r] || (reinterpret_cast<uintptr_t>(ptrs[r]) & (align - 1)))
            return set_error(-1, "%s: peer pointer %d is null or not %zu-byte aligned", what, r, align);
        d.base[r] = ptrs[r];
    }
    return 0;
}

// this rank's granules inside [lo, hi); false when it owns none
static bool dp_range(DpRange& rg, long long lo, long long hi, int gshift, int world, int rank) {
    rg.lo = lo; rg.hi = hi; rg.gshift = gshift; rg.world = world;
    const long long qa = lo >> gshift, qb = (hi - 1) >> gshift;        // first / last granule touched
    long long q0 = qa + ((rank - qa % world) % world + world) % world;
    rg.q0 = q0;
    rg.nq = q0 > qb ? 0 : (qb - q0) / world + 1;
    return rg.nq > 0;
}

extern "C" long long vitae_dp_owned_elems(long long lo, long long hi, int granule_shift, int world, int rank) {
    if (hi <= lo || world < 1 || rank < 0 || rank >= world || granule_shift < 6) return 0;
    DpRange rg;
    if (!dp_range(rg, lo, hi, granule_shift, world, rank)) return 0;
    long long n = 0;
    for (long long k = 0; k < rg.nq; ++k) {
        const long long a = std::max(lo, (rg.q0 + k * world) << granule_shift);
        const long long b = std::min(hi, ((rg.q0 + k * world) + 1) << granule_shift);
        n += b - a;
    }
    return n;
}

extern "C" int vitae_dp_reduce_shard_blocks(long long lo, long long hi, int granule_shift, int world, int rank, int max_blocks) {
    DpRange rg;
    if (hi <= lo || !dp_range(rg, lo, hi, granule_shift, world, rank)) return 0;
    const int cap = max_blocks > 0 ? std::min(max_blocks, OPT_NORM_BLOCKS) : OPT_NORM_BLOCKS;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(ceil_div<long long>(rg.nq << granule_shift, 2048), cap)));
}

extern "C" int vitae_dp_reduce_shard(void* const* grad_ptrs, int world, int rank, long long lo, long long hi, int granule_shift,
                                     float inv_world, float* partials, int max_blocks, void* stream) {
    VITAE_REQUIRE(grad_ptrs && partials && world >= 2 && world <= 8 && rank >= 0 && rank < world, "dp_reduce_shard: bad arguments");
    VITAE_REQUIRE(hi > lo && lo >= 0 && lo % 64 == 0 && hi % 64 == 0 && granule_shift >= 6 && granule_shift <= 30,
                  "dp_reduce_shard: [%lld, %lld) must be 64-aligned, granule shift %d in 6..30", lo, hi, granule_shift);
    DpPeers peers;
    if (int rc = fill_peers(peers, grad_ptrs, world, 16, "dp_reduce_shard")) return rc;
    DpRange rg;
    if (!dp_range(rg, lo, hi, granule_shift, world, rank)) return 0;       // nothing of this slice belongs to this rank
    const int blocks = vitae_dp_reduce_shard_blocks(lo, hi, granule_shift, world, rank, max_blocks);
    cudaStream_t st = as_stream(stream);
#define CALL(W) launch_dp_reduce<W>(blocks, st, peers, rank, rg, inv_world, partials)
    DP_DISPATCH_WORLD(world, CALL)
#undef CALL
    VITAE_CHECK_LAUNCH("dp_reduce_shard");
    return 0;
}

extern "C" int vitae_adamw_shard(void* const* param_ptrs, void* const* param_bf16_ptrs, int world, int rank, long long lo,
                                 long long hi, int granule_shift, const float* grad, float* exp_avg, float* exp_avg_sq,
                                 const unsigned char* group_of_chunk, const unsigned char* f32_chunk, const float* hyper,
                                 int ngroups, const float* ctl, int max_blocks, void* stream) {
    VITAE_REQUIRE(param_ptrs && param_bf16_ptrs && grad && exp_avg && exp_avg_sq && group_of_chunk && f32_chunk && hyper && ctl,
                  "adamw_shard: null pointer");
    VITAE_REQUIRE(world >= 2 && world <= 8 && rank >= 0 && rank < world, "adamw_shard: world=%d rank=%d", world, rank);
    VITAE_REQUIRE(hi > lo && lo >= 0 && lo % 64 == 0 && hi % 64 == 0 && granule_shift >= 6 && granule_shift <= 30 && ngroups > 0 &&
                      ngroups <= 8,
                  "adamw_shard: [%lld, %lld) must be 64-aligned, granule shift %d in 6..30, 1..8 groups", lo, hi, granule_shift);
    DpPeers p32s, p16s;
    if (int rc = fill_peers(p32s, param_ptrs, world, 16, "adamw_shard(master)")) return rc;
    if (int rc = fill_peers(p16s, param_bf16_ptrs, world, 16, "adamw_shard(shadow)")) return rc;
    DpRange rg;
    if (!dp_range(rg, lo, hi, granule_shift, world, rank)) return 0;
    const int cap = max_blocks > 0 ? max_blocks : 148 * 8;
    const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(ceil_div<long long>(rg.nq << granule_shift, 2048), cap)));
    AdamHyper hy;
    memset(&hy, 0, sizeof(hy));
    memcpy(hy.h, hyper, sizeof(float) * 8 * ngroups);
    cudaStream_t st = as_stream(stream);
#define CALL(W) launch_adamw_shard<W>(blocks, st, p32s, p16s, rank, rg, grad, exp_avg, exp_avg_sq, group_of_chunk, f32_chunk, hy, ngroups, ctl)
    DP_DISPATCH_WORLD(world, CALL)
#undef CALL
    VITAE_CHECK_LAUNCH("adamw_shard");
    return 0;
}
