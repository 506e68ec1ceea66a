// bf16 x bf16 -> fp32 GEMM on the Blackwell 5th-gen tensor cores.
//
//   * operands staged by TMA (cp.async.bulk.tensor.2d) into 128B-swizzled shared-memory tiles; a pipeline stage
//     holds SPB sub-blocks of 64 K-elements, so one mbarrier round trip (producer: wait / expect_tx, MMA issuer:
//     wait / commit) is amortised over SPB * 4 tcgen05.mma -- the GEMMs of this path are short (K = 512 .. 3072)
//     and the per-round-trip latency of the single producer / issuer threads, not bandwidth, bounded them
//     (tools/tma_bench.cu: 330 cycles per 64-wide block with one block per stage, 220 = the ~110 B/clk/SM ingest
//     ceiling with two);
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) issued by one elected thread, accumulator in TMEM;
//   * epilogue, compile-time specialised per output kind: each of the four epilogue warps reads its TMEM lane
//     quadrant with tcgen05.ld (lane = accumulator row), applies alpha / bias / residual addend / GELU' / GELU in
//     registers, writes 32-row x 128-byte boxes into 128B-swizzled staging buffers (the drained pipeline stages) and
//     one elected lane hands them to the TMA store engine (cp.async.bulk.tensor store, or fp32 reduce-add when
//     accumulating): ~150 instructions per warp instead of per-element address arithmetic, and ragged M / N edges
//     are clipped by the tensor map;
//   * both operand majors (K-major = nn.Linear layout, MN-major = transposed) so that forward, dgrad and wgrad of
//     every Linear on the path (see include/vitae_b200.h) are the same kernel with different tensor maps;
//   * split-K and the rarely used epilogue features (row scatter / gather maps, unusual output combinations) go
//     through fp32 slabs: every split stores its partial tile (same TMA-store epilogue), a finalize kernel reduces
//     the slabs in fixed order (deterministic) and applies the generic epilogue.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.h"
#include "ptx.cuh"

namespace vitae {

constexpr int BM = 128;
constexpr int BK = 64;             // K elements per sub-block (= one 128-byte swizzled row)
#ifndef VITAE_EPI_WARPS
#define VITAE_EPI_WARPS 8          // 8: two warps per TMEM lane quadrant, each takes half of the tile's columns (BN >= 128)
#endif
constexpr int EPI_THREADS = VITAE_EPI_WARPS * 32;
// Warp roles: the epilogue warps come FIRST, the TMA producer and the MMA issuer LAST.  The warp scheduler favours the highest
// warp id of a sub-partition; as warps 0 / 1 the two single-thread roles queued behind the epilogue warps that poll the
// accumulator barrier next to them (measured on the attention kernels, profiles/r02e_attention_event_log.txt: ~200 cycles
// per tcgen05.mma / commit / barrier wait of a warp-0/1 thread).
constexpr int GEMM_THREADS = 64 + EPI_THREADS;
constexpr int GEMM_PROD_WARP = VITAE_EPI_WARPS, GEMM_MMA_WARP = VITAE_EPI_WARPS + 1;

// generic epilogue description (finalize kernel); mirrors vitae_gemm_epilogue
struct EpiParams {
    float alpha;
    const float* alpha_ptr;
    const float* bias;
    const float* addend;
    const int* add_rows;
    int ldadd;
    const __nv_bfloat16* dgelu_src;
    int ld_dgelu;
    float* out_f32;
    int ld_f32;
    int accumulate;
    __nv_bfloat16* out_bf16;
    __nv_bfloat16* out_gelu_bf16;
    int ld_bf16;
    const int* out_rows;
};

// what the in-kernel epilogue needs besides the output tensor maps
struct EpiLite {
    float alpha;
    const float* alpha_ptr;
    const float* bias;
    const float* addend;              // fp32 [M, ldadd], same rows as the output (EPI_ADD_F32)
    int ldadd;
    const __nv_bfloat16* dgelu_src;   // bf16 [M, ld_dgelu] (EPI_DGELU_BF16)
    int ld_dgelu;
    int accumulate;                   // EPI_F32: add into the destination (TMA reduce-add)
    int a_atoms, b_atoms;             // MN-major operand loaded through the 3-D "atom" tensor map: one TMA per sub-block
    // EPI_PRED_MSE (decoder_pred GEMM + masked patch-reconstruction loss, vitae_gemm_pred_mse)
    const float* mse_vol;             // fp32 volume [B, 4, V, V, V], read in place
    const float* mse_mask;            // fp32 [B, L], 1 = removed patch
    float* mse_part;                  // [CTAs][epilogue warps] partial sums of squared errors
    int mse_V, mse_p, mse_g, mse_L;
    float mse_coef;                   // 2 / (P * sum(mask)): d recon / d pred before the upstream gradient
};

// in-kernel epilogue kinds (everything else is "generic": slabs + finalize kernel)
enum EpiKind : int {
    EPI_BF16 = 0,        // out_bf16 = bf16(v)                                   qkv / decoder_pred forward, dgrads
    EPI_BF16_GELU = 1,   // out_bf16 = bf16(v), out1 = bf16(gelu(v))             fc1 forward (pre-activation kept for GELU')
    EPI_ADD_F32 = 2,     // out_f32 = v + addend                                 proj / fc2 forward (+ residual stream)
    EPI_BF16_F32 = 3,    // out_bf16 = bf16(v), out1(f32) = v                    decoder_pred forward with an fp32 copy
    EPI_F32 = 4,         // out_f32 (+)= v                                       wgrads, split-K slabs
    EPI_DGELU_BF16 = 5,  // out_bf16 = bf16(v * gelu'(src))                      fc2 dgrad
    EPI_PRED_MSE = 6,    // out_bf16 = bf16(v), out1 = bf16(coef * (v - target) * mask), row partials of (v - target)^2
};

#ifdef VITAE_EPI_NOGELU
__device__ __forceinline__ float gelu_erf(float x) { return x; }
__device__ __forceinline__ float dgelu_erf(float x) { return x; }
#else
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
#endif

// Phase tracing (debug builds only: VITAE_TRACE=1 python -m vit_ae_plus_plus_b200.build): per CTA, %globaltimer stamps of
// the kernel phases are written to a device buffer registered with vitae_debug_set_gemm_trace (tools/gemm_trace.py).
#ifdef VITAE_GEMM_TRACE
__device__ unsigned long long* g_gemm_trace = nullptr;
#define GEMM_TRACE(slot)                                                                                      \
    do {                                                                                                      \
        if (g_gemm_trace) {                                                                                   \
            const int cta__ = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;                 \
            unsigned long long t__;                                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                           \
            g_gemm_trace[cta__ * 16 + (slot)] = t__;                                                          \
        }                                                                                                     \
    } while (0)
#else
#define GEMM_TRACE(slot) do { } while (0)
#endif

template <int BN, int SPB, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = BM * BK * 2;               // one sub-block of A
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int SUB_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGE_BYTES = SPB * SUB_BYTES;
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = PIPE_BYTES;             // full[STAGES], empty[STAGES], tmem_full, tmem slot
    static constexpr int BIAS_OFFSET = BAR_OFFSET + (2 * STAGES + 2) * 8;
    static constexpr int TOTAL = BIAS_OFFSET + BN * 4 + 1024; // + alignment slack
    // epilogue staging (reuses the drained pipeline): 8 warps x 2 x 4 KB (one output: two alternating buffers; two
    // outputs: one buffer each)
    static_assert(PIPE_BYTES >= 8 * 2 * 4096, "epilogue staging does not fit into the pipeline stages");
};

// 16-byte chunk j of row r inside a 32-row x 128-byte SWIZZLE_128B box
__device__ __forceinline__ uint32_t swz128(int r, int j) { return static_cast<uint32_t>(r * 128 + ((j ^ (r & 7)) << 4)); }

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

template <int BN, int SPB, int STAGES, bool A_MN, bool B_MN, int KIND>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1, int M, int N,
                         int num_sub, int sub_per_split, EpiLite ep) {
    using S = GemmSmem<BN, SPB, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* sm = smem_raw + (base - raw_u32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::BAR_OFFSET);
    const uint32_t full_bar = base + S::BAR_OFFSET;
    const uint32_t empty_bar = full_bar + STAGES * 8;
    const uint32_t tmem_full_bar = empty_bar + STAGES * 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * STAGES + 1);
    float* bias_s = reinterpret_cast<float*>(sm + S::BIAS_OFFSET);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int sub_begin = blockIdx.z * sub_per_split;
    const int nsub = min(num_sub, sub_begin + sub_per_split) - sub_begin;
    const int niter = (nsub + SPB - 1) / SPB;
    if (threadIdx.x == 0) GEMM_TRACE(0);

    if (warp == GEMM_PROD_WARP && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO0);
        if (KIND == EPI_BF16_GELU || KIND == EPI_BF16_F32 || KIND == EPI_PRED_MSE) tma_prefetch_desc(&tmO1);
    }
    if (warp == GEMM_MMA_WARP && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s * 8, 1);
            mbar_init(empty_bar + s * 8, 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // PDL: our first global access (TMA loads, epilogue operands) waits for the predecessor grid; the next kernel in the
    // stream may start its prologue once every CTA has issued its last MMA (PDL_TRIGGER_LATE below; ptx.cuh).
    if (threadIdx.x == 0) GEMM_TRACE(1);
    PDL_TRIGGER_EARLY();
    pdl_wait();
    if (threadIdx.x == 0) GEMM_TRACE(2);

    if (warp == GEMM_PROD_WARP) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int it = 0; it < niter; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int cnt = min(SPB, nsub - it * SPB);
                mbar_wait(empty_bar + s * 8, ph ^ 1);
                mbar_arrive_expect_tx(full_bar + s * 8, cnt * S::SUB_BYTES);
#pragma unroll
                for (int u = 0; u < SPB; ++u) {
                    if (u < cnt) {
                        const uint32_t sa = base + s * S::STAGE_BYTES + u * S::SUB_BYTES;
                        const uint32_t sb = sa + S::A_BYTES;
                        const int k0 = (sub_begin + it * SPB + u) * BK;
                        if (A_MN) {
                            if (ep.a_atoms) {
                                tma_load_3d(sa, &tmA, full_bar + s * 8, 0, k0, m0 >> 6);
                            } else {
#pragma unroll
                                for (int j = 0; j < BM / 64; ++j)
                                    tma_load_2d(sa + j * (BK * 128), &tmA, full_bar + s * 8, m0 + 64 * j, k0);
                            }
                        } else {
                            tma_load_2d(sa, &tmA, full_bar + s * 8, k0, m0);
                        }
                        if (B_MN) {
                            if (ep.b_atoms) {
                                tma_load_3d(sb, &tmB, full_bar + s * 8, 0, k0, n0 >> 6);
                            } else {
#pragma unroll
                                for (int j = 0; j < BN / 64; ++j)
                                    tma_load_2d(sb + j * (BK * 128), &tmB, full_bar + s * 8, n0 + 64 * j, k0);
                            }
                        } else {
                            tma_load_2d(sb, &tmB, full_bar + s * 8, k0, n0);
                        }
                    }
                }
                if (it == 0) GEMM_TRACE(3);
            }
            GEMM_TRACE(4);
        }
        __syncwarp();
    } else if (warp == GEMM_MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer (single thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
            for (int it = 0; it < niter; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int cnt = min(SPB, nsub - it * SPB);
                mbar_wait(full_bar + s * 8, ph);
                tc_fence_after();
                if (it == 0) GEMM_TRACE(5);
#pragma unroll
                for (int u = 0; u < SPB; ++u) {
                    if (u < cnt) {
                        const uint32_t sa = base + s * S::STAGE_BYTES + u * S::SUB_BYTES;
                        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            // K-major: 16 bf16 = 32 B along the swizzled row; MN-major: 16 K-rows of 128 B
                            const uint64_t da = A_MN ? umma_smem_desc_sw128(sa + kk * 2048, BK * 128, 1024)
                                                     : umma_smem_desc_sw128(sa + kk * 32, 16, 1024);
                            const uint64_t db = B_MN ? umma_smem_desc_sw128(sb + kk * 2048, BK * 128, 1024)
                                                     : umma_smem_desc_sw128(sb + kk * 32, 16, 1024);
                            umma_bf16(tmem_base, da, db, idesc, (it | u | kk) != 0 ? 1u : 0u);
                        }
                    }
                }
                umma_commit(empty_bar + s * 8);  // frees the smem stage once these MMAs retire
            }
            umma_commit(tmem_full_bar);
            PDL_TRIGGER_LATE();
            GEMM_TRACE(6);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int q = warp & 3;                    // TMEM lane quadrant this warp may access
        const int half = warp >> 2;                // which half of the tile's columns (BN >= 128); BN = 64: half 1 idles
        const int et = threadIdx.x;                // 0..255: the epilogue threads come first
        const int mrow = m0 + q * 32 + lane;       // this lane's accumulator row
        // bias slice of this tile -> shared memory (read by every lane for every row)
        for (int c = et; c < BN; c += EPI_THREADS) bias_s[c] = (ep.bias && n0 + c < N) ? ep.bias[n0 + c] : 0.f;
        const float alpha = ep.alpha * (ep.alpha_ptr ? *ep.alpha_ptr : 1.0f);
        named_bar_sync(1, EPI_THREADS);
        constexpr int CW = (BN >= 128 && VITAE_EPI_WARPS == 8) ? BN / 2 : BN;      // columns per epilogue warp
        constexpr bool TWO_OUT = KIND == EPI_BF16_GELU || KIND == EPI_BF16_F32 || KIND == EPI_PRED_MSE;
        // EPI_PRED_MSE: this row's patch (model/vit_autoenc.py:100-113 within-patch order (pz, py, px, c), c = 4 channels)
        bool mse_live = false;
        const float* mse_base = nullptr;
        float mse_rs = 0.f;
        if (KIND == EPI_PRED_MSE && mrow < M) {
            const int L1 = ep.mse_L + 1;
            const int bb = mrow / L1, tt = mrow - bb * L1;
            if (tt > 0 && ep.mse_mask[static_cast<size_t>(bb) * ep.mse_L + tt - 1] != 0.f) {
                mse_live = true;
                const int l = tt - 1, g = ep.mse_g, p = ep.mse_p, V = ep.mse_V;
                const int gz = l / (g * g), gy = (l / g) % g, gx = l % g;
                mse_base = ep.mse_vol + static_cast<size_t>(bb) * 4 * V * V * V + (static_cast<size_t>(gz * p) * V + gy * p) * V + gx * p;
            }
        }
        // staging per warp: 8 KB = two 32-row x 128-byte boxes; one output: they alternate, two outputs: one each
        const uint32_t stg = base + warp * 8192;
        const bool active = (BN >= 128 && VITAE_EPI_WARPS == 8) || half == 0;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 0) GEMM_TRACE(7);

        constexpr bool OUT0_BF16 = KIND != EPI_ADD_F32 && KIND != EPI_F32;   // primary output element type
        int nbox = 0;   // bf16 boxes committed so far
        const int c_begin = half * CW;
#pragma unroll 1
        for (int c = c_begin; active && c < c_begin + CW; c += 32) {
            if (n0 + c >= N) break;
            uint32_t raw[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, raw);
            // the staging buffer about to be rewritten must have been read by the TMA engine
            if (lane == 0) {
                if (TWO_OUT) tma_store_wait_read<0>();
                else tma_store_wait_read<1>();
            }
            // operands of this chunk that come from global memory (lane = row: 128 / 64 contiguous bytes per lane)
            float4 add4[KIND == EPI_ADD_F32 ? 8 : 1];
            uint4 dg4[KIND == EPI_DGELU_BF16 ? 4 : 1];
            if (KIND == EPI_ADD_F32) {
                const float* arow = ep.addend + static_cast<size_t>(mrow) * ep.ldadd + n0 + c;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    add4[j] = (mrow < M && n0 + c + 4 * j < N) ? *reinterpret_cast<const float4*>(arow + 4 * j)
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (KIND == EPI_DGELU_BF16) {
                const __nv_bfloat16* drow = ep.dgelu_src + static_cast<size_t>(mrow) * ep.ld_dgelu + n0 + c;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dg4[j] = (mrow < M && n0 + c + 8 * j < N) ? *reinterpret_cast<const uint4*>(drow + 8 * j)
                                                              : make_uint4(0u, 0u, 0u, 0u);
            }
            tmem_ld_wait();
            __syncwarp();   // lane 0's wait_group.read above now covers the whole warp
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(bias_s + c + 4 * j);
                v[4 * j + 0] = __uint_as_float(raw[4 * j + 0]) * alpha + b.x;
                v[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) * alpha + b.y;
                v[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) * alpha + b.z;
                v[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) * alpha + b.w;
            }
            if (KIND == EPI_ADD_F32) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[4 * j + 0] += add4[j].x; v[4 * j + 1] += add4[j].y; v[4 * j + 2] += add4[j].z; v[4 * j + 3] += add4[j].w;
                }
            }
            if (KIND == EPI_DGELU_BF16) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&dg4[j]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __bfloat1622float2(h[i]);
                        v[8 * j + 2 * i] *= dgelu_erf(f.x);
                        v[8 * j + 2 * i + 1] *= dgelu_erf(f.y);
                    }
                }
            }
            float gq[KIND == EPI_PRED_MSE ? 32 : 1];
            if (KIND == EPI_PRED_MSE) {
                if (mse_live) {
                    // 32 columns = 8 voxels along x times 4 channels: 32-byte segments of the four channel planes
                    const int p = ep.mse_p, V = ep.mse_V;
                    const int vox0 = (n0 + c) >> 2;
                    const int px0 = vox0 % p, py = (vox0 / p) % p, pz = vox0 / (p * p);
                    const float* a = mse_base + (static_cast<size_t>(pz) * V + py) * V + px0;
                    const size_t V3 = static_cast<size_t>(V) * V * V;
                    float tg[4][8];
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const float4 lo = __ldg(reinterpret_cast<const float4*>(a + ch * V3));
                        const float4 hi = __ldg(reinterpret_cast<const float4*>(a + ch * V3) + 1);
                        tg[ch][0] = lo.x; tg[ch][1] = lo.y; tg[ch][2] = lo.z; tg[ch][3] = lo.w;
                        tg[ch][4] = hi.x; tg[ch][5] = hi.y; tg[ch][6] = hi.z; tg[ch][7] = hi.w;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = v[j] - tg[j & 3][j >> 2];
                        mse_rs = fmaf(d, d, mse_rs);
                        gq[j] = d * ep.mse_coef;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) gq[j] = 0.f;
                }
            }
            const int hc = (c >> 5) & 1;   // even / odd 32-column chunk
            // buffer assignment: one output -> boxes alternate between the two buffers; two outputs -> buffer 0 = primary
            // (bf16), buffer 1 = secondary
            const uint32_t buf0 = stg + ((!TWO_OUT && ((OUT0_BF16 ? nbox : hc) & 1)) ? 4096 : 0);
            const uint32_t buf1 = stg + 4096;
            if (OUT0_BF16) {
                // a 32-row x 64-column bf16 box spans two 32-column chunks: even chunk -> 16B chunks 0..3, odd -> 4..7
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sts128(buf0 + swz128(lane, 4 * hc + j), pack_bf16(v[8 * j + 0], v[8 * j + 1]),
                           pack_bf16(v[8 * j + 2], v[8 * j + 3]), pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                           pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                if (KIND == EPI_PRED_MSE) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        sts128(buf1 + swz128(lane, 4 * hc + j), pack_bf16(gq[8 * j + 0], gq[8 * j + 1]),
                               pack_bf16(gq[8 * j + 2], gq[8 * j + 3]), pack_bf16(gq[8 * j + 4], gq[8 * j + 5]),
                               pack_bf16(gq[8 * j + 6], gq[8 * j + 7]));
                }
                if (KIND == EPI_BF16_GELU) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        sts128(buf1 + swz128(lane, 4 * hc + j),
                               pack_bf16(gelu_erf(v[8 * j + 0]), gelu_erf(v[8 * j + 1])),
                               pack_bf16(gelu_erf(v[8 * j + 2]), gelu_erf(v[8 * j + 3])),
                               pack_bf16(gelu_erf(v[8 * j + 4]), gelu_erf(v[8 * j + 5])),
                               pack_bf16(gelu_erf(v[8 * j + 6]), gelu_erf(v[8 * j + 7])));
                }
            }
            if (!OUT0_BF16 || KIND == EPI_BF16_F32) {
                // fp32 box: 32 rows x 32 columns (128 B per row), one per chunk
                const uint32_t buf = KIND == EPI_BF16_F32 ? buf1 : buf0;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts128(buf + swz128(lane, j), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
            }
            fence_proxy_async();
            __syncwarp();
            const bool box_done = OUT0_BF16 && (hc == 1 || n0 + c + 32 >= N);
            if (lane == 0) {
                if (!OUT0_BF16) {
                    if (KIND == EPI_F32 && ep.accumulate) tma_reduce_add_3d(&tmO0, buf0, n0 + c, m0 + q * 32, blockIdx.z);
                    else tma_store_3d(&tmO0, buf0, n0 + c, m0 + q * 32, blockIdx.z);
                }
                if (KIND == EPI_BF16_F32) tma_store_3d(&tmO1, buf1, n0 + c, m0 + q * 32, 0);
                if (box_done) {
                    tma_store_3d(&tmO0, buf0, n0 + (c & ~63), m0 + q * 32, 0);
                    if (KIND == EPI_BF16_GELU || KIND == EPI_PRED_MSE) tma_store_3d(&tmO1, buf1, n0 + (c & ~63), m0 + q * 32, 0);
                }
                if (!OUT0_BF16 || KIND == EPI_BF16_F32 || box_done) tma_store_commit();
            }
            if (box_done) ++nbox;
        }
        if (KIND == EPI_PRED_MSE) {                // one partial per epilogue warp: its 32 rows over its columns, fixed order
            float r = (active && mrow < M) ? mse_rs : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (lane == 0)
                ep.mse_part[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * VITAE_EPI_WARPS + warp] = r;
        }
        if (lane == 0) tma_store_wait_read<0>();   // the engine must have read our staging buffers before the CTA exits
        __syncwarp();
        tc_fence_before();
        if (threadIdx.x == 0) GEMM_TRACE(8);
    }
    __syncthreads();
    if (threadIdx.x == 0) GEMM_TRACE(9);
    if (warp == 0) tmem_dealloc(tmem_base, BN);
}

// ---------------------------------------------------------------------------------------------------- generic epilogue
__device__ __forceinline__ void generic_epilogue4(const EpiParams& ep, float alpha, int m, int n, float4 acc) {
    float v[4] = {acc.x * alpha, acc.y * alpha, acc.z * alpha, acc.w * alpha};
    if (ep.bias) {
        const float4 b = *reinterpret_cast<const float4*>(ep.bias + n);
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
    }
    if (ep.addend) {
        const int ra = ep.add_rows ? ep.add_rows[m] : m;
        const float4 a = *reinterpret_cast<const float4*>(ep.addend + static_cast<size_t>(ra) * ep.ldadd + n);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
    }
    if (ep.dgelu_src) {
        const uint2 raw = *reinterpret_cast<const uint2*>(ep.dgelu_src + static_cast<size_t>(m) * ep.ld_dgelu + n);
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        v[0] *= dgelu_erf(lo.x); v[1] *= dgelu_erf(lo.y); v[2] *= dgelu_erf(hi.x); v[3] *= dgelu_erf(hi.y);
    }
    const int r = ep.out_rows ? ep.out_rows[m] : m;
    if (ep.out_f32) {
        float* o = ep.out_f32 + static_cast<size_t>(r) * ep.ld_f32 + n;
        float4 o4 = make_float4(v[0], v[1], v[2], v[3]);
        if (ep.accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(o);
            o4.x += p.x; o4.y += p.y; o4.z += p.z; o4.w += p.w;
        }
        *reinterpret_cast<float4*>(o) = o4;
    }
    if (ep.out_bf16) {
        uint2 pk;
        pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2*>(ep.out_bf16 + static_cast<size_t>(r) * ep.ld_bf16 + n) = pk;
    }
    if (ep.out_gelu_bf16) {
        uint2 pk;
        pk.x = pack_bf16(gelu_erf(v[0]), gelu_erf(v[1])); pk.y = pack_bf16(gelu_erf(v[2]), gelu_erf(v[3]));
        *reinterpret_cast<uint2*>(ep.out_gelu_bf16 + static_cast<size_t>(r) * ep.ld_bf16 + n) = pk;
    }
}

// Split-K / generic finalize: sum the slabs in fixed order, then the generic epilogue.  One thread per (row, 4 columns).
__global__ void __launch_bounds__(256)
gemm_splitk_finalize_kernel(const float* __restrict__ slabs, int splits, int M, int N, EpiParams ep) {
    pdl_trigger();
    pdl_wait();
    const int groups_per_row = N / 4;
    const long long total = static_cast<long long>(M) * groups_per_row;
    const float alpha = ep.alpha * (ep.alpha_ptr ? *ep.alpha_ptr : 1.0f);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int m = static_cast<int>(idx / groups_per_row);
        const int n = static_cast<int>(idx % groups_per_row) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < splits; ++s) {
            const float4 p = *reinterpret_cast<const float4*>(slabs + (static_cast<size_t>(s) * M + m) * N + n);
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        generic_epilogue4(ep, alpha, m, n, a);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct TmapKey {
    const void* ptr;
    uint64_t d0, d1, d2, ld;
    uint32_t b0, b1, esize;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld == o.ld && b0 == o.b0 && b1 == o.b1 &&
               esize == o.esize;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.d0); mix(k.d1); mix(k.d2); mix(k.ld); mix(k.b0); mix(k.b1); mix(k.esize);
        return h;
    }
};

// Tensor map over a row-major matrix, or (d2 > 0) a stack of d2 matrices of d1 rows (pitch ld elements, matrix pitch
// d1*ld): inner dim d0 (contiguous), box b0 x b1 (x 1), 128B swizzle, element size 2 (bf16) or 4 (fp32).
// Out-of-bounds elements read as zero and are clipped on stores, so ragged M / N / K need no special casing.
int make_tmap(CUtensorMap* out, const void* ptr, uint32_t esize, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld,
              uint32_t b0, uint32_t b1) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    const TmapKey key{ptr, d0, d1, d2, ld, b0, b1, esize};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return set_error(-4, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    const cuuint32_t rank = d2 ? 3 : 2;
    const cuuint64_t dims[3] = {d0, d1, d2 ? d2 : 1};
    const cuuint64_t strides[2] = {ld * esize, d1 * ld * esize};
    const cuuint32_t box[3] = {b0, b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                           const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(-5, "cuTensorMapEncodeTiled failed (%d) ptr=%p esize=%u dims=%llu,%llu,%llu ld=%llu box=%u,%u", (int)r,
                         ptr, esize, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                         (unsigned long long)ld, b0, b1);
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

// MN-major operand [K rows, MN columns] (pitch ld) seen as (64 columns, K rows, MN/64 column atoms): ONE box
// (64, BK, atoms) lands in shared memory as [atom][k][64] -- the 128-byte-swizzled MN-major layout the MMA descriptors
// expect -- where the 2-D map needs one TMA per 64-column atom (6 per sub-block for a 128 x 256 tile; an issue of the
// single producer thread costs ~100 cycles, about the transfer time of one 8 KB atom).  Needs MN % 64 == 0 (the view would
// otherwise read past the row); atoms past MN / 64 are out of bounds of the map and read as zero.
static int make_tmap_atoms(CUtensorMap* out, const void* ptr, uint64_t MN, uint64_t K, uint64_t ld, uint32_t atoms) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    const TmapKey key{ptr, MN, K, 0xA70115ull, ld, atoms, BK, 2};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return -4;
    const cuuint64_t dims[3] = {64, K, MN / 64};
    const cuuint64_t strides[2] = {ld * 2, 128};
    const cuuint32_t box[3] = {64, BK, atoms};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -5;       // the caller falls back to one 2-D box per atom
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

struct GemmLaunch {
    CUtensorMap ta, tb, to0, to1;
    int M, N, num_sub, splits;
    EpiLite ep;
    cudaStream_t stream;
};

template <int BN, int SPB, int STAGES, bool A_MN, bool B_MN, int KIND>
static int launch_gemm(const GemmLaunch& g) {
    using S = GemmSmem<BN, SPB, STAGES>;
    auto kern = gemm_bf16_tcgen05_kernel<BN, SPB, STAGES, A_MN, B_MN, KIND>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(smem=%d): %s", S::TOTAL, cudaGetErrorString(e));
        attr_set = true;
    }
    const int sub_per_split = ceil_div(g.num_sub, g.splits);
    const int eff_splits = ceil_div(g.num_sub, sub_per_split);  // every z-slice gets >= 1 sub-block
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), eff_splits);
    launch_kernel(kern, grid, dim3(GEMM_THREADS), S::TOTAL, g.stream, g.ta, g.tb, g.to0, g.to1, g.M, g.N, g.num_sub,
                  sub_per_split, g.ep);
    VITAE_CHECK_LAUNCH("gemm_bf16_tcgen05");
    return eff_splits;
}

// tile configurations <BN, SPB, STAGES>: "deep" = the whole shared memory for one CTA per SM, "shallow" = two CTAs per SM
template <bool A_MN, bool B_MN, int KIND>
static int dispatch_tile(int bn, bool deep, const GemmLaunch& g) {
    if (bn == 64) return deep ? launch_gemm<64, 2, 4, A_MN, B_MN, KIND>(g) : launch_gemm<64, 2, 2, A_MN, B_MN, KIND>(g);
    if (bn == 128) return deep ? launch_gemm<128, 2, 3, A_MN, B_MN, KIND>(g) : launch_gemm<128, 1, 3, A_MN, B_MN, KIND>(g);
    return deep ? launch_gemm<256, 2, 2, A_MN, B_MN, KIND>(g) : launch_gemm<256, 1, 2, A_MN, B_MN, KIND>(g);
}

// only the (operand major, epilogue kind) pairs the training step uses are instantiated; anything else runs as EPI_F32
// slabs + the generic finalize kernel
static int dispatch(bool amn, bool bmn, int kind, int bn, bool deep, const GemmLaunch& g) {
    if (!amn && !bmn) {
        switch (kind) {
            case EPI_BF16: return dispatch_tile<false, false, EPI_BF16>(bn, deep, g);
            case EPI_BF16_GELU: return dispatch_tile<false, false, EPI_BF16_GELU>(bn, deep, g);
            case EPI_ADD_F32: return dispatch_tile<false, false, EPI_ADD_F32>(bn, deep, g);
            case EPI_BF16_F32: return dispatch_tile<false, false, EPI_BF16_F32>(bn, deep, g);
            default: return dispatch_tile<false, false, EPI_F32>(bn, deep, g);
        }
    }
    if (!amn && bmn) {
        switch (kind) {
            case EPI_BF16: return dispatch_tile<false, true, EPI_BF16>(bn, deep, g);
            case EPI_DGELU_BF16: return dispatch_tile<false, true, EPI_DGELU_BF16>(bn, deep, g);
            default: return dispatch_tile<false, true, EPI_F32>(bn, deep, g);
        }
    }
    if (amn && bmn) return dispatch_tile<true, true, EPI_F32>(bn, deep, g);
    return dispatch_tile<true, false, EPI_F32>(bn, deep, g);
}

static bool kind_available(bool amn, bool bmn, int kind) {
    if (kind == EPI_F32) return true;
    if (!amn && !bmn) return kind == EPI_BF16 || kind == EPI_BF16_GELU || kind == EPI_ADD_F32 || kind == EPI_BF16_F32;
    if (!amn && bmn) return kind == EPI_BF16 || kind == EPI_DGELU_BF16;
    return false;
}

// in-kernel epilogue kind for this epilogue description, or -1 when it needs the generic (slab + finalize) path
static int classify(const vitae_gemm_epilogue* e, bool amn, bool bmn) {
    if (e->out_rows || e->add_rows) return -1;
    const bool f32 = e->out_f32 != nullptr, b16 = e->out_bf16 != nullptr, gl = e->out_gelu_bf16 != nullptr;
    const bool add = e->addend != nullptr, dg = e->dgelu_src != nullptr;
    int kind = -1;
    if (b16 && !f32 && !gl && !add && !dg) kind = EPI_BF16;
    else if (b16 && gl && !f32 && !add && !dg) kind = EPI_BF16_GELU;
    else if (f32 && !b16 && !gl && add && !dg && !e->accumulate) kind = EPI_ADD_F32;
    else if (b16 && f32 && !gl && !add && !dg && !e->accumulate) kind = EPI_BF16_F32;
    else if (f32 && !b16 && !gl && !add && !dg) kind = EPI_F32;
    else if (b16 && dg && !f32 && !gl && !add) kind = EPI_DGELU_BF16;
    if (kind >= 0 && !kind_available(amn, bmn, kind)) kind = -1;
    return kind;
}

}  // namespace vitae

using namespace vitae;

#ifdef VITAE_GEMM_TRACE
extern "C" int vitae_debug_set_gemm_trace(void* buf) {
    cudaError_t e = cudaMemcpyToSymbol(g_gemm_trace, &buf, sizeof(buf));
    return e == cudaSuccess ? 0 : set_error(-3, "set_gemm_trace: %s", cudaGetErrorString(e));
}
#endif

extern "C" size_t vitae_gemm_workspace_bytes(int M, int N, int split_k) {
    if (split_k <= 1) return 0;
    return static_cast<size_t>(split_k) * M * N * sizeof(float);
}

extern "C" size_t vitae_gemm_workspace_bytes_for(const vitae_gemm_epilogue* e, int a_mn_major, int b_mn_major, int M, int N,
                                                 int split_k) {
    const int splits = split_k > 1 ? split_k : 1;
    if (splits == 1 && e && classify(e, a_mn_major != 0, b_mn_major != 0) >= 0) return 0;
    return static_cast<size_t>(splits) * M * N * sizeof(float);
}

extern "C" int vitae_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M,
                               int N, int K, const vitae_gemm_epilogue* e, int tile_n, int split_k, void* workspace,
                               size_t workspace_bytes, void* stream) {
    VITAE_REQUIRE(A && B && e, "gemm: null operand");
    VITAE_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    VITAE_REQUIRE(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
    VITAE_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: leading dims must be multiples of 8 (lda=%d ldb=%d)", lda, ldb);
    VITAE_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                  "gemm: operands must be 16-byte aligned");
    VITAE_REQUIRE(e->out_f32 || e->out_bf16 || e->out_gelu_bf16, "gemm: no output");
    VITAE_REQUIRE(!e->out_f32 || e->ld_f32 % 4 == 0, "gemm: ld_f32 %% 4");
    VITAE_REQUIRE(!(e->out_bf16 || e->out_gelu_bf16) || e->ld_bf16 % 8 == 0, "gemm: ld_bf16 %% 8");
    VITAE_REQUIRE(!e->addend || e->ldadd % 4 == 0, "gemm: ldadd %% 4");
    VITAE_REQUIRE(!e->dgelu_src || e->ld_dgelu % 8 == 0, "gemm: ld_dgelu %% 8");
    VITAE_REQUIRE(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K), "gemm: leading dim too small");
    VITAE_REQUIRE((reinterpret_cast<uintptr_t>(e->out_f32) & 15) == 0 && (reinterpret_cast<uintptr_t>(e->out_bf16) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(e->out_gelu_bf16) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(e->bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(e->addend) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(e->dgelu_src) & 15) == 0,
                  "gemm: epilogue pointers must be 16-byte aligned");

    const bool amn = a_mn_major != 0, bmn = b_mn_major != 0;
    int bn = tile_n;
    if (bn == 0) {
        const long long tiles128 = static_cast<long long>(ceil_div(M, BM)) * ceil_div(N, 128) * (split_k > 1 ? split_k : 1);
        bn = (tiles128 < 120) ? 64 : 128;
    }
    VITAE_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: tile_n must be 0/64/128/256 (got %d)", tile_n);
    const int num_sub = ceil_div(K, BK);
    int splits = split_k > 1 ? split_k : 1;
    if (splits > num_sub) splits = num_sub;
    int kind = classify(e, amn, bmn);
    const bool use_slabs = splits > 1 || kind < 0;
    float* slabs = nullptr;
    if (use_slabs) {
        const size_t need = static_cast<size_t>(splits) * M * N * sizeof(float);
        VITAE_REQUIRE(workspace && workspace_bytes >= need,
                      "gemm: split_k=%d / generic epilogue needs %zu workspace bytes, got %zu "
                      "(vitae_gemm_workspace_bytes_for)", splits, need, workspace_bytes);
        VITAE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "gemm: workspace must be 16-byte aligned");
        slabs = static_cast<float*>(workspace);
        kind = EPI_F32;
    }

    GemmLaunch g;
    g.M = M; g.N = N; g.num_sub = num_sub; g.splits = splits; g.stream = as_stream(stream);
    int rc;
    static const bool use_atoms = [] { const char* e = getenv("VITAE_GEMM_ATOM_TMA"); return !(e && e[0] == '0'); }();
    g.ep.a_atoms = g.ep.b_atoms = 0;
    if (amn && use_atoms && M % 64 == 0 && make_tmap_atoms(&g.ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM / 64) == 0) {
        g.ep.a_atoms = 1;
        rc = 0;
    } else if (amn) rc = make_tmap(&g.ta, A, 2, (uint64_t)M, (uint64_t)K, 0, (uint64_t)lda, 64, BK);
    else            rc = make_tmap(&g.ta, A, 2, (uint64_t)K, (uint64_t)M, 0, (uint64_t)lda, BK, BM);
    if (rc) return rc;
    if (bmn && use_atoms && N % 64 == 0 && make_tmap_atoms(&g.tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)bn / 64) == 0) {
        g.ep.b_atoms = 1;
        rc = 0;
    } else if (bmn) rc = make_tmap(&g.tb, B, 2, (uint64_t)N, (uint64_t)K, 0, (uint64_t)ldb, 64, BK);
    else            rc = make_tmap(&g.tb, B, 2, (uint64_t)K, (uint64_t)N, 0, (uint64_t)ldb, BK, (uint32_t)bn);
    if (rc) return rc;

    g.ep.alpha = e->alpha; g.ep.alpha_ptr = e->alpha_ptr; g.ep.bias = e->bias; g.ep.addend = e->addend; g.ep.ldadd = e->ldadd;
    g.ep.dgelu_src = static_cast<const __nv_bfloat16*>(e->dgelu_src); g.ep.ld_dgelu = e->ld_dgelu;
    g.ep.accumulate = e->accumulate;
    if (use_slabs) {
        // raw partial sums: alpha 1, no bias; the finalize kernel applies the real epilogue
        g.ep.alpha = 1.0f; g.ep.alpha_ptr = nullptr; g.ep.bias = nullptr; g.ep.addend = nullptr; g.ep.dgelu_src = nullptr;
        g.ep.accumulate = 0;
        rc = make_tmap(&g.to0, slabs, 4, (uint64_t)N, (uint64_t)M, (uint64_t)splits, (uint64_t)N, 32, 32);
        if (rc) return rc;
        g.to1 = g.to0;
    } else {
        const bool out0_bf16 = kind != EPI_ADD_F32 && kind != EPI_F32;
        if (out0_bf16) rc = make_tmap(&g.to0, e->out_bf16, 2, (uint64_t)N, (uint64_t)M, 1, (uint64_t)e->ld_bf16, 64, 32);
        else           rc = make_tmap(&g.to0, e->out_f32, 4, (uint64_t)N, (uint64_t)M, 1, (uint64_t)e->ld_f32, 32, 32);
        if (rc) return rc;
        g.to1 = g.to0;
        if (kind == EPI_BF16_GELU)
            rc = make_tmap(&g.to1, e->out_gelu_bf16, 2, (uint64_t)N, (uint64_t)M, 1, (uint64_t)e->ld_bf16, 64, 32);
        if (kind == EPI_BF16_F32)
            rc = make_tmap(&g.to1, e->out_f32, 4, (uint64_t)N, (uint64_t)M, 1, (uint64_t)e->ld_f32, 32, 32);
        if (rc) return rc;
    }

    // Pipeline depth: a grid that fits one CTA per SM gets the deep ring (all of shared memory for one CTA); larger
    // grids get the shallow one so that two CTAs share an SM (one's epilogue overlaps the other's main loop).
    const long long ctas = static_cast<long long>(ceil_div(M, BM)) * ceil_div(N, bn) * splits;
    // VITAE_GEMM_SHALLOW=1 (experiment): always the two-CTAs-per-SM rings, so that GEMMs of the two backward lanes can share SMs
    static const bool force_shallow = [] { const char* e = getenv("VITAE_GEMM_SHALLOW"); return e && e[0] == '1'; }();
    const bool deep = ctas <= 148 && !force_shallow;
    const int eff_splits = dispatch(amn, bmn, kind, bn, deep, g);
    if (eff_splits < 0) return eff_splits;
    if (use_slabs) {
        EpiParams ep;
        ep.alpha = e->alpha; ep.alpha_ptr = e->alpha_ptr; ep.bias = e->bias; ep.addend = e->addend;
        ep.add_rows = e->add_rows; ep.ldadd = e->ldadd;
        ep.dgelu_src = static_cast<const __nv_bfloat16*>(e->dgelu_src); ep.ld_dgelu = e->ld_dgelu;
        ep.out_f32 = e->out_f32; ep.ld_f32 = e->ld_f32; ep.accumulate = e->accumulate;
        ep.out_bf16 = static_cast<__nv_bfloat16*>(e->out_bf16);
        ep.out_gelu_bf16 = static_cast<__nv_bfloat16*>(e->out_gelu_bf16); ep.ld_bf16 = e->ld_bf16;
        ep.out_rows = e->out_rows;
        const long long total = static_cast<long long>(M) * (N / 4);
        const int blocks = static_cast<int>(std::min<long long>(ceil_div<long long>(total, 256), 148 * 8));
        launch_kernel(gemm_splitk_finalize_kernel, dim3(blocks), dim3(256), 0, g.stream, static_cast<const float*>(slabs),
                      eff_splits, M, N, ep);
        VITAE_CHECK_LAUNCH("gemm_splitk_finalize");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- decoder_pred + loss
// decoder_pred (model/vit_autoenc.py:198) with the masked patch-reconstruction loss (:226-227) evaluated in its epilogue:
// the accumulator tile (row = token, 32 columns = 8 voxels x 4 channels of the token's patch) is compared with the raw
// volume in place; besides pred (bf16) the kernel writes g = 2 (pred - target) mask / (P sum(mask)) (bf16: the gradient of
// the loss w.r.t. pred before the upstream factor, which the backward GEMMs apply through alpha_ptr) and per-row partial
// sums of squared errors that vitae_pred_mse_finalize adds up in fixed order.  Replaces masked_mse_fwd + masked_mse_bwd:
// pred is not re-read twice and the volume not read twice (468 -> 234 MB of HBM traffic per 4 x 128^3 x 4 batch).
__global__ void __launch_bounds__(256)
pred_mse_finalize_kernel(const float* __restrict__ part, long long n, float inv_pm, float mask_sum, float* __restrict__ loss_out) {
    pdl_trigger();
    pdl_wait();
    __shared__ double sh[256];
    double s = 0.0;
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(part) + i);
        s += static_cast<double>((v.x + v.y) + (v.z + v.w));
    }
    if (threadIdx.x == 0)
        for (long long k = n4 << 2; k < n; ++k) s += static_cast<double>(part[k]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        loss_out[0] = static_cast<float>(sh[0] * static_cast<double>(inv_pm));
        loss_out[1] = mask_sum;
    }
}

extern "C" size_t vitae_pred_mse_partial_floats(int M, int P, int tile_n) {
    return static_cast<size_t>(ceil_div(M, BM)) * ceil_div(P, tile_n) * VITAE_EPI_WARPS;
}

extern "C" int vitae_gemm_pred_mse(const void* hN, const void* W, const float* bias, int B, int L, int Dd, const float* vol,
                                   const float* mask, int C, int V, int p, float mask_sum, void* pred_bf16, void* g_bf16,
                                   float* partials, int tile_n, void* stream) {
    VITAE_REQUIRE(hN && W && vol && mask && pred_bf16 && g_bf16 && partials, "gemm_pred_mse: null pointer");
    VITAE_REQUIRE(C == 4 && p % 8 == 0 && p > 0 && V % p == 0, "gemm_pred_mse: needs 4 channels and patch %% 8 == 0 (C=%d p=%d V=%d)", C, p, V);
    VITAE_REQUIRE(tile_n == 128 || tile_n == 256, "gemm_pred_mse: tile_n must be 128 or 256");
    const int g = V / p, P = p * p * p * C, M = B * (L + 1);
    VITAE_REQUIRE(L == g * g * g && Dd % 8 == 0 && P % tile_n == 0 && mask_sum > 0.f, "gemm_pred_mse: bad geometry L=%d g=%d P=%d", L, g, P);
    VITAE_REQUIRE(((reinterpret_cast<uintptr_t>(hN) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(vol) |
                    reinterpret_cast<uintptr_t>(pred_bf16) | reinterpret_cast<uintptr_t>(g_bf16) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
                  "gemm_pred_mse: pointers must be 16-byte aligned");
    GemmLaunch gl;
    gl.M = M; gl.N = P; gl.num_sub = ceil_div(Dd, BK); gl.splits = 1; gl.stream = as_stream(stream);
    int rc = make_tmap(&gl.ta, hN, 2, (uint64_t)Dd, (uint64_t)M, 0, (uint64_t)Dd, BK, BM);
    if (rc) return rc;
    rc = make_tmap(&gl.tb, W, 2, (uint64_t)Dd, (uint64_t)P, 0, (uint64_t)Dd, BK, (uint32_t)tile_n);
    if (rc) return rc;
    rc = make_tmap(&gl.to0, pred_bf16, 2, (uint64_t)P, (uint64_t)M, 1, (uint64_t)P, 64, 32);
    if (rc) return rc;
    rc = make_tmap(&gl.to1, g_bf16, 2, (uint64_t)P, (uint64_t)M, 1, (uint64_t)P, 64, 32);
    if (rc) return rc;
    memset(&gl.ep, 0, sizeof(gl.ep));
    gl.ep.alpha = 1.0f; gl.ep.bias = bias;
    gl.ep.mse_vol = vol; gl.ep.mse_mask = mask; gl.ep.mse_part = partials;
    gl.ep.mse_V = V; gl.ep.mse_p = p; gl.ep.mse_g = g; gl.ep.mse_L = L;
    gl.ep.mse_coef = 2.0f / (static_cast<float>(P) * mask_sum);
    const long long ctas = static_cast<long long>(ceil_div(M, BM)) * ceil_div(P, tile_n);
    const int eff = dispatch_tile<false, false, EPI_PRED_MSE>(tile_n, ctas <= 148, gl);
    return eff < 0 ? eff : 0;
}

extern "C" int vitae_pred_mse_finalize(const float* partials, long long n, int P, float mask_sum, float* loss_out, void* stream) {
    VITAE_REQUIRE(partials && loss_out && n > 0 && P > 0 && mask_sum > 0.f, "pred_mse_finalize: bad arguments");
    VITAE_REQUIRE((reinterpret_cast<uintptr_t>(partials) & 15) == 0, "pred_mse_finalize: partials must be 16-byte aligned");
    launch_kernel(pred_mse_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), partials, n,
                  1.0f / (static_cast<float>(P) * mask_sum), mask_sum, loss_out);
    VITAE_CHECK_LAUNCH("pred_mse_finalize");
    return 0;
}

// Times every (tile_n, split_k) candidate for this GEMM on the device and returns the fastest.  The GEMMs of the training
// step are short and latency / ingest bound, which makes the best tiling a property of the exact shape -- and of the cache
// state: in the step every weight matrix is read once per pass and comes from HBM, so a candidate is timed COLD when a
// flush buffer is given (>= 2x L2, overwritten before every timed launch; CUDA events around each launch): a tiling with a
// handful of CTAs looks best on L2-resident operands and is the worst one on cold weights.  Without a flush buffer the
// launches run back to back (warm).  Candidates whose slabs do not fit into `workspace` are skipped.  The outputs are
// overwritten repeatedly with the same values (do not use with accumulate).  Synchronises `stream`.
extern "C" int vitae_gemm_autotune(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M,
                                   int N, int K, const vitae_gemm_epilogue* e, void* workspace, size_t workspace_bytes,
                                   void* flush_buf, size_t flush_bytes, void* stream, int* best_tile_n, int* best_split_k) {
    VITAE_REQUIRE(e && best_tile_n && best_split_k, "gemm_autotune: null argument");
    VITAE_REQUIRE(!e->accumulate, "gemm_autotune: accumulate epilogues cannot be re-run");
    cudaStream_t st = as_stream(stream);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    VITAE_REQUIRE(cap == cudaStreamCaptureStatusNone, "gemm_autotune: stream is capturing");
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)
        return set_error(-3, "gemm_autotune: cudaEventCreate failed");
    const int num_sub = ceil_div(K, BK), tiles_m = ceil_div(M, BM);
    const int tiles_n_opts[3] = {64, 128, 256};
    const int split_opts[6] = {1, 2, 3, 4, 6, 8};
    const bool cold = flush_buf != nullptr && flush_bytes > 0;
    float best = 1e30f;
    int rc = 0;
    *best_tile_n = 0;
    *best_split_k = 1;
    for (int ti = 0; ti < 3 && rc == 0; ++ti) {
        const int tn = tiles_n_opts[ti];
        if (tn > 64 && N < tn) continue;
        const long long tiles = static_cast<long long>(tiles_m) * ceil_div(N, tn);
        for (int si = 0; si < 6 && rc == 0; ++si) {
            const int sk = split_opts[si];
            if (sk > 1 && (num_sub < 4 * sk || tiles * sk > 3 * 148)) continue;
            if (vitae_gemm_workspace_bytes_for(e, a_mn_major, b_mn_major, M, N, sk) > workspace_bytes) continue;
            for (int w = 0; w < 2 && rc == 0; ++w)
                rc = vitae_gemm_bf16(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, e, tn, sk, workspace, workspace_bytes, stream);
            float t = 1e30f;
            if (cold) {
                constexpr int REPS = 6;
                float sum = 0.f;
                for (int r = 0; r < REPS && rc == 0; ++r) {
                    cudaMemsetAsync(flush_buf, r & 0xff, flush_bytes, st);      // evicts operands and outputs from L2
                    cudaEventRecord(e0, st);
                    rc = vitae_gemm_bf16(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, e, tn, sk, workspace, workspace_bytes, stream);
                    cudaEventRecord(e1, st);
                    if (cudaEventSynchronize(e1) != cudaSuccess) rc = set_error(-3, "gemm_autotune: %s", cudaGetErrorString(cudaGetLastError()));
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (r > 0) sum += ms;                                        // the first cold run also warms the instruction cache
                }
                t = sum / (REPS - 1);
            } else {
                constexpr int REPS = 12;
                for (int trial = 0; trial < 2 && rc == 0; ++trial) {
                    cudaEventRecord(e0, st);
                    for (int r = 0; r < REPS && rc == 0; ++r)
                        rc = vitae_gemm_bf16(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, e, tn, sk, workspace, workspace_bytes, stream);
                    cudaEventRecord(e1, st);
                    if (cudaEventSynchronize(e1) != cudaSuccess) rc = set_error(-3, "gemm_autotune: %s", cudaGetErrorString(cudaGetLastError()));
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, e0, e1);
                    t = ms < t ? ms : t;
                }
            }
            if (rc == 0 && t < best) {
                best = t;
                *best_tile_n = tn;
                *best_split_k = sk;
            }
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc == 0 && *best_tile_n == 0) return set_error(-2, "gemm_autotune: no candidate fits the workspace");
    return rc;
}
