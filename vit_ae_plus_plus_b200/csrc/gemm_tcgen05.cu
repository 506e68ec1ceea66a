// bf16 x bf16 -> fp32 GEMM on the Blackwell 5th-gen tensor cores.
//
//   * operands staged by TMA (cp.async.bulk.tensor.2d) into 128B-swizzled shared-memory tiles, STAGES-deep
//     mbarrier ring (full/empty), one producer thread;
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) issued by one elected thread, accumulator in TMEM;
//   * four epilogue warps read the accumulator with tcgen05.ld (32x32b.x32) and apply the fused epilogue
//     (alpha, bias, residual / positional addend with row gather, GELU', GELU, row-scattered fp32 / bf16 stores);
//   * both operand majors (K-major = nn.Linear layout, MN-major = transposed) so that forward, dgrad and wgrad of
//     every Linear on the path (see include/vitae_b200.h) are the same kernel with different tensor maps;
//   * optional split-K: each split writes its fp32 partial tile to a slab, a finalize kernel reduces the slabs in
//     fixed order (deterministic) and applies the epilogue.
#include <mutex>
#include <unordered_map>

#include "common.h"
#include "ptx.cuh"

namespace vitae {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue

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

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Epilogue for 8 consecutive columns [n, n+8) of logical row m (all in range; N % 8 == 0).
__device__ __forceinline__ void epilogue_group8(const EpiParams& ep, float alpha, int m, int n, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= alpha;
    if (ep.bias) {
        const float4 b0 = *reinterpret_cast<const float4*>(ep.bias + n);
        const float4 b1 = *reinterpret_cast<const float4*>(ep.bias + n + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (ep.addend) {
        const int ra = ep.add_rows ? ep.add_rows[m] : m;
        const float* a = ep.addend + static_cast<size_t>(ra) * ep.ldadd + n;
        const float4 a0 = *reinterpret_cast<const float4*>(a);
        const float4 a1 = *reinterpret_cast<const float4*>(a + 4);
        v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
        v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
    }
    if (ep.dgelu_src) {
        const uint4 raw = *reinterpret_cast<const uint4*>(ep.dgelu_src + static_cast<size_t>(m) * ep.ld_dgelu + n);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            v[2 * i] *= dgelu_erf(f.x);
            v[2 * i + 1] *= dgelu_erf(f.y);
        }
    }
    const int r = ep.out_rows ? ep.out_rows[m] : m;
    if (ep.out_f32) {
        float* o = ep.out_f32 + static_cast<size_t>(r) * ep.ld_f32 + n;
        float4 o0 = make_float4(v[0], v[1], v[2], v[3]);
        float4 o1 = make_float4(v[4], v[5], v[6], v[7]);
        if (ep.accumulate) {
            const float4 p0 = *reinterpret_cast<const float4*>(o);
            const float4 p1 = *reinterpret_cast<const float4*>(o + 4);
            o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
            o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
        }
        *reinterpret_cast<float4*>(o) = o0;
        *reinterpret_cast<float4*>(o + 4) = o1;
    }
    if (ep.out_bf16) {
        uint4 pk;
        pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]);
        pk.z = pack_bf16(v[4], v[5]); pk.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(ep.out_bf16 + static_cast<size_t>(r) * ep.ld_bf16 + n) = pk;
    }
    if (ep.out_gelu_bf16) {
        uint4 pk;
        pk.x = pack_bf16(gelu_erf(v[0]), gelu_erf(v[1])); pk.y = pack_bf16(gelu_erf(v[2]), gelu_erf(v[3]));
        pk.z = pack_bf16(gelu_erf(v[4]), gelu_erf(v[5])); pk.w = pack_bf16(gelu_erf(v[6]), gelu_erf(v[7]));
        *reinterpret_cast<uint4*>(ep.out_gelu_bf16 + static_cast<size_t>(r) * ep.ld_bf16 + n) = pk;
    }
}

// Phase tracing (debug builds only: VITAE_TRACE=1 python -m vit_ae_plus_plus_b200.build): per CTA, SM-clock stamps of the
// kernel phases are written to a device buffer registered with vitae_debug_set_gemm_trace (tools/gemm_trace.py).
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

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                         int num_kb, int kb_per_split, EpiParams ep, float* __restrict__ slabs) {
    using S = GemmSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* sm = smem_raw + (base - raw_u32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::BAR_OFFSET);
    const uint32_t full_bar = base + S::BAR_OFFSET;
    const uint32_t empty_bar = full_bar + STAGES * 8;
    const uint32_t tmem_full_bar = empty_bar + STAGES * 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int kb_begin = blockIdx.z * kb_per_split;
    const int kb_end = min(num_kb, kb_begin + kb_per_split);
    const int nkb = kb_end - kb_begin;
    if (threadIdx.x == 0) GEMM_TRACE(0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s * 8, 1);
            mbar_init(empty_bar + s * 8, 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // PDL: this CTA now holds everything it will ever acquire (smem, TMEM columns), so the next kernel in the stream may
    // start its own prologue; our first global access (TMA loads, epilogue operands) waits for the predecessor grid.
    if (threadIdx.x == 0) GEMM_TRACE(1);
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) GEMM_TRACE(2);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(empty_bar + s * 8, ph ^ 1);
                mbar_arrive_expect_tx(full_bar + s * 8, S::STAGE_BYTES);
                const uint32_t sa = base + s * S::STAGE_BYTES;
                const uint32_t sb = sa + S::A_BYTES;
                const int k0 = (kb_begin + i) * BK;
                if (A_MN) {
#pragma unroll
                    for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, full_bar + s * 8, m0 + 64 * j, k0);
                } else {
                    tma_load_2d(sa, &tmA, full_bar + s * 8, k0, m0);
                }
                if (B_MN) {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, full_bar + s * 8, n0 + 64 * j, k0);
                } else {
                    tma_load_2d(sb, &tmB, full_bar + s * 8, k0, n0);
                }
                if (i == 0) GEMM_TRACE(3);
            }
            GEMM_TRACE(4);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (single thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(full_bar + s * 8, ph);
                tc_fence_after();
                if (i == 0) GEMM_TRACE(5);
                const uint32_t sa = base + s * S::STAGE_BYTES;
                const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                    // K-major: 16 bf16 = 32 B along the swizzled row; MN-major: 16 K-rows of 128 B
                    const uint64_t da = A_MN ? umma_smem_desc_sw128(sa + kk * 2048, BK * 128, 1024)
                                             : umma_smem_desc_sw128(sa + kk * 32, 16, 1024);
                    const uint64_t db = B_MN ? umma_smem_desc_sw128(sb + kk * 2048, BK * 128, 1024)
                                             : umma_smem_desc_sw128(sb + kk * 32, 16, 1024);
                    umma_bf16(tmem_base, da, db, idesc, (i | kk) != 0 ? 1u : 0u);
                }
                umma_commit(empty_bar + s * 8);  // frees the smem stage once these MMAs retire
            }
            umma_commit(tmem_full_bar);
            GEMM_TRACE(6);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps (TMEM -> regs -> HBM)
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 64) GEMM_TRACE(7);
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const int m = m0 + q * 32 + lane;
        const float alpha = ep.alpha * (ep.alpha_ptr ? *ep.alpha_ptr : 1.0f);
        float* slab = slabs ? slabs + static_cast<size_t>(blockIdx.z) * M * N : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            if (n0 + c >= N) break;
            uint32_t raw[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, raw);
            tmem_ld_wait();
            if (m < M) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int n = n0 + c + g * 8;
                    if (n < N) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[g * 8 + i]);
                        if (slab) {
                            float* o = slab + static_cast<size_t>(m) * N + n;
                            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                        } else {
                            epilogue_group8(ep, alpha, m, n, v);
                        }
                    }
                }
            }
        }
        tc_fence_before();
        if (threadIdx.x == 64) GEMM_TRACE(8);
    }
    __syncthreads();
    if (threadIdx.x == 0) GEMM_TRACE(9);
    if (warp == 2) tmem_dealloc(tmem_base, BN);
}

// Split-K finalize: sum the slabs in fixed order, then the fused epilogue.
__global__ void gemm_splitk_finalize_kernel(const float* __restrict__ slabs, int splits, int M, int N, EpiParams ep) {
    if (threadIdx.x == 0) GEMM_TRACE(1);
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) GEMM_TRACE(2);
    const int groups_per_row = N / 8;
    const long long total = static_cast<long long>(M) * groups_per_row;
    const float alpha = ep.alpha * (ep.alpha_ptr ? *ep.alpha_ptr : 1.0f);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int m = static_cast<int>(idx / groups_per_row);
        const int n = static_cast<int>(idx % groups_per_row) * 8;
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int s = 0; s < splits; ++s) {
            const float* p = slabs + (static_cast<size_t>(s) * M + m) * N + n;
            const float4 a = *reinterpret_cast<const float4*>(p);
            const float4 b = *reinterpret_cast<const float4*>(p + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
            v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        epilogue_group8(ep, alpha, m, n, v);
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
    uint64_t d0, d1, ld;
    uint32_t b0, b1;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.d0); mix(k.d1); mix(k.ld); mix(k.b0); mix(k.b1);
        return h;
    }
};

// 2-D bf16 tensor map: inner dim d0 (contiguous), outer dim d1 with row pitch ld elements, box b0 x b1, 128B swizzle,
// out-of-bounds elements read as zero (so ragged M/N/K need no special casing in the kernel).
static int make_tmap(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    const TmapKey key{ptr, d0, d1, ld, b0, b1};
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
    const cuuint64_t dims[2] = {d0, d1};
    const cuuint64_t strides[1] = {ld * 2};
    const cuuint32_t box[2] = {b0, b1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(-5, "cuTensorMapEncodeTiled failed (%d) ptr=%p dims=%llu,%llu ld=%llu box=%u,%u", (int)r, ptr,
                         (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int num_kb, int splits, const EpiParams& ep,
                       float* slabs, cudaStream_t stream) {
    using S = GemmSmem<BN, STAGES>;
    auto kern = gemm_bf16_tcgen05_kernel<BN, STAGES, A_MN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(smem=%d): %s", S::TOTAL, cudaGetErrorString(e));
        attr_set = true;
    }
    const int kb_per_split = ceil_div(num_kb, splits);
    const int eff_splits = ceil_div(num_kb, kb_per_split);  // every z-slice gets >= 1 k-block
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), eff_splits);
    launch_kernel(kern, dim3(grid), dim3(GEMM_THREADS), S::TOTAL, stream, ta, tb, M, N, num_kb, kb_per_split, ep, slabs);
    VITAE_CHECK_LAUNCH("gemm_bf16_tcgen05");
    if (slabs) {
        const long long total = static_cast<long long>(M) * (N / 8);
        const int blocks = static_cast<int>(std::min<long long>(ceil_div<long long>(total, 256), 148 * 8));
        launch_kernel(gemm_splitk_finalize_kernel, dim3(blocks), dim3(256), 0, stream, slabs, eff_splits, M, N, ep);
        VITAE_CHECK_LAUNCH("gemm_splitk_finalize");
    }
    return 0;
}

template <int BN, int STAGES>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int num_kb,
                          int splits, const EpiParams& ep, float* slabs, cudaStream_t stream) {
    if (!a_mn && !b_mn) return launch_gemm<BN, STAGES, false, false>(ta, tb, M, N, num_kb, splits, ep, slabs, stream);
    if (!a_mn && b_mn) return launch_gemm<BN, STAGES, false, true>(ta, tb, M, N, num_kb, splits, ep, slabs, stream);
    if (a_mn && b_mn) return launch_gemm<BN, STAGES, true, true>(ta, tb, M, N, num_kb, splits, ep, slabs, stream);
    return launch_gemm<BN, STAGES, true, false>(ta, tb, M, N, num_kb, splits, ep, slabs, stream);
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

    int bn = tile_n;
    if (bn == 0) {
        const long long tiles128 = static_cast<long long>(ceil_div(M, BM)) * ceil_div(N, 128) * (split_k > 1 ? split_k : 1);
        bn = (tiles128 < 120) ? 64 : 128;
    }
    VITAE_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: tile_n must be 0/64/128/256 (got %d)", tile_n);
    const int num_kb = ceil_div(K, BK);
    int splits = split_k > 1 ? split_k : 1;
    if (splits > num_kb) splits = num_kb;
    float* slabs = nullptr;
    if (splits > 1) {
        VITAE_REQUIRE(workspace && workspace_bytes >= vitae_gemm_workspace_bytes(M, N, splits),
                      "gemm: split_k=%d needs %zu workspace bytes, got %zu", splits, vitae_gemm_workspace_bytes(M, N, splits),
                      workspace_bytes);
        slabs = static_cast<float*>(workspace);
    }

    CUtensorMap ta, tb;
    int rc;
    if (a_mn_major) rc = make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
    else            rc = make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
    if (rc) return rc;
    if (b_mn_major) rc = make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
    else            rc = make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)bn);
    if (rc) return rc;

    EpiParams ep;
    ep.alpha = e->alpha; ep.alpha_ptr = e->alpha_ptr; ep.bias = e->bias; ep.addend = e->addend;
    ep.add_rows = e->add_rows; ep.ldadd = e->ldadd;
    ep.dgelu_src = static_cast<const __nv_bfloat16*>(e->dgelu_src); ep.ld_dgelu = e->ld_dgelu;
    ep.out_f32 = e->out_f32; ep.ld_f32 = e->ld_f32; ep.accumulate = e->accumulate;
    ep.out_bf16 = static_cast<__nv_bfloat16*>(e->out_bf16);
    ep.out_gelu_bf16 = static_cast<__nv_bfloat16*>(e->out_gelu_bf16); ep.ld_bf16 = e->ld_bf16;
    ep.out_rows = e->out_rows;

    cudaStream_t st = as_stream(stream);
    const bool amn = a_mn_major != 0, bmn = b_mn_major != 0;
    if (bn == 64) return dispatch_major<64, 4>(amn, bmn, ta, tb, M, N, num_kb, splits, ep, slabs, st);
    if (bn == 128) return dispatch_major<128, 3>(amn, bmn, ta, tb, M, N, num_kb, splits, ep, slabs, st);
    return dispatch_major<256, 4>(amn, bmn, ta, tb, M, N, num_kb, splits, ep, slabs, st);
}
