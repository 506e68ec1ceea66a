// Fused multi-head self-attention on the Blackwell tensor cores: softmax(q k^T * scale) v and its backward with the
// score tiles in tensor memory (model/vit.py:112-121 materialises a [B,H,N,N] fp32 score tensor).
//
// Layouts (unchanged C ABI, include/vitae_b200.h): qkv bf16 [B, N, 3, H, hd] = the qkv Linear's output as is,
// out / dout bf16 [B, N, H*hd], lse / delta fp32 [B, H, N].  hd in {16, 32, 64}; N ragged.
//
// One CTA = one 128-row tile of one (batch, head): 6 warps.
//   warp 0     TMA producer.  Operand tiles are 64-row x 128-byte boxes of a 3-D tensor map (columns, tokens, batch) in
//              SWIZZLE_128B form; rows past a sample's N read as zero.  A head narrower than 64 columns shares its box
//              with its neighbours: the MMA descriptors start (h*hd % 64) * 2 bytes into the swizzled row (K-major
//              operands), or the MMA runs over all 64 columns and the epilogue keeps this head's (MN-major operands).
//   warp 1     one thread issues tcgen05.mma (M = 128, N = 64 -- or the valid part of the last tile rounded up to 16 --,
//              K = 16 per instruction), accumulators in TMEM, completion through tcgen05.commit -> mbarrier.
//   warps 2-5  softmax: each thread owns one row (TMEM lane), reads its scores with tcgen05.ld -- no shuffles --, and
//              writes P (bf16) into a swizzled shared-memory tile that is the A operand of the next MMA.
//
// forward   two passes over the kv tiles (the scores of a 128 x 513 tile do not fit the 256 TMEM columns a CTA gets when
//           two CTAs share an SM, and recomputing q k^T costs ~2 % of the softmax time): pass 1 row maxima (FMNMX3),
//           pass 2 p = ex2(s * scale*log2e - m) (one FFMA + one MUFU per score), row sums, O += P V accumulated in TMEM
//           with no rescaling.  The column mask exists on the last kv tile only.
// backward  dQ kernel (CTA per query tile: S and dP = dO V^T in TMEM, dS -> shared memory, dQ += dS K accumulated in
//           TMEM; also writes delta = rowsum(dO * O)) and dK/dV kernel (CTA per kv tile: S^T and dP^T in TMEM,
//           P^T / dS^T -> shared memory, dV += P^T dO and dK += dS^T Q in TMEM).  No atomics: deterministic.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vitae {

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;
constexpr int AT_THREADS = 192;
constexpr int QT = 128;                 // rows per CTA (UMMA M)
constexpr int KT = 64;                  // columns per score tile
constexpr uint32_t BOX_BYTES = 64 * 128;   // one 64-row x 64-column bf16 box
constexpr uint32_t ROWT_BYTES = 2 * BOX_BYTES;   // a 128-row operand tile = two boxes

// Phase tracing (debug build: VITAE_ATTN_TRACE=1): per CTA, %globaltimer stamps and accumulated barrier-wait cycles go to a
// device buffer registered with vitae_debug_set_attn_trace (tools/attn_trace.py).  Compiled out by default.
#ifdef VITAE_ATTN_TRACE
__device__ unsigned long long* g_attn_trace = nullptr;
#define AT_CTA() ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x)
#define AT_STAMP(slot)                                                                        \
    do {                                                                                      \
        if (g_attn_trace) {                                                                   \
            unsigned long long t__;                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                           \
            g_attn_trace[AT_CTA() * 16 + (slot)] = t__;                                       \
        }                                                                                     \
    } while (0)
#define AT_PUT(slot, v) do { if (g_attn_trace) g_attn_trace[AT_CTA() * 16 + (slot)] = (v); } while (0)
#define AT_WAIT(acc, bar, par) do { const long long c0__ = clock64(); mbar_wait(bar, par); (acc) += clock64() - c0__; } while (0)
// event log of CTA 0: who = 0 MMA thread, 1 one softmax thread, 2 TMA thread; 256 (id, clock) pairs each, behind the per-CTA slots
#define AT_EV(who, id)                                                                                    \
    do {                                                                                                  \
        if (g_attn_trace && AT_CTA() == 0 && ev_n__ < 256) {                                              \
            unsigned long long* e__ = g_attn_trace + (1 << 18) + ((who) * 256 + ev_n__) * 2;             \
            e__[0] = (id);                                                                                \
            e__[1] = clock64();                                                                           \
            ++ev_n__;                                                                                     \
        }                                                                                                 \
    } while (0)
#define AT_EV_DECL int ev_n__ = 0
#else
#define AT_EV(who, id) do { } while (0)
#define AT_EV_DECL do { } while (0)
#define AT_STAMP(slot) do { } while (0)
#define AT_PUT(slot, v) do { } while (0)
#define AT_WAIT(acc, bar, par) do { (void)(acc); mbar_wait(bar, par); } while (0)
#endif

__device__ __forceinline__ uint32_t at_swz(int r, int j) { return static_cast<uint32_t>(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ void at_sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// K-major operand (rows x 64 bf16 in 128-byte swizzled rows), 16-element K step kk, starting sub_bytes into the row
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, uint32_t sub_bytes, int kk) {
    return umma_smem_desc_sw128(tile + sub_bytes + kk * 32, 16, 1024);
}
// MN-major operand (K rows x 64 contiguous bf16): K step kk = 16 rows of 128 bytes
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int kk) {
    return umma_smem_desc_sw128(tile + kk * 2048, BOX_BYTES, 1024);
}

// ------------------------------------------------------------------------------------------------------ forward
// CFG 0: two CTAs per SM (256 TMEM columns each: two score tiles + O; two P buffers; 6 kv slots; 97 KB).
// CFG 1: three CTAs per SM (128 TMEM columns: one score tile + O; one P buffer; 4 kv slots; 65 KB) -- every buffer single,
//        the latency of one CTA's MMA / barrier round trips is covered by the other two; 444 CTA slots hold the 320 CTAs of
//        the decoder shape (64 heads x 5 query tiles, one of them a single-row tail) in ONE wave, 296 do not.
template <int CFG>
struct FwdSmem {
    static constexpr int NSLOT = CFG == 0 ? 6 : 4;     // ring of kv boxes
    static constexpr int NS = CFG == 0 ? 2 : 1;        // score tiles in TMEM
    static constexpr int NP = CFG == 0 ? 2 : 1;        // P tiles in shared memory
    static constexpr uint32_t Q = 0;
    static constexpr uint32_t KV = Q + ROWT_BYTES;
    static constexpr uint32_t P = KV + NSLOT * BOX_BYTES;
    static constexpr uint32_t BAR = P + NP * ROWT_BYTES;
    // barriers: q_full, kv_full[NSLOT], kv_empty[NSLOT], s_full[NS], s_free[NS], p_full[NP], p_free[NP], o_full
    static constexpr int NBAR = 1 + 2 * NSLOT + 2 * NS + 2 * NP + 1;
    static constexpr uint32_t TOTAL = BAR + NBAR * 8 + 16 + 1024;
    static constexpr uint32_t O_COL = NS * KT;                          // O accumulator behind the score tiles
    static constexpr uint32_t TMEM_COLS = CFG == 0 ? 256 : 128;
};

template <int HD, int CFG>
__global__ void __launch_bounds__(AT_THREADS, CFG == 0 ? 2 : 3)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N,
                   int H, float scale_log2) {
    using S = FwdSmem<CFG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_u32);
    const uint32_t q_full = base + S::BAR;
    const uint32_t kv_full = q_full + 8, kv_empty = kv_full + S::NSLOT * 8;
    const uint32_t s_full = kv_empty + S::NSLOT * 8, s_free = s_full + S::NS * 8;
    const uint32_t p_full = s_free + S::NS * 8, p_free = p_full + S::NP * 8, o_full = p_free + S::NP * 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + S::BAR + S::NBAR * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * QT;
    const int D = H * HD;
    const int T = ceil_div(N, KT);
    const int nv_last = N - (T - 1) * KT;                   // valid columns of the last kv tile
    const int n16_last = (nv_last + 15) & ~15;
    const int qcol = h * HD, kcol = D + h * HD, vcol = 2 * D + h * HD;

    if (threadIdx.x == 0) AT_STAMP(0);
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < S::NSLOT; ++i) { mbar_init(kv_full + i * 8, 1); mbar_init(kv_empty + i * 8, 1); }
        for (int i = 0; i < S::NS; ++i) { mbar_init(s_full + i * 8, 1); mbar_init(s_free + i * 8, 128); }
        for (int i = 0; i < S::NP; ++i) { mbar_init(p_full + i * 8, 128); mbar_init(p_free + i * 8, 1); }
        mbar_init(o_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), S::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PDL_TRIGGER_EARLY();
    pdl_wait();
    if (threadIdx.x == 0) {
        AT_STAMP(1);
        unsigned smid__;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid__));
        AT_PUT(9, smid__);
    }

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, ROWT_BYTES);
            tma_load_3d(base + S::Q, &tmQKV, q_full, qcol & ~63, q0, b);
            tma_load_3d(base + S::Q + BOX_BYTES, &tmQKV, q_full, qcol & ~63, q0 + 64, b);
            uint32_t seq = 0;
            AT_EV_DECL;
            auto load = [&](int col, int t) {
                const uint32_t slot = seq % S::NSLOT;
                mbar_wait(kv_empty + slot * 8, ((seq / S::NSLOT) & 1) ^ 1);
                AT_EV(2, 100 + seq);
                mbar_arrive_expect_tx(kv_full + slot * 8, BOX_BYTES);
                tma_load_3d(base + S::KV + slot * BOX_BYTES, &tmQKV, kv_full + slot * 8, col & ~63, t * KT, b);
                ++seq;
            };
            // the order the MMA warp consumes them in: K tiles of pass 1, then K_t / V_(t-1) interleaved
            for (int t = 0; t < T; ++t) load(kcol, t);
            for (int t = 0; t < T; ++t) {
                load(kcol, t);
                if (t >= 1) load(vcol, t - 1);
            }
            load(vcol, T - 1);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            const uint32_t subq = (qcol & 63) * 2, subk = (kcol & 63) * 2;
            constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, 64, 0u, 1u);
            uint32_t seq = 0;
            long long w_kv = 0, w_sfree = 0, w_pfull = 0;
            AT_EV_DECL;
            auto kv_wait = [&]() -> uint32_t {
                const uint32_t slot = seq % S::NSLOT;
                AT_WAIT(w_kv, kv_full + slot * 8, (seq / S::NSLOT) & 1);
                AT_EV(0, 100 + seq);          // kv tile seq landed
                ++seq;
                return slot;
            };
            auto issue_s = [&](int it, int t) {          // S tile number `it` (both passes counted) = Q K_t^T
                const uint32_t slot = kv_wait();
                const int sb = it % S::NS;
                AT_WAIT(w_sfree, s_free + sb * 8, ((it / S::NS) & 1) ^ 1);
                AT_EV(0, 200 + it);           // score buffer free
                tc_fence_after();
                const uint32_t idesc = umma_idesc_bf16(QT, t == T - 1 ? n16_last : KT, 0u, 0u);
                const uint32_t sk = base + S::KV + slot * BOX_BYTES;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)
                    umma_bf16(tmem_base + sb * KT, desc_kmajor(base + S::Q, subq, kk), desc_kmajor(sk, subk, kk), idesc, kk != 0);
                AT_EV(0, 300 + it);           // S MMAs issued
                umma_commit(kv_empty + slot * 8);
                umma_commit(s_full + sb * 8);
                AT_EV(0, 400 + it);           // commits issued
            };
            auto issue_pv = [&](int u) {                 // O += P_u V_u
                const uint32_t slot = kv_wait();
                const int pb = u % S::NP;
                AT_WAIT(w_pfull, p_full + pb * 8, (u / S::NP) & 1);
                AT_EV(0, 500 + u);            // P tile ready
                tc_fence_after();
                const int nk = (u == T - 1 ? n16_last : KT) / 16;
                const uint32_t sv = base + S::KV + slot * BOX_BYTES, sp = base + S::P + pb * ROWT_BYTES;
#pragma unroll 1
                for (int kk = 0; kk < nk; ++kk)
                    umma_bf16(tmem_base + S::O_COL, desc_kmajor(sp, 0, kk), desc_mnmajor(sv, kk), idesc_pv, (u | kk) != 0);
                umma_commit(kv_empty + slot * 8);
                umma_commit(p_free + pb * 8);
                AT_EV(0, 600 + u);            // P V issued + committed
            };
            mbar_wait(q_full, 0);
            AT_STAMP(2);
            for (int t = 0; t < T; ++t) issue_s(t, t);
            AT_STAMP(3);
            for (int t = 0; t < T; ++t) {
                issue_s(T + t, t);
                if (t >= 1) issue_pv(t - 1);
            }
            issue_pv(T - 1);
            umma_commit(o_full);
            AT_STAMP(5);
            AT_PUT(13, w_sfree); AT_PUT(14, w_pfull); AT_PUT(15, w_kv);
            PDL_TRIGGER_LATE();
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- softmax warps: one row per thread
        const int q = warp & 3;                              // TMEM lane quadrant of this warp
        const int row = q * 32 + lane;
        const int grow = q0 + row;
        const bool warp_valid = q0 + q * 32 < N;             // warps without a valid row only keep the barriers going
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float mraw = -INFINITY;
        long long w_s1 = 0, w_s2 = 0, w_pf = 0;
        AT_EV_DECL;
        for (int t = 0; t < T; ++t) {                        // pass 1: row maxima of the raw scores
            const int sb = t % S::NS;
            AT_WAIT(w_s1, s_full + sb * 8, (t / S::NS) & 1);
            if (threadIdx.x == 64) AT_EV(1, 100 + t);        // scores of tile t visible
            tc_fence_after();
            if (warp_valid) {
                const bool last = t == T - 1;
                const int nch = last ? n16_last / 16 : 4;
                uint32_t v[4][16];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < nch) tmem_ld_32x16(t_lane + sb * KT + c * 16, v[c]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c >= nch) continue;
                    if (last && (c + 1) * 16 > nv_last) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c * 16 + j < nv_last) mraw = fmaxf(mraw, __uint_as_float(v[c][j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            mraw = fmax3(mraw, __uint_as_float(v[c][j]), __uint_as_float(v[c][j + 1]));
                    }
                }
            }
            if (threadIdx.x == 64) AT_EV(1, 200 + t);        // maxima done
            tc_fence_before();
            mbar_arrive(s_free + sb * 8);
        }
        if (threadIdx.x == 64) AT_STAMP(4);
        const float msc = mraw * scale_log2;
        float l = 0.f;
        for (int t = 0; t < T; ++t) {                        // pass 2: p = 2^(s * scale_log2 - msc), row sums, P -> smem
            const int it = T + t, sb = it % S::NS, pb = t % S::NP;
            const bool last = t == T - 1;
            const int nch = last ? n16_last / 16 : 4;
            AT_WAIT(w_s2, s_full + sb * 8, (it / S::NS) & 1);
            if (threadIdx.x == 64) AT_EV(1, 300 + t);        // pass 2: scores visible
            tc_fence_after();
            uint32_t v[4][16];
            if (warp_valid) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < nch) tmem_ld_32x16(t_lane + sb * KT + c * 16, v[c]);
                tmem_ld_wait();
            }
            if (threadIdx.x == 64) AT_EV(1, 400 + t);        // scores in registers
            tc_fence_before();
            mbar_arrive(s_free + sb * 8);                    // the scores are in registers: the MMA warp may reuse the tile
            uint32_t pk[4][8];                               // P as packed bf16 pairs, computed before the buffer is needed
            if (warp_valid) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c >= nch) continue;
                    float p[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        p[j] = ex2_approx(fmaf(__uint_as_float(v[c][j]), scale_log2, -msc));
                        if (last && c * 16 + j >= nv_last) p[j] = 0.f;
                        l += p[j];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) pk[c][j] = pack_bf16(p[2 * j], p[2 * j + 1]);
                }
            }
            if (threadIdx.x == 64) AT_EV(1, 500 + t);        // exponentials done
            AT_WAIT(w_pf, p_free + pb * 8, ((t / S::NP) & 1) ^ 1);   // the P V that read this buffer last has completed
            if (threadIdx.x == 64) AT_EV(1, 600 + t);        // P buffer free
            if (warp_valid) {
                const uint32_t sp = base + S::P + pb * ROWT_BYTES;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c >= nch) continue;
                    at_sts128(sp + at_swz(row, 2 * c), pk[c][0], pk[c][1], pk[c][2], pk[c][3]);
                    at_sts128(sp + at_swz(row, 2 * c + 1), pk[c][4], pk[c][5], pk[c][6], pk[c][7]);
                }
            }
            fence_proxy_async();
            mbar_arrive(p_full + pb * 8);
            if (threadIdx.x == 64) AT_EV(1, 700 + t);        // P stored, fenced, signalled
        }
        if (threadIdx.x == 64) AT_STAMP(6);
        mbar_wait(o_full, 0);
        tc_fence_after();
        if (threadIdx.x == 64) {
            AT_STAMP(7);
            AT_PUT(10, w_s1); AT_PUT(11, w_s2); AT_PUT(12, w_pf);
        }
        if (warp_valid) {
            const int subv = vcol & 63;
            uint32_t o[HD / 16][16];
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) tmem_ld_32x16(t_lane + S::O_COL + subv + c * 16, o[c]);
            tmem_ld_wait();
            if (grow < N) {
                const float inv = 1.f / l;
                __nv_bfloat16* orow = out + (static_cast<size_t>(b) * N + grow) * D + h * HD;
#pragma unroll
                for (int c = 0; c < HD / 16; ++c) {
                    uint4 lo, hi;
                    lo.x = pack_bf16(__uint_as_float(o[c][0]) * inv, __uint_as_float(o[c][1]) * inv);
                    lo.y = pack_bf16(__uint_as_float(o[c][2]) * inv, __uint_as_float(o[c][3]) * inv);
                    lo.z = pack_bf16(__uint_as_float(o[c][4]) * inv, __uint_as_float(o[c][5]) * inv);
                    lo.w = pack_bf16(__uint_as_float(o[c][6]) * inv, __uint_as_float(o[c][7]) * inv);
                    hi.x = pack_bf16(__uint_as_float(o[c][8]) * inv, __uint_as_float(o[c][9]) * inv);
                    hi.y = pack_bf16(__uint_as_float(o[c][10]) * inv, __uint_as_float(o[c][11]) * inv);
                    hi.z = pack_bf16(__uint_as_float(o[c][12]) * inv, __uint_as_float(o[c][13]) * inv);
                    hi.w = pack_bf16(__uint_as_float(o[c][14]) * inv, __uint_as_float(o[c][15]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 16) = lo;
                    *reinterpret_cast<uint4*>(orow + c * 16 + 8) = hi;
                }
                lse[(static_cast<size_t>(b) * H + h) * N + grow] = (msc + log2f(l)) * LN2_F;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) AT_STAMP(8);
    if (warp == 2) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------ forward, wide
// Second forward layout (default): score tiles of 128 columns (5 instead of 9 barrier round trips per pass at N = 513) and
// EIGHT softmax warps -- two per TMEM lane quadrant, each taking one 64-column half of every tile -- so that every SM
// sub-partition interleaves four warps when two CTAs share the SM (measured with the one-warp-per-quadrant layout above:
// a softmax warp spends 1250 cycles per 64-column tile against a 512-cycle MUFU floor, all of it serial latency --
// tcgen05.ld round trip, fence, barrier -- that only another warp can cover; profiles/r02c_attention_trace.txt).
// The two halves of a row exchange their maxima / sums through shared memory once per pass.  The linear block index is
// decoded tile-major, so the ragged last query tile of every head is scheduled after all full tiles.
constexpr int AT8_THREADS = 64 + 256;
// backward kernels: eight softmax warps, the TMA producer, and one issuing thread PER MMA STREAM (an issuing thread pays
// 100-200 cycles per tcgen05.mma / commit / barrier wait: with one thread for everything the dQ kernel needed ~15 such
// operations per 64-column tile and was bound by them)
constexpr int ATB_DQ_THREADS = 256 + 3 * 32;      // + producer, S/dP issuer, dQ issuer
constexpr int ATB_DKV_THREADS = 256 + 4 * 32;     // + producer, S^T/dP^T issuer, dV issuer, dK issuer
constexpr int KT8 = 128;
struct Fwd8Smem {
    static constexpr int NSLOT = 3;                               // ring of 128-row kv tiles
    static constexpr uint32_t Q = 0;
    static constexpr uint32_t KV = Q + ROWT_BYTES;
    static constexpr uint32_t P = KV + NSLOT * ROWT_BYTES;        // 128 x 128 bf16 = two K-major atoms
    static constexpr uint32_t RED = P + 2 * ROWT_BYTES;           // float [2][128]: row maxima / sums of the two halves
    static constexpr uint32_t BAR = RED + 2 * 128 * 4;
    // barriers: q_full, kv_full[NSLOT], kv_empty[NSLOT], s_full, s_free, p_full, p_free, o_full
    static constexpr int NBAR = 1 + 2 * NSLOT + 5;
    static constexpr uint32_t TOTAL = BAR + NBAR * 8 + 16 + 1024;
    static constexpr uint32_t O_COL = KT8;
    static constexpr uint32_t TMEM_COLS = 256;
};

template <int HD>
__global__ void __launch_bounds__(AT8_THREADS, 2)
attn_fwd_tc8_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N,
                    int H, int B, float scale_log2) {
    using S = Fwd8Smem;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_u32);
    const uint32_t q_full = base + S::BAR;
    const uint32_t kv_full = q_full + 8, kv_empty = kv_full + S::NSLOT * 8;
    const uint32_t s_full = kv_empty + S::NSLOT * 8, s_free = s_full + 8, p_full = s_free + 8, p_free = p_full + 8, o_full = p_free + 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + S::BAR + S::NBAR * 8);
    float* s_red = reinterpret_cast<float*>(sm + S::RED);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile-major decode: all full query tiles of all heads first, the ragged last tiles at the end of the grid
    const int pairs = H * B;
    const int pair = blockIdx.x % pairs, qt = blockIdx.x / pairs;
    const int h = pair % H, b = pair / H, q0 = qt * QT;
    const int D = H * HD;
    const int T = ceil_div(N, KT8);
    const int nv_last = N - (T - 1) * KT8;
    const int n16_last = (nv_last + 15) & ~15;
    const int qcol = h * HD, kcol = D + h * HD, vcol = 2 * D + h * HD;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < S::NSLOT; ++i) { mbar_init(kv_full + i * 8, 1); mbar_init(kv_empty + i * 8, 1); }
        mbar_init(s_full, 1);
        mbar_init(s_free, 256);
        mbar_init(p_full, 256);
        mbar_init(p_free, 1);
        mbar_init(o_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), S::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PDL_TRIGGER_EARLY();
    pdl_wait();

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer (128-row boxes)
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, ROWT_BYTES);
            tma_load_3d(base + S::Q, &tmQKV, q_full, qcol & ~63, q0, b);
            uint32_t seq = 0;
            auto load = [&](int col, int t) {
                const uint32_t slot = seq % S::NSLOT;
                mbar_wait(kv_empty + slot * 8, ((seq / S::NSLOT) & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full + slot * 8, ROWT_BYTES);
                tma_load_3d(base + S::KV + slot * ROWT_BYTES, &tmQKV, kv_full + slot * 8, col & ~63, t * KT8, b);
                ++seq;
            };
            for (int t = 0; t < T; ++t) load(kcol, t);
            for (int t = 0; t < T; ++t) {
                load(kcol, t);
                if (t >= 1) load(vcol, t - 1);
            }
            load(vcol, T - 1);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            const uint32_t subq = (qcol & 63) * 2, subk = (kcol & 63) * 2;
            constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, 64, 0u, 1u);
            uint32_t seq = 0;
            auto kv_wait = [&]() -> uint32_t {
                const uint32_t slot = seq % S::NSLOT;
                mbar_wait(kv_full + slot * 8, (seq / S::NSLOT) & 1);
                ++seq;
                return slot;
            };
            auto issue_s = [&](int it, int t) {          // score tile number `it` (both passes counted) = Q K_t^T
                const uint32_t slot = kv_wait();
                mbar_wait(s_free, (it & 1) ^ 1);
                tc_fence_after();
                const uint32_t idesc = umma_idesc_bf16(QT, t == T - 1 ? n16_last : KT8, 0u, 0u);
                const uint32_t sk = base + S::KV + slot * ROWT_BYTES;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)
                    umma_bf16(tmem_base, desc_kmajor(base + S::Q, subq, kk), desc_kmajor(sk, subk, kk), idesc, kk != 0);
                umma_commit(kv_empty + slot * 8);
                umma_commit(s_full);
            };
            auto issue_pv = [&](int u) {                 // O += P_u V_u, K = the tile's columns in steps of 16
                const uint32_t slot = kv_wait();
                mbar_wait(p_full, u & 1);
                tc_fence_after();
                const int nk = (u == T - 1 ? n16_last : KT8) / 16;
                const uint32_t sv = base + S::KV + slot * ROWT_BYTES, sp = base + S::P;
#pragma unroll 1
                for (int kk = 0; kk < nk; ++kk)
                    umma_bf16(tmem_base + S::O_COL, desc_kmajor(sp + (kk >> 2) * ROWT_BYTES, 0, kk & 3), desc_mnmajor(sv, kk), idesc_pv,
                              (u | kk) != 0);
                umma_commit(kv_empty + slot * 8);
                umma_commit(p_free);
            };
            mbar_wait(q_full, 0);
            for (int t = 0; t < T; ++t) issue_s(t, t);
            for (int t = 0; t < T; ++t) {
                issue_s(T + t, t);
                if (t >= 1) issue_pv(t - 1);
            }
            issue_pv(T - 1);
            umma_commit(o_full);
            PDL_TRIGGER_LATE();
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- softmax: thread = (row, 64-column half)
        const int q = warp & 3;                              // TMEM lane quadrant of this warp
        const int half = (warp - 2) >> 2;                    // which 64 columns of every 128-column tile
        const int row = q * 32 + lane;
        const int grow = q0 + row;
        const bool warp_valid = q0 + q * 32 < N;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 64;
        float mraw = -INFINITY;
        for (int t = 0; t < T; ++t) {                        // pass 1: row maxima of the raw scores
            const bool last = t == T - 1;
            const int nv = last ? nv_last - half * 64 : 64;                  // valid columns of this half (may be <= 0)
            const int nch = last ? min(4, max(0, (n16_last - half * 64) / 16)) : 4;
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            if (warp_valid && nch > 0) {
                uint32_t v[4][16];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < nch) tmem_ld_32x16(t_lane + c * 16, v[c]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c >= nch) continue;
                    if (last && (c + 1) * 16 > nv) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c * 16 + j < nv) mraw = fmaxf(mraw, __uint_as_float(v[c][j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            mraw = fmax3(mraw, __uint_as_float(v[c][j]), __uint_as_float(v[c][j + 1]));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(s_free);
        }
        s_red[half * 128 + row] = mraw;
        named_bar_sync(1, 256);
        mraw = fmaxf(mraw, s_red[(half ^ 1) * 128 + row]);
        const float msc = mraw * scale_log2;
        float l = 0.f;
        for (int t = 0; t < T; ++t) {                        // pass 2: p = 2^(s * scale_log2 - msc), row sums, P -> smem
            const int it = T + t;
            const bool last = t == T - 1;
            const int nv = last ? nv_last - half * 64 : 64;
            const int nch = last ? min(4, max(0, (n16_last - half * 64) / 16)) : 4;
            mbar_wait(s_full, it & 1);
            tc_fence_after();
            uint32_t pk[4][8];
            // two sub-halves of 32 columns: 32 score registers live at a time (the kernel must fit 2 x 320 threads per SM)
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                uint32_t v[2][16];
                if (warp_valid) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if (2 * sub + c < nch) tmem_ld_32x16(t_lane + (2 * sub + c) * 16, v[c]);
                    tmem_ld_wait();
                }
                if (sub == 1) {
                    tc_fence_before();
                    mbar_arrive(s_free);                     // the scores are in registers: the MMA warp may overwrite the tile
                }
                if (warp_valid) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int cc = 2 * sub + c;
                        if (cc >= nch) continue;
                        float p[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            p[j] = ex2_approx(fmaf(__uint_as_float(v[c][j]), scale_log2, -msc));
                            if (last && cc * 16 + j >= nv) p[j] = 0.f;
                            l += p[j];
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) pk[cc][j] = pack_bf16(p[2 * j], p[2 * j + 1]);
                    }
                }
            }
            mbar_wait(p_free, (t & 1) ^ 1);                  // P V of the previous tile has read the P buffer
            if (warp_valid) {
                const uint32_t sp = base + S::P + half * ROWT_BYTES;     // this half's K-major atom
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c >= nch) continue;
                    at_sts128(sp + at_swz(row, 2 * c), pk[c][0], pk[c][1], pk[c][2], pk[c][3]);
                    at_sts128(sp + at_swz(row, 2 * c + 1), pk[c][4], pk[c][5], pk[c][6], pk[c][7]);
                }
            }
            fence_proxy_async();
            mbar_arrive(p_full);
        }
        named_bar_sync(1, 256);                              // every thread has read its partner's maximum
        s_red[half * 128 + row] = l;
        named_bar_sync(1, 256);
        l += s_red[(half ^ 1) * 128 + row];
        mbar_wait(o_full, 0);
        tc_fence_after();
        if (warp_valid) {
            // the 16-column chunks of this head's O alternate between the two halves
            const uint32_t o_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + S::O_COL + (vcol & 63);
            constexpr int NCH = HD / 16;
            uint32_t o[(NCH + 1) / 2][16];
#pragma unroll
            for (int c = 0; c < (NCH + 1) / 2; ++c)
                if (2 * c + half < NCH) tmem_ld_32x16(o_lane + (2 * c + half) * 16, o[c]);
            tmem_ld_wait();
            if (grow < N) {
                const float inv = 1.f / l;
                __nv_bfloat16* orow = out + (static_cast<size_t>(b) * N + grow) * D + h * HD;
#pragma unroll
                for (int c = 0; c < (NCH + 1) / 2; ++c) {
                    if (2 * c + half >= NCH) continue;
                    uint4 lo, hi;
                    lo.x = pack_bf16(__uint_as_float(o[c][0]) * inv, __uint_as_float(o[c][1]) * inv);
                    lo.y = pack_bf16(__uint_as_float(o[c][2]) * inv, __uint_as_float(o[c][3]) * inv);
                    lo.z = pack_bf16(__uint_as_float(o[c][4]) * inv, __uint_as_float(o[c][5]) * inv);
                    lo.w = pack_bf16(__uint_as_float(o[c][6]) * inv, __uint_as_float(o[c][7]) * inv);
                    hi.x = pack_bf16(__uint_as_float(o[c][8]) * inv, __uint_as_float(o[c][9]) * inv);
                    hi.y = pack_bf16(__uint_as_float(o[c][10]) * inv, __uint_as_float(o[c][11]) * inv);
                    hi.z = pack_bf16(__uint_as_float(o[c][12]) * inv, __uint_as_float(o[c][13]) * inv);
                    hi.w = pack_bf16(__uint_as_float(o[c][14]) * inv, __uint_as_float(o[c][15]) * inv);
                    *reinterpret_cast<uint4*>(orow + (2 * c + half) * 16) = lo;
                    *reinterpret_cast<uint4*>(orow + (2 * c + half) * 16 + 8) = hi;
                }
                if (half == 0) lse[(static_cast<size_t>(b) * H + h) * N + grow] = (msc + log2f(l)) * LN2_F;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------ forward, resident
// Third forward layout (default for N <= 520): ALL scores of a 128-row query tile stay in tensor memory.  Why: the event log
// of the layouts above (profiles/r02e_attention_event_log.txt) shows that every operation of the single issuing thread --
// tcgen05.mma, tcgen05.commit, an mbarrier wait that passes at once -- costs 100-200 cycles, so a kernel that hands
// 64-column tiles back and forth between the MMA thread and the softmax warps spends ~1000 cycles of MMA-thread time per
// tile.  Here the MMA thread issues the whole S = Q K^T with N = 256 keys per instruction, once, and then only the P V
// products (K = 16 per instruction: N / 16 of them, the minimum).  Both softmax passes read the scores from TMEM (no
// recompute, no K reload); every P tile has a buffer of its own (no reuse, no back-pressure); O accumulates in the columns
// of the first score tile once that tile has been consumed; V tiles land in the K tiles' slots.
//   BIG   (N <= 512 + 8): 512 TMEM columns, one CTA per SM, 16 softmax warps (4 per lane quadrant, 32-column slices)
//   SMALL (N <= 256 + 8): 256 TMEM columns, two CTAs per SM, 8 softmax warps (2 per lane quadrant, 64-column slices)
// Key columns past the tensor-core part (N = 513 = 512 patches + cls: ONE column) are evaluated on the CUDA cores by the
// row's first slice instead of costing a 128 x 16 MMA tile and a barrier round trip.
constexpr int ATR_MAX_TAIL_COLS = 8;
template <bool BIG>
struct ResSmem {
    static constexpr int TILES = BIG ? 4 : 2;                      // 128-key tiles
    static constexpr int SLICES = BIG ? 4 : 2;                     // softmax warps per TMEM lane quadrant
    static constexpr int SOFT = SLICES * 128;                      // softmax threads
    static constexpr int THREADS = 64 + SOFT;
    static constexpr uint32_t TMEM_COLS = BIG ? 512 : 256;
    static constexpr uint32_t Q = 0;                               // dead once the score MMAs have completed ...
    static constexpr uint32_t P = 0;                               // ... so the first P atom lives there; per tile two K-major atoms
    static constexpr uint32_t KV = P + TILES * 2 * ROWT_BYTES;     // K tiles first, V tiles afterwards
    static constexpr uint32_t RED = KV + TILES * ROWT_BYTES;       // float [SLICES][128]: partial row maxima, then sums
    static constexpr uint32_t TAILP = RED + SLICES * 128 * 4;      // float [8][128]: p of the key columns past the MMA part
    static constexpr uint32_t BAR = TAILP + ATR_MAX_TAIL_COLS * 128 * 4;
    // barriers: qk_full, s_full, v_full[4], p_full[4], o_full
    static constexpr int NBAR = 11;
    static constexpr uint32_t TOTAL = BAR + NBAR * 8 + 16 + 1024;
};

// 8 consecutive bf16 (one uint4) times 8 floats
__device__ __forceinline__ float dot8(const uint4& a, const float* q) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        s = fmaf(f.x, q[2 * i], s);
        s = fmaf(f.y, q[2 * i + 1], s);
    }
    return s;
}
__device__ __forceinline__ void unpack8(const uint4& a, float* out) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        out[2 * i] = f.x;
        out[2 * i + 1] = f.y;
    }
}

// ---- ragged last rows on the CUDA cores.  Every sequence of this model is 128 k + 1 tokens long (patches + cls), so a
// 128-row tiling leaves ONE row per head; as a tensor-core tile it costs a whole CTA slot for the full CTA duration (the
// per-thread latency chain does not shrink with the row count) and turns the decoder's 256-CTA grids into 320 CTAs = one
// wave more.  Instead the last `pairs` CTAs of each grid are "light": no TMEM, no TMA, no MMA -- all threads of the CTA
// share the row's N dot products (thread t: items t and t + NTHR), block-wide reductions through shared memory.  Scheduled
// last (tile-major block order), they fill the slots the last wave of full tiles leaves free.  N % 128 <= ATR_MAX_TAIL_ROWS.
constexpr int ATR_MAX_TAIL_ROWS = 2;

struct TailScratch {
    float red[64];            // two block reductions of up to 32 warps
    float vec[20 * 64];       // per-warp partial vectors (<= 20 warps x 64 columns)
};

template <int NTHR>
__device__ __forceinline__ float block_allreduce(float v, bool is_max, float* scr, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, w) : v + w;
    }
    if ((tid & 31) == 0) scr[tid >> 5] = v;
    __syncthreads();
    float r = scr[0];
#pragma unroll
    for (int w = 1; w < NTHR / 32; ++w) r = is_max ? fmaxf(r, scr[w]) : r + scr[w];
    return r;
}

// sum over the block of HD values per thread -> thread d < HD returns component d (others 0)
template <int HD, int NTHR>
__device__ __forceinline__ float block_vecsum(float (&v)[HD], float* vec, int tid) {
#pragma unroll
    for (int d = 0; d < HD; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[d] += __shfl_xor_sync(0xffffffffu, v[d], o);
    }
    __syncthreads();                      // previous readers of vec are done
    if ((tid & 31) == 0) {
#pragma unroll
        for (int d = 0; d < HD; ++d) vec[(tid >> 5) * HD + d] = v[d];
    }
    __syncthreads();
    float r = 0.f;
    if (tid < HD) {
#pragma unroll
        for (int w = 0; w < NTHR / 32; ++w) r += vec[w * HD + tid];
    }
    return r;
}

// dot product of two bf16 rows of HD elements (16-byte aligned)
template <int HD>
__device__ __forceinline__ float row_dot(const __nv_bfloat16* a, const __nv_bfloat16* b) {
    const uint4* pa = reinterpret_cast<const uint4*>(a);
    const uint4* pb = reinterpret_cast<const uint4*>(b);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
        float x[8];
        unpack8(__ldg(pa + c), x);
        acc += dot8(__ldg(pb + c), x);
    }
    return acc;
}
// acc[0..HD) += w * row
template <int HD>
__device__ __forceinline__ void row_axpy(float (&acc)[HD], float w, const __nv_bfloat16* row) {
    const uint4* pr = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
        float x[8];
        unpack8(__ldg(pr + c), x);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[8 * c + e] = fmaf(w, x[e], acc[8 * c + e]);
    }
}

// forward of query row gi: out row and lse
template <int HD, int NTHR>
__device__ __noinline__ void tail_row_fwd(const __nv_bfloat16* __restrict__ qkv_b, size_t pitch, int qcol, int kcol, int vcol, int gi,
                                          int N, float scale_log2, __nv_bfloat16* __restrict__ orow, float* __restrict__ lse_out) {
    __shared__ TailScratch ts;
    const int tid = threadIdx.x;
    const __nv_bfloat16* qr = qkv_b + static_cast<size_t>(gi) * pitch + qcol;
    float sc[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int j = tid + k * NTHR;
        sc[k] = j < N ? row_dot<HD>(qr, qkv_b + static_cast<size_t>(j) * pitch + kcol) : -INFINITY;
    }
    const float msc = block_allreduce<NTHR>(fmaxf(sc[0], sc[1]), true, ts.red, tid) * scale_log2;
    float ov[HD], psum = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) ov[d] = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int j = tid + k * NTHR;
        if (j < N) {
            const float p = ex2_approx(fmaf(sc[k], scale_log2, -msc));
            psum += p;
            row_axpy<HD>(ov, p, qkv_b + static_cast<size_t>(j) * pitch + vcol);
        }
    }
    const float l = block_allreduce<NTHR>(psum, false, ts.red + 32, tid);
    const float od = block_vecsum<HD, NTHR>(ov, ts.vec, tid);
    if (tid < HD) orow[tid] = __float2bfloat16(od / l);
    if (tid == 0) *lse_out = (msc + log2f(l)) * LN2_F;
    __syncthreads();
}

// dQ of query row gi (and its delta)
template <int HD, int NTHR>
__device__ __noinline__ void tail_row_dq(const __nv_bfloat16* __restrict__ qkv_b, size_t pitch, int qcol, int kcol, int vcol,
                                         const __nv_bfloat16* __restrict__ orow, const __nv_bfloat16* __restrict__ dorow, float lse_i,
                                         int gi, int N, float scale, float scale_log2, __nv_bfloat16* __restrict__ dqrow,
                                         float* __restrict__ delta_out) {
    __shared__ TailScratch ts;
    const int tid = threadIdx.x;
    const __nv_bfloat16* qr = qkv_b + static_cast<size_t>(gi) * pitch + qcol;
    const float dl = row_dot<HD>(orow, dorow);
    const float lse2 = lse_i * LOG2E_F;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int j = tid + k * NTHR;
        if (j < N) {
            const __nv_bfloat16* kr = qkv_b + static_cast<size_t>(j) * pitch + kcol;
            const float p = ex2_approx(fmaf(row_dot<HD>(qr, kr), scale_log2, -lse2));
            const float ds = p * (row_dot<HD>(dorow, qkv_b + static_cast<size_t>(j) * pitch + vcol) - dl);
            row_axpy<HD>(acc, ds, kr);
        }
    }
    const float r = block_vecsum<HD, NTHR>(acc, ts.vec, tid);
    if (tid < HD) dqrow[tid] = __float2bfloat16(r * scale);
    if (tid == 0) *delta_out = dl;
    __syncthreads();
}

// dK and dV of key row gj
template <int HD, int NTHR>
__device__ __noinline__ void tail_row_dkv(const __nv_bfloat16* __restrict__ qkv_b, size_t pitch, int qcol, int kcol, int vcol,
                                          const __nv_bfloat16* __restrict__ dout_b, int D, const float* __restrict__ lrow,
                                          const float* __restrict__ drow, int gj, int N, float scale, float scale_log2,
                                          __nv_bfloat16* __restrict__ dkrow, __nv_bfloat16* __restrict__ dvrow, int hcol) {
    __shared__ TailScratch ts;
    const int tid = threadIdx.x;
    const __nv_bfloat16* kr = qkv_b + static_cast<size_t>(gj) * pitch + kcol;
    const __nv_bfloat16* vr = qkv_b + static_cast<size_t>(gj) * pitch + vcol;
    float pk[2], dsk[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = tid + k * NTHR;
        pk[k] = dsk[k] = 0.f;
        if (i < N) {
            const __nv_bfloat16* qi = qkv_b + static_cast<size_t>(i) * pitch + qcol;
            const __nv_bfloat16* doi = dout_b + static_cast<size_t>(i) * D + hcol;
            pk[k] = ex2_approx(fmaf(row_dot<HD>(qi, kr), scale_log2, -lrow[i] * LOG2E_F));
            dsk[k] = pk[k] * (row_dot<HD>(doi, vr) - drow[i]);
        }
    }
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = tid + k * NTHR;
        if (i < N) row_axpy<HD>(acc, pk[k], dout_b + static_cast<size_t>(i) * D + hcol);
    }
    float r = block_vecsum<HD, NTHR>(acc, ts.vec, tid);
    if (tid < HD) dvrow[tid] = __float2bfloat16(r);
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = tid + k * NTHR;
        if (i < N) row_axpy<HD>(acc, dsk[k], qkv_b + static_cast<size_t>(i) * pitch + qcol);
    }
    r = block_vecsum<HD, NTHR>(acc, ts.vec, tid);
    if (tid < HD) dkrow[tid] = __float2bfloat16(r * scale);
    __syncthreads();
}

template <int HD, bool BIG>
__global__ void __launch_bounds__(ResSmem<BIG>::THREADS, BIG ? 1 : 2)
attn_fwd_res_kernel(const __grid_constant__ CUtensorMap tmQKV, const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                    float* __restrict__ lse, int N, int H, int B, int tail_rows, float scale_log2) {
    using S = ResSmem<BIG>;
    constexpr int SLICES = S::SLICES;
    constexpr int CS = 128 / SLICES;                  // columns of a tile per slice (32 or 64)
    constexpr int NCH = CS / 16;                      // 16-column chunks per slice and tile
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_u32);
    const uint32_t qk_full = base + S::BAR, s_full = qk_full + 8, v_full = s_full + 8, p_full = v_full + 32, o_full = p_full + 32;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + S::BAR + S::NBAR * 8);
    float* s_red = reinterpret_cast<float*>(sm + S::RED);
    float* s_tailp = reinterpret_cast<float*>(sm + S::TAILP);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Warp roles: softmax warps first, TMA producer and MMA issuer LAST -- the warp scheduler prefers the highest warp id of
    // a sub-partition, and as warps 0 / 1 the two single-thread roles were starved by the softmax warps they feed (every
    // tcgen05.mma / commit / barrier wait took ~200 cycles; profiles/r02e_attention_event_log.txt).
    constexpr int PROD_WARP = S::SOFT / 32, MMA_WARP = S::SOFT / 32 + 1;
    const int pairs = H * B;                          // tile-major: ragged last query tiles are scheduled last
    const int pair = blockIdx.x % pairs, qt = blockIdx.x / pairs;
    const int h = pair % H, b = pair / H, q0 = qt * QT;
    const int D = H * HD;
    const int NM = min(N, S::TILES * 128);            // key columns on the tensor cores
    const int NT = N - NM;                            // key columns past them: CUDA cores (<= ATR_MAX_TAIL_COLS)
    const int n16M = (NM + 15) & ~15;
    const int T = ceil_div(NM, 128);                  // 128-column P / V tiles
    const int qcol = h * HD, kcol = D + h * HD, vcol = 2 * D + h * HD;
    const size_t pitch = static_cast<size_t>(3) * D;
    const __nv_bfloat16* qkv_b = qkv + static_cast<size_t>(b) * N * pitch;

    if (tail_rows > 0 && q0 + tail_rows == N) {       // a light CTA: the ragged last rows of this head, CUDA cores only
        pdl_wait();
        for (int i = 0; i < tail_rows; ++i)
            tail_row_fwd<HD, S::THREADS>(qkv_b, pitch, qcol, kcol, vcol, q0 + i, N, scale_log2,
                                         out + (static_cast<size_t>(b) * N + q0 + i) * D + h * HD,
                                         lse + (static_cast<size_t>(b) * H + h) * N + q0 + i);
        return;
    }
    if (threadIdx.x == 0) AT_STAMP(0);
    if (warp == PROD_WARP && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp == MMA_WARP && lane == 0) {
        mbar_init(qk_full, 1);
        mbar_init(s_full, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(v_full + i * 8, 1); mbar_init(p_full + i * 8, S::SOFT); }
        mbar_init(o_full, ceil_div(min(N, S::TILES * 128), 128) > 2 ? 2 : 1);      // one arrival per P V issuer (below)
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), S::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PDL_TRIGGER_EARLY();
    pdl_wait();
    if (threadIdx.x == 0) {
        AT_STAMP(1);
        unsigned smid__;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid__));
        AT_PUT(9, smid__);
    }

    // O (+)= P_t V_t for tiles [t0, t1) into the accumulator at `acc`: K = 16 keys per instruction, the minimum count; the
    // descriptors of a tile differ only in their 14-bit start-address field, so they are formed by one add each
    auto issue_pv = [&](int t0, int t1, uint32_t acc) {
        constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, 64, 0u, 1u);
        for (int t = t0; t < t1; ++t) {
            mbar_wait(v_full + t * 8, 0);
            mbar_wait(p_full + t * 8, 0);
            tc_fence_after();
            const int nk = min(128, n16M - 128 * t) / 16;
            const uint64_t dp = desc_kmajor(base + S::P + t * 2 * ROWT_BYTES, 0, 0);
            const uint64_t dv = desc_mnmajor(base + S::KV + t * ROWT_BYTES, 0);
            if (nk == 8) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_bf16(acc, dp + (((kk >> 2) * ROWT_BYTES + (kk & 3) * 32) >> 4), dv + ((kk * 2048) >> 4), idesc_pv,
                              (t != t0 || kk != 0) ? 1u : 0u);
            } else {
#pragma unroll 1
                for (int kk = 0; kk < nk; ++kk)
                    umma_bf16(acc, dp + (((kk >> 2) * ROWT_BYTES + (kk & 3) * 32) >> 4), dv + ((kk * 2048) >> 4), idesc_pv,
                              (t != t0 || kk != 0) ? 1u : 0u);
            }
        }
    };

    if (warp == PROD_WARP) {
        // ---------------------------------------------------------------- TMA producer: Q and every K tile, later every V tile
        if (lane == 0) {
            mbar_arrive_expect_tx(qk_full, (1 + T) * ROWT_BYTES);
            tma_load_3d(base + S::Q, &tmQKV, qk_full, qcol & ~63, q0, b);
            for (int t = 0; t < T; ++t) tma_load_3d(base + S::KV + t * ROWT_BYTES, &tmQKV, qk_full, kcol & ~63, t * 128, b);
            mbar_wait(s_full, 0);                     // the score MMAs have read K: its tiles become the V tiles
            for (int t = 0; t < T; ++t) {
                mbar_arrive_expect_tx(v_full + t * 8, ROWT_BYTES);
                tma_load_3d(base + S::KV + t * ROWT_BYTES, &tmQKV, v_full + t * 8, vcol & ~63, t * 128, b);
            }
            // its loads are out: this thread becomes the second P V issuer (tiles 2, 3 into a second accumulator)
            if (T > 2) {
                issue_pv(2, T, tmem_base + 64);
                umma_commit(o_full);
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            const uint32_t subq = (qcol & 63) * 2, subk = (kcol & 63) * 2;
            mbar_wait(qk_full, 0);
            AT_STAMP(2);
            tc_fence_after();
            for (int j = 0; j * 256 < NM; ++j) {      // S[:, 256 j ...] = Q K^T, up to 256 keys per instruction
                const uint32_t idesc = umma_idesc_bf16(QT, min(256, n16M - 256 * j), 0u, 0u);
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)
                    umma_bf16(tmem_base + 256 * j, desc_kmajor(base + S::Q, subq, kk),
                              desc_kmajor(base + S::KV + j * 2 * ROWT_BYTES, subk, kk), idesc, kk != 0);
            }
            umma_commit(s_full);
            issue_pv(0, min(T, 2), tmem_base);        // O += P_t V_t, O in the columns of the (consumed) first score tile
            umma_commit(o_full);
            AT_STAMP(5);
            PDL_TRIGGER_LATE();
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- softmax: thread = (row, CS-column slice of each tile)
        const int q = warp & 3;
        const int slice = warp >> 2;
        const int row = q * 32 + lane;
        const int grow = q0 + row;
        const bool warp_valid = q0 + q * 32 < N;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float mraw = -INFINITY;
        float s_tail[ATR_MAX_TAIL_COLS];
        mbar_wait(s_full, 0);
        if (threadIdx.x == 32) AT_STAMP(3);
        tc_fence_after();
        if (NT > 0 && slice == 0 && warp_valid) {     // scores of the key columns past the MMA part: q_row . k_j
            float qv[HD];
            const int chunk0 = (qcol & 63) >> 3;
#pragma unroll
            for (int c = 0; c < HD / 8; ++c)
                unpack8(*reinterpret_cast<const uint4*>(sm + S::Q + at_swz(row, chunk0 + c)), qv + 8 * c);
#pragma unroll
            for (int j = 0; j < ATR_MAX_TAIL_COLS; ++j) {
                s_tail[j] = 0.f;
                if (j < NT) {
                    const uint4* kr = reinterpret_cast<const uint4*>(qkv_b + static_cast<size_t>(NM + j) * pitch + kcol);
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < HD / 8; ++c) acc += dot8(__ldg(kr + c), qv + 8 * c);
                    s_tail[j] = acc;
                    mraw = fmaxf(mraw, acc);
                }
            }
        }
        for (int t = 0; t < T; ++t) {                 // pass 1: row maxima
            const int col0 = t * 128 + slice * CS;
            if (!warp_valid || col0 >= n16M) continue;
            const int nch = min(NCH, (n16M - col0) / 16);
#pragma unroll
            for (int c2 = 0; c2 < NCH; c2 += 2) {     // two chunks (32 score registers) at a time
                uint32_t v[2][16];
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if (c2 + c < nch) tmem_ld_32x16(t_lane + col0 + (c2 + c) * 16, v[c]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c2 + c >= nch) continue;
                    const int cb = col0 + (c2 + c) * 16;
                    if (cb + 16 > NM) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (cb + j < NM) mraw = fmaxf(mraw, __uint_as_float(v[c][j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            mraw = fmax3(mraw, __uint_as_float(v[c][j]), __uint_as_float(v[c][j + 1]));
                    }
                }
            }
        }
        s_red[slice * 128 + row] = mraw;
        named_bar_sync(1, S::SOFT);
        if (threadIdx.x == 32) AT_STAMP(4);
#pragma unroll
        for (int i = 0; i < SLICES; ++i) mraw = fmaxf(mraw, s_red[i * 128 + row]);
        const float msc = mraw * scale_log2;
        float l = 0.f;
        for (int t = 0; t < T; ++t) {                 // pass 2: p = 2^(s * scale_log2 - msc), P tile t -> its own buffer
            const int col0 = t * 128 + slice * CS;
            if (warp_valid && col0 < n16M) {
                const int nch = min(NCH, (n16M - col0) / 16);
                const uint32_t sp = base + S::P + t * 2 * ROWT_BYTES + ((slice * CS) >> 6) * ROWT_BYTES;
                const int ch0 = ((slice * CS) & 63) >> 3;                   // first 16-byte chunk of this slice in its atom
#pragma unroll
                for (int c2 = 0; c2 < NCH; c2 += 2) {
                    uint32_t v[2][16];
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if (c2 + c < nch) tmem_ld_32x16(t_lane + col0 + (c2 + c) * 16, v[c]);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c2 + c >= nch) continue;
                        const int cb = col0 + (c2 + c) * 16;
                        float p[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            p[j] = ex2_approx(fmaf(__uint_as_float(v[c][j]), scale_log2, -msc));
                            if (cb + j >= NM) p[j] = 0.f;
                            l += p[j];
                        }
                        const int ch = ch0 + 2 * (c2 + c);
                        at_sts128(sp + at_swz(row, ch), pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]),
                                  pack_bf16(p[6], p[7]));
                        at_sts128(sp + at_swz(row, ch + 1), pack_bf16(p[8], p[9]), pack_bf16(p[10], p[11]), pack_bf16(p[12], p[13]),
                                  pack_bf16(p[14], p[15]));
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(p_full + t * 8);
        }
        if (NT > 0 && slice == 0 && warp_valid) {
#pragma unroll
            for (int j = 0; j < ATR_MAX_TAIL_COLS; ++j)
                if (j < NT) {
                    const float p = ex2_approx(fmaf(s_tail[j], scale_log2, -msc));
                    l += p;
                    s_tailp[j * 128 + row] = p;
                }
        }
        if (threadIdx.x == 32) AT_STAMP(6);
        named_bar_sync(1, S::SOFT);                   // every thread has read the maxima: the buffer takes the sums
        s_red[slice * 128 + row] = l;
        named_bar_sync(1, S::SOFT);
        l = 0.f;
#pragma unroll
        for (int i = 0; i < SLICES; ++i) l += s_red[i * 128 + row];
        mbar_wait(o_full, 0);
        if (threadIdx.x == 32) AT_STAMP(7);
        tc_fence_after();
        for (int ch = slice; ch < HD / 16; ch += SLICES) {        // 16-column chunks of this head's O, dealt to the slices
            if (!warp_valid) break;
            uint32_t o[16], o2[16];
            tmem_ld_32x16(t_lane + (vcol & 63) + ch * 16, o);
            if (T > 2) tmem_ld_32x16(t_lane + 64 + (vcol & 63) + ch * 16, o2);     // tiles 2, 3 accumulated separately
            tmem_ld_wait();
            if (grow < N) {
                float of[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) of[j] = __uint_as_float(o[j]) + (T > 2 ? __uint_as_float(o2[j]) : 0.f);
                for (int j = 0; j < NT; ++j) {        // + p_j v_j of the key columns past the MMA part
                    const float p = s_tailp[j * 128 + row];
                    const uint4* vr = reinterpret_cast<const uint4*>(qkv_b + static_cast<size_t>(NM + j) * pitch + vcol + ch * 16);
                    float vv[16];
                    unpack8(__ldg(vr), vv);
                    unpack8(__ldg(vr + 1), vv + 8);
#pragma unroll
                    for (int d = 0; d < 16; ++d) of[d] = fmaf(p, vv[d], of[d]);
                }
                const float inv = 1.f / l;
                __nv_bfloat16* orow = out + (static_cast<size_t>(b) * N + grow) * D + h * HD + ch * 16;
                uint4 lo, hi;
                lo.x = pack_bf16(of[0] * inv, of[1] * inv); lo.y = pack_bf16(of[2] * inv, of[3] * inv);
                lo.z = pack_bf16(of[4] * inv, of[5] * inv); lo.w = pack_bf16(of[6] * inv, of[7] * inv);
                hi.x = pack_bf16(of[8] * inv, of[9] * inv); hi.y = pack_bf16(of[10] * inv, of[11] * inv);
                hi.z = pack_bf16(of[12] * inv, of[13] * inv); hi.w = pack_bf16(of[14] * inv, of[15] * inv);
                *reinterpret_cast<uint4*>(orow) = lo;
                *reinterpret_cast<uint4*>(orow + 8) = hi;
            }
        }
        if (slice == 0 && grow < N) lse[(static_cast<size_t>(b) * H + h) * N + grow] = (msc + log2f(l)) * LN2_F;
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) AT_STAMP(8);
    if (warp == 0) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------ backward: dQ
struct DqSmem {
    static constexpr int NSLOT = 4;
    static constexpr uint32_t Q = 0;
    static constexpr uint32_t DO = Q + ROWT_BYTES;
    static constexpr uint32_t KV = DO + ROWT_BYTES;
    static constexpr uint32_t DS = KV + NSLOT * BOX_BYTES;
    static constexpr uint32_t BAR = DS + ROWT_BYTES;
    // barriers: qdo_full, kv_full[NSLOT], kv_empty[NSLOT], sdp_full, sdp_free, ds_full, ds_free, dq_full
    static constexpr int NBAR = 1 + 2 * NSLOT + 5;
    static constexpr uint32_t TOTAL = BAR + NBAR * 8 + 16 + 1024;
    static constexpr uint32_t TMEM_COLS = 256;      // S at 0, dP at 64, dQ at 128
};

template <int HD>
__global__ void __launch_bounds__(ATB_DQ_THREADS, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                      const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                      const float* __restrict__ lse, float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int N, int H,
                      int B, int tail_rows, float scale, float scale_log2) {
    using S = DqSmem;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_u32);
    const uint32_t qdo_full = base + S::BAR;
    const uint32_t kv_full = qdo_full + 8, kv_empty = kv_full + S::NSLOT * 8;
    const uint32_t sdp_full = kv_empty + S::NSLOT * 8, sdp_free = sdp_full + 8, ds_full = sdp_free + 8, ds_free = ds_full + 8,
                   dq_full = ds_free + 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + S::BAR + S::NBAR * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pairs = H * B;                       // tile-major: the ragged last tiles are scheduled last
    const int pair = blockIdx.x % pairs;
    const int h = pair % H, b = pair / H, q0 = (blockIdx.x / pairs) * QT;
    const int D = H * HD;
    const int T = ceil_div(N, KT);
    const int nv_last = N - (T - 1) * KT;
    const int n16_last = (nv_last + 15) & ~15;
    const int qcol = h * HD, kcol = D + h * HD, vcol = 2 * D + h * HD;

    if (tail_rows > 0 && q0 + tail_rows == N) {       // a light CTA: the ragged last query rows of this head, CUDA cores only
        pdl_wait();
        const size_t pitch = static_cast<size_t>(3) * D;
        const __nv_bfloat16* qkv_b = qkv + static_cast<size_t>(b) * N * pitch;
        for (int i = 0; i < tail_rows; ++i) {
            const size_t r = static_cast<size_t>(b) * N + q0 + i;
            tail_row_dq<HD, ATB_DQ_THREADS>(qkv_b, pitch, qcol, kcol, vcol, out + r * D + h * HD, dout + r * D + h * HD,
                                         lse[(static_cast<size_t>(b) * H + h) * N + q0 + i], q0 + i, N, scale, scale_log2,
                                         dqkv + r * pitch + qcol, delta + (static_cast<size_t>(b) * H + h) * N + q0 + i);
        }
        return;
    }
    constexpr int PROD_WARP = 8, MMA_WARP = 9, DQ_WARP = 10;   // after the eight softmax warps: the scheduler prefers high warp ids
    if (warp == PROD_WARP && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
    }
    if (warp == MMA_WARP && lane == 0) {
        mbar_init(qdo_full, 1);
        for (int i = 0; i < S::NSLOT; ++i) { mbar_init(kv_full + i * 8, 1); mbar_init(kv_empty + i * 8, 1); }
        mbar_init(sdp_full, 1);
        mbar_init(sdp_free, 256);
        mbar_init(ds_full, 256);
        mbar_init(ds_free, 1);
        mbar_init(dq_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), S::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PDL_TRIGGER_EARLY();
    pdl_wait();

    if (warp == PROD_WARP) {
        if (lane == 0) {
            mbar_arrive_expect_tx(qdo_full, 2 * ROWT_BYTES);
            tma_load_3d(base + S::Q, &tmQKV, qdo_full, qcol & ~63, q0, b);
            tma_load_3d(base + S::Q + BOX_BYTES, &tmQKV, qdo_full, qcol & ~63, q0 + 64, b);
            tma_load_3d(base + S::DO, &tmDO, qdo_full, qcol & ~63, q0, b);
            tma_load_3d(base + S::DO + BOX_BYTES, &tmDO, qdo_full, qcol & ~63, q0 + 64, b);
            uint32_t seq = 0;
            for (int t = 0; t < T; ++t) {
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const uint32_t slot = seq % S::NSLOT;
                    mbar_wait(kv_empty + slot * 8, ((seq / S::NSLOT) & 1) ^ 1);
                    mbar_arrive_expect_tx(kv_full + slot * 8, BOX_BYTES);
                    tma_load_3d(base + S::KV + slot * BOX_BYTES, &tmQKV, kv_full + slot * 8, (which ? vcol : kcol) & ~63, t * KT, b);
                    ++seq;
                }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            const uint32_t subq = (qcol & 63) * 2, subk = (kcol & 63) * 2, subv = (vcol & 63) * 2;
            uint32_t seq = 0;
            auto kv_wait = [&]() -> uint32_t {
                const uint32_t slot = seq % S::NSLOT;
                mbar_wait(kv_full + slot * 8, (seq / S::NSLOT) & 1);
                ++seq;
                return slot;
            };
            mbar_wait(qdo_full, 0);
            for (int t = 0; t < T; ++t) {
                const uint32_t slot_k = kv_wait();
                const uint32_t slot_v = kv_wait();
                mbar_wait(sdp_free, (t & 1) ^ 1);
                tc_fence_after();
                const uint32_t idesc = umma_idesc_bf16(QT, t == T - 1 ? n16_last : KT, 0u, 0u);
                const uint32_t sk = base + S::KV + slot_k * BOX_BYTES, sv = base + S::KV + slot_v * BOX_BYTES;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)     // S = Q K^T
                    umma_bf16(tmem_base, desc_kmajor(base + S::Q, subq, kk), desc_kmajor(sk, subk, kk), idesc, kk != 0);
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)     // dP = dO V^T
                    umma_bf16(tmem_base + 64, desc_kmajor(base + S::DO, subq, kk), desc_kmajor(sv, subv, kk), idesc, kk != 0);
                umma_commit(kv_empty + slot_v * 8);
                umma_commit(sdp_full);
            }
            PDL_TRIGGER_LATE();
        }
        __syncwarp();
    } else if (warp == DQ_WARP) {
        // ---------------------------------------------------------------- second issuer: dQ += dS_t K_t
        if (lane == 0) {
            constexpr uint32_t idesc_dq = umma_idesc_bf16(QT, 64, 0u, 1u);
            const uint64_t d_ds = desc_kmajor(base + S::DS, 0, 0);
            for (int t = 0; t < T; ++t) {
                const uint32_t slot_k = (2 * t) % S::NSLOT;              // K_t is the (2 t)-th box of the ring
                mbar_wait(ds_full, t & 1);                               // implies S_t completed, i.e. K_t has landed
                tc_fence_after();
                const int nk = (t == T - 1 ? n16_last : KT) / 16;
                const uint64_t d_k = desc_mnmajor(base + S::KV + slot_k * BOX_BYTES, 0);
                if (nk == 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base + 128, d_ds + ((kk * 32) >> 4), d_k + ((kk * 2048) >> 4), idesc_dq, (t | kk) != 0);
                } else {
#pragma unroll 1
                    for (int kk = 0; kk < nk; ++kk)
                        umma_bf16(tmem_base + 128, d_ds + ((kk * 32) >> 4), d_k + ((kk * 2048) >> 4), idesc_dq, (t | kk) != 0);
                }
                umma_commit(kv_empty + slot_k * 8);
                umma_commit(ds_free);
            }
            umma_commit(dq_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int half = warp >> 2;                    // two warps per TMEM lane quadrant: 32 of a tile's 64 columns each
        const int row = q * 32 + lane;
        const int grow = q0 + row;
        const bool warp_valid = q0 + q * 32 < N;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // this row's log-sum-exp (base 2) and delta = rowsum(dO * O); delta also goes to global memory for the dK/dV kernel
        float lse2 = 0.f, dl = 0.f;
        if (grow < N) {
            lse2 = lse[(static_cast<size_t>(b) * H + h) * N + grow] * LOG2E_F;
            const uint4* po = reinterpret_cast<const uint4*>(out + (static_cast<size_t>(b) * N + grow) * D + h * HD);
            const uint4* pg = reinterpret_cast<const uint4*>(dout + (static_cast<size_t>(b) * N + grow) * D + h * HD);
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) {
                const uint4 a = po[i], g = pg[i];
                const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
                const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 af = __bfloat1622float2(ah[j]), gf = __bfloat1622float2(gh[j]);
                    dl = fmaf(af.x, gf.x, dl);
                    dl = fmaf(af.y, gf.y, dl);
                }
            }
            if (half == 0) delta[(static_cast<size_t>(b) * H + h) * N + grow] = dl;
        }
        for (int t = 0; t < T; ++t) {
            const bool last = t == T - 1;
            const int nch = last ? n16_last / 16 : 4;
            mbar_wait(sdp_full, t & 1);
            tc_fence_after();
            {
                // chunk-serial (a 16-column chunk of S and of dP at a time: 80 registers per thread with 352 threads x 2 CTAs):
                // chunk 0 is reduced to eight packed registers before chunk 1 is read, and the score tiles are released to the
                // S/dP issuer as soon as chunk 1 is in registers
                uint32_t pk[2][8];
                bool have[2] = {false, false};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int cc = 2 * half + c;
                    have[c] = warp_valid && cc < nch;
                    uint32_t sv[16], dv[16];
                    if (have[c]) {
                        tmem_ld_32x16(t_lane + cc * 16, sv);
                        tmem_ld_32x16(t_lane + 64 + cc * 16, dv);
                        tmem_ld_wait();
                    }
                    if (c == 1) {
                        tc_fence_before();
                        mbar_arrive(sdp_free);                   // this thread's part of both score tiles is in registers
                    }
                    if (have[c]) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            float p0 = ex2_approx(fmaf(__uint_as_float(sv[j]), scale_log2, -lse2));
                            float p1 = ex2_approx(fmaf(__uint_as_float(sv[j + 1]), scale_log2, -lse2));
                            if (last && cc * 16 + j >= nv_last) p0 = 0.f;
                            if (last && cc * 16 + j + 1 >= nv_last) p1 = 0.f;
                            pk[c][j >> 1] = pack_bf16(p0 * (__uint_as_float(dv[j]) - dl), p1 * (__uint_as_float(dv[j + 1]) - dl));
                        }
                    }
                }
                mbar_wait(ds_free, (t & 1) ^ 1);                 // dQ MMA of the previous tile has read the dS buffer
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int cc = 2 * half + c;
                    if (have[c]) {
                        at_sts128(base + S::DS + at_swz(row, 2 * cc), pk[c][0], pk[c][1], pk[c][2], pk[c][3]);
                        at_sts128(base + S::DS + at_swz(row, 2 * cc + 1), pk[c][4], pk[c][5], pk[c][6], pk[c][7]);
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(ds_full);
        }
        mbar_wait(dq_full, 0);
        tc_fence_after();
        if (warp_valid) {
            const int subk = kcol & 63;          // dQ = dS K: the MMA ran over the 64 columns of K's box
            constexpr int NCH = HD / 16;         // 16-column chunks of this head alternate between the two halves
            uint32_t o[(NCH + 1) / 2][16];
#pragma unroll
            for (int c = 0; c < (NCH + 1) / 2; ++c)
                if (2 * c + half < NCH) tmem_ld_32x16(t_lane + 128 + subk + (2 * c + half) * 16, o[c]);
            tmem_ld_wait();
            if (grow < N) {
                __nv_bfloat16* drow = dqkv + (static_cast<size_t>(b) * N + grow) * (3 * D) + h * HD;
#pragma unroll
                for (int c = 0; c < (NCH + 1) / 2; ++c) {
                    if (2 * c + half >= NCH) continue;
                    uint4 lo, hi;
                    lo.x = pack_bf16(__uint_as_float(o[c][0]) * scale, __uint_as_float(o[c][1]) * scale);
                    lo.y = pack_bf16(__uint_as_float(o[c][2]) * scale, __uint_as_float(o[c][3]) * scale);
                    lo.z = pack_bf16(__uint_as_float(o[c][4]) * scale, __uint_as_float(o[c][5]) * scale);
                    lo.w = pack_bf16(__uint_as_float(o[c][6]) * scale, __uint_as_float(o[c][7]) * scale);
                    hi.x = pack_bf16(__uint_as_float(o[c][8]) * scale, __uint_as_float(o[c][9]) * scale);
                    hi.y = pack_bf16(__uint_as_float(o[c][10]) * scale, __uint_as_float(o[c][11]) * scale);
                    hi.z = pack_bf16(__uint_as_float(o[c][12]) * scale, __uint_as_float(o[c][13]) * scale);
                    hi.w = pack_bf16(__uint_as_float(o[c][14]) * scale, __uint_as_float(o[c][15]) * scale);
                    *reinterpret_cast<uint4*>(drow + (2 * c + half) * 16) = lo;
                    *reinterpret_cast<uint4*>(drow + (2 * c + half) * 16 + 8) = hi;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------ backward: dK / dV
struct DkvSmem {
    static constexpr int NSLOT = 4;
    static constexpr uint32_t K = 0;
    static constexpr uint32_t V = K + ROWT_BYTES;
    static constexpr uint32_t RING = V + ROWT_BYTES;             // Q_t / dO_t boxes
    static constexpr uint32_t PT = RING + NSLOT * BOX_BYTES;
    static constexpr uint32_t DST = PT + ROWT_BYTES;
    static constexpr uint32_t STATS = DST + ROWT_BYTES;          // lse2[2][64], delta[2][64]
    static constexpr uint32_t BAR = STATS + 4 * 64 * 4;
    // barriers: kv_full, r_full[NSLOT], r_empty[NSLOT], sdp_full, sdp_free, pds_full, pds_free, dkv_full
    static constexpr int NBAR = 1 + 2 * NSLOT + 5;
    static constexpr uint32_t TOTAL = BAR + NBAR * 8 + 16 + 1024;
    static constexpr uint32_t TMEM_COLS = 256;      // S^T at 0, dP^T at 64, dK at 128, dV at 192
};

template <int HD>
__global__ void __launch_bounds__(ATB_DKV_THREADS, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                       const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                       const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int N, int H, int B, int tail_rows,
                       float scale, float scale_log2) {
    using S = DkvSmem;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_u32);
    const uint32_t kv_full = base + S::BAR;
    const uint32_t r_full = kv_full + 8, r_empty = r_full + S::NSLOT * 8;
    const uint32_t sdp_full = r_empty + S::NSLOT * 8, sdp_free = sdp_full + 8, pds_full = sdp_free + 8, pds_free = pds_full + 8,
                   dkv_full = pds_free + 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + S::BAR + S::NBAR * 8);
    float* s_lse = reinterpret_cast<float*>(sm + S::STATS);      // [2][64]
    float* s_del = s_lse + 128;                                  // [2][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pairs = H * B;
    const int pair = blockIdx.x % pairs;
    const int h = pair % H, b = pair / H, kv0 = (blockIdx.x / pairs) * QT;
    const int D = H * HD;
    const int T = ceil_div(N, KT);                   // query tiles of 64 rows
    const int nv_last = N - (T - 1) * KT;
    const int n16_last = (nv_last + 15) & ~15;
    const int qcol = h * HD, kcol = D + h * HD, vcol = 2 * D + h * HD;

    if (tail_rows > 0 && kv0 + tail_rows == N) {      // a light CTA: the ragged last key rows of this head, CUDA cores only
        pdl_wait();
        const size_t pitch = static_cast<size_t>(3) * D;
        const __nv_bfloat16* qkv_b = qkv + static_cast<size_t>(b) * N * pitch;
        for (int i = 0; i < tail_rows; ++i) {
            const size_t r = static_cast<size_t>(b) * N + kv0 + i;
            tail_row_dkv<HD, ATB_DKV_THREADS>(qkv_b, pitch, qcol, kcol, vcol, dout + static_cast<size_t>(b) * N * D, D,
                                          lse + (static_cast<size_t>(b) * H + h) * N, delta + (static_cast<size_t>(b) * H + h) * N,
                                          kv0 + i, N, scale, scale_log2, dqkv + r * pitch + kcol, dqkv + r * pitch + vcol, h * HD);
        }
        return;
    }
    constexpr int PROD_WARP = 8, MMA_WARP = 9, DV_WARP = 10, DK_WARP = 11;   // after the eight softmax warps
    if (warp == PROD_WARP && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
    }
    if (warp == MMA_WARP && lane == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < S::NSLOT; ++i) { mbar_init(r_full + i * 8, 1); mbar_init(r_empty + i * 8, 1); }
        mbar_init(sdp_full, 1);
        mbar_init(sdp_free, 256);
        mbar_init(pds_full, 256);
        mbar_init(pds_free, 2);      // the dV and the dK issuer
        mbar_init(dkv_full, 2);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), S::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PDL_TRIGGER_EARLY();
    pdl_wait();

    if (warp == PROD_WARP) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * ROWT_BYTES);
            tma_load_3d(base + S::K, &tmQKV, kv_full, kcol & ~63, kv0, b);
            tma_load_3d(base + S::K + BOX_BYTES, &tmQKV, kv_full, kcol & ~63, kv0 + 64, b);
            tma_load_3d(base + S::V, &tmQKV, kv_full, vcol & ~63, kv0, b);
            tma_load_3d(base + S::V + BOX_BYTES, &tmQKV, kv_full, vcol & ~63, kv0 + 64, b);
            uint32_t seq = 0;
            for (int t = 0; t < T; ++t) {
#pragma unroll
                for (int which = 0; which < 2; ++which) {     // Q_t, then dO_t
                    const uint32_t slot = seq % S::NSLOT;
                    mbar_wait(r_empty + slot * 8, ((seq / S::NSLOT) & 1) ^ 1);
                    mbar_arrive_expect_tx(r_full + slot * 8, BOX_BYTES);
                    tma_load_3d(base + S::RING + slot * BOX_BYTES, which ? &tmDO : &tmQKV, r_full + slot * 8, qcol & ~63, t * KT, b);
                    ++seq;
                }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            const uint32_t subq = (qcol & 63) * 2, subk = (kcol & 63) * 2, subv = (vcol & 63) * 2;
            uint32_t seq = 0;
            auto r_wait = [&]() -> uint32_t {
                const uint32_t slot = seq % S::NSLOT;
                mbar_wait(r_full + slot * 8, (seq / S::NSLOT) & 1);
                ++seq;
                return slot;
            };
            mbar_wait(kv_full, 0);
            for (int t = 0; t < T; ++t) {
                const uint32_t slot_q = r_wait();
                const uint32_t slot_do = r_wait();
                mbar_wait(sdp_free, (t & 1) ^ 1);
                tc_fence_after();
                const uint32_t idesc = umma_idesc_bf16(QT, t == T - 1 ? n16_last : KT, 0u, 0u);
                const uint32_t sq = base + S::RING + slot_q * BOX_BYTES, sdo = base + S::RING + slot_do * BOX_BYTES;
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)     // S^T = K Q_t^T
                    umma_bf16(tmem_base, desc_kmajor(base + S::K, subk, kk), desc_kmajor(sq, subq, kk), idesc, kk != 0);
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)     // dP^T = V dO_t^T
                    umma_bf16(tmem_base + 64, desc_kmajor(base + S::V, subv, kk), desc_kmajor(sdo, subq, kk), idesc, kk != 0);
                umma_commit(sdp_full);
            }
            PDL_TRIGGER_LATE();
        }
        __syncwarp();
    } else if (warp == DV_WARP || warp == DK_WARP) {
        // ---------------------------------------------------------------- dV += P_t^T dO_t   |   dK += dS_t^T Q_t
        if (lane == 0) {
            constexpr uint32_t idesc_acc = umma_idesc_bf16(QT, 64, 0u, 1u);
            const bool is_dv = warp == DV_WARP;
            const uint64_t d_a = desc_kmajor(base + (is_dv ? S::PT : S::DST), 0, 0);
            const uint32_t acc = tmem_base + (is_dv ? 192 : 128);
            for (int t = 0; t < T; ++t) {
                const uint32_t slot = (2 * t + (is_dv ? 1 : 0)) % S::NSLOT;     // ring order: Q_t, dO_t
                mbar_wait(pds_full, t & 1);                // implies S^T_t / dP^T_t completed, i.e. Q_t and dO_t have landed
                tc_fence_after();
                const int nk = (t == T - 1 ? n16_last : KT) / 16;
                const uint64_t d_b = desc_mnmajor(base + S::RING + slot * BOX_BYTES, 0);
                if (nk == 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(acc, d_a + ((kk * 32) >> 4), d_b + ((kk * 2048) >> 4), idesc_acc, (t | kk) != 0);
                } else {
#pragma unroll 1
                    for (int kk = 0; kk < nk; ++kk)
                        umma_bf16(acc, d_a + ((kk * 32) >> 4), d_b + ((kk * 2048) >> 4), idesc_acc, (t | kk) != 0);
                }
                umma_commit(r_empty + slot * 8);
                umma_commit(pds_free);
            }
            umma_commit(dkv_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int half = warp >> 2;                // two warps per TMEM lane quadrant: 32 of a tile's 64 query columns each
        const int row = q * 32 + lane;                   // kv row of this thread
        const int grow = kv0 + row;
        const bool warp_valid = kv0 + q * 32 < N;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int st = threadIdx.x;                      // 0..255: the softmax threads come first
        const float* lrow = lse + (static_cast<size_t>(b) * H + h) * N;
        const float* drow = delta + (static_cast<size_t>(b) * H + h) * N;
        auto stage = [&](int t) {                        // statistics of query tile t -> shared memory (buffer t & 1)
            const int c = st & 63, qi = t * KT + c;
            if (st < 64) s_lse[(t & 1) * 64 + c] = qi < N ? lrow[qi] * LOG2E_F : INFINITY;   // padded query rows: p = 0
            else if (st < 128) s_del[(t & 1) * 64 + c] = qi < N ? drow[qi] : 0.f;
        };
        stage(0);
        named_bar_sync(1, 256);
        for (int t = 0; t < T; ++t) {
            const bool last = t == T - 1;
            const int nch = last ? n16_last / 16 : 4;
            if (t + 1 < T) stage(t + 1);
            const float* sl = s_lse + (t & 1) * 64;
            const float* sd = s_del + (t & 1) * 64;
            mbar_wait(sdp_full, t & 1);
            tc_fence_after();
            {
                uint32_t pp[2][8], dd8[2][8];      // chunk-serial, as in the dQ kernel
                bool have[2] = {false, false};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int cc = 2 * half + c;
                    have[c] = warp_valid && cc < nch;
                    uint32_t sv[16], dv[16];
                    if (have[c]) {
                        tmem_ld_32x16(t_lane + cc * 16, sv);
                        tmem_ld_32x16(t_lane + 64 + cc * 16, dv);
                        tmem_ld_wait();
                    }
                    if (c == 1) {
                        tc_fence_before();
                        mbar_arrive(sdp_free);
                    }
                    if (have[c]) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 l4 = *reinterpret_cast<const float4*>(sl + cc * 16 + 4 * j4);
                            const float4 d4 = *reinterpret_cast<const float4*>(sd + cc * 16 + 4 * j4);
                            const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
                            float p[4], ds[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int j = 4 * j4 + e;
                                p[e] = ex2_approx(fmaf(__uint_as_float(sv[j]), scale_log2, -lv[e]));
                                ds[e] = p[e] * (__uint_as_float(dv[j]) - dd[e]);
                            }
                            pp[c][2 * j4] = pack_bf16(p[0], p[1]); pp[c][2 * j4 + 1] = pack_bf16(p[2], p[3]);
                            dd8[c][2 * j4] = pack_bf16(ds[0], ds[1]); dd8[c][2 * j4 + 1] = pack_bf16(ds[2], ds[3]);
                        }
                    }
                }
                mbar_wait(pds_free, (t & 1) ^ 1);        // the dK / dV MMAs of the previous tile have read P^T / dS^T
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int cc = 2 * half + c;
                    if (have[c]) {
                        at_sts128(base + S::PT + at_swz(row, 2 * cc), pp[c][0], pp[c][1], pp[c][2], pp[c][3]);
                        at_sts128(base + S::PT + at_swz(row, 2 * cc + 1), pp[c][4], pp[c][5], pp[c][6], pp[c][7]);
                        at_sts128(base + S::DST + at_swz(row, 2 * cc), dd8[c][0], dd8[c][1], dd8[c][2], dd8[c][3]);
                        at_sts128(base + S::DST + at_swz(row, 2 * cc + 1), dd8[c][4], dd8[c][5], dd8[c][6], dd8[c][7]);
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(pds_full);
            named_bar_sync(1, 256);      // next tile's statistics are staged; this tile's are no longer read
        }
        mbar_wait(dkv_full, 0);
        tc_fence_after();
        if (warp_valid) {
            const int subq = qcol & 63;      // both accumulators span the 64 columns of the Q / dO boxes
            const int which = half;          // half 0: dK (x scale) -> k section, half 1: dV -> v section
            uint32_t o[HD / 16][16];
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) tmem_ld_32x16(t_lane + 128 + which * 64 + subq + c * 16, o[c]);
            tmem_ld_wait();
            if (grow < N) {
                const float f = which ? 1.0f : scale;
                __nv_bfloat16* orow = dqkv + (static_cast<size_t>(b) * N + grow) * (3 * D) + (which ? vcol : kcol);
#pragma unroll
                for (int c = 0; c < HD / 16; ++c) {
                    uint4 lo, hi;
                    lo.x = pack_bf16(__uint_as_float(o[c][0]) * f, __uint_as_float(o[c][1]) * f);
                    lo.y = pack_bf16(__uint_as_float(o[c][2]) * f, __uint_as_float(o[c][3]) * f);
                    lo.z = pack_bf16(__uint_as_float(o[c][4]) * f, __uint_as_float(o[c][5]) * f);
                    lo.w = pack_bf16(__uint_as_float(o[c][6]) * f, __uint_as_float(o[c][7]) * f);
                    hi.x = pack_bf16(__uint_as_float(o[c][8]) * f, __uint_as_float(o[c][9]) * f);
                    hi.y = pack_bf16(__uint_as_float(o[c][10]) * f, __uint_as_float(o[c][11]) * f);
                    hi.z = pack_bf16(__uint_as_float(o[c][12]) * f, __uint_as_float(o[c][13]) * f);
                    hi.w = pack_bf16(__uint_as_float(o[c][14]) * f, __uint_as_float(o[c][15]) * f);
                    *reinterpret_cast<uint4*>(orow + c * 16) = lo;
                    *reinterpret_cast<uint4*>(orow + c * 16 + 8) = hi;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

template <typename Kern>
static int set_smem(Kern kern, uint32_t bytes, bool& done) {
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        if (e != cudaSuccess) return set_error(-3, "attention: cudaFuncSetAttribute(smem=%u): %s", bytes, cudaGetErrorString(e));
        done = true;
    }
    return 0;
}

static int fwd_cfg() {       // VITAE_ATTN_FWD_CFG=0: two CTAs per SM with double buffers (A/B timing); default 1 (three per SM)
    static const int cfg = [] { const char* e = getenv("VITAE_ATTN_FWD_CFG"); return e && e[0] == '0' ? 0 : 1; }();
    return cfg;
}

template <int HD, int CFG>
static int launch_fwd_cfg(const CUtensorMap& tq, void* out, float* lse, int B, int N, int H, float sl2, cudaStream_t st) {
    static bool attr = false;
    if (int rc = set_smem(attn_fwd_tc_kernel<HD, CFG>, FwdSmem<CFG>::TOTAL, attr)) return rc;
    launch_kernel(attn_fwd_tc_kernel<HD, CFG>, dim3(ceil_div(N, QT), H, B), dim3(AT_THREADS), FwdSmem<CFG>::TOTAL, st, tq,
                  static_cast<__nv_bfloat16*>(out), lse, N, H, sl2);
    VITAE_CHECK_LAUNCH("attention_fwd");
    return 0;
}

template <int HD>
static int launch_fwd(const CUtensorMap& tq, void* out, float* lse, int B, int N, int H, float sl2, cudaStream_t st) {
    return fwd_cfg() == 0 ? launch_fwd_cfg<HD, 0>(tq, out, lse, B, N, H, sl2, st) : launch_fwd_cfg<HD, 1>(tq, out, lse, B, N, H, sl2, st);
}

template <int HD>
static int launch_fwd8(const CUtensorMap& tq128, void* out, float* lse, int B, int N, int H, float sl2, cudaStream_t st) {
    static bool attr = false;
    if (int rc = set_smem(attn_fwd_tc8_kernel<HD>, Fwd8Smem::TOTAL, attr)) return rc;
    launch_kernel(attn_fwd_tc8_kernel<HD>, dim3(ceil_div(N, QT) * H * B), dim3(AT8_THREADS), Fwd8Smem::TOTAL, st, tq128,
                  static_cast<__nv_bfloat16*>(out), lse, N, H, B, sl2);
    VITAE_CHECK_LAUNCH("attention_fwd");
    return 0;
}

// query tiles on the tensor cores and ragged last rows on the CUDA cores for a sequence of N tokens
// Light tail CTAs are opt-in (VITAE_ATTN_LIGHT_TAILS=1): measured on B200 (profiles/r02j_attention_light_tails.txt) their
// chain of dependent global loads makes them the longest CTAs of the small encoder grids (N = 129: backward 21 -> 28 us) and
// buys the decoder grid 4 us of 60; by default a ragged last row is one more (mostly empty) tensor-core tile.
static void res_tiling(int N, int* q_tiles, int* tail_rows) {
    static const bool light = [] { const char* e = getenv("VITAE_ATTN_LIGHT_TAILS"); return e && e[0] == '1'; }();
    const int full = N / QT, r = N % QT;
    *tail_rows = (light && full > 0 && r > 0 && r <= ATR_MAX_TAIL_ROWS && N <= 2 * AT8_THREADS) ? r : 0;   // two items per thread
    *q_tiles = full + ((r > *tail_rows) ? 1 : 0);
}

template <int HD, bool BIG>
static int launch_fwd_res(const CUtensorMap& tq128, const void* qkv, void* out, float* lse, int B, int N, int H, float sl2,
                          cudaStream_t st) {
    static bool attr = false;
    using S = ResSmem<BIG>;
    if (int rc = set_smem(attn_fwd_res_kernel<HD, BIG>, S::TOTAL, attr)) return rc;
    int q_tiles, tail_rows;
    res_tiling(N, &q_tiles, &tail_rows);
    launch_kernel(attn_fwd_res_kernel<HD, BIG>, dim3((q_tiles + (tail_rows ? 1 : 0)) * H * B), dim3(S::THREADS), S::TOTAL, st, tq128,
                  static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, N, H, B, tail_rows, sl2);
    VITAE_CHECK_LAUNCH("attention_fwd");
    return 0;
}

template <int HD>
static int dispatch_fwd_res(const CUtensorMap& tq128, const void* qkv, void* out, float* lse, int B, int N, int H, float sl2,
                            cudaStream_t st) {
    return N <= 256 + ATR_MAX_TAIL_COLS ? launch_fwd_res<HD, false>(tq128, qkv, out, lse, B, N, H, sl2, st)
                                        : launch_fwd_res<HD, true>(tq128, qkv, out, lse, B, N, H, sl2, st);
}

template <int HD>
static int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tdo, const void* qkv, const void* out, const void* dout,
                      const float* lse, float* delta, void* dqkv, int B, int N, int H, float scale, float sl2, cudaStream_t st) {
    static bool attr_dq = false, attr_dkv = false;
    if (int rc = set_smem(attn_bwd_dq_tc_kernel<HD>, DqSmem::TOTAL, attr_dq)) return rc;
    if (int rc = set_smem(attn_bwd_dkv_tc_kernel<HD>, DkvSmem::TOTAL, attr_dkv)) return rc;
    int q_tiles, tail_rows;
    res_tiling(N, &q_tiles, &tail_rows);
    const dim3 grid((q_tiles + (tail_rows ? 1 : 0)) * H * B);
    launch_kernel(attn_bwd_dq_tc_kernel<HD>, grid, dim3(ATB_DQ_THREADS), DqSmem::TOTAL, st, tq, tdo,
                  static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
                  static_cast<const __nv_bfloat16*>(dout), lse, delta, static_cast<__nv_bfloat16*>(dqkv), N, H, B, tail_rows, scale, sl2);
    VITAE_CHECK_LAUNCH("attention_bwd_dq");
    launch_kernel(attn_bwd_dkv_tc_kernel<HD>, grid, dim3(ATB_DKV_THREADS), DkvSmem::TOTAL, st, tq, tdo,
                  static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dout), lse,
                  static_cast<const float*>(delta), static_cast<__nv_bfloat16*>(dqkv), N, H, B, tail_rows, scale, sl2);
    VITAE_CHECK_LAUNCH("attention_bwd_dkv");
    return 0;
}

}  // namespace vitae

using namespace vitae;

// the round-1 mma.sync kernels (attention.cu), kept for A/B timing: VITAE_ATTN_LEGACY=1
namespace vitae {
int attention_fwd_legacy(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, void* stream);
int attention_bwd_legacy(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int B,
                         int N, int H, int hd, float scale, void* stream);
}  // namespace vitae

#ifdef VITAE_ATTN_TRACE
extern "C" int vitae_debug_set_attn_trace(void* buf) {
    cudaError_t e = cudaMemcpyToSymbol(g_attn_trace, &buf, sizeof(buf));
    return e == cudaSuccess ? 0 : set_error(-3, "set_attn_trace: %s", cudaGetErrorString(e));
}
#endif

// VITAE_ATTN_LEGACY=1: the round-1 mma.sync kernels for forward and backward; =bwd: for the backward only (A/B timing)
static bool use_legacy(bool backward) {
    static const int mode = [] {
        const char* e = getenv("VITAE_ATTN_LEGACY");
        return !e ? 0 : (e[0] == '1' ? 3 : (e[0] == 'b' ? 2 : (e[0] == 'f' ? 1 : 0)));
    }();
    return backward ? (mode & 2) != 0 : (mode & 1) != 0;
}

extern "C" int vitae_attention_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale, void* stream) {
    if (use_legacy(false)) return attention_fwd_legacy(qkv, out, lse, B, N, H, hd, scale, stream);
    VITAE_REQUIRE(qkv && out && lse, "attention_fwd: null pointer");
    VITAE_REQUIRE(B > 0 && N > 0 && H > 0 && (hd == 16 || hd == 32 || hd == 64), "attention_fwd: unsupported shape B=%d N=%d H=%d hd=%d", B, N, H, hd);
    VITAE_REQUIRE((H * hd) % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "attention_fwd: qkv / out must be 16-byte aligned and H*hd a multiple of 8");
    const int D = H * hd;
    const float sl2 = scale * LOG2E_F;
    cudaStream_t st = as_stream(stream);
    CUtensorMap tq;
    static const bool narrow = [] { const char* e = getenv("VITAE_ATTN_FWD"); return e && e[0] == 'v' && e[1] == '1'; }();
    static const bool streaming = [] { const char* e = getenv("VITAE_ATTN_FWD"); return e && e[0] == 'v' && e[1] == '2'; }();
    if (!narrow && !streaming && N <= 512 + ATR_MAX_TAIL_COLS) {     // default: all scores resident in TMEM
        if (int rc = make_tmap(&tq, qkv, 2, 3ull * D, (uint64_t)N, (uint64_t)B, 3ull * D, 64, 128)) return rc;
        if (hd == 64) return dispatch_fwd_res<64>(tq, qkv, out, lse, B, N, H, sl2, st);
        if (hd == 32) return dispatch_fwd_res<32>(tq, qkv, out, lse, B, N, H, sl2, st);
        return dispatch_fwd_res<16>(tq, qkv, out, lse, B, N, H, sl2, st);
    }
    if (!narrow) {      // longer sequences (or VITAE_ATTN_FWD=v2): 128-column score tiles streamed through TMEM, eight softmax warps
        if (int rc = make_tmap(&tq, qkv, 2, 3ull * D, (uint64_t)N, (uint64_t)B, 3ull * D, 64, 128)) return rc;
        if (hd == 64) return launch_fwd8<64>(tq, out, lse, B, N, H, sl2, st);
        if (hd == 32) return launch_fwd8<32>(tq, out, lse, B, N, H, sl2, st);
        return launch_fwd8<16>(tq, out, lse, B, N, H, sl2, st);
    }
    if (int rc = make_tmap(&tq, qkv, 2, 3ull * D, (uint64_t)N, (uint64_t)B, 3ull * D, 64, 64)) return rc;
    if (hd == 64) return launch_fwd<64>(tq, out, lse, B, N, H, sl2, st);
    if (hd == 32) return launch_fwd<32>(tq, out, lse, B, N, H, sl2, st);
    return launch_fwd<16>(tq, out, lse, B, N, H, sl2, st);
}

extern "C" int vitae_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                                   void* dqkv, int B, int N, int H, int hd, float scale, void* stream) {
    if (use_legacy(true)) return attention_bwd_legacy(qkv, out, dout, lse, delta, dqkv, B, N, H, hd, scale, stream);
    VITAE_REQUIRE(qkv && out && dout && lse && delta && dqkv, "attention_bwd: null pointer");
    VITAE_REQUIRE(B > 0 && N > 0 && H > 0 && (hd == 16 || hd == 32 || hd == 64), "attention_bwd: unsupported shape B=%d N=%d H=%d hd=%d", B, N, H, hd);
    VITAE_REQUIRE((H * hd) % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(dout) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
                  "attention_bwd: tensors must be 16-byte aligned and H*hd a multiple of 8");
    const int D = H * hd;
    CUtensorMap tq, tdo;
    if (int rc = make_tmap(&tq, qkv, 2, 3ull * D, (uint64_t)N, (uint64_t)B, 3ull * D, 64, 64)) return rc;
    if (int rc = make_tmap(&tdo, dout, 2, (uint64_t)D, (uint64_t)N, (uint64_t)B, (uint64_t)D, 64, 64)) return rc;
    const float sl2 = scale * LOG2E_F;
    cudaStream_t st = as_stream(stream);
    if (hd == 64) return launch_bwd<64>(tq, tdo, qkv, out, dout, lse, delta, dqkv, B, N, H, scale, sl2, st);
    if (hd == 32) return launch_bwd<32>(tq, tdo, qkv, out, dout, lse, delta, dqkv, B, N, H, scale, sl2, st);
    return launch_bwd<16>(tq, tdo, qkv, out, dout, lse, delta, dqkv, B, N, H, scale, sl2, st);
}
