// Micro-benchmark of the TMA -> shared-memory load path on B200 (no MMA): how many bytes per clock one SM can ingest
// as a function of box shape, ring depth, CTAs per SM, number of issuing threads and how many CTAs share a tile.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_bench tools/tma_bench.cu -lcuda && /tmp/tma_bench
// The numbers size the GEMM tiles ("what bounds the small GEMMs").
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../vit_ae_plus_plus_b200/csrc/ptx.cuh"

using namespace vitae;

constexpr int BK = 64;

// Each CTA streams `nsub` 64-wide k-sub-blocks (an A box rows_a x 64 and a B box rows_b x 64 each) through a ring of
// `stages` stages of `spb` sub-blocks.  `nprod` producer threads (lane 0 of warps 0..nprod-1) share the issue work:
//   nprod <= spb (spb % nprod == 0): every producer loads spb/nprod sub-blocks of every stage (full barrier count nprod)
//   nprod >  spb (spb == 1)        : producer j owns stages j, j+nprod, ...           (full barrier count 1)
// The consumer (lane 0 of warp 4) waits for a stage and releases it immediately.
template <int STAGES, int SPB, int NPROD>
__global__ void __launch_bounds__(160)
tma_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int rows_a, int rows_b,
                  int nsub, int share_a, int share_b, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const int a_bytes = rows_a * BK * 2, b_bytes = rows_b * BK * 2;
    const int sub_bytes = a_bytes + b_bytes;
    const int stage_bytes = sub_bytes * SPB;
    const uint32_t full_bar = base + STAGES * stage_bytes;
    const uint32_t empty_bar = full_bar + STAGES * 8;
    constexpr int FULL_COUNT = NPROD <= SPB ? NPROD : 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s * 8, FULL_COUNT);
            mbar_init(empty_bar + s * 8, 1);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    const int a_row0 = (blockIdx.x / share_a) * rows_a;
    const int b_row0 = (blockIdx.x % share_b) * rows_b;
    const int nstage_iters = nsub / SPB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    if (warp < NPROD && lane == 0) {
        if (NPROD <= SPB) {
            constexpr int MINE = SPB / (NPROD <= SPB ? NPROD : 1);
            for (int it = 0; it < nstage_iters; ++it) {
                const int s = it % STAGES;
                mbar_wait(empty_bar + s * 8, ((it / STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(full_bar + s * 8, MINE * sub_bytes);
#pragma unroll
                for (int u = 0; u < MINE; ++u) {
                    const int sub = warp * MINE + u;
                    const uint32_t dst = base + s * stage_bytes + sub * sub_bytes;
                    const int k0 = (it * SPB + sub) * BK;
                    tma_load_2d(dst, &tmA, full_bar + s * 8, k0, a_row0);
                    tma_load_2d(dst + a_bytes, &tmB, full_bar + s * 8, k0, b_row0);
                }
            }
        } else {
            for (int it = warp; it < nstage_iters; it += NPROD) {
                const int s = it % STAGES;
                mbar_wait(empty_bar + s * 8, ((it / STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(full_bar + s * 8, sub_bytes);
                const uint32_t dst = base + s * stage_bytes;
                tma_load_2d(dst, &tmA, full_bar + s * 8, it * BK, a_row0);
                tma_load_2d(dst + a_bytes, &tmB, full_bar + s * 8, it * BK, b_row0);
            }
        }
    } else if (warp == 4 && lane == 0) {
        for (int it = 0; it < nstage_iters; ++it) {
            const int s = it % STAGES;
            mbar_wait(full_bar + s * 8, (it / STAGES) & 1);
            mbar_arrive(empty_bar + s * 8);
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* ptr, uint64_t K, uint64_t rows, uint32_t box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {K, rows};
    const cuuint64_t strides[1] = {K * 2};
    const cuuint32_t box[2] = {BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        exit(1);
    }
    return m;
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    const uint64_t K = 4096, ROWS = 148 * 2 * 256;   // 620 MB >> L2 when every CTA has its own rows
    __nv_bfloat16 *A, *B;
    cudaMalloc(&A, K * ROWS * 2);
    cudaMalloc(&B, K * ROWS * 2);
    cudaMemset(A, 0, K * ROWS * 2);
    cudaMemset(B, 0, K * ROWS * 2);
    unsigned long long* cyc;
    cudaMalloc(&cyc, 1024 * 8);
    printf("%-7s %-7s %-7s %-4s %-6s %-6s | %9s %10s %10s %10s\n", "rows_a", "rows_b", "stages", "spb", "nprod", "grid", "us",
           "cyc/sub", "B/clk/CTA", "TB/s");
    struct Cfg { int ra, rb, grid, nsub; };
    const Cfg shapes[] = {{128, 64, 136, 64}, {128, 64, 296, 64}, {128, 128, 148, 64}, {128, 256, 148, 64}};
    for (const Cfg& c : shapes) {
        CUtensorMap ta = make_map(enc, A, K, ROWS, c.ra), tb = make_map(enc, B, K, ROWS, c.rb);
        const int sub_bytes = (c.ra + c.rb) * BK * 2;
        auto run = [&](auto kern, int stages, int spb, int nprod) {
            const int smem = stages * spb * sub_bytes + stages * 16 + 2048;
            if (smem > 227 * 1024) return;
            if (c.grid > 148 && 2 * smem > 227 * 1024) return;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                kern<<<c.grid, 160, smem>>>(ta, tb, c.ra, c.rb, c.nsub, 8, 8, cyc);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                if (err != cudaSuccess) {
                    printf("kernel failed: %s\n", cudaGetErrorString(err));
                    exit(1);
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            unsigned long long h[1024];
            cudaMemcpy(h, cyc, c.grid * 8, cudaMemcpyDeviceToHost);
            unsigned long long sum = 0;
            for (int i = 0; i < c.grid; ++i) sum += h[i];
            const double med = double(sum) / c.grid;
            const double bytes_cta = double(sub_bytes) * c.nsub;
            printf("%-7d %-7d %-7d %-4d %-6d %-6d | %9.2f %10.1f %10.1f %10.2f\n", c.ra, c.rb, stages, spb, nprod, c.grid,
                   best * 1e3, med / c.nsub, bytes_cta / med, bytes_cta * c.grid / (best * 1e-3) / 1e12);
        };
        run(tma_stream_kernel<4, 1, 1>, 4, 1, 1);
        run(tma_stream_kernel<8, 1, 1>, 8, 1, 1);
        run(tma_stream_kernel<4, 1, 2>, 4, 1, 2);
        run(tma_stream_kernel<8, 1, 2>, 8, 1, 2);
        run(tma_stream_kernel<4, 1, 4>, 4, 1, 4);
        run(tma_stream_kernel<8, 1, 4>, 8, 1, 4);
        run(tma_stream_kernel<2, 2, 1>, 2, 2, 1);
        run(tma_stream_kernel<4, 2, 1>, 4, 2, 1);
        run(tma_stream_kernel<4, 2, 2>, 4, 2, 2);
        run(tma_stream_kernel<2, 4, 1>, 2, 4, 1);
        run(tma_stream_kernel<2, 4, 2>, 2, 4, 2);
        run(tma_stream_kernel<2, 4, 4>, 2, 4, 4);
        run(tma_stream_kernel<3, 2, 2>, 3, 2, 2);
    }
    return 0;
}
