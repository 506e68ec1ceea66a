/*
 * vitae_b200.h -- C ABI of libvitae_b200.so: the sm_100a kernels behind the 3D ViT masked-autoencoder
 * training path of ViT-AE++ (reference: chinmay5/vit_ae_plus_plus, pure PyTorch).
 *
 * The reference has no native interface: its "FFI" for this path is torch's ATen dispatch (nn.Conv3d,
 * nn.Linear, nn.LayerNorm, softmax, nn.GELU, argsort/gather, elementwise loss).  Each entry point below names the
 * reference call site(s) it replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * binding a reference maintainer adds.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is a DEVICE pointer borrowed for the duration of the enqueue;
 *   - `stream` is a cudaStream_t passed as void*; work is only enqueued -- no allocation, no synchronisation,
 *     CUDA-graph capturable;
 *   - returns 0 on success, negative on error; vitae_last_error() returns a thread-local message;
 *   - row-major contiguous layouts unless a leading dimension is given; bf16 = __nv_bfloat16 bit pattern;
 *   - "rows" index arrays are int32.
 */
#ifndef VITAE_B200_H
#define VITAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITAE_ABI_VERSION 1

int vitae_abi_version(void);
const char* vitae_last_error(void);
/* 0 when the current device is compute capability 10.x (B200); negative otherwise. */
int vitae_check_device(void);
/* Number of kernels this library has enqueued in this process (diagnostic; bench.py's gpu_launches). */
long long vitae_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Dense contraction on the 5th-gen tensor cores (tcgen05.mma, TMA-staged 128B-swizzled tiles, TMEM accumulator).
 * Replaces every nn.Linear on the path -- model/vit.py:107,114 (qkv), :109,122 (proj), :84-86,91-95 (fc1/fc2),
 * model/vit_autoenc.py:40,181 (decoder_embed), :52,198 (decoder_pred) -- the im2col form of the Conv3d patch embed
 * (model/vit.py:65,72) and their autograd backward (dgrad / wgrad) run by loss.backward() (utils/misc.py:258).
 *
 *   acc[m,n] = sum_k A(m,k) * B(n,k)                      bf16 x bf16 -> fp32
 *   A: a_mn_major == 0: stored [M, K] row-major, leading dim lda;   == 1: stored [K, M] row-major (A transposed)
 *   B: b_mn_major == 0: stored [N, K] row-major (nn.Linear weight); == 1: stored [K, N] row-major
 *   forward  y = x W^T : A=x (0), B=W (0)           dgrad dx = dy W : A=dy (0), B=W (1)
 *   wgrad    dW = dy^T x: A=dy (1), B=x (1)
 *
 * Epilogue, per element (r = out_rows ? out_rows[m] : m,  ra = add_rows ? add_rows[m] : m):
 *   v = alpha * (alpha_ptr ? *alpha_ptr : 1) * acc + (bias ? bias[n] : 0) + (addend ? addend[ra*ldadd + n] : 0)
 *   if dgelu_src: v *= gelu'(dgelu_src[m*ld_dgelu + n])            (erf GELU derivative)
 *   if out_f32 : out_f32[r*ld_f32 + n]  = (accumulate ? old : 0) + v
 *   if out_bf16: out_bf16[r*ld_bf16 + n] = bf16(v)
 *   if out_gelu_bf16: out_gelu_bf16[r*ld_bf16 + n] = bf16(gelu(v))   (erf GELU, model/vit.py:81,92)
 * Requirements: N % 8 == 0, leading dims % 8 == 0, 16-byte aligned base pointers (operands, outputs, bias, addend).
 * The common epilogues of the training step (bf16 out; bf16 + GELU twin; fp32 out + same-row addend; bf16 + fp32 out;
 * fp32 out / accumulate; bf16 out with GELU') run inside the GEMM kernel (TMA-store epilogue).  split_k > 1, row maps
 * (out_rows / add_rows) and any other combination go through fp32 slabs and a finalize kernel (deterministic fixed-order
 * reduction, no atomics) and need workspace >= vitae_gemm_workspace_bytes_for(...) bytes, 16-byte aligned.
 */
typedef struct vitae_gemm_epilogue {
    float alpha;
    const float* alpha_ptr;
    const float* bias;
    const float* addend;
    const int32_t* add_rows;
    int32_t ldadd;
    const void* dgelu_src; /* bf16 */
    int32_t ld_dgelu;
    float* out_f32;
    int32_t ld_f32;
    int32_t accumulate;
    void* out_bf16;      /* bf16 */
    void* out_gelu_bf16; /* bf16 */
    int32_t ld_bf16;
    const int32_t* out_rows;
} vitae_gemm_epilogue;

int vitae_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N,
                    int K, const vitae_gemm_epilogue* ep, int tile_n, int split_k, void* workspace,
                    size_t workspace_bytes, void* stream);
size_t vitae_gemm_workspace_bytes(int M, int N, int split_k);
/* workspace bytes vitae_gemm_bf16 needs for this epilogue / operand layout / split (0 = none) */
size_t vitae_gemm_workspace_bytes_for(const vitae_gemm_epilogue* ep, int a_mn_major, int b_mn_major, int M, int N,
                                      int split_k);

/* Times the (tile_n, split_k) candidates of this GEMM on the device and returns the fastest (CUDA events; the GEMMs of
 * the step are short and latency bound, so the best tiling depends on the exact shape).  With flush_buf (>= 2x L2, e.g.
 * 256 MB) every timed launch runs on a flushed L2 -- the state the step's weight matrices are in; NULL: back-to-back warm
 * launches.  Re-runs the GEMM many times: outputs are overwritten with identical values; not allowed for accumulate
 * epilogues or on a capturing stream.  Candidates whose slabs exceed `workspace_bytes` are skipped.  Synchronises `stream`. */
int vitae_gemm_autotune(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N,
                        int K, const vitae_gemm_epilogue* ep, void* workspace, size_t workspace_bytes, void* flush_buf,
                        size_t flush_bytes, void* stream, int* best_tile_n, int* best_split_k);

/* ------------------------------------------------------------------------------------------------------------
 * LayerNorm (biased variance, eps inside sqrt) -- nn.LayerNorm at model/vit.py:131,135,140-143 and
 * model/vit_autoenc.py:36,51,174,195.  x fp32 [rows, D]; y bf16 [rows, D] (GEMM operand) and/or y_f32; mean/rstd
 * fp32 [rows] saved for backward.  Warp-shuffle row statistics.
 */
int vitae_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                        float* mean, float* rstd, int rows, int D, float eps, void* stream);
/* dx_out = (dx_in ? dx_in : 0) + LN'(dy); dy is bf16 (dy_bf16) or fp32 (dy_f32) or, when both are given, their sum (two
 * consumers of the normalised tensor: decoder + contrastive predictor, model/vit_autoenc.py:272-283); also emits a bf16 copy of dx_out
 * (operand of the next dgrad/wgrad GEMM) when dx_out_bf16 != NULL.  Row-wise only (it is on the dgrad critical path). */
int vitae_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma,
                        const float* mean, const float* rstd, const float* dx_in, float* dx_out, void* dx_out_bf16,
                        int rows, int D, void* stream);
/* Column reductions of the same backward, off the critical path, in one launch: dgamma = sum dy*xhat, dbeta = sum dy,
 * dbias = column sums of dx_out (= gradient of the bias that was added to this residual stream: proj.bias / fc2.bias,
 * model/vit.py:142-143); NULL outputs are skipped; accumulate adds into them.  workspace:
 * vitae_layernorm_param_grads_workspace_bytes(rows, D) bytes, ZERO-FILLED before its first use (ticket counters; every call
 * leaves them zero) and not shared by calls that may run concurrently.  Deterministic (fixed-order slice reduction). */
int vitae_layernorm_param_grads(const void* dy_bf16, const float* dy_f32, const float* x, const float* mean,
                                const float* rstd, const float* dx_out, void* workspace, float* dgamma, float* dbeta,
                                float* dbias, int accumulate, int rows, int D, void* stream);
size_t vitae_layernorm_param_grads_workspace_bytes(int rows, int D);
int vitae_layernorm_bwd_blocks(int rows);
/* out_k[c] = (accumulate ? out_k[c] : 0) + sum_blk partials[k][blk][c] for k = 0..2; NULL outputs are skipped. */
int vitae_reduce_partials(const float* partials, int nblk, int D, float* out0, float* out1, float* out2,
                          int accumulate, void* stream);

/* Column sums in one launch: out[c] = (accumulate ? out[c] : 0) + sum_r in[r*ld + c]  (bias gradients of
 * nn.Linear, model/vit.py:84-86,107-109).  Exactly one of in_bf16 / in_f32 is non-NULL; cols % 8 == 0, ld % 8 == 0.
 * workspace: vitae_colsum_workspace_bytes(rows, cols) bytes, ZERO-FILLED before its first use (ticket counters;
 * every call leaves them zero again) and not shared by calls that may run concurrently.  Deterministic.
 * scale_ptr (optional): device scalar multiplied into the sums (an upstream loss gradient that lives on the device). */
int vitae_colsum(const void* in_bf16, const float* in_f32, int rows, int cols, int ld, float* out, int accumulate,
                 void* workspace, const float* scale_ptr, void* stream);
size_t vitae_colsum_workspace_bytes(int rows, int cols);

/* All column reductions of one transformer block's backward in one launch: each job is a plain column sum
 * (x == NULL: out0[c] = sum_r a[r*ld + c], a bf16 or fp32) or a LayerNorm affine-gradient reduction (out0 = dgamma =
 * sum_r dy*xhat, out1 = dbeta = sum_r dy, with dy = a (+ a2) and xhat = (x - mean) * rstd; x fp32 [rows, cols]).  Up to 6
 * jobs over matrices with the same number of rows; jobs is a HOST array read during the call; accumulate adds into the
 * outputs.  workspace: vitae_block_colreduce_workspace_bytes(rows, sum of the jobs' cols each rounded up to 256) bytes,
 * ZERO-FILLED before its first use (ticket counters, self-resetting), not shared by concurrent calls.  Deterministic. */
typedef struct vitae_col_job {
    const void* a;
    const float* a2;
    const float* x;
    const float* mean;
    const float* rstd;
    float* out0;
    float* out1;
    int32_t cols, ld, a_is_bf16, reserved;
} vitae_col_job;
int vitae_block_colreduce(const vitae_col_job* jobs, int njobs, int rows, int accumulate, void* workspace,
                          size_t workspace_bytes, void* stream);
size_t vitae_block_colreduce_workspace_bytes(int rows, int total_cols_padded);

/* ------------------------------------------------------------------------------------------------------------
 * Fused multi-head self-attention, flash-style (scores never leave the SM) -- model/vit.py:112-121:
 *   qkv bf16 [B, N, 3, H, hd] (the qkv Linear output as-is), out bf16 [B, N, H*hd], lse fp32 [B, H, N].
 *   softmax(q k^T * scale) v, online softmax in fp32 with warp-shuffle row reductions.  hd in {16, 32, 64}.
 */
int vitae_attention_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, float scale,
                        void* stream);
/* dqkv bf16 [B, N, 3, H, hd]; delta fp32 [B, H, N] is scratch (rowsum(dout*out)). */
int vitae_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                        void* dqkv, int B, int N, int H, int hd, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Per-sample random masking -- model/vit_autoenc.py:130-155: ids_shuffle = argsort(noise) (stable), ids_restore =
 * argsort(ids_shuffle), mask[b, l] = 1 if patch l is removed.  noise fp32 [B, L]; outputs int32 / fp32 [B, L].
 */
int vitae_random_masking(const float* noise, int32_t* ids_shuffle, int32_t* ids_restore, float* mask, int B, int L,
                         int len_keep, void* stream);

/* Row maps derived from ids_shuffle that express the reference's cat / gather / repeat token shuffling
 * (model/vit_autoenc.py:147 keep-gather, :168-170 cls prepend, :184-190 mask tokens + unshuffle + pos) as row
 * scatters in GEMM epilogues.  Ne = keep+1, Nd = L+1; all outputs int32:
 *   enc_tok_rows[b*keep+j] = b*Ne+1+j              enc_cls_rows[b] = b*Ne
 *   pe_pos_rows[b*keep+j]  = 1+ids_shuffle[b,j]    (row of pos_embed added to kept patch j)
 *   dec_rows_of_enc[b*Ne+t]     = t==0 ? b*Nd : b*Nd+1+ids_shuffle[b,t-1]   (decoder row of encoder token t)
 *   dec_pos_rows_of_enc[b*Ne+t] = t==0 ? 0    : 1+ids_shuffle[b,t-1]        (row of decoder_pos_embed)
 *   masked_dec_rows[b*(L-keep)+i] = b*Nd+1+ids_shuffle[b,keep+i], masked_pos_rows[..] = 1+ids_shuffle[b,keep+i] */
int vitae_build_row_maps(const int32_t* ids_shuffle, int B, int L, int keep, int32_t* enc_tok_rows,
                         int32_t* enc_cls_rows, int32_t* pe_pos_rows, int32_t* dec_rows_of_enc,
                         int32_t* dec_pos_rows_of_enc, int32_t* masked_dec_rows, int32_t* masked_pos_rows,
                         void* stream);

/* Patch gather for the Conv3d(k=s=p) patch embed -- model/vit.py:65,72 + the torch.gather of kept tokens at
 * model/vit_autoenc.py:147: row (b*keep + j) of `cols` (bf16 [B*keep, C*p^3], K order (c,pz,py,px) = conv weight
 * order) is patch ids_shuffle[b, j] of volume b.  vol fp32 [B, C, V, V, V]. */
int vitae_im2col_patches(const float* vol, const int32_t* ids_shuffle, void* cols_bf16, int B, int C, int V, int p,
                         int L, int keep, void* stream);

/* Token assembly helpers (model/vit_autoenc.py:168-170 cls prepend, :184-190 mask tokens + unshuffle + pos). */
/* dst[row_idx[i] (or i), :] = src0[(src0_rows ? src0_rows[i] : 0), :] + src1[(src1_rows ? src1_rows[i] : 0), :] */
int vitae_fill_rows(float* dst, const int32_t* row_idx, int nrows, int D, const float* src0, const int32_t* src0_rows,
                    const float* src1, const int32_t* src1_rows, void* stream);
/* dst_bf16[i, :] = bf16(src[row_idx[i], :])  (gathers gradient rows into a GEMM operand); dst_f32 optional */
int vitae_gather_rows(const float* src, const int32_t* row_idx, int nrows, int D, void* dst_bf16, float* dst_f32,
                      void* stream);
/* out[:] = (accumulate ? out : 0) + sum_i src[row_idx[i], :]  (cls_token / mask_token gradients), single block */
int vitae_sum_rows(const float* src, const int32_t* row_idx, int nrows, int D, float* out, int accumulate,
                   void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Masked patch-reconstruction loss -- model/vit_autoenc.py:100-113 (patchify) + :226-227 (masked MSE), without
 * materialising the patchified target: pred [B, Nd=L+1, P] (row 0 of each sample = cls, ignored; bf16 or fp32),
 * vol fp32 [B, C, V, V, V], P = p^3*C ordered (pz,py,px,c).  loss_out[0] = sum_masked mean_P (pred-target)^2 / sum(mask),
 * loss_out[1] = sum(mask) (consumed by the backward as mask_sum).  patch_sums: fp32 scratch [B*L].
 */
int vitae_masked_mse_fwd(const void* pred, int pred_is_bf16, const float* vol, const float* mask, float* patch_sums,
                         float* loss_out, int B, int C, int V, int p, void* stream);
/* dpred[b, 1+l, :] = mask[b,l] * 2 (pred - target) / (P * sum(mask)) * (*dloss); cls rows and kept patches are
 * written as zeros.  dpred bf16 [B, Nd, P]. */
int vitae_masked_mse_bwd(const void* pred, int pred_is_bf16, const float* vol, const float* mask,
                         const float* mask_sum, const float* dloss, void* dpred_bf16, int B, int C, int V, int p,
                         void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Edge-map loss (shipped default use_edge_map = yes) -- model/vit_autoenc.py:221-224 with
 * model/model_utils/sobel_filter.py:10-45 and gaussian_filter.py:5-26:
 *   raw_edge = mean_{b,v} ( sum_c |sobel(unpatchify(pred)_c)|(v) - sum_c |sobel(blur(target_c))|(v) )^2
 * scratch: vitae_edge_scratch_floats(B, C, V) floats, shared by the three calls of one step (target, fwd, bwd).
 * vitae_edge_target: E_tgt fp32 [B, V^3] from the fp32 volume [B, C, V, V, V]; taps: HOST array of the ntaps (odd, <= 16)
 *   normalised 1-D Gaussian taps (the reference's dense ks^3 kernel is their outer product).
 * vitae_edge_loss_fwd: loss_out[0] = raw_edge for pred bf16 [B, L+1, P] (cls row first); writes the residual
 *   D = E_pred - E_tgt to resid fp32 [B, V^3] and keeps F = D * g / |g| (bf16, 3 per channel voxel) in scratch.
 * vitae_edge_loss_bwd: dpred_bf16 [B, L+1, P] += (*upstream) * d raw_edge / d pred  (|grad| = 0 contributes 0); reads F
 *   from scratch (resid is only checked for NULL).  V and p multiples of 4, C in {1, 2, 4}. */
size_t vitae_edge_scratch_floats(int B, int C, int V);
int vitae_edge_target(const float* vol, const float* taps, int ntaps, float* scratch, float* E_tgt, int B, int C, int V,
                      void* stream);
int vitae_edge_loss_fwd(const void* pred_bf16, const float* E_tgt, float* scratch, float* resid, float* loss_out, int B,
                        int C, int V, int p, void* stream);
int vitae_edge_loss_bwd(const float* resid, const float* scratch, const float* upstream, void* dpred_bf16, int B, int C,
                        int V, int p, void* stream);

/* Pulls up to 12 device regions into L2 (cp.async.bulk.prefetch.L2): the next transformer block's weights and saved
 * activations while the current block computes (every kernel of the path is a few microseconds long and would otherwise
 * start with a cold HBM load).  ptrs / bytes: HOST arrays read during the call.  No reference counterpart (performance
 * plumbing of the B200 path). */
int vitae_prefetch_l2(const void* const* ptrs, const size_t* bytes, int n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Parameter plumbing: fp32 master weights -> flat bf16 shadow (one launch for all tensors).
 * table: device array of ntensors records {uint64 src_ptr, uint64 dst_elem_offset, uint64 numel}. */
int vitae_cast_params_bf16(const void* table, int ntensors, void* dst_bf16, long long total_elems, void* stream);

/* Fused AdamW over a flat fp32 parameter/gradient/moment buffer (torch.optim.AdamW semantics,
 * k_fold_cross_valid_combined_brats.py:168-169): grad is first multiplied by (*inv_scale) (GradScaler unscale);
 * if (*found_inf != 0) the step is skipped.  wd_mask[i]==0 => no weight decay for element block. */
int vitae_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                     long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                     float bias_corr1, float bias_corr2, const float* inv_scale, const float* found_inf, void* stream);

/* Optimizer step over the flat buffers in three launches, replacing GradScaler.unscale_ /
 * get_grad_norm_ / GradScaler.step(AdamW) / GradScaler.update (utils/misc.py:257-292) and torch.optim.AdamW built at
 * k_fold_cross_valid_combined_brats.py:168-169.
 * ctl: device fp32[8] = {[0] loss scale, [1] growth tracker, [2] found_inf of this step, [3] 1/scale used by this
 * step, [4] unscaled global gradient L2 norm, [5] optimizer steps taken (skipped steps do not count)}.
 * vitae_optim_prepare: one pass over grad (and the optional second region grad2: parameters that live outside the main flat
 * buffer, e.g. the contrastive predictor) -> ctl[2..4]; then GradScaler.update on ctl[0..1] (use_scaler != 0) and
 * ctl[5] += 1 unless the step is skipped.  workspace: vitae_optim_workspace_bytes() bytes.
 * vitae_adamw_flat: AdamW over [0, n) (n % 64 == 0); group_of_chunk[i >> 6] = parameter group of element i
 * (>= ngroups: frozen / padding, left untouched); hyper: HOST fp32 [ngroups][8] = {lr, beta1, beta2, eps,
 * weight_decay, 0, 0, 0}, copied into the launch parameters during the call (ngroups <= 8); bias corrections use ctl[5]; the step is skipped when ctl[2] != 0; also writes the bf16
 * shadow param_bf16 (GEMM operands) when non-NULL.  All pointers may be offset to a 64-aligned sub-range of the flat
 * buffers (group_of_chunk offset by start/64): the step can be issued layer group by layer group.  max_blocks > 0 caps the
 * grid (a step that overlaps the next forward must leave SM slots to it); 0 = default. */
int vitae_optim_prepare(const float* grad, long long n, const float* grad2, long long n2, float* ctl, float* workspace,
                        float growth_factor, float backoff_factor, int growth_interval, int use_scaler, void* stream);
size_t vitae_optim_workspace_bytes(void);
/* The two halves of vitae_optim_prepare on their own.  vitae_grad_sqnorm: vitae_grad_sqnorm_blocks(n, max_blocks)
 * partial sums of squares of a gradient slice -> partials[] (the backward computes them per stage, underneath the later
 * stages, so that the step does not start with an 80 us pass over all gradients, utils/misc.py:280-292).
 * vitae_optim_finalize: ctl[2..5] and the GradScaler.update of ctl[0..1] from npartials partial sums (any layout). */
int vitae_grad_sqnorm_blocks(long long n, int max_blocks);
int vitae_grad_sqnorm(const float* grad, long long n, float* partials, int max_blocks, void* stream);
int vitae_optim_finalize(const float* partials, int npartials, float* ctl, float growth_factor, float backoff_factor,
                         int growth_interval, int use_scaler, void* stream);
int vitae_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                     long long n, const unsigned char* group_of_chunk, const float* hyper, int ngroups,
                     const float* ctl, int max_blocks, void* stream);

/* Data parallel, sharded optimizer step over NVLink peer memory (replaces "all-reduce the gradients, then run
 * the same AdamW on every rank" -- the reference itself has no distributed step: its scripts never wrap the model in DDP,
 * k_fold_cross_valid_combined_brats.py:154).  The flat gradient / master / shadow buffers of all ranks are symmetric
 * allocations mapped into every rank's address space; *_ptrs[r] = rank r's base address of the buffer (2 <= world <= 8).
 * Ownership is block-cyclic and never changes: granule q = elements [q << granule_shift, (q + 1) << granule_shift) belongs to
 * rank q % world, so every slice [lo, hi) of the flat index space (multiples of 64; the backward's stage slices) splits
 * evenly.  All calls below work on rank `rank`'s part of [lo, hi):
 *   vitae_dp_owned_elems: how many elements that is;
 *   vitae_dp_reduce_shard: own gradient copy := inv_world * sum over ranks (peer copies pulled with plain loads over NVLink,
 *     summed in rank order); partials[vitae_dp_reduce_shard_blocks(...)] = per-block sums of squares of the result (0 blocks:
 *     the rank owns nothing of the slice and nothing is launched);
 *   vitae_sum_partials: out[0] = sum of n partials (accumulated in double): each rank publishes one value;
 *   vitae_optim_finalize_peers: vitae_optim_finalize over npartials values from EACH rank's partial buffer
 *     (partial_ptrs[r], read over NVLink, summed rank-major in double: every rank computes the same control block);
 *   vitae_adamw_shard: vitae_adamw_flat (grad / exp_avg / exp_avg_sq / group_of_chunk / f32_chunk: this rank's whole flat
 *     buffers, element 0), then the all-gather from inside the same kernel: the new bf16 shadow is stored into every rank's
 *     shadow buffer, the new fp32 master into this rank's and -- for the 64-element chunks with f32_chunk[chunk] != 0 (tensors
 *     the kernels read in fp32: biases, LayerNorm affine, tokens) -- into every peer's.  The moments, and the fp32 master of
 *     the other chunks, stay with the owner.
 * The caller orders these against the other ranks (a cross-rank barrier before the reduce, before the finalize and after
 * the update). */
long long vitae_dp_owned_elems(long long lo, long long hi, int granule_shift, int world, int rank);
int vitae_dp_reduce_shard_blocks(long long lo, long long hi, int granule_shift, int world, int rank, int max_blocks);
int vitae_dp_reduce_shard(void* const* grad_ptrs, int world, int rank, long long lo, long long hi, int granule_shift,
                          float inv_world, float* partials, int max_blocks, void* stream);
int vitae_sum_partials(const float* partials, int n, float* out, void* stream);
int vitae_optim_finalize_peers(void* const* partial_ptrs, int world, int npartials, float* ctl, float growth_factor,
                               float backoff_factor, int growth_interval, int use_scaler, void* stream);
int vitae_adamw_shard(void* const* param_ptrs, void* const* param_bf16_ptrs, int world, int rank, long long lo, long long hi,
                      int granule_shift, const float* grad, float* exp_avg, float* exp_avg_sq,
                      const unsigned char* group_of_chunk, const unsigned char* f32_chunk, const float* hyper, int ngroups,
                      const float* ctl, int max_blocks, void* stream);

/* decoder_pred (model/vit_autoenc.py:198) with the masked patch-reconstruction loss (:226-227, utils/custom_loss.py's role in
 * the north star) fused into its epilogue.  hN bf16 [B*(L+1), Dd] (decoder_norm output, cls row first per sample), W bf16
 * [P, Dd], bias fp32 [P], P = p^3 * C with C == 4 and p % 8 == 0; vol fp32 [B, C, V, V, V] read in place; mask fp32 [B, L];
 * mask_sum = sum(mask) (known on the host: B * (L - keep)).  Writes pred bf16 [B*(L+1), P], g bf16 (same shape) =
 * 2 (pred - target) mask / (P mask_sum) with zero rows for kept patches and the cls token (the backward GEMMs multiply the
 * upstream gradient in through alpha_ptr) and partials fp32 [vitae_pred_mse_partial_floats(M, P, tile_n)].
 * vitae_pred_mse_finalize: loss_out[0] = sum(partials) / (P mask_sum), loss_out[1] = mask_sum, fixed order. */
size_t vitae_pred_mse_partial_floats(int M, int P, int tile_n);
int vitae_gemm_pred_mse(const void* hN, const void* W, const float* bias, int B, int L, int Dd, const float* vol,
                        const float* mask, int C, int V, int p, float mask_sum, void* pred_bf16, void* g_bf16,
                        float* partials, int tile_n, void* stream);
int vitae_pred_mse_finalize(const float* partials, long long n, int P, float mask_sum, float* loss_out, void* stream);

/* fp32 <-> bf16 copies of a flat gradient slice: the data-parallel exchange (one all-reduce of the trainable parameters'
 * gradients per optimizer step; the reference scripts never wrap the model in DDP,
 * k_fold_cross_valid_combined_brats.py:154) moves bf16 over NVLink and the slice is widened again afterwards.
 * max_blocks > 0 caps the grid (the copies run beside the backward); 0 = default. */
int vitae_cast_f32_to_bf16(const float* src, void* dst_bf16, long long n, int max_blocks, void* stream);
int vitae_cast_bf16_to_f32(const void* src_bf16, float* dst, long long n, int max_blocks, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * On-device intensity normalisation of raw volumes, replacing Dataset._normalize_data
 * (dataset/egd_dataset/egd.py:44-50, dataset/brats_dataset/brats.py:26-32), so that the host ships the storage type.
 * raw: [B, C, voxels] of raw_type (0 f32, 1 f16, 2 bf16, 3 u16, 4 i16, 5 u8); out: fp32, same shape.
 * mode 0: z-score per channel with the unbiased variance (egd.py:45-47); 1: z-score per sample (brats.py:27-29);
 * 2: min-max of the sample to [-1, 1] (egd.py:48-50).  stats (optional): fp32 [groups][2] = {offset a, scale s} of
 * out = (x - a) * s per group (B*C groups in mode 0, B otherwise).  voxels % 8 == 0.
 * workspace: vitae_ingest_workspace_bytes(B, C) bytes.  Deterministic (fixed reduction order). */
size_t vitae_ingest_workspace_bytes(int B, int C);
int vitae_ingest_normalize(const void* raw, int raw_type, float* out, int B, int C, long long voxels, int mode,
                           void* workspace, float* stats, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Contrastive head: predictor = Linear(no bias) -> BatchNorm1d -> ReLU -> Linear
 * (model/vit_autoenc.py:263-268) and the loop's cosine loss (utils/train_one_epoch.py:32,113-114).  The Linear layers
 * are vitae_gemm_bf16 calls.
 * vitae_bn_relu_fwd: h fp32 [M, D] (output of the first Linear) -> act bf16 = relu(batchnorm(h)) with the batch
 *   statistics of the M rows (training mode, biased variance, eps); mean / rstd fp32 [D] are kept for the backward;
 *   running_mean / running_var (both or neither) are updated with `momentum` and the unbiased variance, like torch.
 * vitae_bn_relu_bwd: dact bf16 (gradient w.r.t. act) -> dh bf16 (gradient w.r.t. h); dgamma / dbeta fp32 [D] (+)=.
 * vitae_cosine_loss_fwd: loss[0] = -weight * (mean_i cos(p1_i, z2_i) + mean_i cos(p2_i, z1_i)) / 2 over fp32 [M, D]
 *   rows, cos with torch.nn.CosineSimilarity's eps = 1e-8; workspace: vitae_cosine_loss_workspace_floats(M) floats,
 *   kept for the backward.  vitae_cosine_loss_bwd: dp1 / dp2 fp32 [M, D] = upstream[0] * d loss / d p (z detached). */
int vitae_bn_relu_fwd(const float* h, const float* gamma, const float* beta, float eps, void* act_bf16, float* mean,
                      float* rstd, float* running_mean, float* running_var, float momentum, int M, int D, void* stream);
int vitae_bn_relu_bwd(const void* dact_bf16, const float* h, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, void* dh_bf16, float* dgamma, float* dbeta, int accumulate, int M, int D,
                      void* stream);
size_t vitae_cosine_loss_workspace_floats(int M);
int vitae_cosine_loss_fwd(const float* p1, const float* z2, const float* p2, const float* z1, int M, int D, float weight,
                          float* workspace, float* loss, void* stream);
int vitae_cosine_loss_bwd(const float* p1, const float* z2, const float* p2, const float* z1, int M, int D, float weight,
                          const float* workspace, const float* upstream, float* dp1, float* dp2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITAE_B200_H */
