"""tcgen05 GEMM (C ABI vitae_gemm_bf16) vs torch fp32 matmul on the same bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def _check(out, ref, tol=2e-2):
    err = (out.float() - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-6
    assert err / den < tol, f"max abs err {err} vs ref max {den}"


SHAPES = [(128, 128, 64), (128, 64, 128), (256, 256, 256), (516, 2304, 768), (516, 768, 3072), (200, 72, 200),
          (2052, 1536, 512), (33, 8, 24), (1000, 1000, 1000 - 8)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tile_n", [64, 128, 256])
def test_forward_layout(M, N, K, tile_n):
    from vit_ae_plus_plus_b200 import ops
    a, w = _mk((M, K), 1), _mk((N, K), 2)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, M, N, K, out_f32=out, tile_n=tile_n, split_k=1)
    torch.cuda.synchronize()
    _check(out, a.float() @ w.float().t(), 1e-3)


@pytest.mark.parametrize("M,N,K", [(516, 768, 2304), (2052, 512, 2048), (200, 72, 136), (128, 128, 64)])
@pytest.mark.parametrize("tile_n", [64, 128])
def test_dgrad_layout(M, N, K, tile_n):
    # dx[M, N=in] = dy[M, K=out] @ W[K=out, N=in]  -> B stored [K, N] (MN-major)
    from vit_ae_plus_plus_b200 import ops
    dy, w = _mk((M, K), 3), _mk((K, N), 4)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(dy, w, M, N, K, b_mn_major=True, out_f32=out, tile_n=tile_n, split_k=1)
    torch.cuda.synchronize()
    _check(out, dy.float() @ w.float(), 1e-3)


@pytest.mark.parametrize("M,N,K", [(2304, 768, 516), (512, 2048, 2052), (768, 16384, 512), (72, 136, 200), (128, 128, 64)])
@pytest.mark.parametrize("tile_n", [64, 128, 256])
def test_wgrad_layout(M, N, K, tile_n):
    # dW[M=out, N=in] = dy[K=tok, M]^T @ x[K=tok, N] -> both operands MN-major
    from vit_ae_plus_plus_b200 import ops
    dy, x = _mk((K, M), 5), _mk((K, N), 6)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(dy, x, M, N, K, a_mn_major=True, b_mn_major=True, out_f32=out, tile_n=tile_n, split_k=1)
    torch.cuda.synchronize()
    _check(out, dy.float().t() @ x.float(), 1e-3)


def test_a_mn_b_k_layout():
    from vit_ae_plus_plus_b200 import ops
    M, N, K = 192, 136, 264
    a, b = _mk((K, M), 7), _mk((N, K), 8)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, M, N, K, a_mn_major=True, out_f32=out, tile_n=64, split_k=1)
    torch.cuda.synchronize()
    _check(out, a.float().t() @ b.float().t(), 1e-3)


@pytest.mark.parametrize("split", [2, 3, 7, 16])
def test_split_k(split):
    from vit_ae_plus_plus_b200 import ops
    M, N, K = 512, 768, 4096
    a, w = _mk((M, K), 9), _mk((N, K), 10)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, M, N, K, bias=bias, out_f32=out, tile_n=128, split_k=split)
    torch.cuda.synchronize()
    _check(out, a.float() @ w.float().t() + bias, 1e-3)


def test_epilogue_bias_residual_rows_and_bf16():
    from vit_ae_plus_plus_b200 import ops
    M, N, K = 300, 256, 192
    a, w = _mk((M, K), 11), _mk((N, K), 12)
    bias = torch.randn(N, device="cuda")
    table = torch.randn(50, N, device="cuda")
    add_rows = torch.randint(0, 50, (M,), device="cuda", dtype=torch.int32)
    perm = torch.randperm(M + 20, device="cuda")[:M].to(torch.int32)
    out = torch.zeros(M + 20, N, device="cuda")
    out16 = torch.zeros(M + 20, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, M, N, K, bias=bias, addend=table, add_rows=add_rows, out_f32=out, out_bf16=out16, out_rows=perm,
             alpha=0.5)
    torch.cuda.synchronize()
    ref = 0.5 * (a.float() @ w.float().t()) + bias + table[add_rows.long()]
    _check(out[perm.long()], ref, 1e-3)
    _check(out16[perm.long()], ref, 1e-2)
    untouched = torch.ones(M + 20, dtype=torch.bool, device="cuda")
    untouched[perm.long()] = False
    assert out[untouched].abs().max().item() == 0


def test_epilogue_gelu_pair_dgelu_accumulate_alpha_ptr():
    from vit_ae_plus_plus_b200 import ops
    M, N, K = 260, 512, 128
    a, w = _mk((M, K), 13, 0.5), _mk((N, K), 14, 0.2)
    bias = torch.randn(N, device="cuda") * 0.1
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    act = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, M, N, K, bias=bias, out_bf16=pre, out_gelu_bf16=act)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    _check(pre, ref, 1e-2)
    _check(act, torch.nn.functional.gelu(ref), 1e-2)
    # dgelu multiply + residual accumulate + device-scalar alpha
    src = _mk((M, N), 15)
    acc = torch.randn(M, N, device="cuda")
    acc0 = acc.clone()
    scal = torch.tensor([3.0], device="cuda")
    ops.gemm(a, w, M, N, K, dgelu_src=src, out_f32=acc, accumulate=True, alpha_ptr=scal)
    torch.cuda.synchronize()
    x = src.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ref2 = acc0 + 3.0 * (a.float() @ w.float().t()) * x.grad
    _check(acc, ref2, 1e-3)


def test_auto_config_matches():
    from vit_ae_plus_plus_b200 import ops
    for (M, N, K) in [(512, 768, 16384), (2052, 512, 16384), (516, 768, 768)]:
        a, w = _mk((M, K), 16, 0.1), _mk((N, K), 17, 0.1)
        out = torch.empty(M, N, device="cuda")
        ops.gemm(a, w, M, N, K, out_f32=out)
        torch.cuda.synchronize()
        _check(out, a.float() @ w.float().t(), 2e-3)


@pytest.mark.parametrize("B,V,p,keep,tile_n", [(2, 32, 8, 16, 128), (1, 64, 16, 16, 256), (3, 32, 8, 40, 128)])
def test_pred_gemm_with_fused_masked_mse(B, V, p, keep, tile_n):
    """vitae_gemm_pred_mse: decoder_pred (model/vit_autoenc.py:198) + masked patch-reconstruction loss (:226-227) in one
    kernel, against torch fp32 on the same bf16 operands: pred, the unscaled gradient g, and the loss value."""
    from vit_ae_plus_plus_b200 import ops
    C, Dd = 4, 64
    g_, L, P = V // p, (V // p) ** 3, p ** 3 * 4
    M = B * (L + 1)
    hN, W = _mk((M, Dd), 11), _mk((P, Dd), 12, 0.2)
    bias = torch.randn(P, device="cuda") * 0.1
    vol = torch.randn(B, C, V, V, V, device="cuda")
    mask = torch.zeros(B, L, device="cuda")
    for b in range(B):
        mask[b, torch.randperm(L, device="cuda")[:L - keep]] = 1.0
    mask_sum = float(B * (L - keep))
    pred = torch.empty(M, P, device="cuda", dtype=torch.bfloat16)
    gq = torch.full((M, P), float("nan"), device="cuda", dtype=torch.bfloat16)
    part = torch.empty(ops.pred_mse_partial_floats(M, P, tile_n), device="cuda")
    loss_out = torch.empty(2, device="cuda")
    ops.gemm_pred_mse(hN, W, bias, B, L, Dd, vol, mask, p, mask_sum, pred, gq, part, tile_n)
    ops.pred_mse_finalize(part, P, mask_sum, loss_out)
    torch.cuda.synchronize()
    ref = hN.float() @ W.float().t() + bias                                            # [M, P]
    _check(pred, ref, 1e-2)
    target = vol.reshape(B, C, g_, p, g_, p, g_, p).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(B, L, P)   # vit_autoenc.py:100-113
    rp = ref.reshape(B, L + 1, P)[:, 1:, :]
    loss_ref = (((rp - target) ** 2).mean(-1) * mask).sum() / mask.sum()
    assert abs(loss_out[0].item() - loss_ref.item()) < 1e-4 * abs(loss_ref.item()) and loss_out[1].item() == mask_sum
    g_ref = torch.zeros(B, L + 1, P, device="cuda")
    g_ref[:, 1:, :] = 2.0 * (rp - target) * mask[:, :, None] / (P * mask_sum)
    _check(gq.reshape(B, L + 1, P), g_ref, 1e-2)
    assert torch.isfinite(gq.float()).all() and (gq.reshape(B, L + 1, P)[:, 0].float() == 0).all()
