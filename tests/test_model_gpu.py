"""End-to-end parity of the B200 path (MaskedAutoencoderViT -> engine -> C ABI kernels) against the CPU oracle and the
golden vectors produced by the unmodified reference (tests/golden/*.npz, oracle/make_golden.py).

Same seeded parameters, inputs and mask noise on both sides.  Tolerances (BASELINE.json north_star): the kernels compute
with bf16 tensor-core operands and fp32 accumulation -> 1e-2 relative for loss / pred / gradients (max-norm relative
for tensors); mask / ids are index work -> bit-exact."""
import argparse
import os
from functools import partial

import numpy as np
import pytest
import torch
from torch import nn

from oracle import mae_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-2


def build(cfg, P=None):
    from vit_ae_plus_plus_b200.model.vit_autoenc import MaskedAutoencoderViT
    m = MaskedAutoencoderViT(**cfg, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                             args=argparse.Namespace(perceptual_weight=0, use_imagenet=False))
    if P is not None:
        m.load_state_dict(P, strict=True)
    return m.cuda()


def relmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def relnorm(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).norm().item() / (b.norm().item() + 1e-30)


def run_ours(cfg, P, x, noise, mask_ratio, edge_w=0.0):
    m = build(cfg, P)
    losses, pred, mask = m(x.cuda(), mask_ratio=mask_ratio, edge_map_weight=edge_w, noise=noise)
    losses[0].backward()
    grads = {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters() if p.grad is not None}
    return [l.detach().cpu() for l in losses], pred.detach().float().cpu(), mask.cpu(), grads, m


def _record(tag, per_param, extra=None):
    """Per-parameter gradient errors of the BASELINE-sized runs go to gpurun_out/ (merged back from the GPU box; the
    summaries under profiles/ are made from these files)."""
    import json
    out_dir = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        rows = sorted(((e, n) for n, e in per_param.items()), reverse=True)
        with open(os.path.join(out_dir, f"grad_errors_{tag}.json"), "w") as f:
            json.dump({"tag": tag, "metric": "||g - g_oracle||_2 / ||g_oracle||_2 per parameter tensor", "worst": rows[:12],
                       "n_params": len(rows), "n_above_1e-2": sum(e > 1e-2 for e, _ in rows),
                       "median": rows[len(rows) // 2][0], **(extra or {})}, f, indent=1)
    except OSError:
        pass


# North-star tolerance for the bf16 path: 1e-2.  Tensors allowed more, and why:
#   * ``attn.qkv.bias``: its K third has an exactly-zero true gradient (softmax is invariant to a per-query shift of the
#     scores, model/vit.py:117-119), so what any implementation produces there is rounding noise; it is compared on the
#     Q and V thirds only (``_signal``).
def _signal(name, t):
    if name.endswith("attn.qkv.bias"):
        d = t.shape[0] // 3
        return torch.cat([t[:d], t[2 * d:]])
    return t


def check_against_oracle(cfg, seed, B, mask_ratio, edge_w=0.0, grad_tol=1e-2, tag=None):
    P = O.init_params(cfg, seed)
    V, C = cfg["volume_size"], cfg["in_chans"]
    x = torch.randn(B, C, V, V, V, generator=torch.Generator().manual_seed(seed + 1))
    _, L, _ = O.geometry(cfg)
    torch.manual_seed(seed + 2)
    noise = torch.rand(B, L)
    l_ref, pred_ref, mask_ref, g_ref = O.forward_backward(x, P, cfg, mask_ratio, noise, edge_w, with_edge=edge_w != 0)
    l, pred, mask, g, _ = run_ours(cfg, P, x, noise, mask_ratio, edge_w)
    assert torch.equal(mask, mask_ref)                                   # bit-exact index work
    assert abs(l[2].item() - l_ref[2].item()) <= TOL * abs(l_ref[2].item()), (l[2], l_ref[2])
    assert abs(l[0].item() - l_ref[0].item()) <= TOL * abs(l_ref[0].item())
    assert relmax(pred, pred_ref) < TOL, relmax(pred, pred_ref)
    assert sorted(g) == sorted(g_ref)                                    # every trainable parameter gets a gradient
    per_param = {n: relnorm(_signal(n, g[n]), _signal(n, g_ref[n])) for n in g}
    if tag:
        _record(tag, per_param, {"batch": B, "mask_ratio": mask_ratio, "recon_rel_err": abs(l[2].item() - l_ref[2].item()) / abs(l_ref[2].item()),
                                 "pred_relmax": relmax(pred, pred_ref)})
    worst = max((e, n) for n, e in per_param.items())
    assert worst[0] < grad_tol, (worst, sorted(((e, n) for n, e in per_param.items()), reverse=True)[:6])
    return worst


@pytest.mark.parametrize("name,B,ratio", [("tiny", 2, 0.75), ("tiny", 3, 0.5), ("tiny", 1, 0.25), ("small", 2, 0.75)])
def test_small_configs_match_oracle(name, B, ratio):
    check_against_oracle(O.CONFIGS[name], 11, B, ratio)


def test_tiny_with_edge_map_term_matches_oracle():
    # SURVEY row f-1 (interim torch ops on the kernels' pred): loss[0] = w*edge + recon, gradient flows through pred
    check_against_oracle(O.CONFIGS["tiny"], 5, 2, 0.75, edge_w=0.05, grad_tol=1.5e-2)


@pytest.mark.parametrize("name", ["tiny", "tiny_r50", "small"])
def test_matches_reference_golden(name):
    """Against outputs of the *unmodified reference* (not the oracle): tests/golden/mae_<name>.npz."""
    g = np.load(os.path.join(GOLD, f"mae_{name}.npz"))
    cfg = O.CONFIGS[str(g["config_name"])]
    pseed, iseed, nseed, batch, stride, _ = [int(v) for v in g["meta"]]
    V = cfg["volume_size"]
    x = torch.randn(batch, cfg["in_chans"], V, V, V, generator=torch.Generator().manual_seed(iseed))
    _, L, _ = O.geometry(cfg)
    torch.manual_seed(nseed)
    noise = torch.rand(batch, L)
    losses, pred, mask, grads, _ = run_ours(cfg, O.init_params(cfg, pseed), x, noise, float(g["mask_ratio"]))
    np.testing.assert_array_equal(mask.numpy(), g["mask"])
    assert abs(losses[2].item() - float(g["losses"][2])) <= TOL * abs(float(g["losses"][2]))
    ref = torch.from_numpy(g["pred_sample"])
    assert relmax(pred.reshape(-1)[::stride], ref) < TOL
    for n in [str(s) for s in g["grad_names"]]:
        ref_norm = float(g[f"gnorm/{n}"])
        assert abs(grads[n].double().norm().item() - ref_norm) <= 2e-2 * ref_norm + 1e-9, n


def test_vit_base_one_volume_matches_oracle():
    """configs[1] geometry (ViT-B/16, 128^3 x 4) at batch 1: the oracle's fp32 CPU forward+backward takes seconds."""
    worst = check_against_oracle(O.CONFIGS["vit_base_128"], 3, 1, 0.75, tag="vit_base_128_r75")
    print("worst grad rel err", worst)


def test_vit_large_96_one_volume_matches_oracle():
    """configs[3] geometry (ViT-L/16, 96^3 x 4: D = 1024, 24 blocks, L = 216, keep = 54) at batch 1."""
    worst = check_against_oracle(O.CONFIGS["vit_large_96"], 4, 1, 0.75, tag="vit_large_96_r75")
    print("worst grad rel err", worst)


@pytest.mark.parametrize("ratio", [0.25, 0.50])
def test_vit_base_mask_ratio_sweep_matches_oracle(ratio):
    """configs[4]: mask-ratio sweep on ViT-B 128^3 (keep = 384 / 256 kept patches, ragged encoder lengths 385 / 257):
    loss, pred, mask AND every parameter gradient of one volume against the oracle."""
    check_against_oracle(O.CONFIGS["vit_base_128"], 6, 1, ratio, tag=f"vit_base_128_r{int(ratio * 100)}")


def test_gradient_accumulation_and_zero_grad():
    cfg = O.CONFIGS["tiny"]
    P = O.init_params(cfg, 2)
    m = build(cfg, P)
    x = torch.randn(2, 1, 32, 32, 32, generator=torch.Generator().manual_seed(1)).cuda()
    noise = torch.rand(2, 64, generator=torch.Generator().manual_seed(2))
    m(x, noise=noise)[0][0].backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m(x, noise=noise)[0][0].backward()                       # accumulates in place (accum_iter > 1 in the reference loop)
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert relmax(p.grad, 2 * g1[n]) < 1e-5, n
    for p in m.parameters():
        p.grad = None                                        # optimizer.zero_grad(set_to_none=True)
    m(x, noise=noise)[0][0].backward()
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert torch.equal(p.grad, g1[n]), n             # deterministic kernels: bitwise repeatable
    # grad scaling (GradScaler multiplies the loss by a device scalar, utils/misc.py:258)
    for p in m.parameters():
        p.grad = None
    (m(x, noise=noise)[0][0] * 1024.0).backward()
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert relmax(p.grad, 1024.0 * g1[n]) < 1e-5, n


def test_no_cpu_fallback():
    from vit_ae_plus_plus_b200._lib import VitaeError
    from vit_ae_plus_plus_b200.model.vit_autoenc import MaskedAutoencoderViT
    m = MaskedAutoencoderViT(**O.CONFIGS["tiny"], args=argparse.Namespace(perceptual_weight=0))
    with pytest.raises(VitaeError):
        m(torch.zeros(1, 1, 32, 32, 32))
    with pytest.raises(VitaeError):
        m.blocks[0](torch.zeros(1, 17, 128))


def test_api_helpers_roundtrip():
    cfg = O.CONFIGS["tiny"]
    P = O.init_params(cfg, 4)
    m = build(cfg, P)
    x = torch.randn(2, 1, 32, 32, 32, generator=torch.Generator().manual_seed(3))
    assert torch.equal(m.unpatchify(m.patchify(x)), x)                       # custom_operation_checks.py:16-20
    assert torch.equal(m.patchify(x), O.patchify(x, 8))
    noise = torch.rand(2, 64, generator=torch.Generator().manual_seed(9))
    latent, mask, ids_restore = m.forward_encoder(x.cuda(), 0.75, noise=noise)
    lat_ref, mask_ref, ids_ref = O.forward_encoder(x, P, cfg, 0.75, noise)
    assert torch.equal(mask.cpu(), mask_ref) and torch.equal(ids_restore.cpu(), ids_ref)
    assert relmax(latent, lat_ref) < TOL
    pred = m.forward_decoder(latent, ids_restore)
    pred_ref = O.forward_decoder(lat_ref, P, cfg, ids_ref)
    assert relmax(pred, pred_ref) < TOL
    losses = m.forward_loss(x.cuda(), pred, mask)
    ref = O.masked_mse(pred.cpu(), O.patchify(x, 8), mask_ref)
    assert abs(losses[2].item() - ref.item()) < 1e-4 * ref.item()
    with torch.no_grad():
        l2, p2, m2 = m(x.cuda(), noise=noise)
    assert relmax(p2, pred_ref) < TOL and p2.dtype == torch.float32


def test_factory_and_named_ctors():
    from vit_ae_plus_plus_b200.model import model_factory, vit_autoenc
    args = argparse.Namespace(model="mae_vit_base_patch16", volume_size=32, in_channels=2, patch_size=16,
                              perceptual_weight=0, use_imagenet=False)
    m = model_factory.get_models("autoenc", args)
    assert isinstance(m, vit_autoenc.MaskedAutoencoderViT)
    ref_keys = set(O.param_shapes(dict(O.CONFIGS["vit_base_128"], volume_size=32, in_chans=2)))
    assert set(m.state_dict()) == ref_keys
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) > 100e6
    with pytest.raises(NotImplementedError):
        model_factory.get_models("contrastive", args)      # fine-tuning classifier: outside the B200 path, says so


# ------------------------------------------------------------------------------------------------ contrastive wrapper
def _contr_case(seed=0):
    cfg = O.CONFIGS["small"]
    g = np.load(os.path.join(GOLD, "contr_small.npz"))
    pseed, iseed, nseed, batch, stride, _ = [int(v) for v in g["meta"]]
    V, C = cfg["volume_size"], cfg["in_chans"]
    gen = torch.Generator().manual_seed(iseed)
    x1 = torch.randn(batch, C, V, V, V, generator=gen)
    x2 = x1 + 0.1 * torch.randn(batch, C, V, V, V, generator=gen)
    _, L, _ = O.geometry(cfg)
    torch.manual_seed(nseed)
    n1, n2 = torch.rand(batch, L), torch.rand(batch, L)
    P = dict(O.init_params(cfg, pseed))
    P.update(O.init_predictor_params(cfg, pseed))
    return cfg, g, P, x1, x2, n1, n2, stride


@pytest.mark.parametrize("graphs", [False, True])
def test_contrastive_wrapper_matches_oracle_and_reference_golden(graphs):
    """ContrastiveMAEViT (model/vit_autoenc.py:241-285): 7-tuple forward, the loop's contrastive loss
    (utils/train_one_epoch.py:113-114), backward of the sum through both encoder passes -- against the oracle (pinned to
    the reference by tests/test_oracle_golden.py::test_contrastive_wrapper_matches_reference) and the reference golden."""
    from vit_ae_plus_plus_b200.model.vit_autoenc import ContrastiveMAEViT
    cfg, g, P, x1, x2, n1, n2, stride = _contr_case()
    w = float(g["contr_weight"])
    m = ContrastiveMAEViT(**cfg, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                          args=argparse.Namespace(perceptual_weight=0, use_imagenet=False))
    missing, unexpected = m.load_state_dict(P, strict=False)
    assert not unexpected and all(k.startswith("predictor.1.") for k in missing), (missing, unexpected)
    m = m.cuda().train()
    m.use_cuda_graph = graphs
    crit = torch.nn.CosineSimilarity(dim=1)
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in P.items()}
    lo, pred_o, mask_o, p1o, p2o, z1o, z2o = O.forward_contrastive(x1, x2, leaves, cfg, 0.75, n1, n2, 0.0, with_edge=False)
    (lo[0] + O.contrastive_loss(p1o, p2o, z1o, z2o, w)).backward()
    for rep in range(3 if graphs else 1):           # eager, capture, replay
        m.zero_grad(set_to_none=True)
        losses, pred, mask, p1, p2, z1, z2 = m(x1.cuda(), x2.cuda(), mask_ratio=0.75, edge_map_weight=0, noise=n1, noise2=n2)
        contr = w * (-(crit(p1, z2).mean() + crit(p2, z1).mean()) * 0.5)
        (losses[0] + contr).backward()
        torch.cuda.synchronize()
        assert torch.equal(mask.cpu(), mask_o)
        assert abs(losses[2].item() - float(g["losses"][2])) < TOL * float(g["losses"][2])      # reference golden
        assert abs(contr.item() - float(g["contr"])) < 2e-2 * abs(float(g["contr"]))
        assert p1.shape == p1o.shape and z2.shape == z2o.shape and not z1.requires_grad
        assert relmax(p1, p1o) < 2e-2 and relmax(p2, p2o) < 2e-2 and relmax(z2, z2o) < TOL
        np.testing.assert_allclose(p1.detach().float().cpu().reshape(-1)[::stride].numpy(), g["p1_sample"], rtol=0,
                                   atol=2e-2 * np.abs(g["p1_sample"]).max())
        errs = []
        for n, p in m.named_parameters():
            if n in O.FROZEN:
                assert p.grad is None
                continue
            ref = leaves[n].grad
            assert p.grad is not None, n
            errs.append((relnorm(p.grad, ref), n))
        print("worst gradient errors", sorted(errs)[-4:])
        for n, p in m.named_parameters():
            if n in O.FROZEN:
                continue
            # The predictor's BatchNorm1d normalises over only B*Ne = 14 token rows here and feeds a cosine loss: that
            # chain amplifies bf16 rounding of the encoder activations ~30x (CPU experiment: rounding the block inputs to
            # bf16 in the fp32 oracle alone moves these gradients by 6 %, against 0.2 % for the MAE-only loss).  The
            # plumbing itself is checked at 3e-2 by test_latent_gradient_paths_match_oracle below.
            assert relnorm(p.grad, leaves[n].grad) < 0.12, (n, sorted(errs)[-4:])
            assert abs(p.grad.double().norm().item() - float(g[f"gnorm/{n}"])) < 0.1 * float(g[f"gnorm/{n}"]) + 1e-6, n


@pytest.mark.parametrize("graphs", [False, True])
def test_latent_gradient_paths_match_oracle(graphs):
    """The engine plumbing behind the contrastive model, without the ill-conditioned predictor: loss = recon(view 1) +
    <G1, latent(view 1)> + <G2, latent(view 2)>.  Exercises the second activation arena, the latent gradient added at the
    encoder norm (LayerNorm backward with two upstream gradients), the encoder-only backward and its accumulation."""
    from vit_ae_plus_plus_b200.model import vit_autoenc as VA
    cfg, g, P, x1, x2, n1, n2, _ = _contr_case()
    m = VA.ContrastiveMAEViT(**cfg, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                             args=argparse.Namespace(perceptual_weight=0, use_imagenet=False))
    m.load_state_dict(P, strict=False)
    m = m.cuda().train()
    m.use_cuda_graph = graphs
    eng = m.engine()
    keep = m._len_keep(0.75)
    gen = torch.Generator().manual_seed(77)
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in P.items()}
    lat1o, mask_o, ids = O.forward_encoder(x1, leaves, cfg, 0.75, n1)
    pred_o = O.forward_decoder(lat1o, leaves, cfg, ids)
    recon_o = O.masked_mse(pred_o, O.patchify(x1, cfg["patch_size"]), mask_o)
    lat2o, _, _ = O.forward_encoder(x2, leaves, cfg, 0.75, n2)
    G1 = torch.randn(lat1o.shape[0] * lat1o.shape[1], lat1o.shape[2], generator=gen) * 0.02
    G2 = torch.randn(G1.shape, generator=gen) * 0.02
    (recon_o + (lat1o.reshape(G1.shape) * G1).sum() + (lat2o.reshape(G2.shape) * G2).sum()).backward()
    for rep in range(3 if graphs else 1):
        m.zero_grad(set_to_none=True)
        eng.use_graphs = graphs
        recon, _edge, pred, mask, lat1, lat2, _p1, _p2 = VA._ContrastiveStep.apply(m.cls_token, m, x1.cuda(), x2.cuda(),
                                                                                   n1.cuda(), n2.cuda(), keep, False)
        (recon + (lat1 * G1.cuda()).sum() + (lat2 * G2.cuda()).sum()).backward()
        torch.cuda.synchronize()
        assert relmax(lat1, lat1o.reshape(G1.shape)) < TOL and relmax(lat2, lat2o.reshape(G2.shape)) < TOL
        worst = max((relnorm(p.grad, leaves[n].grad), n) for n, p in m.named_parameters()
                    if n not in O.FROZEN and not n.startswith("predictor."))
        assert worst[0] < 3e-2, worst


def test_contrastive_model_through_the_training_loop_api():
    """contr_mae_vit_base_patch16 via get_models + train_one_stage_epoch's call sequence on a down-sized volume (the 7-tuple
    branch of the loop, utils/train_one_epoch.py:51-58).  The predictor runs on the kernels and its parameters live in the
    engine's flat buffers; the fused optimizer's result matches the reference's unfused GradScaler + torch AdamW sequence."""
    from vit_ae_plus_plus_b200.model import model_factory
    from vit_ae_plus_plus_b200.utils import misc, train_one_epoch as T
    args = argparse.Namespace(model="contr_mae_vit_base_patch16", volume_size=32, in_channels=1, patch_size=16,
                              perceptual_weight=0, use_imagenet=False, mask_ratio=0.5, accum_iter=1, contr_weight=0.001,
                              lr=1e-4, min_lr=0.0, warmup_epochs=0, epochs=2)
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(2, 1, 32, 32, 32, generator=g), torch.randn(2, 1, 32, 32, 32, generator=g), torch.zeros(2))
               for _ in range(3)]
    results = {}
    for fused in (True, False):
        torch.manual_seed(11)                                  # same init and the same on-device mask noise sequence
        model = model_factory.get_models("autoenc_contr", args).cuda()
        opt = torch.optim.AdamW(misc.add_weight_decay(model, 0.05), lr=1e-4, betas=(0.9, 0.95))
        scaler = misc.NativeScalerWithGradNormCount()
        scaler.allow_fused = fused
        before = model.predictor[0].weight.detach().clone()
        stats = T.train_one_stage_epoch(model, batches, opt, torch.device("cuda"), 0, scaler, log_writer=None, args=args,
                                        edge_map_weight=0.01)
        assert set(stats) >= {"lr", "edge_map_loss", "reconstruction_loss", "perceptual_loss", "contr_loss", "loss"}
        assert np.isfinite(stats["loss"]) and stats["contr_loss"] != 0.0 and stats["edge_map_loss"] > 0.0
        assert not torch.equal(before, model.predictor[0].weight.detach())
        assert (scaler._fused is not None) == fused
        results[fused] = (stats, {k: v.detach().clone() for k, v in model.state_dict().items()})
        if fused:
            fo = model.engine().fused_optimizer()
            # the predictor's parameters live in the engine's flat buffers like all others: no second ("extras") buffer
            flat = model.engine().flat
            assert fo.ex_total == 0 and model.predictor[0].weight.data_ptr() == flat.v32["predictor.0.weight"].data_ptr()
            st = opt.state[model.predictor[3].bias]
            assert st["exp_avg"].abs().sum().item() > 0          # optimizer state stays visible in torch's layout
    (sf, pf), (su, pu) = results[True], results[False]
    assert abs(sf["loss"] - su["loss"]) < 1e-4 * abs(su["loss"])
    for k in pu:
        if pu[k].dtype.is_floating_point and pu[k].numel() > 1 and "running" not in k:
            d = (pf[k].double() - pu[k].double()).abs().max().item()
            # elements whose true gradient is ~0 are moved by +-lr per step in a direction set by rounding noise (Adam
            # normalises it): allow the 3 steps x lr = 3e-4 such elements can drift apart, on top of 5e-3 relative
            assert d < 5e-3 * (pu[k].double().abs().max().item() + 1e-12) + 3.5e-4, k


# ------------------------------------------------------------------------------------------------ encoder-only inference (f-3)
@pytest.mark.parametrize("name,cfgname,gp", [("small_gp", "small", True), ("tiny_cls", "tiny", False)])
def test_vit_feature_extractor_matches_oracle_and_reference_golden(name, cfgname, gp):
    """VisionTransformer3D.forward_features / forward (model/vit.py:265-297) on the kernels, all patches, no masking."""
    from vit_ae_plus_plus_b200.model.vit import VisionTransformer3D
    g = np.load(os.path.join(GOLD, "vit_features.npz"))
    cfg = O.CONFIGS[cfgname]
    P = O.init_vit_params(cfg, 2, gp, seed=0)
    m = VisionTransformer3D(volume_size=cfg["volume_size"], in_chans=cfg["in_chans"], num_classes=2, patch_size=cfg["patch_size"],
                            embed_dim=cfg["embed_dim"], depth=cfg["depth"], num_heads=cfg["num_heads"], mlp_ratio=cfg["mlp_ratio"],
                            global_pool=gp, norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_path_rate=0.1)
    assert sorted(m.state_dict()) == sorted(O.vit_param_names(cfg, gp))              # the reference's state_dict keys
    m.load_state_dict(P, strict=True)
    m = m.cuda().eval()
    V, C = cfg["volume_size"], cfg["in_chans"]
    x = torch.randn(3, C, V, V, V, generator=torch.Generator().manual_seed(1))
    with pytest.raises(Exception):
        m.forward_features(x.cuda())                 # autograd enabled: inference-only path says so
    with torch.no_grad():
        f = m.forward_features(x.cuda())
        y = m(x.cuda())
    ref = torch.from_numpy(g[f"feat/{name}"])
    assert relmax(f, ref) < TOL, relmax(f, ref)
    assert relmax(f, O.vit_forward_features(x, P, cfg, gp)) < TOL
    assert relmax(y, torch.from_numpy(g[f"logits/{name}"])) < 2e-2


def test_mae_checkpoint_hand_off_to_the_feature_extractor():
    """The post-training flow of k_fold_cross_valid_combined_brats.py:219-253: get_models('vit'), load the MAE checkpoint's
    'model' dict with strict=False, the asserted missing-key set, then forward_features under no_grad."""
    from vit_ae_plus_plus_b200.model import model_factory
    args = argparse.Namespace(model="mae_vit_base_patch16", volume_size=32, in_channels=1, patch_size=16, perceptual_weight=0,
                              use_imagenet=False, nb_classes=2, global_pool=True, drop_path=0.1)
    mae = model_factory.get_models("autoenc", args).cuda()
    ckpt = {k: v.detach().cpu().clone() for k, v in mae.state_dict().items()}
    vit = model_factory.get_models("vit", args)
    msg = vit.load_state_dict(ckpt, strict=False)
    assert set(msg.missing_keys) == {"head.weight", "head.bias", "fc_norm.weight", "fc_norm.bias"}
    assert all(k.startswith(("decoder_", "mask_token", "norm.")) for k in msg.unexpected_keys)
    vit = vit.cuda().eval()
    x = torch.randn(2, 1, 32, 32, 32, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.no_grad():
        f = vit.forward_features(x)
    assert f.shape == (2, 768) and torch.isfinite(f).all()
    # same encoder weights: the MAE's own encoder with nothing masked sees the same token stream
    P = {k: v.detach().cpu() for k, v in vit.state_dict().items()}
    cfg = dict(volume_size=32, patch_size=16, in_chans=1, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4)
    assert relmax(f, O.vit_forward_features(x.cpu(), P, cfg, True)) < TOL
