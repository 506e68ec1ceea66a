"""Pins oracle/mae_oracle.py (the CPU checker) against golden vectors produced by the unmodified reference
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import mae_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["tiny", "tiny_edge", "tiny_r50", "small"]


def _load(name):
    return np.load(os.path.join(GOLD, f"mae_{name}.npz"))


def _inputs(g):
    cfg = O.CONFIGS[str(g["config_name"])]
    pseed, iseed, nseed, batch, stride, steps = [int(v) for v in g["meta"]]
    V = cfg["volume_size"]
    x = torch.randn(batch, cfg["in_chans"], V, V, V, generator=torch.Generator().manual_seed(iseed))
    _, L, _ = O.geometry(cfg)
    torch.manual_seed(nseed)
    noise = torch.rand(batch, L)
    return cfg, O.init_params(cfg, pseed), x, noise, nseed, stride, steps


@pytest.mark.parametrize("name", CASES)
def test_forward_backward_matches_reference(name):
    g = _load(name)
    cfg, P, x, noise, _, stride, _ = _inputs(g)
    ew = float(g["edge_map_weight"])
    losses, pred, mask, grads = O.forward_backward(x, P, cfg, float(g["mask_ratio"]), noise, ew, with_edge=True)
    np.testing.assert_allclose([float(l) for l in losses], g["losses"], rtol=2e-5, atol=1e-6)
    np.testing.assert_array_equal(mask.numpy(), g["mask"])                       # bit-exact (index work)
    np.testing.assert_allclose(pred.reshape(-1)[::stride].numpy(), g["pred_sample"], rtol=1e-4, atol=2e-5)
    assert abs(pred.double().sum().item() - float(g["pred_sum"])) <= 1e-3 * max(1.0, abs(float(g["pred_sum"])))
    np.testing.assert_allclose((pred.double() ** 2).sum().item(), float(g["pred_sqsum"]), rtol=1e-5)
    names = [str(n) for n in g["grad_names"]]
    assert sorted(names) == sorted(grads.keys())          # every trainable param gets a gradient
    for n in names:
        gr = grads[n].double().reshape(-1)
        ref_norm = float(g[f"gnorm/{n}"])
        assert abs(gr.norm().item() - ref_norm) <= 2e-4 * ref_norm + 1e-9, n
        np.testing.assert_allclose(gr[:16].numpy(), g[f"ghead/{n}"], rtol=2e-3, atol=2e-5 * max(ref_norm, 1e-6), err_msg=n)


@pytest.mark.parametrize("name", ["tiny", "tiny_edge"])
def test_adamw_loss_curve_matches_reference(name):
    g = _load(name)
    cfg, P, x, _, nseed, _, steps = _inputs(g)
    ew = float(g["edge_map_weight"])
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in P.items()}
    opt = torch.optim.AdamW(O.weight_decay_groups(list(leaves.items()), 0.05), lr=1e-3, betas=(0.9, 0.95))
    _, L, _ = O.geometry(cfg)
    curve = []
    for step in range(steps):
        torch.manual_seed(nseed + 1 + step)
        noise = torch.rand(x.shape[0], L)
        losses, _, _, _ = O.forward(x, leaves, cfg, float(g["mask_ratio"]), noise, ew, with_edge=True)
        opt.zero_grad(set_to_none=True)
        losses[0].backward()
        opt.step()
        curve.append(float(losses[0]))
    np.testing.assert_allclose(curve, g["curve"], rtol=2e-4)


def _contr_inputs(g):
    cfg = O.CONFIGS[str(g["config_name"])]
    pseed, iseed, nseed, batch, stride, _ = [int(v) for v in g["meta"]]
    V, C = cfg["volume_size"], cfg["in_chans"]
    gen = torch.Generator().manual_seed(iseed)
    x1 = torch.randn(batch, C, V, V, V, generator=gen)
    x2 = x1 + 0.1 * torch.randn(batch, C, V, V, V, generator=gen)
    _, L, _ = O.geometry(cfg)
    torch.manual_seed(nseed)
    noise1, noise2 = torch.rand(batch, L), torch.rand(batch, L)      # view 1 first, then view 2 (vit_autoenc.py:272,277)
    P = dict(O.init_params(cfg, pseed))
    P.update(O.init_predictor_params(cfg, pseed))
    return cfg, P, x1, x2, noise1, noise2, stride


def test_contrastive_wrapper_matches_reference():
    """ContrastiveMAEViT forward (7-tuple) + the loop's contrastive loss + backward of the sum, against the unmodified
    reference (golden: oracle/make_golden.py::run_reference_contrastive)."""
    g = np.load(os.path.join(GOLD, "contr_small.npz"))
    cfg, P, x1, x2, n1, n2, stride = _contr_inputs(g)
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in P.items()}
    losses, pred, mask, p1, p2, z1, z2 = O.forward_contrastive(x1, x2, leaves, cfg, float(g["mask_ratio"]), n1, n2,
                                                               0.0, with_edge=True)
    contr = O.contrastive_loss(p1, p2, z1, z2, float(g["contr_weight"]))
    (losses[0] + contr).backward()
    np.testing.assert_allclose([float(l) for l in losses], g["losses"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(float(contr), float(g["contr"]), rtol=1e-4)
    np.testing.assert_array_equal(mask.numpy(), g["mask"])
    np.testing.assert_allclose(p1.detach().reshape(-1)[::stride].numpy(), g["p1_sample"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(p2.detach().reshape(-1)[::stride].numpy(), g["p2_sample"], rtol=1e-3, atol=1e-4)
    assert not z1.requires_grad and not z2.requires_grad
    names = [str(n) for n in g["grad_names"]]
    got = {k for k, v in leaves.items() if v.grad is not None}
    assert sorted(names) == sorted(got)
    for n in names:
        gr = leaves[n].grad.double().reshape(-1)
        ref_norm = float(g[f"gnorm/{n}"])
        assert abs(gr.norm().item() - ref_norm) <= 5e-4 * ref_norm + 1e-9, n
        np.testing.assert_allclose(gr[:16].numpy(), g[f"ghead/{n}"], rtol=5e-3, atol=5e-5 * max(ref_norm, 1e-6), err_msg=n)


@pytest.mark.parametrize("name,cfgname,gp", [("small_gp", "small", True), ("tiny_cls", "tiny", False)])
def test_vit_feature_extraction_matches_reference(name, cfgname, gp):
    """Encoder-only inference (VisionTransformer3D.forward_features, model/vit.py:265-284) against the reference golden."""
    g = np.load(os.path.join(GOLD, "vit_features.npz"))
    cfg = O.CONFIGS[cfgname]
    P = O.init_vit_params(cfg, 2, gp, seed=0)
    V, C = cfg["volume_size"], cfg["in_chans"]
    x = torch.randn(3, C, V, V, V, generator=torch.Generator().manual_seed(1))
    np.testing.assert_allclose(O.vit_forward_features(x, P, cfg, gp).numpy(), g[f"feat/{name}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.vit_forward(x, P, cfg, gp).numpy(), g[f"logits/{name}"], rtol=1e-5, atol=1e-6)
    assert sorted(P) == sorted(O.vit_param_names(cfg, gp))


def test_pos_embed_matches_reference():
    g = np.load(os.path.join(GOLD, "pos_embed.npz"))
    for key in g.files:
        _, d, gs = key.split("_")
        pe = O.sincos_pos_embed_3d(int(d), int(gs), cls_token=True)
        ref = g[key]
        got = pe if ref.ndim == 2 else pe.reshape(-1)[::13]
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
        assert np.all(pe[0] == 0)                        # cls row is zeros (vit_helpers.py:27-28)


def test_patchify_roundtrip_and_mask_invariants():
    # invariants lifted from visualization/custom_operation_checks.py:16-20 and vit_autoenc.py:136-155
    x = torch.randn(2, 3, 24, 24, 24)
    assert torch.equal(O.unpatchify(O.patchify(x, 8), 8), x)
    noise = torch.rand(4, 64)
    t = torch.randn(4, 64, 8)
    for r in (0.25, 0.5, 0.75):
        xm, mask, ids_restore = O.random_masking(t, r, noise)
        keep = int(64 * (1 - r))
        assert xm.shape[1] == keep
        assert torch.all(mask.sum(1) == 64 - keep)
        assert torch.equal(torch.sort(ids_restore, dim=1).values, torch.arange(64).expand(4, -1))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference only exists in the build container")
def test_live_reference_small_batch3():
    """Live check against the unmodified reference on an input that is NOT in the golden set."""
    from oracle import ref_shim
    cfg = O.CONFIGS["tiny"]
    model = ref_shim.build_reference_model(cfg)
    P = O.init_params(cfg, 7)
    model.load_state_dict(P, strict=False)
    x = torch.randn(3, 1, 32, 32, 32, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(99)
    losses_ref, pred_ref, mask_ref = model(x, mask_ratio=0.6, edge_map_weight=0.02)
    torch.manual_seed(99)
    noise = torch.rand(3, 64)
    losses, pred, mask, _ = O.forward(x, P, cfg, 0.6, noise, 0.02, with_edge=True)
    assert torch.equal(mask, mask_ref)
    torch.testing.assert_close(pred, pred_ref, rtol=1e-5, atol=1e-5)
    for a, b in zip(losses, losses_ref):
        torch.testing.assert_close(a, b.detach().reshape(()), rtol=1e-5, atol=1e-6)


def _reference_method(path, name):
    """Compiles ONE method of a reference file without importing the module (its imports need torchio / the datasets):
    the function's own source lines run unmodified in a namespace that holds what the file imports for it."""
    import ast
    src = open(path).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "sqrt": torch.sqrt, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference only exists in the build container")
@pytest.mark.parametrize("dataset,z,mode", [("egd_dataset/egd.py", True, "z_score_channel"), ("egd_dataset/egd.py", False, "min_max"),
                                            ("brats_dataset/brats.py", True, "z_score_sample"),
                                            ("brats_dataset/brats.py", False, "min_max")])
def test_normalize_volume_matches_reference_dataset_code(dataset, z, mode):
    """oracle.normalize_volume against Dataset._normalize_data of the unmodified reference (egd.py:44-50, brats.py:26-32)."""
    import types
    ref = _reference_method(os.path.join("/root/reference/dataset", dataset), "_normalize_data")
    g = torch.Generator().manual_seed(3)
    vol = (torch.rand(4, 16, 16, 16, generator=g) * 4000).round()          # scanner-like intensities
    vol[1] = vol[1] * 0.25 + 100
    want = ref(types.SimpleNamespace(use_z_score=z), vol.clone())
    got = O.normalize_volume(vol, mode)
    torch.testing.assert_close(got, want, rtol=0, atol=0)
