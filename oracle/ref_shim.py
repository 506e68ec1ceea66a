"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (chinmay5/vit_ae_plus_plus).

Imports the reference's ``model.vit_autoenc`` from ``/root/reference`` (read-only, only present in
the build container, never on the GPU box) so that ``oracle/make_golden.py`` can generate the golden
vectors under ``tests/golden/`` and so that container-only tests can pin ``oracle/mae_oracle.py``
against the real thing.  Nothing in the product package imports this file.

Two shims are needed, no source edits:
  1. ``timm`` is not installed; ``model/vit.py:8-9`` and ``model/model_utils/vit_helpers.py:6`` import four
     helpers from it that are never executed on the MAE path -> stub modules in ``sys.modules``.
  2. ``vgg_perceptual_loss.__init__`` loads ``model/ckp-399.pth`` (``model/model_utils/perceptual_loss.py:20-23``)
     which is not shipped -> ``torch.load`` of that path returns ``{}`` (it is loaded with strict=False; the
     perceptual weight is 0 in every config we pin).
"""
import argparse
import importlib
import math
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("VITAE_REF_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "vit_autoenc.py"))


def _install_timm_stub():
    if "timm" in sys.modules:
        return

    def named_apply(fn, module, name="", depth_first=True, include_root=False):
        if not depth_first and include_root:
            fn(module=module, name=name)
        for child_name, child in module.named_children():
            child_name = ".".join((name, child_name)) if name else child_name
            named_apply(fn, child, child_name, depth_first, True)
        if depth_first and include_root:
            fn(module=module, name=name)
        return module

    def lecun_normal_(t):
        fan_in = t.shape[1] if t.dim() > 1 else t.shape[0]
        return torch.nn.init.normal_(t, std=math.sqrt(1.0 / fan_in))

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    helpers = types.ModuleType("timm.models.helpers")
    layers = types.ModuleType("timm.models.layers")
    weight_init = types.ModuleType("timm.models.layers.weight_init")
    helpers.named_apply = named_apply
    helpers.adapt_input_conv = lambda in_chans, w: w
    weight_init.trunc_normal_ = torch.nn.init.trunc_normal_
    weight_init.lecun_normal_ = lecun_normal_
    layers.weight_init = weight_init
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    models.helpers = helpers
    models.layers = layers
    timm.models = models
    for name, mod in [("timm", timm), ("timm.models", models), ("timm.models.helpers", helpers),
                      ("timm.models.layers", layers), ("timm.models.layers.weight_init", weight_init)]:
        sys.modules[name] = mod


_REF_MODULES = {}


def load_reference():
    """Returns the reference's ``model.vit_autoenc`` module (imported under its own package names)."""
    if "vit_autoenc" in _REF_MODULES:
        return _REF_MODULES["vit_autoenc"]
    if not reference_available():
        raise FileNotFoundError(f"reference not found at {REF_ROOT} (it only exists in the build container)")
    _install_timm_stub()
    # the reference uses top-level package names 'model', 'utils', 'environment_setup'
    for name in ("model", "utils"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REF_ROOT):
            raise RuntimeError(f"a non-reference module named {name!r} is already imported")
    sys.path.insert(0, REF_ROOT)
    real_load = torch.load

    def patched_load(f, *a, **k):
        if isinstance(f, (str, os.PathLike)) and str(f).endswith("ckp-399.pth"):
            return {}
        return real_load(f, *a, **k)

    torch.load = patched_load
    try:
        mod = importlib.import_module("model.vit_autoenc")
    finally:
        sys.path.remove(REF_ROOT)
    mod._vitae_patched_load = patched_load
    mod._vitae_real_load = real_load
    _REF_MODULES["vit_autoenc"] = mod
    return mod


def reference_args(**over):
    ns = argparse.Namespace(use_imagenet=False, perceptual_weight=0)
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


def build_reference_model(cfg: dict, contrastive: bool = False):
    """cfg: same keys as oracle.mae_oracle.CONFIGS entries."""
    from functools import partial
    mod = load_reference()
    real = torch.load
    torch.load = mod._vitae_patched_load
    try:
        m = (mod.ContrastiveMAEViT if contrastive else mod.MaskedAutoencoderViT)(
            volume_size=cfg["volume_size"], patch_size=cfg["patch_size"], in_chans=cfg["in_chans"],
            embed_dim=cfg["embed_dim"], depth=cfg["depth"], num_heads=cfg["num_heads"],
            decoder_embed_dim=cfg["decoder_embed_dim"], decoder_depth=cfg["decoder_depth"],
            decoder_num_heads=cfg["decoder_num_heads"], mlp_ratio=cfg["mlp_ratio"],
            norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), args=reference_args())
    finally:
        torch.load = real
    return m


def build_reference_vit(cfg: dict, num_classes: int, global_pool: bool):
    """The reference's VisionTransformer3D as model_factory.get_models('vit', args) builds it (model/model_factory.py:19-22),
    with the encoder geometry of ``cfg``."""
    from functools import partial
    load_reference()
    import importlib
    vit = importlib.import_module("model.vit")
    return vit.VisionTransformer3D(volume_size=cfg["volume_size"], in_chans=cfg["in_chans"], num_classes=num_classes,
                                   patch_size=cfg["patch_size"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                                   num_heads=cfg["num_heads"], mlp_ratio=cfg["mlp_ratio"], global_pool=global_pool,
                                   norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_path_rate=0.1)
