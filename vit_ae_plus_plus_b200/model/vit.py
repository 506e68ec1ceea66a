"""Drop-in for the part of the reference's ``model/vit.py`` that the pre-training scripts use after training: the
``VisionTransformer3D`` feature extractor.  ``k_fold_cross_valid_combined_brats.py:219-253`` builds it
with ``get_models('vit', args)``, loads the MAE checkpoint into it (``strict=False`` with an asserted missing-key set) and
calls ``utils.feature_extraction.generate_features`` -> ``model.forward_features(images)`` under ``torch.no_grad()``.

Same constructor signature, parameter names / shapes (state_dict ABI) and ``forward_features`` / ``forward`` return values as
model/vit.py:147-297; the encoder runs on the sm_100a kernels of libvitae_b200.so (all patches, no masking: Conv3d patch
embed as im2col + tcgen05 GEMM with the positional table in the epilogue, the transformer blocks, the final LayerNorm).
Inference only: ``forward`` / ``forward_features`` with autograd enabled raise -- fine-tuning (post_training_utils/) is outside
the B200 path.  No CPU / PyTorch fallback.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Dict, Optional

import torch
from torch import nn

from .. import ops
from .._lib import VitaeError
from ..engine import FlatParams, MAEEngine, MAEPlan, StackSpec, _Arena, _Lanes, block_param_names
from .vit_autoenc import Affine, Block, Dense, PatchEmbed3D, _ParamHolder

_F32, _BF16, _I32 = torch.float32, torch.bfloat16, torch.int32


class _EncPlan:
    """Activation arena of the encoder-only pass for one batch size."""

    def __init__(self, eng: "EncoderEngine", B: int):
        a = _Arena(eng.device)
        L, D = eng.L, eng.enc.dim
        self.B, self.N, self.M = B, L + 1, B * (L + 1)
        self.ids = torch.arange(L, dtype=_I32, device=eng.device).repeat(B, 1).contiguous()   # identity: nothing is masked
        self.maps = ops.build_row_maps(self.ids, L)
        self.cols = a.new((B * L, eng.Kpe), _BF16)
        self.enc = MAEPlan._stack(a, eng.enc, self.M, B, self.N)
        self.pooled = a.new((B, D), _F32)
        self.feat = a.new((B, D), _F32)
        self.tok_rows = [torch.arange(b * self.N + 1, (b + 1) * self.N, dtype=_I32, device=eng.device) for b in range(B)]
        self.nbytes = a.nbytes


class EncoderEngine:
    """Flat parameters + kernel orchestration of the encoder-only forward (model/vit.py:265-284).  Borrows the block
    forward of MAEEngine (same kernels, same activation buffers)."""

    def __init__(self, cfg: dict, named_params: Dict[str, torch.nn.Parameter], pos_embed: torch.nn.Parameter, ln_eps: float,
                 final_norm: str):
        dev = pos_embed.device
        if dev.type != "cuda":
            raise VitaeError("VisionTransformer3D runs on a B200 only; there is no CPU / PyTorch fallback")
        ops._lib.check(ops._lib.load().vitae_check_device(), "vitae_check_device")
        self.cfg, self.device = cfg, dev
        V, p, C = cfg["volume_size"], cfg["patch_size"], cfg["in_chans"]
        self.V, self.p, self.C = V, p, C
        self.L = (V // p) ** 3
        self.Kpe = C * p ** 3
        self.eps = float(ln_eps)
        D = cfg["embed_dim"]
        self.enc = StackSpec("blocks", D, cfg["num_heads"], int(D * cfg["mlp_ratio"]), cfg["depth"])
        if self.enc.head_dim not in (16, 32, 64) or D % 8 or D > 1024 or p % 4:
            raise VitaeError(f"unsupported encoder geometry (width {D}, head_dim {self.enc.head_dim}, patch {p})")
        order = ["patch_embed.proj.weight", "patch_embed.proj.bias", "cls_token"]
        for i in range(cfg["depth"]):
            order += block_param_names("blocks", i)
        order += [f"{final_norm}.weight", f"{final_norm}.bias"]
        self.final_norm = final_norm
        self.flat = FlatParams({n: named_params[n] for n in order}, order, dev)
        self.pos_param = pos_embed
        self.plans: Dict[int, _EncPlan] = {}
        self.lanes = _Lanes(dev)
        self.ws_main = ops.GrowBuf(dev)
        self.use_side_lane = False
        self.use_l2_prefetch = False

    # the attribute / method set MAEEngine._stack_fwd relies on
    def _w(self, name):
        return self.flat.v16[name]

    def _p(self, name):
        return self.flat.v32[name]

    def _need(self, name):
        pass

    def _prefetch(self, tensors):
        pass

    def _prefetch_block(self, st, sb, i, with_acts):
        pass

    def features(self, vol: torch.Tensor, global_pool: bool) -> torch.Tensor:
        """vol fp32 [B, C, V, V, V] -> fp32 [B, D]: mean of the patch tokens through fc_norm (global_pool) or the cls row of
        the final norm."""
        B, D, L = vol.shape[0], self.enc.dim, self.L
        pl = self.plans.get(B)
        if pl is None:
            pl = self.plans[B] = _EncPlan(self, B)
        self.flat.refresh_shadow()
        pos = self.pos_param.detach().reshape(L + 1, D)            # a learned parameter here (model/vit.py:190)
        ops.im2col_patches(vol, pl.ids, pl.cols, self.p, L)
        x0 = pl.enc.x[0]
        ops.gemm(pl.cols, self._w("patch_embed.proj.weight"), B * L, D, self.Kpe, bias=self._p("patch_embed.proj.bias"),
                 addend=pos, add_rows=pl.maps["pe_pos_rows"], ldadd=D, out_f32=x0, out_rows=pl.maps["enc_tok_rows"],
                 workspace=self.ws_main)
        ops.fill_rows(x0, pl.maps["enc_cls_rows"], B, D, self._p("cls_token"), None, pos, None)
        MAEEngine._stack_fwd(self, self.enc, pl.enc, pl.M, B, pl.N)
        x = pl.enc.x[-1]
        gamma, beta = self._p(f"{self.final_norm}.weight"), self._p(f"{self.final_norm}.bias")
        if global_pool:     # x[:, 1:, :].mean(dim=1) -> fc_norm (model/vit.py:277-279)
            for b in range(B):
                ops.sum_rows(x, pl.tok_rows[b], L, D, pl.pooled[b], accumulate=False)
            pl.pooled.mul_(1.0 / L)
        else:               # norm(x)[:, 0] (model/vit.py:281-282): LayerNorm is row-wise, so only the cls rows are normalised
            ops.gather_rows(x, pl.maps["enc_cls_rows"], B, D, None, pl.pooled)
        ops.layernorm_fwd(pl.pooled, gamma, beta, None, None, None, self.eps, y_f32=pl.feat)
        return pl.feat.clone()


class VisionTransformer3D(nn.Module):
    """Constructor signature of model/vit.py:157-160; see the module docstring for what runs where."""

    def __init__(self, volume_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=True, representation_size=None, distilled=False,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., embed_layer=None, norm_layer=None,
                 act_layer=None, weight_init='', global_pool=False):
        super().__init__()
        if not qkv_bias or representation_size or distilled or drop_rate or attn_drop_rate:
            raise VitaeError("VisionTransformer3D: only the configuration the k-fold scripts build is supported "
                             "(qkv_bias=True, no representation layer / distillation / dropout)")
        if weight_init not in ('', 'jax', 'jax_nlhb', 'nlhb'):
            raise AssertionError(weight_init)
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        eps = float(getattr(norm_layer(8), "eps", 1e-6))
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 1
        self.patch_embed = PatchEmbed3D(volume_size, patch_size, in_chans, embed_dim)
        L = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = None
        self.pos_embed = nn.Parameter(torch.zeros(1, L + 1, embed_dim))
        self.pos_drop = nn.Identity()            # p = 0 (drop_rate), model/vit.py:191
        # drop_path is accepted and ignored by the reference's Block as well (model/vit.py:128-134,140-141)
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, eps) for _ in range(depth)])
        self.global_pool = global_pool
        if global_pool:
            self.fc_norm = Affine(embed_dim, eps)          # and no ``norm`` (deleted at model/vit.py:222)
        else:
            self.norm = Affine(embed_dim, eps)
        self.pre_logits = nn.Identity()
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.head_dist = None
        self.ln_eps = eps
        self.cfg = dict(volume_size=self.patch_embed.volume_size[0], patch_size=self.patch_embed.patch_size[0],
                        in_chans=in_chans, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio)
        self._engine: Optional[EncoderEngine] = None
        self.init_weights(weight_init)

    def init_weights(self, mode=''):
        # model/vit.py:228-241 with mode '' (what the factory passes): trunc_normal(.02) for pos_embed / cls_token / every
        # Linear weight, zero biases, LayerNorm (1, 0); the Conv3d keeps torch's default init
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        for m in self.modules():
            if isinstance(m, Dense):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                nn.init.zeros_(m.bias)
        fan_in = self.cfg["in_chans"] * self.cfg["patch_size"] ** 3
        nn.init.kaiming_uniform_(self.patch_embed.proj.weight.data.view(self.embed_dim, -1), a=math.sqrt(5))
        nn.init.uniform_(self.patch_embed.proj.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'dist_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def engine(self) -> EncoderEngine:
        eng = self._engine
        if eng is not None and eng.flat.still_aliased() and eng.device == self.pos_embed.device:
            return eng
        if self.cls_token.device.type != "cuda":
            raise VitaeError(f"VisionTransformer3D runs on a B200 only (module is on {self.cls_token.device}); there is no "
                             "CPU / PyTorch fallback")
        skip = ("pos_embed", "head.weight", "head.bias")
        named = {n: p for n, p in self.named_parameters() if n not in skip}
        self._engine = EncoderEngine(self.cfg, named, self.pos_embed, self.ln_eps, "fc_norm" if self.global_pool else "norm")
        return self._engine

    def forward_features(self, x):
        """model/vit.py:265-284 -> fp32 [B, embed_dim].  Inference only."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise VitaeError("VisionTransformer3D.forward_features is inference-only on the B200 path: call it under "
                             "torch.no_grad() (utils/feature_extraction.py:9 does); fine-tuning uses the reference's model.vit")
        V = self.cfg["volume_size"]
        if x.dim() != 5 or tuple(x.shape[1:]) != (self.cfg["in_chans"], V, V, V):
            raise VitaeError(f"expected a (N, {self.cfg['in_chans']}, {V}, {V}, {V}) volume, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise VitaeError("input volume must be a CUDA tensor (no CPU path)")
        return self.engine().features(x.contiguous().float(), self.global_pool)

    def forward(self, x):
        # model/vit.py:286-297 without the distillation head; the classifier head is a [num_classes, D] product on B rows
        f = self.forward_features(x)
        return self.head(f) if isinstance(self.head, nn.Linear) else f
