"""Drop-in for the reference's ``model/vit_autoenc.py``: same constructors, forward signature / return tuple and
state_dict keys, but the arithmetic runs in the hand-written sm_100a kernels of
libvitae_b200.so, orchestrated by ``vit_ae_plus_plus_b200.engine.MAEEngine``.

The sub-modules below (``PatchEmbed3D.proj``, ``Block.attn.qkv`` ...) are *parameter containers*: they exist so that
``named_parameters()`` / ``state_dict()`` / ``str(model)`` look like the reference's (model/vit_autoenc.py:18-63,
model/vit.py:52-144) and checkpoints round-trip; none of them computes anything -- calling one raises.  There is no
CPU or torch fallback: ``forward`` on a non-CUDA module raises ``VitaeError``.

Differences a caller can observe (all documented in DESIGN.md):
  * gradients are accumulated into ``param.grad`` by the backward kernels themselves (views of one flat buffer);
  * ``forward`` accepts an optional trailing ``noise=`` (the ``torch.rand(N, L)`` draw of vit_autoenc.py:139) so that
    tests can feed the same mask to the reference; default behaviour draws it exactly like the reference does;
  * ties in ``noise`` are ordered stably (the reference's un-stable ``argsort`` leaves them unspecified);
  * ``pred`` / ``mask`` returned by ``forward`` are views of per-shape workspaces that the next ``forward`` overwrites;
  * ``loss_list[1]`` (raw edge-map loss, fused Sobel / Gaussian stencil kernels) is only evaluated when ``edge_map_weight != 0`` or ``report_edge_loss`` is
    set; the VGG perceptual term (never differentiable in the reference, perceptual_loss.py:68-69) is supported for
    ``perceptual_weight == 0`` only (the shipped default, config.ini:34).
"""
from __future__ import annotations

import math
import weakref
from functools import partial
from typing import Optional

import torch
from torch import nn

from .. import ops
from .._lib import VitaeError
from ..engine import MAEEngine
from .model_utils import sincos_pos_embed_3d


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


class _ParamHolder(nn.Module):
    def forward(self, *a, **k):
        raise VitaeError(f"{type(self).__name__} is a parameter container; the fused B200 kernels compute it "
                         "(call the MaskedAutoencoderViT, not its sub-modules)")


class Dense(_ParamHolder):
    """Holds weight [out, in] (+ bias) of an nn.Linear site (model/vit.py:84-86,107-109; vit_autoenc.py:40,52)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        nn.init.xavier_uniform_(self.weight)                      # vit_autoenc.py:90-95

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"


class Affine(_ParamHolder):
    """Holds the LayerNorm affine (weight=1, bias=0 at init; vit_autoenc.py:96-98) and its eps."""

    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.normalized_shape, self.eps = (dim,), eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def extra_repr(self):
        return f"{self.normalized_shape}, eps={self.eps}"


class PatchProj(_ParamHolder):
    """Holds the Conv3d(k=s=patch) weight (D, C, p, p, p) + bias of model/vit.py:65."""

    def __init__(self, in_chans: int, embed_dim: int, patch: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(embed_dim, in_chans, patch, patch, patch))
        self.bias = nn.Parameter(torch.empty(embed_dim))
        fan_in = in_chans * patch ** 3
        nn.init.xavier_uniform_(self.weight.data.view(embed_dim, -1))     # vit_autoenc.py:80-82
        nn.init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))   # torch Conv3d default, untouched by :90-98

    def extra_repr(self):
        d, c, p = self.weight.shape[:3]
        return f"{c}, {d}, kernel_size=({p}, {p}, {p}), stride=({p}, {p}, {p})"


class PatchEmbed3D(_ParamHolder):
    def __init__(self, volume_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.volume_size, self.patch_size = _triple(volume_size), _triple(patch_size)
        if len(set(self.volume_size)) != 1 or len(set(self.patch_size)) != 1:
            raise VitaeError("cubic volumes / patches only (the reference's patchify asserts the same, vit_autoenc.py:106)")
        self.grid_size = tuple(v // p for v, p in zip(self.volume_size, self.patch_size))
        self.num_patches = self.grid_size[0] * self.grid_size[1] * self.grid_size[2]
        self.proj = PatchProj(in_chans, embed_dim, self.patch_size[0])


class Attention(_ParamHolder):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = Dense(dim, dim * 3, bias=True)
        self.proj = Dense(dim, dim)


class Mlp3D(_ParamHolder):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = Dense(dim, hidden)
        self.fc2 = Dense(hidden, dim)


class Block(_ParamHolder):
    def __init__(self, dim, num_heads, mlp_ratio, eps):
        super().__init__()
        self.norm1 = Affine(dim, eps)
        self.attn = Attention(dim, num_heads)
        self.norm2 = Affine(dim, eps)
        self.mlp = Mlp3D(dim, int(dim * mlp_ratio))


class _MAEStep(torch.autograd.Function):
    """One autograd node for the whole step: forward enqueues the forward kernels, backward the backward kernels.
    ``anchor`` is any trainable parameter (it only makes the outputs require grad); parameter gradients are written
    by the kernels straight into ``param.grad`` (views of the flat gradient buffer), so backward returns None."""

    @staticmethod
    def forward(ctx, anchor, module, vol, noise, keep, want_edge):
        eng = module._engine
        pl = eng.forward(vol, noise, keep, want_loss=True, pred_f32=module.pred_dtype == torch.float32, want_edge=want_edge)
        pl.step_id += 1
        ctx.module, ctx.pl, ctx.step_id, ctx.want_edge = module, pl, pl.step_id, want_edge
        ctx.set_materialize_grads(False)
        recon = pl.loss_out[0].clone()
        raw_edge = pl.edge_out[0].clone() if want_edge else torch.zeros((), device=vol.device)
        mask = pl.mask.clone()
        pred = pl.pred_view(module.pred_dtype)
        ctx.mark_non_differentiable(mask)
        return recon, raw_edge, pred, mask

    @staticmethod
    def backward(ctx, drecon, dedge, dpred, _dmask):
        module, pl = ctx.module, ctx.pl
        if pl.step_id != ctx.step_id:
            raise VitaeError("backward() after a later forward() of the same shape: the activation workspace was reused")
        module._backward(pl, drecon, dpred, dedge=dedge if ctx.want_edge else None)
        return None, None, None, None, None, None


class MaskedAutoencoderViT(nn.Module):
    """Masked autoencoder with a 3-D ViT backbone -- constructor signature of model/vit_autoenc.py:18-21."""

    def __init__(self, volume_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4., norm_layer=nn.LayerNorm,
                 norm_pix_loss=False, args=None):
        super().__init__()
        eps = float(getattr(norm_layer(8), "eps", 1e-5))     # the reference passes partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = PatchEmbed3D(volume_size, patch_size, in_chans, embed_dim)
        L = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, L + 1, embed_dim), requires_grad=False)
        self.embed_dim = embed_dim
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, eps) for _ in range(depth)])
        self.norm = Affine(embed_dim, eps)
        self.decoder_embed = Dense(embed_dim, decoder_embed_dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, L + 1, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.ModuleList([Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, eps)
                                             for _ in range(decoder_depth)])
        self.decoder_norm = Affine(decoder_embed_dim, eps)
        self.decoder_pred = Dense(decoder_embed_dim, self.patch_embed.patch_size[0] ** 3 * in_chans)
        self.args = args
        self.perceptual_weight = 0 if args is None else getattr(args, "perceptual_weight", 0)
        print(f"Using perceptual weight of {self.perceptual_weight}")
        if norm_pix_loss:
            raise VitaeError("norm_pix_loss=True is dead code in the reference (model_factory.py:12 never forwards it) "
                             "and is not implemented")
        self.norm_pix_loss = False
        self.ln_eps = eps
        self.cfg = dict(volume_size=self.patch_embed.volume_size[0], patch_size=self.patch_embed.patch_size[0],
                        in_chans=in_chans, embed_dim=embed_dim, depth=depth, num_heads=num_heads,
                        decoder_embed_dim=decoder_embed_dim, decoder_depth=decoder_depth,
                        decoder_num_heads=decoder_num_heads, mlp_ratio=mlp_ratio)
        self.pred_dtype = torch.float32      # dtype of the returned ``pred`` (set to torch.bfloat16 to skip the fp32 copy)
        self.report_edge_loss = False        # evaluate loss_list[1] even when edge_map_weight == 0
        self.require_backward_grad_sync = True   # data-parallel: all-reduce gradients in backward (see no_sync())
        self.use_cuda_graph = True               # replay the forward / backward kernel sequences as CUDA graphs
        self._engine: Optional[MAEEngine] = None
        self.initialize_weights()

    # ------------------------------------------------------------------------------------------------ init
    def initialize_weights(self):
        g = round(self.patch_embed.num_patches ** (1 / 3))
        with torch.no_grad():
            self.pos_embed.copy_(torch.from_numpy(sincos_pos_embed_3d(self.pos_embed.shape[-1], g)).float().unsqueeze(0))
            self.decoder_pos_embed.copy_(
                torch.from_numpy(sincos_pos_embed_3d(self.decoder_pos_embed.shape[-1], g)).float().unsqueeze(0))
        nn.init.normal_(self.cls_token, std=.02)
        nn.init.normal_(self.mask_token, std=.02)
        # Dense / Affine / PatchProj initialise themselves with the reference's distributions (vit_autoenc.py:80-98)

    # ------------------------------------------------------------------------------------------------ engine plumbing
    def _trainable(self):
        """Parameters that live in the engine's flat buffers: everything of the MAE and the contrastive predictor (the
        never-called projection head of ``use_proj`` keeps its own storage)."""
        return {n: p for n, p in self.named_parameters()
                if n not in ("pos_embed", "decoder_pos_embed") and not n.startswith("projection_head.")}

    def engine(self) -> MAEEngine:
        """Builds (once per device placement) the flat parameter buffers; the nn.Parameters become views of them."""
        eng = self._engine
        if eng is not None and eng.flat.still_aliased() and eng.pos.device == self.pos_embed.device:
            return eng
        if self.cls_token.device.type != "cuda":
            raise VitaeError("MaskedAutoencoderViT runs on a B200 only (module is on "
                             f"{self.cls_token.device}); there is no CPU / PyTorch fallback")
        self._engine = MAEEngine(self.cfg, self._trainable(), self.pos_embed, self.decoder_pos_embed, self.ln_eps)
        self._engine.use_graphs = self.use_cuda_graph
        ref = weakref.ref(self._engine)
        for p in self._engine.flat.params.values():
            p._vitae_engine = ref          # lets utils.misc.NativeScalerWithGradNormCount find the fused optimizer
        self._engine.broadcast_parameters()
        self._broadcast_extra_parameters()
        return self._engine

    def _broadcast_extra_parameters(self):
        """Data parallel: parameters outside the flat buffers (none in the plain MAE) follow rank 0 as well."""

    @property
    def graph_replayed_launches(self) -> int:
        """Kernels executed through CUDA-graph replays (the library's launch counter only sees direct enqueues)."""
        return 0 if self._engine is None else self._engine.graph_replayed_launches

    def no_sync(self):
        """Context manager: skip the data-parallel gradient all-reduce (gradient-accumulation micro-steps)."""
        module = self

        class _NoSync:
            def __enter__(self):
                self.prev = module.require_backward_grad_sync
                module.require_backward_grad_sync = False

            def __exit__(self, *exc):
                module.require_backward_grad_sync = self.prev
        return _NoSync()

    def _accumulating(self) -> bool:
        """True when this backward adds to gradients already sitting in the aliased ``.grad`` views (accumulation
        micro-step) rather than starting from zero."""
        flat = self._engine.flat
        state = flat.grads_alias()
        if state is None:
            state = flat.grads_alias(thorough=True)
        return state is True and not flat.overwrite_grads

    def _backward(self, pl, drecon, dpred, dlatent=None, second=None, dedge=None):
        eng = self._engine
        flat = eng.flat
        state = flat.grads_alias()
        if state is None:
            state = flat.grads_alias(thorough=True)
        saved = None
        if state is None:   # foreign / partially set .grad tensors: keep them and add ours afterwards
            saved = {n: flat.params[n].grad.clone() for n in flat.order if flat.params[n].grad is not None}
        # accumulate into aliased .grad tensors unless the fused optimizer has consumed them since the last backward
        acc = state is True and not flat.overwrite_grads
        flat.overwrite_grads = False
        eng.use_graphs = self.use_cuda_graph
        # the gradient exchange overlaps the backward stages, except when foreign gradients or a second (encoder-only,
        # contrastive view 2) backward still have to be added first
        sync = self.require_backward_grad_sync
        if sync and eng.defer_exchange is not None:
            # the caller's fused optimizer step exchanges the gradients itself (dp.ShardedStep): "reduce" = it follows this
            # backward (overlap the owner-side reduce with the stages when nothing else has to be added first), "skip" =
            # gradient accumulation, no exchange yet
            eng.grads_local = True
            if eng.defer_exchange == "skip" or saved is not None or second is not None:
                sync = False
        overlap_sync = sync and saved is None and second is None
        eng.backward(pl, drecon, dpred_extra=dpred, accumulate=acc, sync_grads=overlap_sync, dlatent=dlatent, dedge=dedge)
        if second is not None and second[1] is not None:
            eng.backward(second[0], None, accumulate=True, dlatent=second[1], encoder_only=True)
        if saved is not None:
            eng.norm_partials = 0        # foreign gradients are added below: the per-stage norm partials no longer cover them
        if state is not True:
            for n in flat.order:
                p = flat.params[n]
                if saved is not None and n in saved:
                    flat.vg[n].add_(saved[n])
                if p.requires_grad:
                    p.grad = flat.vg[n]
        if sync and not overlap_sync:
            eng.allreduce_gradients()
        if sync:
            self._sync_extra_grads()

    def _sync_extra_grads(self):
        """Data parallel: gradients of parameters outside the flat buffers (none in the plain MAE)."""

    def state_dict(self, *args, **kwargs):
        if self._engine is not None:
            self._engine.wait_params()       # an overlapped optimizer step may still be writing the parameters
            self._engine.sync_master()       # sharded data-parallel steps: fetch the master of the parts other ranks own
        return super().state_dict(*args, **kwargs)

    # frozen helper modules of the reference model whose buffers / weights ride along in its checkpoints
    # (model/vit_autoenc.py:54-57: SobelFilter3d, PerceptualLoss); here they are constants inside the kernels
    REFERENCE_ONLY_PREFIXES = ("sobel_filter3D.", "perceptual_loss.")

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts this module's own checkpoints and the reference's (``misc.load_model``, reference misc.py:320): keys of
        the reference's frozen Sobel / VGG helpers are dropped before the (strict) match."""
        sd = {k: v for k, v in state_dict.items() if not k.startswith(self.REFERENCE_ONLY_PREFIXES)}
        if self._engine is not None:
            self._engine.wait_params()
        out = super().load_state_dict(sd, strict=strict, assign=assign)
        if self._engine is not None:
            self._engine.flat.invalidate_shadow()
        return out

    # ------------------------------------------------------------------------------------------------ reference API
    def patchify(self, volume):
        """model/vit_autoenc.py:100-113: (N, C, V, V, V) -> (N, L, p^3*C), within-patch order (pz, py, px, c).
        API helper (a permuted copy); the training step never patchifies -- the loss kernels read the volume."""
        p = self.patch_embed.patch_size[0]
        N, C, V = volume.shape[0], volume.shape[1], volume.shape[2]
        assert volume.shape[2] == volume.shape[3] == volume.shape[4] and V % p == 0
        g = V // p
        x = volume.reshape(N, C, g, p, g, p, g, p).permute(0, 2, 4, 6, 3, 5, 7, 1)
        return x.reshape(N, g ** 3, p ** 3 * C)

    def unpatchify(self, x):
        """model/vit_autoenc.py:115-128: inverse of patchify."""
        p = self.patch_embed.patch_size[0]
        N, L = x.shape[0], x.shape[1]
        g = round(L ** (1 / 3))
        assert g ** 3 == L
        x = x.reshape(N, g, g, g, p, p, p, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
        return x.reshape(N, -1, g * p, g * p, g * p)

    def _check_volume(self, x):
        V = self.cfg["volume_size"]
        if x.dim() != 5 or tuple(x.shape[1:]) != (self.cfg["in_chans"], V, V, V):
            raise VitaeError(f"expected a (N, {self.cfg['in_chans']}, {V}, {V}, {V}) volume, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise VitaeError("input volume must be a CUDA tensor (no CPU path)")
        return x.contiguous().float()

    def _noise(self, x, noise):
        L = self.patch_embed.num_patches
        if noise is None:
            return torch.rand(x.shape[0], L, device=x.device)            # vit_autoenc.py:139
        return noise.to(device=x.device, dtype=torch.float32).contiguous()

    def _len_keep(self, mask_ratio):
        return int(self.patch_embed.num_patches * (1 - mask_ratio))      # vit_autoenc.py:137

    @torch.no_grad()
    def forward_encoder(self, x, mask_ratio, noise=None):
        """model/vit_autoenc.py:157-177 -> (latent fp32 [N, keep+1, D], mask [N, L], ids_restore int64 [N, L])."""
        eng = self.engine()
        x = self._check_volume(x)
        pl = eng.plan(x.shape[0], self._len_keep(mask_ratio))
        eng.refresh_shadow()
        eng.encode(pl, x, self._noise(x, noise))
        latent = pl.latent.float().view(x.shape[0], pl.Ne, self.embed_dim)
        return latent, pl.mask.clone(), pl.ids_restore.long()

    @torch.no_grad()
    def forward_decoder(self, x, ids_restore):
        """model/vit_autoenc.py:179-203 -> pred fp32 [N, L, p^3*C] for an arbitrary (latent, ids_restore) pair."""
        eng = self.engine()
        B, Ne = x.shape[0], x.shape[1]
        pl = eng.plan(B, Ne - 1)
        eng.refresh_shadow()
        pl.latent.copy_(x.reshape(B * Ne, -1))
        pl.ids_shuffle.copy_(torch.argsort(ids_restore, dim=1))
        ops.build_row_maps(pl.ids_shuffle, Ne - 1, pl.maps)
        eng.decode(pl, pred_f32=True)
        return pl.pred32[:, 1:, :].clone()

    def forward_loss(self, imgs, pred, mask, edge_map_weight=0):
        """model/vit_autoenc.py:205-232 -> [loss, raw_edge, recon, percep] for a caller-supplied pred (API helper;
        the training step computes the reconstruction term inside ``forward``).  Not differentiable."""
        eng = self.engine()
        imgs = self._check_volume(imgs)
        B, L, P = pred.shape
        with torch.no_grad():
            full = torch.zeros(B, L + 1, P, device=imgs.device, dtype=torch.float32)
            full[:, 1:] = pred
            sums = torch.empty(B * L, device=imgs.device)
            out = torch.empty(2, device=imgs.device)
            ops.masked_mse_fwd(full, imgs, mask.float().contiguous(), sums, out, eng.p)
            raw_edge = None
            if edge_map_weight != 0 or self.report_edge_loss:
                C, V = eng.C, eng.V
                scratch = torch.empty(ops.edge_scratch_floats(B, C, V), device=imgs.device)
                e_tgt, resid = torch.empty(B, V, V, V, device=imgs.device), torch.empty(B, V, V, V, device=imgs.device)
                eo = torch.empty(1, device=imgs.device)
                ops.edge_target(imgs, eng.edge_taps, scratch, e_tgt)
                ops.edge_loss_fwd(full.to(torch.bfloat16), e_tgt, scratch, resid, eo, B, C, V, eng.p)
                raw_edge = eo[0].clone()
            return self._loss_list(out[0].clone(), raw_edge, edge_map_weight)

    def _loss_list(self, recon, raw_edge, edge_map_weight):
        """[edge_w * raw_edge + recon + percep, raw_edge, recon, percep] (model/vit_autoenc.py:231-232); ``raw_edge`` None:
        the edge-map term was not evaluated (weight 0 and report_edge_loss off) and is reported as 0."""
        if self.perceptual_weight != 0:
            raise VitaeError("perceptual_weight != 0 needs the reference's VGG-16 checkpoint (model/ckp-399.pth, not "
                             "shipped); only the shipped default 0 is supported")
        percep = torch.zeros((), device=recon.device)
        if raw_edge is not None:
            loss = edge_map_weight * raw_edge + recon + percep
        else:
            raw_edge = torch.zeros((), device=recon.device)
            loss = recon
        return [loss, raw_edge, recon, percep]

    def forward(self, sample, mask_ratio=0.75, edge_map_weight=0, noise=None):
        """model/vit_autoenc.py:234-238 -> ([loss, raw_edge, recon, percep], pred [N, L, p^3*C], mask [N, L])."""
        eng = self.engine()
        eng.use_graphs = self.use_cuda_graph
        x = self._check_volume(sample)
        noise = self._noise(x, noise)
        keep = self._len_keep(mask_ratio)
        if keep < 1:
            raise VitaeError(f"mask_ratio={mask_ratio} keeps no patch")
        want_edge = edge_map_weight != 0 or self.report_edge_loss
        if torch.is_grad_enabled() and self.cls_token.requires_grad:
            recon, raw_edge, pred, mask = _MAEStep.apply(self.cls_token, self, x, noise, keep, want_edge)
        else:
            pl = eng.forward(x, noise, keep, want_loss=True, pred_f32=self.pred_dtype == torch.float32, want_edge=want_edge)
            pl.step_id += 1
            recon, pred, mask = pl.loss_out[0].clone(), pl.pred_view(self.pred_dtype), pl.mask.clone()
            raw_edge = pl.edge_out[0].clone() if want_edge else None
        return self._loss_list(recon, raw_edge if want_edge else None, edge_map_weight), pred, mask


class _ContrastiveStep(torch.autograd.Function):
    """One autograd node for both views of the contrastive model: forward = full MAE pass on view 1 + encoder pass on
    view 2, each followed by the predictor (model/vit_autoenc.py:263-268,282-283) on its latent; backward = predictor
    backward of both views (their latent gradients join whatever arrives for the latents directly), full backward of view
    1, then the encoder-only backward of view 2 accumulating into the same gradient buffers.  A single node fixes that order."""

    @staticmethod
    def forward(ctx, anchor, module, vol1, vol2, noise1, noise2, keep, want_edge):
        eng = module._engine
        pl1 = eng.forward(vol1, noise1, keep, want_loss=True, pred_f32=module.pred_dtype == torch.float32,
                          want_edge=want_edge, with_predictor=True)
        pl2 = eng.forward_encoder_only(vol2, noise2, keep, slot=1, with_predictor=True)
        pl1.step_id += 1
        pl2.step_id += 1
        ctx.module, ctx.pl1, ctx.pl2, ctx.ids, ctx.want_edge = module, pl1, pl2, (pl1.step_id, pl2.step_id), want_edge
        ctx.set_materialize_grads(False)
        recon, mask = pl1.loss_out[0].clone(), pl1.mask.clone()
        raw_edge = pl1.edge_out[0].clone() if want_edge else torch.zeros((), device=vol1.device)
        pred = pl1.pred_view(module.pred_dtype)
        latent1, latent2 = pl1.latent32.clone(), pl2.latent32.clone()  # [B*Ne, D], model/vit_autoenc.py:280-281
        p1, p2 = pl1.pb.p.clone(), pl2.pb.p.clone()
        ctx.mark_non_differentiable(mask)
        return recon, raw_edge, pred, mask, latent1, latent2, p1, p2

    @staticmethod
    def backward(ctx, drecon, dedge, dpred, _dmask, dlat1, dlat2, dp1, dp2):
        module, pl1, pl2 = ctx.module, ctx.pl1, ctx.pl2
        if (pl1.step_id, pl2.step_id) != ctx.ids:
            raise VitaeError("backward() after a later forward() of the same shape: the activation workspace was reused")
        eng = module._engine
        acc = module._accumulating()
        eng.use_graphs = module.use_cuda_graph
        for pl, dp, which in ((pl1, dp1, 0), (pl2, dp2, 1)):
            if dp is None:
                continue
            g = eng.predictor_backward(pl, dp, accumulate=acc)
            acc = True
            if which == 0:
                dlat1 = g if dlat1 is None else dlat1 + g
            else:
                dlat2 = g if dlat2 is None else dlat2 + g
        if not acc:                       # the predictor took no part in this loss: its gradient is zero, not last step's
            eng.zero_predictor_grads()
        module._backward(pl1, drecon, dpred, dlatent=dlat1, second=(pl2, dlat2), dedge=dedge if ctx.want_edge else None)
        return None, None, None, None, None, None, None, None


class ContrastiveMAEViT(MaskedAutoencoderViT):
    """MAE + contrastive predictor on the encoder tokens of two views -- model/vit_autoenc.py:241-285, the k-fold scripts'
    default ``--model contr_mae_vit_base_patch16`` (k_fold_cross_valid_combined_brats.py:37).  Both encoder passes, the
    decoder, the loss, the predictor (two Linear layers on the tcgen05 GEMM around a fused BatchNorm1d + ReLU kernel,
    SURVEY row f-2) and all their gradients run in the B200 kernels.  ``self.predictor`` is a parameter / buffer container
    with the reference's state_dict keys; its parameters live in the engine's flat buffers like every other one, so an
    optimizer over ``model.parameters()`` takes the fused AdamW path."""

    def __init__(self, volume_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4., norm_layer=nn.LayerNorm,
                 norm_pix_loss=False, args=None, use_proj=False):
        super().__init__(volume_size=volume_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                         depth=depth, num_heads=num_heads, decoder_embed_dim=decoder_embed_dim,
                         decoder_depth=decoder_depth, decoder_num_heads=decoder_num_heads, mlp_ratio=mlp_ratio,
                         norm_layer=norm_layer, norm_pix_loss=norm_pix_loss, args=args)
        self.use_proj = use_proj
        D = self.embed_dim
        if use_proj:   # built by the reference, never called in its forward (vit_autoenc.py:254-262,270-285)
            self.projection_head = nn.Sequential(nn.Linear(D, D, bias=False), nn.BatchNorm1d(D), nn.ReLU(inplace=True),
                                                 nn.Linear(D, D, bias=False), nn.BatchNorm1d(D), nn.ReLU(inplace=True),
                                                 nn.Linear(D, D, bias=False), nn.BatchNorm1d(D, affine=False))
        # built after the base class' init: keeps torch's default Linear / BatchNorm initialisation (vit_autoenc.py:263-268)
        self.predictor = nn.Sequential(nn.Linear(D, D, bias=False), nn.BatchNorm1d(D), nn.ReLU(inplace=True), nn.Linear(D, D))

    def _extra_modules(self):
        return [self.projection_head] if self.use_proj else []

    def _broadcast_extra_parameters(self):
        from .. import dp
        for m in self._extra_modules():
            for t in list(m.parameters()) + list(m.buffers()):
                dp.broadcast_flat(t.data)
        for t in self.predictor.buffers():           # BatchNorm running statistics (its parameters are in the flat buffer)
            dp.broadcast_flat(t.data)

    def _sync_extra_grads(self):
        # the predictor's backward has already run when the engine node's backward is called (it is downstream of the latents)
        from .. import dp
        for m in self._extra_modules():
            for p in m.parameters():
                if p.grad is not None:
                    dp.allreduce_mean_(p.grad)

    def engine(self):
        eng = super().engine()
        eng.want_latent32 = True       # plans built from now on keep an fp32 copy of the normalised encoder output
        bn = self.predictor[1]
        if not bn.track_running_stats or bn.momentum is None:
            raise VitaeError("the predictor's BatchNorm1d must track running statistics with a fixed momentum (torch defaults)")
        eng.bn_state = (bn.running_mean, bn.running_var, float(bn.eps), float(bn.momentum))
        return eng

    def forward(self, view1, view2, mask_ratio=0.75, edge_map_weight=0, noise=None, noise2=None):
        """-> ([loss, raw_edge, recon, percep], pred, mask, p1, p2, z1, z2) with p / z of shape [B*(keep+1), D]
        (model/vit_autoenc.py:270-285).  ``noise`` / ``noise2``: optional mask noise of view 1 / view 2 (tests)."""
        eng = self.engine()
        eng.use_graphs = self.use_cuda_graph
        x1, x2 = self._check_volume(view1), self._check_volume(view2)
        n1 = self._noise(x1, noise)                 # drawn in the reference's order: view 1 first (:272), then view 2 (:277)
        n2 = self._noise(x2, noise2)
        keep = self._len_keep(mask_ratio)
        if keep < 1:
            raise VitaeError(f"mask_ratio={mask_ratio} keeps no patch")
        want_edge = edge_map_weight != 0 or self.report_edge_loss
        if not self.training:
            raise VitaeError("ContrastiveMAEViT runs its predictor's BatchNorm1d with batch statistics (training mode) only; "
                             "the reference never evaluates this model in eval mode")
        if torch.is_grad_enabled() and self.cls_token.requires_grad:
            recon, raw_edge, pred, mask, lat1, lat2, p1, p2 = _ContrastiveStep.apply(self.cls_token, self, x1, x2, n1, n2, keep,
                                                                                     want_edge)
        else:
            pl1 = eng.forward(x1, n1, keep, want_loss=True, pred_f32=self.pred_dtype == torch.float32, want_edge=want_edge,
                              with_predictor=True)
            pl2 = eng.forward_encoder_only(x2, n2, keep, slot=1, with_predictor=True)
            pl1.step_id += 1
            pl2.step_id += 1
            recon, pred, mask = pl1.loss_out[0].clone(), pl1.pred_view(self.pred_dtype), pl1.mask.clone()
            raw_edge = pl1.edge_out[0].clone() if want_edge else None
            lat1, lat2 = pl1.latent32.clone(), pl2.latent32.clone()
            p1, p2 = pl1.pb.p.clone(), pl2.pb.p.clone()
        self.predictor[1].num_batches_tracked.add_(2)        # two BatchNorm forward calls per step (vit_autoenc.py:282-283)
        return (self._loss_list(recon, raw_edge if want_edge else None, edge_map_weight), pred, mask, p1, p2, lat1.detach(),
                lat2.detach())


def mae_vit_large_patch16_dec512d8b(**kwargs):
    return MaskedAutoencoderViT(embed_dim=1024, depth=24, num_heads=16, decoder_embed_dim=512, decoder_depth=8,
                                decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_base_patch16_dec512d8b(**kwargs):
    return MaskedAutoencoderViT(embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=8,
                                decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def contr_mae_vit_base_patch16_dec512d8b(**kwargs):
    return ContrastiveMAEViT(embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=8,
                             decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


mae_vit_base_patch16 = mae_vit_base_patch16_dec512d8b
mae_vit_large_patch16 = mae_vit_large_patch16_dec512d8b
contr_mae_vit_base_patch16 = contr_mae_vit_base_patch16_dec512d8b
