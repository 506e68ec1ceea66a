"""Replays the call sequence of the reference's k-fold pre-training script on a B200 through the OVERLAY names
(``model.*`` / ``utils.*``, overlay/README.md) -- k_fold_cross_valid_combined_brats.py:78-253:

  misc.init_distributed_mode -> seed -> get_models('autoenc') -> .to(device) -> add_weight_decay + AdamW(betas=(.9,.95))
  -> NativeScaler -> misc.load_model -> [train_one_stage_epoch per epoch, edge_map_weight schedule] -> misc.save_model
  -> get_models('vit') -> torch.load -> interpolate_pos_embed -> load_state_dict(strict=False) + the asserted missing
  keys -> forward_features

and checks, against the CPU oracle (oracle/mae_oracle.py) driven by the same loop semantics (utils/train_one_epoch.py:21-110:
lr schedule per iteration, loss = loss_list[0] + contr_loss, GradScaler + AdamW), the per-epoch meters, plus
save -> load -> continue across the fused optimizer.  Run by tests/test_kfold_replay_gpu.py in a fresh process with
PYTHONPATH=overlay:<repo>; prints ``REPLAY_OK`` on success.

The scripts also need timm / torchio / the datasets, none of which exist here: ``add_weight_decay`` comes from utils.misc
(timm 0.5.4 semantics) and the DataLoader is a list of (augmented, original, label) batches.
"""
import argparse
import builtins
import math
import os
import sys
import tempfile

import torch


def main(model_name: str) -> None:
    from model.model_factory import get_models                                    # brats.py:21
    from utils import misc                                                         # :24
    from utils.misc import NativeScalerWithGradNormCount as NativeScaler           # :14
    from utils.train_one_epoch import train_one_stage_epoch                        # :27
    from oracle import mae_oracle as O
    assert misc.__name__ == "vit_ae_plus_plus_b200.utils.misc" and get_models.__module__.startswith("vit_ae_plus_plus_b200.")

    contrastive = model_name.startswith("contr_")
    tmp = tempfile.mkdtemp(prefix="vitae_kfold_")
    args = argparse.Namespace(
        dist_on_itp=False, dist_url="env://", seed=42, model=model_name, in_channels=2, volume_size=32, patch_size=16,
        perceptual_weight=0, use_imagenet=False, mask_ratio=0.75, contr_weight=0.1, accum_iter=1, batch_size=2,
        lr=None, blr=5e-2, min_lr=0.0, warmup_epochs=1, epochs=3, weight_decay=0.05, resume="", output_dir=tmp,
        start_epoch=0, use_edge_map=True, nb_classes=2, global_pool=True, drop_path=0.1, log_dir=tmp)
    plain_print = builtins.print
    misc.init_distributed_mode(args)                                               # :78
    builtins.print = plain_print          # single process: keep the test log free of time stamps
    device = torch.device("cuda:0")
    torch.manual_seed(args.seed + misc.get_rank())                                 # :86-88

    g = torch.Generator().manual_seed(5)
    B, C, V = args.batch_size, args.in_channels, args.volume_size
    n_iter = 3
    loader = [(torch.randn(B, C, V, V, V, generator=g).pin_memory(), torch.randn(B, C, V, V, V, generator=g).pin_memory(),
               torch.zeros(B)) for _ in range(n_iter)]

    model = get_models(model_name="autoenc", args=args)                            # :150
    model.to(device)
    cfg = dict(model.cfg)
    L = (V // args.patch_size) ** 3
    # the oracle starts from the model's own initial parameters
    P0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    eff_batch_size = args.batch_size * args.accum_iter * misc.get_world_size()
    args.lr = args.blr * eff_batch_size / 256                                      # :159-160
    optimizer = torch.optim.AdamW(misc.add_weight_decay(model, args.weight_decay), lr=args.lr, betas=(0.9, 0.95))   # :168-169
    loss_scaler = NativeScaler()                                                   # :171
    misc.load_model(args=args, model_without_ddp=model, optimizer=optimizer, loss_scaler=loss_scaler)   # :173 (no resume)

    # the mask noise the model will draw (vit_autoenc.py:139: torch.rand(N, L, device=x.device), view 1 then view 2)
    draws_per_step = 2 if contrastive else 1
    torch.manual_seed(1234)
    noises = [torch.rand(B, L, device=device).cpu() for _ in range(args.epochs * n_iter * draws_per_step)]
    torch.manual_seed(1234)

    stats = []
    for epoch in range(args.start_epoch, args.epochs):                             # :181-203
        edge_w = 0.01 * (1 - epoch / args.epochs) if args.use_edge_map else 0
        stats.append(train_one_stage_epoch(model, loader, optimizer, device, epoch, loss_scaler, log_writer=None, args=args,
                                           edge_map_weight=edge_w))
    assert loss_scaler._fused is not None, "the k-fold optimizer must take the fused AdamW path"
    misc.save_model(args=args, model=model, model_without_ddp=model, optimizer=optimizer, loss_scaler=loss_scaler,
                    epoch="min_loss_k_fold_split_0")                               # :198
    ckpt_path = os.path.join(args.output_dir, "checkpoint-min_loss_k_fold_split_0.pth")

    # ---- oracle: same loop semantics on the CPU (fp32)
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN and not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
              for k, v in P0.items() if v.is_floating_point()}
    named = [(k, v) for k, v in leaves.items() if v.requires_grad]
    opt_o = torch.optim.AdamW(O.weight_decay_groups(named, args.weight_decay), lr=args.lr, betas=(0.9, 0.95))
    k = 0
    for epoch in range(args.epochs):
        edge_w = 0.01 * (1 - epoch / args.epochs)
        sums = {"loss": 0.0, "reconstruction_loss": 0.0, "edge_map_loss": 0.0, "contr_loss": 0.0}
        for step, (aug, orig, _) in enumerate(loader):
            lr = O.cosine_lr(step / n_iter + epoch, args.lr, args.min_lr, args.warmup_epochs, args.epochs)
            for grp in opt_o.param_groups:
                grp["lr"] = lr
            if contrastive:
                losses, _, _, p1, p2, z1, z2 = O.forward_contrastive(aug, orig, leaves, cfg, args.mask_ratio, noises[k],
                                                                     noises[k + 1], edge_w, with_edge=True)
                contr = O.contrastive_loss(p1, p2, z1, z2, args.contr_weight)
            else:
                losses, _, _, _ = O.forward(aug, leaves, cfg, args.mask_ratio, noises[k], edge_w, with_edge=True)
                contr = torch.zeros(())
            k += draws_per_step
            loss = losses[0] + contr
            opt_o.zero_grad(set_to_none=True)
            loss.backward()
            opt_o.step()
            sums["loss"] += float(loss); sums["reconstruction_loss"] += float(losses[2])
            sums["edge_map_loss"] += float(losses[1]); sums["contr_loss"] += float(contr)
        got = stats[epoch]
        for name, total in sums.items():
            want = total / n_iter
            err = abs(got[name] - want) / max(abs(want), 1e-3)
            print(f"epoch {epoch} {name}: B200 {got[name]:.6f} oracle {want:.6f} rel {err:.2e}")
            assert err < 1e-2, (epoch, name, got[name], want)
        lrs = [O.cosine_lr(s / n_iter + epoch, args.lr, args.min_lr, args.warmup_epochs, args.epochs) for s in range(n_iter)]
        assert abs(got["lr"] - sum(lrs) / n_iter) < 1e-12          # the returned dict holds global averages (:110)

    # ---- save -> load -> continue == continue (fused optimizer state, loss scale, device-side step count)
    ck = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    steps_done = args.epochs * n_iter
    assert all(float(s["step"]) == steps_done for s in ck["optimizer"]["state"].values()), "optimizer step count not saved"
    assert ck["scaler"]["_growth_tracker"] == steps_done

    def more(m, opt, scaler):
        torch.manual_seed(777)
        return train_one_stage_epoch(m, loader, opt, device, args.epochs - 1, scaler, log_writer=None, args=args, edge_map_weight=0)
    cont = more(model, optimizer, loss_scaler)
    torch.manual_seed(999)                    # a fresh process would start from different initial weights
    model2 = get_models(model_name="autoenc", args=args)
    model2.to(device)
    optimizer2 = torch.optim.AdamW(misc.add_weight_decay(model2, args.weight_decay), lr=args.lr, betas=(0.9, 0.95))
    scaler2 = NativeScaler()
    args.resume = ckpt_path
    misc.load_model(args=args, model_without_ddp=model2, optimizer=optimizer2, loss_scaler=scaler2)        # :173 (resume)
    args.resume = ""
    resumed = more(model2, optimizer2, scaler2)
    for name in ("loss", "reconstruction_loss", "contr_loss"):
        err = abs(resumed[name] - cont[name]) / max(abs(cont[name]), 1e-3)
        print(f"continue vs resume {name}: {cont[name]:.7f} {resumed[name]:.7f} rel {err:.2e}")
        assert err < 2e-4, (name, cont[name], resumed[name])
    sd1, sd2 = model.state_dict(), model2.state_dict()
    worst = max(float((sd1[n].double() - sd2[n].double()).norm() / (sd1[n].double().norm() + 1e-12)) for n in sd1
                if sd1[n].is_floating_point())
    print(f"continue vs resume parameters: worst rel diff {worst:.2e}")
    assert worst < 2e-3, worst
    assert float(next(iter(optimizer2.state_dict()["state"].values()))["step"]) == steps_done + n_iter

    # ---- hand-off to the feature extractor (:213-245)
    del model2
    vit = get_models(model_name="vit", args=args)                                  # :219
    checkpoint_model = torch.load(ckpt_path, map_location="cpu", weights_only=False)["model"]   # :223
    state_dict = vit.state_dict()
    for key in ["head.weight", "head.bias"]:
        if key in checkpoint_model and checkpoint_model[key].shape != state_dict[key].shape:
            del checkpoint_model[key]
    try:
        from model.model_utils.vit_helpers import interpolate_pos_embed            # the reference's, through the overlay
        interpolate_pos_embed(vit, checkpoint_model)                               # same grid: leaves the table untouched
    except ImportError:
        pass                                                                       # GPU box: no reference checkout
    msg = vit.load_state_dict(checkpoint_model, strict=False)                      # :237
    vit.to(device)
    assert set(msg.missing_keys) == {"head.weight", "head.bias", "fc_norm.weight", "fc_norm.bias"}, msg.missing_keys   # :242
    vit.eval()
    with torch.no_grad():
        feats = vit.forward_features(loader[0][0].to(device))                      # utils/feature_extraction.py:31
    Pv = {k: v.detach().cpu() for k, v in vit.state_dict().items()}
    want = O.vit_forward_features(loader[0][0], Pv, cfg, global_pool=True)
    err = float((feats.cpu().double() - want.double()).norm() / want.double().norm())
    print(f"forward_features vs oracle: rel {err:.2e}")
    assert feats.shape == (B, cfg["embed_dim"]) and err < 1e-2, err
    print("REPLAY_OK", model_name)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "mae_vit_base_patch16")
