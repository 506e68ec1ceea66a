"""Drop-in for the reference's ``utils/train_one_epoch.py``: same two entry points, same arguments, same returned
``{meter: global_avg}`` dict, same meters and TensorBoard tags -- restated for a B200-rate step.

What is kept from the reference loop (utils/train_one_epoch.py:21-110): per-iteration lr schedule on accumulation
boundaries (:44-45), ``model(view1=, view2=, mask_ratio=, edge_map_weight=)`` for the contrastive model (:51),
``loss[0] + contr_loss`` (:58), division by ``accum_iter`` and ``loss_scaler(..., update_grad=)`` (:70-72),
``zero_grad`` on update steps (:73-74), meters ``lr / edge_map_loss / reconstruction_loss / perceptual_loss /
contr_loss / loss`` (:60-64,78-81), TensorBoard scalars at ``epoch_1000x`` (:90-101), abort on a non-finite loss (:66-68).

What changes, and why (a ViT-B step is a few ms here; SURVEY.md 3.1 counts >= 11 device syncs per step upstream):
  * the five ``.item()`` reads, ``torch.cuda.synchronize()`` and ``torch.cuda.empty_cache()`` of every step are replaced
    by ONE read of a stacked device tensor every ``print_freq`` steps; consequently the non-finite check fires at the
    next flush instead of the same step (GradScaler already skips the update of a step with inf/nan gradients);
  * the five per-step scalar all-reduces become one all-reduce of the stacked scalars per flush;
  * a plain ``MaskedAutoencoderViT`` (3-tuple forward) is accepted as well: its contrastive term is 0;
  * with several ranks, gradient all-reduce is skipped on accumulation micro-steps (``model.no_sync()``);
  * batches are prefetched to the device on a copy stream one step ahead (``misc.DevicePrefetcher``): the 134 MB
    host -> device copy of a 4 x 4 x 128^3 batch (:47-48) costs more PCIe time than half a B200 step.
"""
from __future__ import annotations

import contextlib
import math
import sys
from typing import Iterable

import torch

from . import lr_sched, misc

PRINT_FREQ = 20


class _CosinePairLoss(torch.autograd.Function):
    """contr_weight * -(cos(p1, z2).mean() + cos(p2, z1).mean()) / 2 in two launches forward and one backward
    (vitae_cosine_loss_fwd / _bwd, csrc/predictor.cu); z1 / z2 are detached by the model (vit_autoenc.py:285)."""

    @staticmethod
    def forward(ctx, p1, p2, z1, z2, weight):
        from .. import ops
        p1, p2, z1, z2 = (t.detach().float().contiguous() for t in (p1, p2, z1, z2))
        ws = torch.empty(ops.cosine_loss_workspace_floats(p1.shape[0]), dtype=torch.float32, device=p1.device)
        loss = torch.empty(1, dtype=torch.float32, device=p1.device)
        ops.cosine_loss_fwd(p1, z2, p2, z1, float(weight), ws, loss)
        ctx.save_for_backward(p1, p2, z1, z2, ws)
        ctx.weight = float(weight)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        from .. import ops
        p1, p2, z1, z2, ws = ctx.saved_tensors
        dp1, dp2 = torch.empty_like(p1), torch.empty_like(p2)
        ops.cosine_loss_bwd(p1, z2, p2, z1, ctx.weight, ws, dloss.detach().float().reshape(1).contiguous(), dp1, dp2)
        return dp1, dp2, None, None, None


def compute_contrastive_loss(args, criterion, p1, p2, z1, z2):
    """-(cos(p1, z2) + cos(p2, z1)) / 2 scaled by ``args.contr_weight`` (reference :113-114).  With the reference's
    criterion (``nn.CosineSimilarity(dim=1)``, :32) on CUDA rows the fused kernels compute it; anything else is evaluated
    with the criterion as given."""
    if (p1.is_cuda and p1.dim() == 2 and type(criterion) is torch.nn.CosineSimilarity and criterion.dim == 1
            and criterion.eps == 1e-8 and not z1.requires_grad and not z2.requires_grad):
        return _CosinePairLoss.apply(p1, p2, z1, z2, args.contr_weight)
    return args.contr_weight * (-(criterion(p1, z2).mean() + criterion(p2, z1).mean()) * 0.5)


class _DeferredScalars:
    """Collects per-step device scalars and turns them into python floats with one sync per flush."""

    def __init__(self, names, metric_logger, log_writer, on_nonfinite):
        self.names, self.logger, self.writer, self.on_nonfinite = names, metric_logger, log_writer, on_nonfinite
        self.rows, self.meta = [], []

    def add(self, tensors, lr, x_axis, log_step):
        self.rows.append(torch.stack([t.detach().float().reshape(()) for t in tensors]))
        self.meta.append((lr, x_axis, log_step))

    def flush(self):
        if not self.rows:
            return
        block = torch.stack(self.rows)
        local = block.tolist()                                    # the one device sync
        reduced = local
        if misc.get_world_size() > 1:      # every rank takes part (the reference all-reduces unconditionally, :78-81)
            torch.distributed.all_reduce(block)
            reduced = (block / misc.get_world_size()).tolist()
        for vals, red, (lr, x_axis, log_step) in zip(local, reduced, self.meta):
            named = dict(zip(self.names, vals))
            if not math.isfinite(named["loss"]):
                self.on_nonfinite(named["loss"])
            self.logger.update(**named)
            self.logger.update(lr=lr)
            if self.writer is not None and log_step:
                rn = dict(zip(self.names, red))
                self.writer.add_scalar("train_loss", rn["loss"], x_axis)
                self.writer.add_scalar("lr", lr, x_axis)
                self.writer.add_scalar("reconstruction_loss", rn["reconstruction_loss"], x_axis)
                self.writer.add_scalar("sobel_loss", rn["edge_map_loss"], x_axis)
                self.writer.add_scalar("perceptual_loss", rn["perceptual_loss"], x_axis)
                self.writer.add_scalar("contr_loss", rn["contr_loss"], x_axis)
        self.rows, self.meta = [], []


def _abort(value):
    print("Loss is {}, stopping training".format(value))
    sys.exit(1)


def train_one_stage_epoch(model: torch.nn.Module, data_loader: Iterable, optimizer: torch.optim.Optimizer,
                          device: torch.device, epoch: int, loss_scaler, log_writer=None, args=None,
                          edge_map_weight=0):
    model.train(True)
    metric_logger = misc.MetricLogger(delimiter="  ")
    metric_logger.add_meter("lr", misc.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    header = "Epoch: [{}]".format(epoch)
    criterion = torch.nn.CosineSimilarity(dim=1).to(device)
    accum_iter = args.accum_iter
    n_iter = len(data_loader)
    names = ["edge_map_loss", "reconstruction_loss", "perceptual_loss", "contr_loss", "loss"]
    deferred = _DeferredScalars(names, metric_logger, log_writer, _abort)
    contrastive = hasattr(model, "predictor")
    zero = None

    optimizer.zero_grad()
    if log_writer is not None:
        print("log_dir: {}".format(log_writer.log_dir))
    # the loop never looks at ``pred`` (utils/train_one_epoch.py:51-53 discards it): skip the fp32 copy of the prediction
    # (201 MB per 4 x 128^3 x 4 batch) for the duration of the epoch; ``model.pred_dtype`` is restored afterwards
    saved_pred_dtype = getattr(model, "pred_dtype", None)
    if saved_pred_dtype is not None:
        model.pred_dtype = torch.bfloat16
    # args.device_normalize (optional, not in the reference's argument set): the loader yields raw volumes in their storage
    # dtype and the normalisation of Dataset._normalize_data runs on the device (misc.DevicePrefetcher)
    batches = (misc.DevicePrefetcher(data_loader, device, normalize=getattr(args, "device_normalize", None))
               if torch.device(device).type == "cuda" else data_loader)
    for step, (sample, original_volume, _) in enumerate(
            metric_logger.log_every(batches, PRINT_FREQ, header, before_print=deferred.flush)):
        if step % accum_iter == 0:
            lr_sched.adjust_learning_rate(optimizer, step / n_iter + epoch, args)
        update = (step + 1) % accum_iter == 0
        sample = sample.to(device, non_blocking=True)
        original_volume = original_volume.to(device, non_blocking=True)

        if contrastive:
            losses, _pred, _mask, p1, p2, z1, z2 = model(view1=sample, view2=original_volume,
                                                         mask_ratio=args.mask_ratio, edge_map_weight=edge_map_weight)
            contr_loss = compute_contrastive_loss(args, criterion, p1, p2, z1, z2)
        else:
            losses, _pred, _mask = model(sample, mask_ratio=args.mask_ratio, edge_map_weight=edge_map_weight)
            if zero is None:
                zero = torch.zeros((), device=losses[0].device)
            contr_loss = zero
        loss = losses[0] + contr_loss
        deferred.add([losses[1], losses[2], losses[3], contr_loss, loss], optimizer.param_groups[0]["lr"],
                     int((step / n_iter + epoch) * 1000), update)

        sync = contextlib.nullcontext() if (update or not hasattr(model, "no_sync")) else model.no_sync()
        with sync:
            loss_scaler(loss / accum_iter, optimizer, parameters=model.parameters(), update_grad=update)
        if update:
            optimizer.zero_grad()

    if saved_pred_dtype is not None:
        model.pred_dtype = saved_pred_dtype
    eng = getattr(model, "_engine", None)
    if eng is not None:
        eng.wait_params()
    deferred.flush()
    metric_logger.synchronize_between_processes()
    print("Averaged stats:", metric_logger)
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}


def train_one_epoch(model, criterion, data_loader, optimizer, device, epoch, loss_scaler, max_norm=0, log_writer=None,
                    args=None):
    """Contrastive-only loop of the reference (utils/train_one_epoch.py:117-180; no script calls it): the model maps
    ``(original, augmented)`` to ``(p1, p2, z1, z2)`` and ``criterion`` is a cosine similarity.  Kept for signature
    compatibility; meters ``lr`` and ``loss``, TensorBoard tags ``loss`` and ``lr``."""
    model.train(True)
    metric_logger = misc.MetricLogger(delimiter="  ")
    metric_logger.add_meter("lr", misc.SmoothedValue(window_size=1, fmt="{value:.6f}"))
    header = "Epoch: [{}]".format(epoch)
    accum_iter = args.accum_iter
    n_iter = len(data_loader)
    pending = []

    def flush():
        if not pending:
            return
        block = torch.stack([p[0] for p in pending])
        local = block.tolist()
        reduced = local
        if misc.get_world_size() > 1:      # every rank takes part, whether or not it owns a SummaryWriter
            torch.distributed.all_reduce(block)
            reduced = (block / misc.get_world_size()).tolist()
        for v, r, (_, lr, x_axis, log_step) in zip(local, reduced, pending):
            if not math.isfinite(v):
                _abort(v)
            metric_logger.update(loss=v)
            metric_logger.update(lr=lr)
            if log_writer is not None and log_step:
                log_writer.add_scalar("loss", r, x_axis)
                log_writer.add_scalar("lr", lr, x_axis)
        pending.clear()

    optimizer.zero_grad()
    if log_writer is not None:
        print("log_dir: {}".format(log_writer.log_dir))
    for step, (augmented, original, _) in enumerate(
            metric_logger.log_every(data_loader, PRINT_FREQ, header, before_print=flush)):
        if step % accum_iter == 0:
            lr_sched.adjust_learning_rate(optimizer, step / n_iter + epoch, args)
        update = (step + 1) % accum_iter == 0
        augmented = augmented.to(device, non_blocking=True)
        original = original.to(device, non_blocking=True)
        p1, p2, z1, z2 = model(original, augmented)
        loss = -(criterion(p1, z2).mean() + criterion(p2, z1).mean()) * 0.5
        max_lr = max(g["lr"] for g in optimizer.param_groups)
        pending.append((loss.detach().float().reshape(()), max_lr, int((step / n_iter + epoch) * 1000), update))
        loss_scaler(loss / accum_iter, optimizer, clip_grad=max_norm, parameters=model.parameters(), create_graph=False,
                    update_grad=update)
        if update:
            optimizer.zero_grad()
    flush()
    metric_logger.synchronize_between_processes()
    print("Averaged stats:", metric_logger)
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}
