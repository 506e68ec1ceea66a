"""Per-iteration learning-rate schedule of the reference (utils/lr_sched.py:9-21): linear warm-up over
``args.warmup_epochs`` then a half cosine from ``args.lr`` down to ``args.min_lr`` at ``args.epochs``; a param group's
optional ``lr_scale`` multiplies it.  Host arithmetic only."""
import math


def lr_at(epoch: float, lr: float, min_lr: float, warmup_epochs: float, epochs: float) -> float:
    if epoch < warmup_epochs:
        return lr * epoch / warmup_epochs
    # operation order as in the reference (pi * elapsed, then / span): the schedule is bit-identical, not 1 ulp off
    phase = math.pi * (epoch - warmup_epochs) / (epochs - warmup_epochs)
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(phase))


def adjust_learning_rate(optimizer, epoch, args):
    lr = lr_at(epoch, args.lr, args.min_lr, args.warmup_epochs, args.epochs)
    for group in optimizer.param_groups:
        group["lr"] = lr * group["lr_scale"] if "lr_scale" in group else lr
    return lr
