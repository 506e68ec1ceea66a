"""utils.train_one_epoch.train_one_epoch (the reference's contrastive-only loop, utils/train_one_epoch.py:117-180) on CPU
against a hand-rolled restatement of that loop's semantics: per-iteration lr on accumulation boundaries, loss =
-(cos(p1, z2) + cos(p2, z1)) / 2, division by accum_iter, clipping at ``max_norm``, zero_grad on update steps, meters
``loss`` / ``lr`` returned as global averages, TensorBoard tags at epoch_1000x."""
import argparse
import copy

import torch

from vit_ae_plus_plus_b200.utils import lr_sched, misc
from vit_ae_plus_plus_b200.utils.train_one_epoch import train_one_epoch


class _Siam(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = torch.nn.Linear(12, 8)
        self.pred = torch.nn.Linear(8, 8)

    def forward(self, a, b):
        z1, z2 = self.enc(a.flatten(1)), self.enc(b.flatten(1))
        return self.pred(z1), self.pred(z2), z1.detach(), z2.detach()


class _Writer:
    log_dir = "mem"

    def __init__(self):
        self.rows = []

    def add_scalar(self, tag, value, x):
        self.rows.append((tag, value, x))


def test_train_one_epoch_matches_loop_semantics():
    torch.manual_seed(0)
    model = _Siam()
    ref = copy.deepcopy(model)
    g = torch.Generator().manual_seed(1)
    loader = [(torch.randn(4, 3, 4, generator=g), torch.randn(4, 3, 4, generator=g), torch.zeros(4)) for _ in range(6)]
    args = argparse.Namespace(accum_iter=2, lr=1e-2, min_lr=1e-4, warmup_epochs=1, epochs=4)
    crit = torch.nn.CosineSimilarity(dim=1)
    opt = torch.optim.AdamW(model.parameters(), lr=args.lr)
    writer = _Writer()
    out = train_one_epoch(model, crit, loader, opt, "cpu", 1, misc.NativeScalerWithGradNormCount(), max_norm=0.5,
                          log_writer=writer, args=args)

    opt_r = torch.optim.AdamW(ref.parameters(), lr=args.lr)
    opt_r.zero_grad()
    losses, lrs = [], []
    for step, (aug, orig, _) in enumerate(loader):
        if step % args.accum_iter == 0:
            lr_sched.adjust_learning_rate(opt_r, step / len(loader) + 1, args)
        p1, p2, z1, z2 = ref(orig, aug)
        loss = -(crit(p1, z2).mean() + crit(p2, z1).mean()) * 0.5
        losses.append(loss.item())
        lrs.append(opt_r.param_groups[0]["lr"])
        (loss / args.accum_iter).backward()
        if (step + 1) % args.accum_iter == 0:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
            opt_r.step()
            opt_r.zero_grad()
    assert abs(out["loss"] - sum(losses) / len(losses)) < 1e-6
    assert abs(out["lr"] - sum(lrs) / len(lrs)) < 1e-12
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=1e-7), (a - b).abs().max()
    # TensorBoard: one ``loss`` and one ``lr`` row per update step, x = int((step / n + epoch) * 1000)
    xs = [int((s / len(loader) + 1) * 1000) for s in range(len(loader)) if (s + 1) % args.accum_iter == 0]
    assert [r[2] for r in writer.rows if r[0] == "loss"] == xs and [r[2] for r in writer.rows if r[0] == "lr"] == xs
    logged = [r[1] for r in writer.rows if r[0] == "loss"]
    assert all(abs(v - losses[s]) < 1e-6 for v, s in zip(logged, [1, 3, 5]))
