"""Step-level behaviour of the B200 path: CUDA-graph replay vs eager launches, the side lane, the fused optimizer
(vitae_optim_prepare / vitae_adamw_flat behind utils.misc.NativeScalerWithGradNormCount) vs the reference's
GradScaler + torch.optim.AdamW sequence (utils/misc.py:257-271), and the 100-step loss curve against the CPU oracle
(BASELINE.json north_star: "100-step loss curve matching the reference to 1e-2")."""
import argparse
import math
from functools import partial

import pytest
import torch
from torch import nn

from oracle import mae_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build(cfg, P):
    from vit_ae_plus_plus_b200.model.vit_autoenc import MaskedAutoencoderViT
    m = MaskedAutoencoderViT(**cfg, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                             args=argparse.Namespace(perceptual_weight=0, use_imagenet=False))
    m.load_state_dict(P, strict=True)
    return m.cuda()


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-30)


def _signal(name, t):
    """Drops the K third of a qkv bias: softmax is invariant to a shift of the scores along the key axis, so that
    gradient is exactly zero in the reference's algebra (model/vit.py:117-119) and pure rounding noise in any
    implementation -- Adam's normalisation turns that noise into +-lr steps that no two code paths reproduce."""
    if name.endswith("attn.qkv.bias"):
        d = t.shape[0] // 3
        return torch.cat([t[:d], t[2 * d:]])
    return t


# ------------------------------------------------------------------------------------------------ optimizer kernels
@pytest.mark.parametrize("use_scaler", [True, False])
def test_optim_prepare_and_adamw_flat_match_torch(use_scaler):
    """Two parameter groups (decay / no decay, timm add_weight_decay at k_fold_..._brats.py:168) interleaved in one flat
    buffer, against torch.optim.AdamW + torch.amp.GradScaler semantics, including a skipped (inf) step."""
    from vit_ae_plus_plus_b200 import ops
    gen = torch.Generator().manual_seed(7)
    sizes = [(0, 640), (640, 128), (768, 64 * 301), (768 + 64 * 301, 192)]          # 64-aligned tensors
    n = sizes[-1][0] + sizes[-1][1]
    group_of = [1, 0, 1, 0]
    p0 = torch.randn(n, generator=gen)
    refs = [torch.nn.Parameter(p0[o:o + k].clone()) for o, k in sizes]
    opt = torch.optim.AdamW([{"params": [refs[1], refs[3]], "weight_decay": 0.0},
                             {"params": [refs[0], refs[2]], "weight_decay": 0.05}], lr=2e-3, betas=(0.9, 0.95))
    p = p0.clone().to(DEV)
    g = torch.zeros(n, device=DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p16 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    gm = torch.full((n // 64,), 255, dtype=torch.uint8)
    for (o, k), gi in zip(sizes, group_of):
        gm[o // 64:(o + k) // 64] = gi
    gm = gm.to(DEV)
    scale0 = 1024.0 if use_scaler else 1.0
    ctl = torch.tensor([scale0, 0, 0, 0, 0, 0, 0, 0], dtype=torch.float32, device=DEV)
    ws = torch.empty(ops.optim_workspace_bytes(), dtype=torch.uint8, device=DEV)
    rows = lambda lr: [(lr, 0.9, 0.95, 1e-8, 0.0), (lr, 0.9, 0.95, 1e-8, 0.05)]
    growth_interval = 3
    scale, tracker, taken = scale0, 0, 0
    for step in range(1, 8):
        lr = 2e-3 * (1 - 0.1 * step)
        for grp in opt.param_groups:
            grp["lr"] = lr
        grad = torch.randn(n, generator=gen)
        inf_step = use_scaler and step == 4
        gdev = (grad * scale).to(DEV)
        if inf_step:
            gdev[12345] = float("inf")
        g.copy_(gdev)
        before = p.clone()
        ops.optim_prepare(g, n, ctl, ws, 2.0, 0.5, growth_interval, use_scaler)
        ops.adamw_flat(p, g, m, v, p16, n, gm, rows(lr), ctl)
        c = ctl.tolist()
        if inf_step:
            assert c[2] == 1.0 and torch.equal(p, before)                 # skipped (utils/misc.py:267)
            scale, tracker = scale * 0.5, 0
        else:
            for r, (o, k) in zip(refs, sizes):
                r.grad = grad[o:o + k].clone()
            opt.step()
            taken += 1
            ref = torch.cat([r.detach() for r in refs])
            assert _rel(p.cpu(), ref) < 2e-6, step
            assert torch.equal(p16, p.bfloat16())
            assert abs(c[4] - grad.double().norm().item()) < 1e-4 * grad.double().norm().item()   # unscaled norm
            if use_scaler:
                tracker += 1
                if tracker == growth_interval:
                    scale, tracker = scale * 2.0, 0
        if use_scaler:
            assert c[0] == scale and c[1] == tracker, (step, c[:2], scale, tracker)
        assert c[5] == taken
    ref_m = torch.cat([opt.state[r]["exp_avg"] for r in refs])
    assert _rel(m.cpu(), ref_m) < 1e-5


# ------------------------------------------------------------------------------------------------ graphs / lanes
def _grads(m):
    return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("name,B", [("tiny", 2), ("small", 2)])
def test_graph_replay_and_side_lane_are_bitwise_equal_to_eager(name, B):
    cfg = O.CONFIGS[name]
    P = O.init_params(cfg, 21)
    V, C = cfg["volume_size"], cfg["in_chans"]
    _, L, _ = O.geometry(cfg)
    x = torch.randn(B, C, V, V, V, generator=torch.Generator().manual_seed(5)).cuda()
    noise = torch.rand(B, L, generator=torch.Generator().manual_seed(6))

    def run(m):
        for p in m.parameters():
            p.grad = None
        losses, pred, mask = m(x, noise=noise)
        losses[0].backward()
        torch.cuda.synchronize()
        return losses[2].detach().clone(), pred.detach().clone(), _grads(m)

    ref_m = build(cfg, P)
    ref_m.use_cuda_graph = False
    ref_m.engine().use_side_lane = False
    l0, pred0, g0 = run(ref_m)                              # plain single-stream eager launches

    from vit_ae_plus_plus_b200 import _lib
    m = build(cfg, P)
    eng = m.engine()
    n0 = _lib.load().vitae_launch_count()
    results = [run(m) for _ in range(4)]                    # call 1 eager, call 2 captures + replays, 3-4 replay
    slots = list(eng.plans[(B, int(L * (1 - 0.75)))].graphs.values())
    assert len(slots) == 2 and all(s.graph is not None and s.launches > 0 for s in slots)
    executed = _lib.load().vitae_launch_count() - n0 + eng.graph_replayed_launches
    assert executed == 4 * sum(s.launches for s in slots) + 1   # bench.py's gpu_launches bookkeeping (+1: bf16 cast)
    for l, pred, g in results:
        assert torch.equal(l, l0) and torch.equal(pred, pred0)
        assert sorted(g) == sorted(g0)
        for n in g:
            assert torch.equal(g[n], g0[n]), n
    # accumulation graph (aliased .grad kept): twice the gradient
    losses, _, _ = m(x, noise=noise)
    losses[0].backward()
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert _rel(p.grad, 2 * g0[n]) < 1e-5, n


def test_many_input_addresses_fall_back_to_a_static_copy():
    from vit_ae_plus_plus_b200 import engine as E
    cfg = O.CONFIGS["tiny"]
    P = O.init_params(cfg, 3)
    m = build(cfg, P)
    noise = torch.rand(2, 64, generator=torch.Generator().manual_seed(1))
    xs = [torch.randn(2, 1, 32, 32, 32, generator=torch.Generator().manual_seed(100 + i)).cuda()
          for i in range(E.MAX_INPUT_ADDRESSES + 3)]
    assert len({x.data_ptr() for x in xs}) == len(xs)
    ref_m = build(cfg, P)
    ref_m.use_cuda_graph = False
    for rep in range(3):
        for x in xs:
            with torch.no_grad():
                a = m(x, noise=noise)[0][2].clone()
                b = ref_m(x, noise=noise)[0][2].clone()
            assert torch.equal(a, b)
    pl = m.engine().plans[(2, 16)]
    assert len({k[1] for k in pl.graphs if k[0] == "fwd"}) == E.MAX_INPUT_ADDRESSES + 1 and pl.vol_static is not None


# ------------------------------------------------------------------------------------------------ training steps
def _train(cfg, P, xs, noises, fused, graphs, steps, lr=1e-3, overlap=False):
    from vit_ae_plus_plus_b200.utils import misc
    m = build(cfg, P)
    m.use_cuda_graph = graphs
    opt = torch.optim.AdamW(misc.add_weight_decay(m, 0.05), lr=lr, betas=(0.9, 0.95))
    scaler = misc.NativeScalerWithGradNormCount()
    scaler.allow_fused = fused
    scaler.overlap_optimizer = overlap
    curve, norms = [], []
    for i in range(steps):
        losses, _, _ = m(xs[i % len(xs)], mask_ratio=0.75, noise=noises[i])
        norms.append(scaler(losses[0], opt, parameters=m.parameters(), update_grad=True))
        opt.zero_grad()
        curve.append(losses[0].detach())
    assert (scaler._fused is not None) == fused
    return m, opt, scaler, torch.stack(curve).cpu(), torch.stack([n.detach().float().reshape(()) for n in norms]).cpu()


def test_fused_optimizer_path_matches_torch_gradscaler_adamw_path():
    cfg = O.CONFIGS["tiny"]
    P = O.init_params(cfg, 8)
    g = torch.Generator().manual_seed(4)
    xs = [torch.randn(2, 1, 32, 32, 32, generator=g).cuda() for _ in range(2)]
    noises = [torch.rand(2, 64, generator=g) for _ in range(6)]
    m_t, opt_t, sc_t, curve_t, norm_t = _train(cfg, P, xs, noises, fused=False, graphs=False, steps=6)
    m_f, opt_f, sc_f, curve_f, norm_f = _train(cfg, P, xs, noises, fused=True, graphs=True, steps=6)
    assert _rel(curve_f, curve_t) < 1e-4, (curve_f, curve_t)
    assert _rel(norm_f, norm_t) < 1e-4
    sd_t, sd_f = m_t.state_dict(), m_f.state_dict()
    for k in sd_t:
        assert _rel(_signal(k, sd_f[k]), _signal(k, sd_t[k])) < 5e-3, k   # Adam amplifies last-bit differences to ~lr steps
    # optimizer / scaler state stays in torch's layout (checkpoints: utils/misc.py:295-312)
    m_f.engine().fused_optimizer().sync_state(opt_f)
    st_t, st_f = opt_t.state_dict()["state"], opt_f.state_dict()["state"]
    assert sorted(st_t) == sorted(st_f)
    # tensors whose true gradient is (nearly) zero hold rounding noise only: compare them on the scale of the others
    floor = 1e-3 * max(float(st_t[i]["exp_avg"].abs().max()) for i in st_t)
    for i in st_t:
        assert float(st_f[i]["step"]) == float(st_t[i]["step"]) == 6.0
        a, b = st_f[i]["exp_avg"].double(), st_t[i]["exp_avg"].double()
        assert (a - b).abs().max().item() < 5e-3 * max(b.abs().max().item(), floor), i
    assert sc_f.state_dict()["scale"] == sc_t.state_dict()["scale"]
    assert sc_f.state_dict()["_growth_tracker"] == sc_t.state_dict()["_growth_tracker"] == 6


def test_switching_between_fused_and_torch_optimizer_paths():
    """exp_avg / exp_avg_sq are views of the flat moment buffers: a plain optimizer.step() continues from them."""
    from vit_ae_plus_plus_b200.utils import misc
    cfg = O.CONFIGS["tiny"]
    P = O.init_params(cfg, 9)
    g = torch.Generator().manual_seed(14)
    xs = [torch.randn(2, 1, 32, 32, 32, generator=g).cuda()]
    noises = [torch.rand(2, 64, generator=g) for _ in range(4)]
    m_t, _, _, curve_t, _ = _train(cfg, P, xs, noises, fused=False, graphs=True, steps=4)
    m = build(cfg, P)
    opt = torch.optim.AdamW(misc.add_weight_decay(m, 0.05), lr=1e-3, betas=(0.9, 0.95))
    scaler = misc.NativeScalerWithGradNormCount()
    curve = []
    for i in range(4):
        scaler.allow_fused = i < 2                       # two fused steps, then two through torch
        losses, _, _ = m(xs[0], mask_ratio=0.75, noise=noises[i])
        scaler(losses[0], opt, parameters=m.parameters(), update_grad=True)
        opt.zero_grad()
        curve.append(losses[0].detach())
    assert _rel(torch.stack(curve).cpu(), curve_t) < 1e-4
    for k, v in m_t.state_dict().items():
        assert _rel(_signal(k, m.state_dict()[k]), _signal(k, v)) < 5e-3, k


def test_loss_curve_100_steps_matches_oracle():
    """Same parameters, volumes and per-step mask noise through 100 AdamW steps on both sides; the oracle is the CPU
    restatement of the reference (pinned to it by tests/test_oracle_golden.py::test_adamw_loss_curve_matches_reference)."""
    cfg = O.CONFIGS["tiny"]
    steps, lr = 100, 1e-3
    P = O.init_params(cfg, 31)
    g = torch.Generator().manual_seed(32)
    xs = [torch.randn(2, 1, 32, 32, 32, generator=g) for _ in range(4)]
    noises = [torch.rand(2, 64, generator=g) for _ in range(steps)]
    # oracle
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in P.items()}
    opt = torch.optim.AdamW(O.weight_decay_groups(list(leaves.items()), 0.05), lr=lr, betas=(0.9, 0.95))
    ref = []
    for i in range(steps):
        losses, _, _, _ = O.forward(xs[i % 4], leaves, cfg, 0.75, noises[i], 0.0, with_edge=False)
        opt.zero_grad(set_to_none=True)
        losses[0].backward()
        opt.step()
        ref.append(float(losses[0].detach()))
    _, _, _, curve, _ = _train(cfg, P, [x.cuda() for x in xs], noises, fused=True, graphs=True, steps=steps, lr=lr)
    ref = torch.tensor(ref)
    err = ((curve - ref).abs() / ref.abs()).max().item()
    assert err < 1e-2, err
    assert curve[-1] < 0.9 * curve[0]                     # and it actually trains


def test_loss_curve_vit_base_128_matches_oracle():
    """The north star quotes the loss curve beside ViT-B 128^3: 30 AdamW steps of configs[1]'s model at batch 1 (the CPU
    oracle needs ~1.5 s per step on the box's cores), same parameters / volumes / per-step mask noise on both sides, 1e-2
    per step.  The curve is written to gpurun_out/loss_curve_vit_base_128.json (summarised under profiles/)."""
    import json
    import os
    cfg = O.CONFIGS["vit_base_128"]
    steps, lr = 30, 1.5e-4
    P = O.init_params(cfg, 41)
    g = torch.Generator().manual_seed(42)
    xs = [torch.randn(1, 4, 128, 128, 128, generator=g) for _ in range(3)]
    noises = [torch.rand(1, 512, generator=g) for _ in range(steps)]
    leaves = {k: v.clone().requires_grad_(k not in O.FROZEN) for k, v in P.items()}
    opt = torch.optim.AdamW(O.weight_decay_groups(list(leaves.items()), 0.05), lr=lr, betas=(0.9, 0.95))
    ref = []
    for i in range(steps):
        losses, _, _, _ = O.forward(xs[i % 3], leaves, cfg, 0.75, noises[i], 0.0, with_edge=False)
        opt.zero_grad(set_to_none=True)
        losses[0].backward()
        opt.step()
        ref.append(float(losses[0].detach()))
    _, _, _, curve, _ = _train(cfg, P, [x.cuda() for x in xs], noises, fused=True, graphs=True, steps=steps, lr=lr)
    ref = torch.tensor(ref)
    rel = ((curve - ref).abs() / ref.abs())
    try:
        out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        json.dump({"config": "vit_base_128 batch 1, AdamW lr 1.5e-4 wd 0.05 betas .9/.95, mask 0.75", "steps": steps,
                   "b200": curve.tolist(), "oracle": ref.tolist(), "max_rel_err": rel.max().item()},
                  open(os.path.join(out, "loss_curve_vit_base_128.json"), "w"), indent=1)
    except OSError:
        pass
    assert rel.max().item() < 1e-2, rel
    assert curve[-1] < curve[0]


def test_device_prefetcher_delivers_batches_in_order_through_rotating_buffers():
    """utils.misc.DevicePrefetcher: batch k+1 is copied on a side stream while batch k is consumed; two device buffers
    per tensor slot are reused, so a wrong event ordering shows up as a batch overwritten before it was read."""
    from vit_ae_plus_plus_b200.utils import misc
    g = torch.Generator().manual_seed(5)
    host = [(torch.randn(2, 1, 32, 32, 32, generator=g).pin_memory(), torch.randn(2, 1, 32, 32, 32, generator=g).pin_memory(),
             torch.tensor([i, i + 1])) for i in range(7)]
    pf = misc.DevicePrefetcher(host, "cuda")
    assert len(pf) == 7
    ptrs, sums = set(), []
    burn = torch.randn(2048, 2048, device=DEV)
    for i, (a, b, lab) in enumerate(pf):
        assert a.is_cuda and b.is_cuda and lab.is_cuda
        ptrs.add(a.data_ptr())
        burn = burn @ burn * 1e-3                      # keep the consumer stream busy while the next copy is in flight
        sums.append((a.double().sum() + 2 * b.double().sum() + lab.sum()).reshape(1))
    torch.cuda.synchronize()
    assert len(ptrs) == 2
    ref = [float(a.double().sum() + 2 * b.double().sum() + lab.sum()) for a, b, lab in host]
    got = torch.cat(sums).cpu().tolist()
    assert all(abs(x - y) < 1e-9 * (1 + abs(y)) for x, y in zip(got, ref))
    assert list(misc.DevicePrefetcher([], "cuda")) == []


@pytest.mark.parametrize("graphs", [True, False])
def test_overlapped_optimizer_step_is_bit_identical(graphs):
    """overlap_optimizer: AdamW runs per layer group on its own stream and the next forward waits group by group
    (external events inside the captured forward).  Same kernels, same order per element -> identical bits; a missing
    wait shows up as a forward that read a half-updated layer."""
    cfg = O.CONFIGS["small"] if "small" in O.CONFIGS else O.CONFIGS["tiny"]
    P = O.init_params(cfg, 21)
    g = torch.Generator().manual_seed(22)
    V, C = cfg["volume_size"], cfg["in_chans"]
    L = (V // cfg["patch_size"]) ** 3
    xs = [torch.randn(2, C, V, V, V, generator=g).cuda() for _ in range(2)]
    noises = [torch.rand(2, L, generator=g) for _ in range(8)]
    m_a, _, _, curve_a, norm_a = _train(cfg, P, xs, noises, fused=True, graphs=graphs, steps=8, overlap=False)
    m_b, _, _, curve_b, norm_b = _train(cfg, P, xs, noises, fused=True, graphs=graphs, steps=8, overlap=True)
    assert torch.equal(curve_a, curve_b) and torch.equal(norm_a, norm_b)
    sd_a, sd_b = m_a.state_dict(), m_b.state_dict()
    for k in sd_a:
        assert torch.equal(sd_a[k], sd_b[k]), k
    eng = m_b.engine()
    assert len(eng.group_ranges) >= 3 and not eng.params_in_flight     # state_dict() waited for the optimizer stream
