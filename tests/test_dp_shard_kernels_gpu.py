"""The kernels of the sharded data-parallel step (include/vitae_b200.h: vitae_dp_reduce_shard, vitae_sum_partials,
vitae_optim_finalize_peers, vitae_adamw_shard) on ONE GPU: the "ranks" are W sets of buffers on the same device -- the
kernels only dereference the base addresses they are given, peer memory or not.  Checked against the replicated step
(vitae_optim_prepare + vitae_adamw_flat on the mean gradient), which tests/test_step_gpu.py pins to torch.optim.AdamW +
GradScaler.  The real multi-GPU run (symmetric allocations, barriers, overlap with the backward) is tests/dp_check.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = 8            # granule shift of the tests: 256 elements, so that small buffers still spread over all ranks


def _layout(n_chunks, seed):
    """group map (two parameter groups + some frozen chunks) and fp32-everywhere flags per 64-element chunk"""
    g = torch.Generator().manual_seed(seed)
    gm = torch.randint(0, 2, (n_chunks,), generator=g).to(torch.uint8)
    gm[torch.rand(n_chunks, generator=g) < 0.05] = 255
    wide = (torch.rand(n_chunks, generator=g) < 0.3).to(torch.uint8)
    return gm, wide


@pytest.mark.parametrize("W", [2, 3, 8])
@pytest.mark.parametrize("use_scaler", [True, False])
def test_sharded_step_equals_replicated_step(W, use_scaler):
    from vit_ae_plus_plus_b200 import ops
    n = 64 * 1237                                  # ragged against the 256-element granules
    slices = [(0, 64 * 100), (64 * 100, 64 * 101), (64 * 101, 64 * 700), (64 * 700, n)]    # "stage slices" of the backward
    gen = torch.Generator().manual_seed(11 * W + use_scaler)
    gm, wide = _layout(n // 64, 5)
    gm_d, wide_d = gm.to(DEV), wide.to(DEV)
    p0 = torch.randn(n, generator=gen)
    rows = [(2e-3, 0.9, 0.95, 1e-8, 0.0), (2e-3, 0.9, 0.95, 1e-8, 0.05)]
    scale0 = 512.0 if use_scaler else 1.0
    # replicated reference state
    p_ref, m_ref, v_ref = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p16_ref = p_ref.bfloat16()
    ctl_ref = torch.tensor([scale0, 0, 0, 0, 0, 0, 0, 0], dtype=torch.float32, device=DEV)
    ws = torch.empty(ops.optim_workspace_bytes(), dtype=torch.uint8, device=DEV)
    # "ranks"
    p32 = [p0.clone().to(DEV) for _ in range(W)]
    p16 = [p32[r].bfloat16() for r in range(W)]
    grads = [torch.zeros(n, device=DEV) for _ in range(W)]
    ms = [torch.zeros(n, device=DEV) for _ in range(W)]
    vs = [torch.zeros(n, device=DEV) for _ in range(W)]
    ctls = [ctl_ref.clone() for _ in range(W)]
    CAP = 148 * 4
    parts = [torch.zeros(len(slices) * CAP + 64, device=DEV) for _ in range(W)]
    t_g, t_p32, t_p16 = (ops.peer_table([t.data_ptr() for t in ts]) for ts in (grads, p32, p16))
    t_tot = ops.peer_table([t.data_ptr() + 4 * len(slices) * CAP for t in parts])
    owner = (torch.arange(n) >> G) % W
    for step in range(1, 5):
        local = [torch.randn(n, generator=gen) * scale0 for _ in range(W)]
        if use_scaler and step == 3:
            local[W - 1][777] = float("inf")          # one rank overflows: every rank must skip
        for r in range(W):
            grads[r].copy_(local[r])
        # ---- replicated: mean in rank order, then the plain kernels
        mean = grads[0].clone()
        for r in range(1, W):
            mean += grads[r]
        mean *= 1.0 / W
        ops.optim_prepare(mean, n, ctl_ref, ws, 2.0, 0.5, 3, use_scaler)
        ops.adamw_flat(p_ref, mean, m_ref, v_ref, p16_ref, n, gm_d, rows, ctl_ref)
        # ---- sharded: every rank reduces its part of every slice ...
        for r in range(W):
            parts[r].zero_()
            for k, (a, b) in enumerate(slices):
                nb = ops.dp_reduce_shard(t_g, W, r, a, b, G, 1.0 / W, parts[r][k * CAP:(k + 1) * CAP], max_blocks=7 if k == 2 else 0)
                assert nb == ops.dp_reduce_shard_blocks(a, b, G, W, r, 7 if k == 2 else 0)
            ops.sum_partials(parts[r], len(slices) * CAP, parts[r][len(slices) * CAP:])
        for r in range(W):
            mine = (owner == r).to(DEV)
            assert torch.equal(grads[r][mine], mean[mine]), (step, r)                 # the owner holds the mean, bit for bit
            assert torch.equal(grads[r][~mine], local[r].to(DEV)[~mine])              # the rest of its copy is untouched
        # ... then the same control block everywhere and the update + all-gather of its part
        for r in range(W):
            ops.optim_finalize_peers(t_tot, W, 1, ctls[r], 2.0, 0.5, 3, use_scaler)
        for r in range(W):
            ops.adamw_shard(t_p32, t_p16, W, r, 0, n, G, grads[r], ms[r], vs[r], gm_d, wide_d, rows, ctls[r])
        torch.cuda.synchronize()
        c_ref = ctl_ref.tolist()
        for r in range(W):
            c = ctls[r].tolist()
            assert c[:4] == c_ref[:4] and c[5] == c_ref[5], (step, r, c, c_ref)
            assert abs(c[4] - c_ref[4]) <= 1e-5 * abs(c_ref[4]) or (c[4] != c[4] and c_ref[4] != c_ref[4]) \
                or c[4] == c_ref[4], (c[4], c_ref[4])
            mine = (owner == r).to(DEV)
            everywhere = (wide_d.repeat_interleave(64) != 0) | mine
            assert torch.equal(p16[r].view(torch.int16), p16_ref.view(torch.int16)), (step, r)       # shadow: complete on every rank
            assert torch.equal(p32[r][everywhere], p_ref[everywhere]), (step, r)                     # master: own part + flagged chunks
            assert torch.equal(ms[r][mine], m_ref[mine]) and torch.equal(vs[r][mine], v_ref[mine])   # moments: own part
        if use_scaler and step == 3:
            assert c_ref[2] == 1.0                                                    # the step was skipped


def test_stale_master_parts_are_exactly_the_unflagged_chunks_of_other_ranks():
    """After one sharded step a rank's fp32 master differs from the owner's only where (a) another rank owns the granule
    and (b) the chunk is not flagged fp32-everywhere: what dp.ShardedStep.sync_master has to pull."""
    from vit_ae_plus_plus_b200 import ops
    W, n = 4, 64 * 64
    gm = torch.zeros(n // 64, dtype=torch.uint8, device=DEV)
    wide = torch.zeros(n // 64, dtype=torch.uint8, device=DEV)
    wide[::3] = 1
    p32 = [torch.ones(n, device=DEV) for _ in range(W)]
    p16 = [torch.ones(n, device=DEV, dtype=torch.bfloat16) for _ in range(W)]
    g = torch.ones(n, device=DEV)
    ctl = torch.tensor([1.0, 0, 0, 1.0, 0, 1.0, 0, 0], device=DEV)
    t_p32, t_p16 = (ops.peer_table([t.data_ptr() for t in ts]) for ts in (p32, p16))
    for r in range(W):
        ops.adamw_shard(t_p32, t_p16, W, r, 0, n, G, g, torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), gm, wide,
                        [(1e-2, 0.9, 0.95, 1e-8, 0.0)], ctl)
    torch.cuda.synchronize()
    owner = ((torch.arange(n) >> G) % W).to(DEV)
    flagged = wide.repeat_interleave(64) != 0
    for r in range(W):
        changed = p32[r] != 1.0
        assert torch.equal(changed, (owner == r) | flagged)
        assert (p16[r].float() != 1.0).all()
