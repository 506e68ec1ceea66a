"""Probe of the NVLS building blocks on the live process group (torchrun, >= 2 GPUs): symmetric allocation + multicast
mapping (torch.distributed._symmetric_memory: plumbing), vitae_dp_reduce_shard and vitae_adamw_flat_mc checked against
NCCL / the local kernels, and their timings next to NCCL's all-reduce.  Rank 0 prints JSON lines."""
import json, os, sys
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_ae_plus_plus_b200 import ops  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(t.item(), 4)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    say = (lambda **kw: print(json.dumps(kw), flush=True)) if rank == 0 else (lambda **kw: None)
    N = 112 * 1024 * 1024 // 64 * 64          # ~ the flat buffer of ViT-B (elements)
    g = symm_mem.empty(N, dtype=torch.float32, device=dev)
    p32 = symm_mem.empty(N, dtype=torch.float32, device=dev)
    p16 = symm_mem.empty(N, dtype=torch.bfloat16, device=dev)
    hg, hp, hh = (symm_mem.rendezvous(t, dist.group.WORLD) for t in (g, p32, p16))
    say(step="rendezvous", multicast=[int(h.multicast_ptr != 0) for h in (hg, hp, hh)], world=hg.world_size, rank=hg.rank,
        signal_pad_size=hg.signal_pad_size)
    if not hg.multicast_ptr:
        say(step="no multicast: stop")
        return
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    g.copy_(torch.randn(N, device=dev, generator=gen))
    ref = g.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    # ---- reduce-scatter through the switch
    per = (N // 64 + world - 1) // world * 64
    a, b = min(N, rank * per), min(N, (rank + 1) * per)
    partials = torch.zeros(148 * 4, device=dev)
    g_work = symm_mem.empty(N, dtype=torch.float32, device=dev)
    hw = symm_mem.rendezvous(g_work, dist.group.WORLD)
    g_work.copy_(g)
    hw.barrier(channel=0)
    nb = ops.dp_reduce_shard(hw.multicast_ptr, g_work, a, b - a, 1.0 / world, partials)
    hw.barrier(channel=0)
    torch.cuda.synchronize()
    err = (g_work[a:b] - ref[a:b]).abs().max().item()
    sq = partials[:nb].double().sum().item()
    sq_ref = ref[a:b].double().pow(2).sum().item()
    say(step="reduce_shard", max_abs_err=err, sqnorm_rel_err=abs(sq - sq_ref) / sq_ref, shard=[a, b], blocks=nb)
    # ---- sharded AdamW with multicast stores
    p32.copy_(torch.randn(N, device=dev, generator=torch.Generator(device=dev).manual_seed(3)))   # same on every rank
    p16.zero_()
    m = torch.zeros(N, device=dev); v = torch.zeros(N, device=dev)
    ctl = torch.tensor([1.0, 0, 0, 1.0, 0, 1.0, 0, 0], device=dev)
    gm = torch.zeros(N // 64, dtype=torch.uint8, device=dev)
    rows = [(1e-3, 0.9, 0.95, 1e-8, 0.05)]
    # expected: the plain kernel over the whole buffer with the mean gradient
    pe, p16e, me, ve = p32.clone(), torch.zeros(N, dtype=torch.bfloat16, device=dev), m.clone(), v.clone()
    ops.adamw_flat(pe, ref, me, ve, p16e, N, gm, rows, ctl)
    hp.barrier(channel=0)
    if b > a:
        ops.adamw_flat_mc(p32, g_work, m, v, hp.multicast_ptr, hh.multicast_ptr, b - a, gm, rows, ctl, start=a)
    hp.barrier(channel=0)
    torch.cuda.synchronize()
    bad32 = (p32 != pe).sum().item()
    bad16 = (p16.view(torch.int16) != p16e.view(torch.int16)).sum().item()
    t = torch.tensor([bad32, bad16], device=dev, dtype=torch.float64)
    dist.all_reduce(t)
    say(step="adamw_mc", mismatching_master=t[0].item(), mismatching_shadow=t[1].item(),
        moments_ok=bool(torch.equal(m[a:b], me[a:b]) and torch.equal(v[a:b], ve[a:b])))
    # ---- timings
    res = {}
    for mb in (50, 85, 450):
        n = min(N, mb * 1000 * 1000 // 4 // 64 * 64)
        res[f"nccl_allreduce_{mb}MB_ms"] = timeit(lambda: dist.all_reduce(g[:n], op=dist.ReduceOp.AVG))
    res["barrier_ms"] = timeit(lambda: hw.barrier(channel=0))
    for cap in (0, 148, 32, 16, 8):
        res[f"reduce_shard_full_blocks{cap}_ms"] = timeit(lambda: ops.dp_reduce_shard(hw.multicast_ptr, g_work, a, b - a, 1.0 / world, partials, max_blocks=cap))
    res["adamw_mc_shard_ms"] = timeit(lambda: ops.adamw_flat_mc(p32, g_work, m, v, hp.multicast_ptr, hh.multicast_ptr, b - a, gm, rows, ctl, start=a))
    res["adamw_local_full_ms"] = timeit(lambda: ops.adamw_flat(pe, ref, me, ve, p16e, N, gm, rows, ctl))
    # ---- micro-benchmarks of the primitives over this rank's shard (GB/s of shard bytes)
    import ctypes
    from vit_ae_plus_plus_b200 import _lib
    lib = _lib.load()
    lib.vitae_debug_mc_reduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.vitae_debug_mc_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    nsh = b - a
    gbs = lambda ms: round(nsh * 4 / ms / 1e6, 1)
    scratch = torch.empty(nsh, device=dev)
    sweep = {}
    for mode in (1, 4, 8, 104, 108):
        for blocks in (148 * 8, 148 * 4, 148 * 2, 148, 64, 32, 16):
            ms = timeit(lambda: lib.vitae_debug_mc_reduce(hw.multicast_ptr + 4 * a, scratch.data_ptr(), nsh, mode, blocks, st()), n=5, warm=2)
            sweep[f"ld_reduce_mode{mode}_blocks{blocks}"] = gbs(ms)
    peer = hw.buffer_ptrs[(rank + 1) % world]
    for mode in (0, 1, 2):
        for blocks in (148 * 8, 148 * 2, 148, 64, 32, 16):
            dst = (hw.multicast_ptr if mode < 2 else peer) + 4 * a
            ms = timeit(lambda: lib.vitae_debug_mc_copy(scratch.data_ptr(), dst, nsh, mode, blocks, st()), n=5, warm=2)
            sweep[f"{'multimem_st' if mode < 2 else 'p2p_st'}_mode{mode}_blocks{blocks}"] = gbs(ms)
    say(step="primitive sweep: GB/s of this rank's shard", shard_mb=round(nsh * 4 / 1e6, 1), **sweep)
    try:
        res["torch_multimem_allreduce_450MB_ms"] = timeit(lambda: torch.ops.symm_mem.multimem_all_reduce_(g_work, "sum", dist.group.WORLD.group_name))
    except Exception as e:      # noqa: BLE001
        res["torch_multimem_allreduce"] = f"unavailable: {type(e).__name__}: {e}"[:200]
    say(step="timings", world=world, elements=N, **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
