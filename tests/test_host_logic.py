"""Host-side logic of the sharded data-parallel step that needs no GPU: the ownership arithmetic exported by the C ABI
(include/vitae_b200.h: vitae_dp_owned_elems, vitae_dp_reduce_shard_blocks) and the switch that keeps the step off outside a
single-node NCCL group."""
import torch


# ------------------------------------------------------------------------------------------------ sharded step: ownership
def test_block_cyclic_ownership_tiles_every_slice():
    """vitae_dp_owned_elems / vitae_dp_reduce_shard_blocks (host arithmetic of the C ABI, no GPU): granule q belongs to
    rank q % world, so the ranks' parts of any 64-aligned slice are disjoint and add up to the slice."""
    import random
    from vit_ae_plus_plus_b200 import ops
    rnd = random.Random(3)
    for _ in range(200):
        shift = rnd.choice([6, 8, 12, 16])
        world = rnd.randint(1, 8)
        lo = 64 * rnd.randint(0, 5000)
        hi = lo + 64 * rnd.randint(1, 20000)
        owned = [ops.dp_owned_elems(lo, hi, shift, world, r) for r in range(world)]
        assert sum(owned) == hi - lo
        # brute force on the granule grid
        G = 1 << shift
        for r in range(world):
            want = sum(min(hi, (q + 1) * G) - max(lo, q * G) for q in range(lo // G, (hi - 1) // G + 1) if q % world == r)
            assert owned[r] == want
            nb = ops.dp_reduce_shard_blocks(lo, hi, shift, world, r, 0)
            assert (nb == 0) == (owned[r] == 0) and nb <= 148 * 4
            assert ops.dp_reduce_shard_blocks(lo, hi, shift, world, r, 5) <= 5
    assert ops.dp_owned_elems(0, 64, 16, 8, 3) == 0 and ops.dp_owned_elems(0, 64, 16, 8, 0) == 64


def test_sharded_step_is_off_without_a_single_node_nccl_group():
    from vit_ae_plus_plus_b200 import dp
    assert not dp.sharded_enabled()                      # no process group here
    t = dp.alloc_flat(128, torch.float32, "cpu", owner=object())
    assert t.device.type == "cpu" and not dp.is_symmetric(t) and float(t.abs().sum()) == 0.0
