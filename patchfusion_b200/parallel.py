"""Multi-GPU execution of the tile path (one process per GPU, torch.distributed over NCCL/NVLink).

The reference only replicates whole images over ranks (`tools/test.py:219-229`); tiles of one image are never
sharded.  Tiles are independent given the per-image coarse outputs and the stitch is a weighted sum, so two
decompositions are offered:

* images-per-rank (bench.py `value`, weak scaling): every rank runs whole images; no data-path collective.
* tiles-per-rank (`PatchFusion.forward(..., shard=(rank, world))`): coarse branch + G2L are replicated
  (1.3 TF, a third of one tile - cheaper than broadcasting 117 MB of taps), tile i of the flattened pass list goes to
  rank i % world, each rank writes its fused predictions into a block [ceil(n / world), ph, pw] and ONE all-gather of
  those blocks (5.7 MB per rank for 4K P49) lets every rank run the deterministic stitch (pf_stitch_gather) over the
  global tile list: the canvas is bit-identical to the single-device one for any world size.
"""
import torch


def shard_indices(n_items, rank, world):
    """Round-robin ownership: item i belongs to rank i % world (balanced to within one item)."""
    assert 0 <= rank < world
    return list(range(rank, n_items, world))


def shard_counts(n_items, world):
    return [len(range(r, n_items, world)) for r in range(world)]


def slot_table(n_items, world):
    """Row of global item i inside the all-gathered blocks [world * ceil(n / world), ...]: rank-major, then the
    item's position in its rank's block."""
    per = max(-(-n_items // world), 1)
    return [(i % world) * per + i // world for i in range(n_items)]


def gather_blocks(block, world, group=None):
    """all_gather the per-rank prediction blocks -> [world * rows, ...] on every rank (the one collective)."""
    import torch.distributed as dist
    assert dist.get_world_size(group) == world
    block = block.contiguous()
    full = torch.empty((world * block.shape[0],) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(full, block, group=group)
    else:                                   # gloo (CPU tests): list form
        dist.all_gather(list(full.view((world,) + tuple(block.shape)).unbind(0)), block, group=group)
    return full


def stitch_reference(full, origins, slots, mask, shape):
    """torch restatement of pf_stitch_gather (CPU gloo test): fixed tile-list order, one accumulation per tile."""
    num, den = torch.zeros(shape), torch.zeros(shape)
    th, tw = mask.shape
    for (y, x), s in zip(origins, slots):
        num[y:y + th, x:x + tw] += mask * full[s]
        den[y:y + th, x:x + tw] += mask
    return num, den
