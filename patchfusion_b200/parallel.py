"""Multi-GPU execution of the tile path (one process per GPU, torch.distributed over NCCL/NVLink).

The reference only replicates whole images over ranks (`tools/test.py:219-229`); tiles of one image are never
sharded.  Tiles are independent given the per-image coarse outputs and the stitch is a weighted sum, so two
decompositions are offered:

* images-per-rank (bench.py default, weak scaling): every rank runs whole images; one `all_gather` assembles the
  batch of depth canvases.
* tiles-per-rank (`PatchFusion.forward(..., shard=(rank, world))`): coarse branch + G2L are replicated (1.3 TF, a
  third of one tile - cheaper than broadcasting 117 MB of taps), tile i of the flattened pass list goes to rank
  i % world, each rank scatter-accumulates its tiles into local (num, den) canvases and ONE all-gather of the
  stacked canvases + a fixed-order sum reproduces the single-GPU canvas up to fp32 summation order.
"""
import torch


def shard_indices(n_items, rank, world):
    """Round-robin ownership: item i belongs to rank i % world (balanced to within one item)."""
    assert 0 <= rank < world
    return list(range(rank, n_items, world))


def shard_counts(n_items, world):
    return [len(range(r, n_items, world)) for r in range(world)]


def gather_canvases(num, den, group=None):
    """all_gather the per-rank (num, den) canvases -> [world, 2, H, W] on every rank (one collective)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.stack([num, den]).contiguous()
    stack = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(stack.view(-1), mine.view(-1), group=group)
    else:                                   # gloo (CPU tests): list form
        dist.all_gather(list(stack.unbind(0)), mine, group=group)
    return stack


def reduce_canvases_reference(stack):
    """torch restatement of pf_stitch_reduce (used by the CPU gloo test): fixed rank-order sum."""
    out = stack[0].clone()
    for r in range(1, stack.shape[0]):
        out += stack[r]
    return out[0], out[1]
