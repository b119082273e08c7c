"""world_size-2 gloo test (CPU) of the tile-sharding host logic: ownership plan, slot table, the single all-gather of
the per-rank prediction blocks and the fixed-order stitch reproduce the single-process canvas BIT-EXACTLY."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from patchfusion_b200.parallel import (block_rows, gather_blocks, shard_counts, shard_indices, slot_table,
                                       stitch_reference, tile_plan)

pytestmark = pytest.mark.timeout(300)


def test_shard_plan():
    for n, w in [(49, 8), (16, 8), (225, 8), (9, 2), (3, 4), (128, 8)]:
        owned = [shard_indices(n, r, w) for r in range(w)]
        flat = sorted(i for o in owned for i in o)
        assert flat == list(range(n))
        c = shard_counts(n, w)
        assert max(c) - min(c) <= 1 and sum(c) == n
        # slot table: item i sits at row (i // w) of rank (i % w)'s block
        per = -(-n // w)
        slots = slot_table(n, w)
        assert len(set(slots)) == n
        for r in range(w):
            assert [slots[i] for i in owned[r]] == [r * per + j for j in range(len(owned[r]))]
    assert shard_counts(49, 8) == [7, 6, 6, 6, 6, 6, 6, 6]


def test_owner_plan():
    """tile_plan: owner_cost = 0 is round-robin; with a cost the owner gets that many fewer items, every item has one
    owner and the slot table is a bijection into the rank-major gathered blocks."""
    for n, w in [(49, 8), (10, 4), (3, 4), (0, 2)]:
        assert tile_plan(n, w) == [i % w for i in range(n)]
        assert slot_table(n, w, tile_plan(n, w)) == slot_table(n, w)
    plan = tile_plan(49, 8, owner_cost=2.7)
    assert shard_counts(49, 8, plan) == [4, 7, 7, 7, 6, 6, 6, 6]
    assert shard_counts(49, 2, tile_plan(49, 2, 2.7)) == [23, 26]
    for n, w, c in [(49, 8, 2.7), (49, 2, 2.7), (353, 8, 2.7), (5, 8, 2.7), (16, 4, 100.0)]:
        plan = tile_plan(n, w, c)
        per = block_rows(n, w, plan)
        owned = [shard_indices(n, r, w, plan) for r in range(w)]
        assert sorted(i for o in owned for i in o) == list(range(n))
        slots = slot_table(n, w, plan)
        assert len(set(slots)) == n
        for r in range(w):
            assert [slots[i] for i in owned[r]] == [r * per + j for j in range(len(owned[r]))]
    assert shard_counts(16, 4, tile_plan(16, 4, 100.0))[0] == 0      # an owner that is too busy gets no tiles


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    th, tw, shape = 24, 32, (48, 64)
    origins = [(0, 0), (0, 32), (24, 0), (24, 32), (0, 16), (24, 16), (12, 0), (12, 32), (12, 16)]
    n = len(origins)
    tiles = torch.rand(n, th, tw, generator=g)
    mask = torch.rand(th, tw, generator=g) + 1e-3
    full_n, full_d = stitch_reference(tiles, origins, list(range(n)), mask, shape)
    same = True
    for plan in (None, tile_plan(n, world, owner_cost=2.7)):  # round-robin and the coarse-owner plan
        own = shard_indices(n, rank, world, plan)
        per = block_rows(n, world, plan)
        block = torch.full((per, th, tw), float('nan'))       # padding rows must never be read
        for j, i in enumerate(own):
            block[j] = tiles[i]
        full = gather_blocks(block, world)
        assert full.shape == (world * per, th, tw)
        num, den = stitch_reference(full, origins, slot_table(n, world, plan), mask, shape)
        same = same and torch.equal(num, full_n) and torch.equal(den, full_d)
    if rank == 0:
        out.put(same)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
