"""world_size-2 gloo test (CPU) of the tile-sharding host logic: ownership plan, the single all-gather of the
(num, den) canvases and the fixed-order reduction reproduce the single-process stitch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from patchfusion_b200.parallel import gather_canvases, reduce_canvases_reference, shard_counts, shard_indices

pytestmark = pytest.mark.timeout(300)


def test_shard_plan():
    for n, w in [(49, 8), (16, 8), (225, 8), (9, 2), (3, 4)]:
        owned = [shard_indices(n, r, w) for r in range(w)]
        flat = sorted(i for o in owned for i in o)
        assert flat == list(range(n))
        c = shard_counts(n, w)
        assert max(c) - min(c) <= 1 and sum(c) == n
    assert shard_counts(49, 8) == [7, 6, 6, 6, 6, 6, 6, 6]


def _stitch(tiles, origins, mask, shape, idx):
    num, den = torch.zeros(shape), torch.zeros(shape)
    th, tw = mask.shape
    for i in idx:
        y, x = origins[i]
        num[y:y + th, x:x + tw] += mask * tiles[i]
        den[y:y + th, x:x + tw] += mask
    return num, den


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    th, tw, shape = 24, 32, (48, 64)
    origins = [(0, 0), (0, 32), (24, 0), (24, 32), (0, 16), (24, 16), (12, 0), (12, 32), (12, 16)]
    tiles = torch.rand(len(origins), th, tw, generator=g)
    mask = torch.rand(th, tw, generator=g) + 1e-3
    num, den = _stitch(tiles, origins, mask, shape, shard_indices(len(origins), rank, world))
    stack = gather_canvases(num, den)
    assert stack.shape == (world, 2) + shape
    n, d = reduce_canvases_reference(stack)
    full_n, full_d = _stitch(tiles, origins, mask, shape, range(len(origins)))
    err = ((n / d) - (full_n / full_d)).abs().max().item()
    if rank == 0:
        out.put(err)
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
    assert out.get(timeout=5) < 1e-6
