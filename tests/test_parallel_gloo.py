"""N>1 host logic on CPU: world_size-2 gloo process group exercising the panel sharding, the final latent
all-gather (uneven shards) and the max-over-ranks timing reduction used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsensei_b200.parallel import gather_latents, max_over_ranks, panel_cost, shard_by_cost, shard_range


def test_shard_range_is_a_contiguous_balanced_partition():
    for total in (0, 1, 5, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(32, 8, 3) == (12, 16)          # cfg4: bs=32 over 8 GPUs -> 4 panels (B=8 under CFG) each


def test_shard_by_cost_balances_var_res_buckets():
    shapes = [(1536, 672), (512, 2048), (1024, 1024), (864, 1216), (704, 1472), (512, 512), (1216, 864), (768, 1344)]
    shapes = shapes * 2
    costs = [panel_cost(h, w) for h, w in shapes]
    parts = shard_by_cost(costs, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(shapes)))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) / min(loads) < 1.25


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]
        a, b = shard_range(total, world, rank)
        full = torch.arange(total * 4 * 2 * 3, dtype=torch.float32).reshape(total, 4, 2, 3)
        got = gather_latents(full[a:b].clone(), counts)
        assert torch.equal(got, full), f"rank {rank}: gathered latents out of order"
        assert max_over_ranks(10.0 + rank, "cpu") == 10.0 + world - 1
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather_and_max():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 5), nprocs=2, join=True)      # 5 panels -> shards of 3 and 2


def test_bench_shards_one_global_batch_over_the_ranks():
    """bench.py --gpus N: every rank builds the same seeded global batch (bs * N panels, CFG-concatenated) and keeps the
    contiguous rows parallel.shard_range gives it; the shards must tile the global batch in panel order."""
    import torch
    import bench
    from diffsensei_b200 import parallel
    from oracle.config import TINY
    world, bs = 4, 2
    glob = bench.synthetic_inputs(TINY, bs * world, 8, 8, 2, "cpu", dialogs=True)
    parts = [bench.shard_inputs(glob, *parallel.shard_range(bs * world, world, r)) for r in range(world)]
    assert torch.equal(torch.cat([p[0] for p in parts]), glob[0])                      # latents, panel order
    for k in (1, 2, 3, 4, 5):                                                           # CFG halves stay paired
        neg = torch.cat([p[k][:bs] for p in parts])
        pos = torch.cat([p[k][bs:] for p in parts])
        assert torch.equal(torch.cat([neg, pos]), glob[k])
