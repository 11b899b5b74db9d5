"""N>1 path on CPU: world_size-2 gloo processes shard prompts, 'generate', and gather in order."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200sd import sharding as Sh


def test_shard_indices_cover_everything_once():
    for n, world in [(64, 8), (7, 2), (3, 4), (0, 2)]:
        seen = sorted(i for r in range(world) for i in Sh.shard_indices(n, r, world))
        assert seen == list(range(n))
    assert Sh.shard_prompts(list("abcdefgh"), 1, 4) == ["b", "f"]
    assert Sh.chunks([1, 2, 3], 2) == [[1, 2], [3, 3]]
    with pytest.raises(ValueError):
        Sh.shard_indices(4, 2, 2)


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prompts = [f"prompt {i}" for i in range(n_items)]
    mine = Sh.shard_prompts(prompts, rank, world)
    local = [f"image({p})@{rank}" for p in mine]           # stand-in for the per-GPU pipeline
    dist.barrier()                                          # the only collective on the data path: none
    t = Sh.max_over_ranks(10.0 + rank, dist)
    out = Sh.gather_in_order(local, n_items, rank, world, dist)
    if rank == 0:
        q.put((t, out))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_items, world, port = 7, 2, 29533
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 11.0  # max over ranks
    assert out == [f"image(prompt {i})@{i % 2}" for i in range(n_items)]
