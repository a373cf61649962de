"""world_size-2 gloo tests (CPU) of the data-parallel host logic: the single flat-bucket all-reduce + 1/world
scale, the rank-0 broadcast of the trainable bucket, and the per-rank data shards."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frozenbilm_b200.optim import allreduce_sum_and_scale, broadcast_flat, shard_seed
    from frozenbilm_b200.config import CONFIG_TINY
    from frozenbilm_b200.synth import synth_batch

    g = torch.full((1003,), float(rank + 1))
    scale = allreduce_sum_and_scale(g)
    mean = g * scale
    p = torch.arange(17, dtype=torch.float32) + 100 * rank
    broadcast_flat(p, 0)
    b = synth_batch(CONFIG_TINY, 2, 16, shard_seed(1234, rank, 0))
    q.put((rank, float(mean[0]), float(mean[-1]), scale, p.tolist(), int(b["input_ids"].sum())))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_broadcast_and_shards():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, m0, m1, scale, plist, ids_sum in res:
        assert scale == 0.5
        assert m0 == m1 == 1.5           # mean of rank+1 over 2 ranks
        assert plist == list(range(17))   # every rank holds rank 0's parameters after the broadcast
    assert res[0][5] != res[1][5]         # different data shard per rank


def test_single_process_is_a_noop():
    from frozenbilm_b200.optim import allreduce_sum_and_scale
    g = torch.ones(5)
    assert allreduce_sum_and_scale(g) == 1.0 and g.sum() == 5
