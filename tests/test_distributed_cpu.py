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
    # the module's own data-parallel hook (host logic, on CPU tensors): per-layer ranges of the flat bucket reduced as
    # the backward reaches them, the rest in the finalizer; result = mean over ranks, identical on every rank
    from frozenbilm_b200.model import DebertaV2ForMaskedLM
    cfg = CONFIG_TINY
    torch.manual_seed(0)
    m = DebertaV2ForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats)
    m._ensure_flat(torch.device("cpu"))
    layers, rest = m._layer_ranges()
    fl = m._flat
    total = fl["total"]
    gen = torch.Generator().manual_seed(100 + rank)
    fl["grad32"][:total] = torch.randn(total, generator=gen)
    local = fl["grad32"][:total].clone()
    m._dp = dict(layers=layers, rest=rest, done=set(), works=[], stream=None)
    for i in sorted(layers, reverse=True)[:-1]:          # every layer but layer 0 arrives through the overlap hook
        m._dp_layer_ready(i, [])
    m._finalize_backward()
    covered = sum(b_ - a_ for a_, b_ in layers.values()) + sum(b_ - a_ for a_, b_ in rest)
    gens = [torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    want = sum(gens) / world
    ok_overlap = bool(torch.allclose(fl["grad32"][:total], want, atol=1e-6)) and covered == total and fl["synced"]
    # several forwards -> one backward: nothing through the hook, the finalizer reduces the whole bucket
    fl["grad32"][:total] = local
    m._dp = dict(layers=layers, rest=rest, done=set(), works=[], stream=None)
    m._finalize_backward()
    ok_whole = bool(torch.allclose(fl["grad32"][:total], want, atol=1e-6))
    q.put((rank, float(mean[0]), float(mean[-1]), scale, p.tolist(), int(b["input_ids"].sum()), ok_overlap, ok_whole,
           len(layers), float(fl["grad32"][:total].double().sum())))
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
    for rank, m0, m1, scale, plist, ids_sum, ok_overlap, ok_whole, n_layers, gsum in res:
        assert scale == 0.5
        assert m0 == m1 == 1.5           # mean of rank+1 over 2 ranks
        assert plist == list(range(17))   # every rank holds rank 0's parameters after the broadcast
        assert ok_overlap and ok_whole and n_layers > 0
    assert res[0][5] != res[1][5]         # different data shard per rank
    assert res[0][9] == res[1][9]         # bit-identical reduced gradients on both ranks


def test_single_process_is_a_noop():
    from frozenbilm_b200.optim import allreduce_sum_and_scale
    g = torch.ones(5)
    assert allreduce_sum_and_scale(g) == 1.0 and g.sum() == 5
