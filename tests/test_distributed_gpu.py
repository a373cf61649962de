"""Data-parallel correctness on real GPUs (SURVEY section 4 iii): two ranks over NCCL, different data shards.

* after every step the flat trainable bucket is BIT-identical on both ranks (gradients are averaged inside backward(),
  overlapped per layer with the backward itself; parameters start from rank 0's values);
* the averaged gradient equals the gradient of a single process run on the concatenated batch (fp32 tolerance: the
  weight-gradient reductions run over the tokens in a different order).
Skipped with fewer than two GPUs (run it with `gpurun --gpus 2`; its log is committed under profiles/)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(cfg, B, Lt, seed, n_lab=4):
    """Every row carries exactly n_lab labels, so that the mean over a rank's labels weighs both shards equally."""
    from frozenbilm_b200.synth import synth_batch
    b = synth_batch(cfg, B, Lt, seed, ragged=False)
    g = torch.Generator().manual_seed(seed + 1)
    lab = torch.full_like(b["input_ids"], -100)
    for r in range(B):
        pos = torch.randperm(Lt, generator=g)[:n_lab]
        lab[r, pos] = b["input_ids"][r, pos]
    b["labels"] = lab
    return b


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from frozenbilm_b200.config import CONFIG_TINY
    from frozenbilm_b200.model import DebertaV2ForMaskedLM
    from frozenbilm_b200.optim import FusedClipAdam, broadcast_trainable
    from frozenbilm_b200.synth import synth_params
    cfg = CONFIG_TINY.replace(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, adapter_dropout=0.0)
    P = synth_params(cfg, 7 + rank)                       # ranks start from DIFFERENT weights (the reference's bug) ...
    frozen = synth_params(cfg, 7)
    for k in P:
        if not (("adapter" in k) or ("LayerNorm" in k) or ("linear_video" in k)):
            P[k] = frozen[k]                              # ... only the trainable tensors differ
    model = DebertaV2ForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats, dropout=0.0)
    model.load_state_dict(P, strict=False)
    model.to(dev).train()
    opt = FusedClipAdam(model, lr=1e-3, betas=(0.9, 0.95), max_norm=0.1)
    broadcast_trainable(model)                            # ... and are made equal here
    sums, grad_first = [], None
    for it in range(3):
        b = _batch(cfg, 2, 40, 100 + 10 * it + rank)      # per-rank shard
        out = model(**{k: v.to(dev) for k, v in b.items()})
        out["loss"].backward()
        p32, g32, _ = model.flat_buffers()
        if it == 0:
            grad_first = g32.clone().cpu()
        opt.step()
        torch.cuda.synchronize()
        sums.append(p32.clone().cpu())
        model.zero_grad()
    q.put((rank, [s.numpy().tobytes() for s in sums], grad_first.numpy().tobytes(), model._flat["total"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_stay_bit_identical_and_match_single_process():
    import numpy as np
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (r0, params0, g0, total), (r1, params1, g1, _) = res
    for it in range(3):
        assert params0[it] == params1[it], f"parameters diverged between the ranks after step {it}"
    assert g0 == g1, "averaged gradients differ between the ranks"
    # single process, concatenated batch, rank 0's starting weights
    from frozenbilm_b200.config import CONFIG_TINY
    from frozenbilm_b200.model import DebertaV2ForMaskedLM
    from frozenbilm_b200.synth import synth_params
    cfg = CONFIG_TINY.replace(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, adapter_dropout=0.0)
    model = DebertaV2ForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats, dropout=0.0)
    model.load_state_dict(synth_params(cfg, 7), strict=False)
    model.to("cuda").train()
    bs = [_batch(cfg, 2, 40, 100 + r) for r in range(world)]
    cat = {k: torch.cat([b[k] for b in bs], 0).to("cuda") for k in bs[0]}
    model(**cat)["loss"].backward()
    _, g32, _ = model.flat_buffers()
    ref = g32[:total].cpu().numpy()
    got = np.frombuffer(g0, dtype=np.float32)[:total]
    assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max(), (np.abs(got - ref).max(), np.abs(ref).max())
