"""SURVEY 8f row 1 / BASELINE configs[3]: answer scoring at the [MASK] token.

``model.answer_scores`` (head GEMM on the [MASK] rows only + fused softmax/top-k, candidates batched with the visual
prefix projected once) must give what the reference scripts compute from ``forward()["logits"]``:
videoqa.py:157-170 (softmax + topk over a realistic answer vocabulary) and mc.py:150-172 (softmax(-1)[:, 0] of C
separate forwards)."""
import numpy as np
import pytest
import torch

from frozenbilm_b200.config import CONFIG_TINY
from frozenbilm_b200.synth import synth_batch, synth_params
from tests.test_model_gpu import build, to_dev

pytestmark = pytest.mark.gpu
DEV = "cuda"
MASK_ID = 7


def _with_one_mask(batch, seed):
    g = torch.Generator().manual_seed(seed)
    ids = batch["input_ids"].clone()
    ids[ids == MASK_ID] = 8
    for b in range(ids.shape[0]):
        n_live = int(batch["attention_mask"][b].sum())
        ids[b, int(torch.randint(1, n_live, (1,), generator=g))] = MASK_ID
    batch = dict(batch)
    batch["input_ids"] = ids
    return batch


def test_softmax_topk_kernel():
    from frozenbilm_b200 import lib
    g = torch.Generator(device=DEV).manual_seed(0)
    for P, V, k in ((5, 2, 1), (17, 3129, 10), (3, 40000, 10), (4, 9, 9)):
        ld = (V + 7) // 8 * 8
        x = torch.randn(P, ld, device=DEV, generator=g) * 3
        probs, tp, ti = lib.softmax_topk(x, V, k, want_probs=True)
        ref = x[:, :V].double().softmax(-1)
        assert (probs.double() - ref).abs().max().item() <= 1e-6
        rp, ri = torch.topk(ref, k, -1)
        assert torch.equal(ti, ri) and (tp.double() - rp).abs().max().item() <= 1e-6


def test_videoqa_scores_match_forward_logits_and_oracle():
    from oracle import frozenbilm_oracle as orc
    cfg = CONFIG_TINY.replace(n_ans=3129)           # a realistic open-ended answer vocabulary
    P = synth_params(cfg, 41)
    b = _with_one_mask(synth_batch(cfg, 4, 30, 42), 1)
    model = build(cfg, P).eval()
    kw = to_dev(b)
    out = model.answer_scores(mask_token_id=MASK_ID, topk=10, return_probs=True, **kw)
    with torch.no_grad():
        full = model(**kw)["logits"]                                   # [B, F+Lt, n_ans], the reference's route
    F = b["video"].shape[1]
    at_mask = full[:, F:][kw["input_ids"] == MASK_ID]
    assert at_mask.shape == (4, 3129) and out["probs"].shape == (4, 3129)
    assert (out["probs"] - at_mask.softmax(-1)).abs().max().item() <= 1e-5
    assert torch.equal(out["top_i"], torch.topk(out["probs"], 10, -1).indices)
    ref = orc.forward(P, cfg, bf16_gemm_inputs=True, **{k: b[k] for k in ("video", "video_mask", "input_ids",
                                                                          "attention_mask")})
    rl = ref["logits"][:, F:][b["input_ids"] == MASK_ID]
    assert (at_mask.cpu() - rl).abs().max().item() <= 0.02 * rl.abs().mean().item()
    assert (out["probs"].cpu() - rl.softmax(-1)).abs().max().item() <= 2e-3


def test_multiple_choice_candidates_batched_equals_per_candidate_forwards():
    cfg = CONFIG_TINY.replace(n_ans=2)
    P = synth_params(cfg, 43)
    model = build(cfg, P).eval()
    B, C, Lt = 3, 5, 26
    vid = synth_batch(cfg, B, Lt, 44)
    cands = [_with_one_mask(synth_batch(cfg, B, Lt, 50 + c), c) for c in range(C)]
    video, vmask = vid["video"].to(DEV), vid["video_mask"].to(DEV)
    # mc.py:150-172: one forward per candidate, softmax(-1)[:, 0] at the mask token
    per = []
    with torch.no_grad():
        for c in range(C):
            ids, am = cands[c]["input_ids"].to(DEV), cands[c]["attention_mask"].to(DEV)
            lg = model(video=video, video_mask=vmask, input_ids=ids, attention_mask=am)["logits"]
            per.append(lg[:, 10:][ids == MASK_ID].softmax(-1)[:, 0])
    per = torch.stack(per, 1)                                           # [B, C]
    ids = torch.stack([cands[c]["input_ids"] for c in range(C)], 1).reshape(B * C, Lt).to(DEV)
    am = torch.stack([cands[c]["attention_mask"] for c in range(C)], 1).reshape(B * C, Lt).to(DEV)
    out = model.answer_scores(ids, am, video=video, video_mask=vmask, mask_token_id=MASK_ID, candidates=C,
                              return_probs=True)
    got = out["probs"][:, 0].view(B, C)
    assert (got - per).abs().max().item() <= 2e-3
    assert torch.equal(got.max(1).indices, per.max(1).indices)
