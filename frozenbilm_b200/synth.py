"""Deterministic synthetic weights and inputs (there are no checkpoints or datasets offline).

The SAME generator feeds the reference (golden generation, container only), the
oracle restatement, the CUDA path and bench.py, so every arm sees identical
weights/inputs for a given (config, seed).  Parameter names/shapes are the
reference's ``named_parameters()``.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .config import FBConfig


def param_spec(cfg: FBConfig) -> "OrderedDict[str, tuple]":
    """name -> shape for every tensor the hot path reads (reference model/deberta.py)."""
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    s = OrderedDict()
    if cfg.arch == "bert":
        e = "bert.embeddings."
        s[e + "word_embeddings.weight"] = (V, H)
        s[e + "position_embeddings.weight"] = (cfg.max_position_embeddings, H)
        s[e + "token_type_embeddings.weight"] = (cfg.type_vocab_size, H)
        s[e + "LayerNorm.weight"] = (H,)
        s[e + "LayerNorm.bias"] = (H,)
        if cfg.features_dim:
            s[e + "linear_video.weight"] = (H, cfg.features_dim)
            s[e + "linear_video.bias"] = (H,)
        for i in range(cfg.num_hidden_layers):
            p = f"bert.encoder.layer.{i}."
            for nm in ("query", "key", "value"):
                s[p + f"attention.self.{nm}.weight"] = (H, H)
                s[p + f"attention.self.{nm}.bias"] = (H,)
            s[p + "attention.output.dense.weight"] = (H, H)
            s[p + "attention.output.dense.bias"] = (H,)
            s[p + "attention.output.LayerNorm.weight"] = (H,)
            s[p + "attention.output.LayerNorm.bias"] = (H,)
            s[p + "intermediate.dense.weight"] = (I, H)
            s[p + "intermediate.dense.bias"] = (I,)
            s[p + "output.dense.weight"] = (H, I)
            s[p + "output.dense.bias"] = (H,)
            s[p + "output.LayerNorm.weight"] = (H,)
            s[p + "output.LayerNorm.bias"] = (H,)
        h = "cls.predictions."
        s[h + "bias"] = (V,)
        s[h + "transform.dense.weight"] = (H, H)
        s[h + "transform.dense.bias"] = (H,)
        s[h + "transform.LayerNorm.weight"] = (H,)
        s[h + "transform.LayerNorm.bias"] = (H,)
        s[h + "decoder.weight"] = (V, H)
        if cfg.n_ans:
            s["answer_embeddings.weight"] = (cfg.n_ans, H)
            s["answer_bias"] = (cfg.n_ans,)
        return s
    e = "deberta.embeddings."
    s[e + "word_embeddings.weight"] = (V, H)
    s[e + "position_embeddings.weight"] = (cfg.max_position_embeddings, H)
    s[e + "LayerNorm.weight"] = (H,)
    s[e + "LayerNorm.bias"] = (H,)
    if cfg.features_dim:
        s[e + "linear_video.weight"] = (H, cfg.features_dim)
        s[e + "linear_video.bias"] = (H,)
    for i in range(cfg.num_hidden_layers):
        p = f"deberta.encoder.layer.{i}."
        for nm in ("query_proj", "key_proj", "value_proj"):
            s[p + f"attention.self.{nm}.weight"] = (H, H)
            s[p + f"attention.self.{nm}.bias"] = (H,)
        s[p + "attention.output.dense.weight"] = (H, H)
        s[p + "attention.output.dense.bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,)
        s[p + "attention.output.LayerNorm.bias"] = (H,)
        if cfg.ds_factor_attn:
            r = cfg.r_attn
            s[p + "attention.output.adapter.down.weight"] = (r, H)
            s[p + "attention.output.adapter.down.bias"] = (r,)
            s[p + "attention.output.adapter.up.weight"] = (H, r)
            s[p + "attention.output.adapter.up.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (I, H)
        s[p + "intermediate.dense.bias"] = (I,)
        s[p + "output.dense.weight"] = (H, I)
        s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,)
        s[p + "output.LayerNorm.bias"] = (H,)
        if cfg.ds_factor_ff:
            r = cfg.r_ff
            s[p + "output.adapter.down.weight"] = (r, H)
            s[p + "output.adapter.down.bias"] = (r,)
            s[p + "output.adapter.up.weight"] = (H, r)
            s[p + "output.adapter.up.bias"] = (H,)
    c = "deberta.encoder."
    s[c + "rel_embeddings.weight"] = (2 * cfg.position_buckets, H)
    s[c + "LayerNorm.weight"] = (H,)
    s[c + "LayerNorm.bias"] = (H,)
    s[c + "conv.conv.weight"] = (H, H, 3)
    s[c + "conv.conv.bias"] = (H,)
    s[c + "conv.LayerNorm.weight"] = (H,)
    s[c + "conv.LayerNorm.bias"] = (H,)
    h = "lm_predictions.lm_head."
    s[h + "dense.weight"] = (H, H)
    s[h + "dense.bias"] = (H,)
    s[h + "LayerNorm.weight"] = (H,)
    s[h + "LayerNorm.bias"] = (H,)
    s[h + "bias"] = (V,)
    if cfg.n_ans:
        s["answer_embeddings.weight"] = (cfg.n_ans, H)
        s["answer_bias"] = (cfg.n_ans,)
    return s


def is_trainable(name: str, cfg: FBConfig) -> bool:
    """Reference freezing rules: model/deberta.py:1152-1158 (encoder), 1334-1339 (head),
    1348-1350 (answer embeddings)."""
    if name.startswith("answer_"):
        return not cfg.freeze_last
    if name.startswith("cls."):
        return not cfg.freeze_mlm   # bert.py:748-750 freezes the WHOLE head, its LayerNorm included
    if name.startswith("lm_predictions."):
        if not cfg.freeze_mlm:
            return True
        return cfg.ft_ln and "LayerNorm" in name
    if not cfg.freeze_lm:
        return True
    if "linear_video" in name or "adapter" in name:
        return True
    return cfg.ft_ln and "LayerNorm" in name


def synth_params(cfg: FBConfig, seed: int = 1234, visible: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """fp32 CPU tensors.  LM matrices ~ N(0, 0.02) like the reference's _init_weights; with
    ``visible`` the adapters, LayerNorm affine and biases are re-randomised to O(0.02-0.1)
    so that a missing adapter / LN-grad term cannot hide under the tolerance."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = OrderedDict()
    for name, shape in param_spec(cfg).items():
        if "LayerNorm.weight" in name:
            t = 1.0 + (0.1 if visible else 0.0) * torch.randn(shape, generator=g)
        elif "LayerNorm.bias" in name:
            t = (0.1 if visible else 0.0) * torch.randn(shape, generator=g)
        elif "adapter" in name:
            std = 0.03 if visible else 1e-3
            t = std * torch.randn(shape, generator=g)
            if not visible:
                t = t.clamp(-2e-3, 2e-3)
        elif name.endswith(".bias") or name == "answer_bias":
            t = (0.02 if visible else 0.0) * torch.randn(shape, generator=g)
        elif "rel_embeddings" in name or "position_embeddings" in name:
            t = (0.5 if visible else 0.02) * torch.randn(shape, generator=g)
        elif "word_embeddings" in name or name == "answer_embeddings.weight" or name == "cls.predictions.decoder.weight":
            t = (0.05 if visible else 0.02) * torch.randn(shape, generator=g)
        elif "token_type_embeddings" in name:
            t = (0.5 if visible else 0.02) * torch.randn(shape, generator=g)
        elif name.endswith("query_proj.weight") or name.endswith("key_proj.weight") or \
                name.endswith("self.query.weight") or name.endswith("self.key.weight"):
            # larger q/k weights give non-flat attention so softmax/bias bugs are visible
            t = (0.06 if visible else 0.02) * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.float().contiguous()
    return out


def synth_batch(cfg: FBConfig, batch: int, n_text: int, seed: int = 1234, *, n_frames: int | None = None,
                ragged: bool = True, label_frac: float = 0.15, mask_token_id: int | None = None):
    """WebVid-shaped synthetic batch: CLIP-like fp16-stored features, random token ids,
    15% MLM labels (>=1 per row).  With ``ragged`` the last row is right-padded and has a
    masked frame; row 0 is always full."""
    F = cfg.max_feats if n_frames is None else n_frames
    g = torch.Generator(device="cpu").manual_seed(seed)
    V = cfg.vocab_size
    video = torch.randn(batch, F, cfg.features_dim, generator=g).half().float()
    input_ids = torch.randint(1, V, (batch, n_text), generator=g, dtype=torch.long)
    attention_mask = torch.ones(batch, n_text, dtype=torch.long)
    video_mask = torch.ones(batch, F, dtype=torch.long)
    if ragged and batch > 1:
        npad = max(1, n_text // 5)
        attention_mask[-1, n_text - npad:] = 0
        input_ids[-1, n_text - npad:] = cfg.pad_token_id
        video_mask[-1, F // 2:] = 0
        video[-1, F // 2:] = 0
    labels = torch.full((batch, n_text), -100, dtype=torch.long)
    pick = torch.rand(batch, n_text, generator=g) < label_frac
    pick &= attention_mask.bool()
    for b in range(batch):
        if not pick[b].any():
            pick[b, 0] = True
    labels[pick] = input_ids[pick]
    if mask_token_id is not None:
        input_ids = input_ids.clone()
        input_ids[pick] = mask_token_id
    return dict(video=video, video_mask=video_mask, input_ids=input_ids,
                attention_mask=attention_mask, labels=labels)
