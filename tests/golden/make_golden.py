"""Generate golden fixtures by running the UNMODIFIED reference (container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Each fixture holds, for one (config, batch shape, seed) case of frozenbilm_b200/synth.py:
loss, a strided slice of the logits, the hidden states' checksums, and for every trainable
tensor its gradient L2 norm plus a strided slice.  The weights/inputs are regenerated from
the seed at test time (they are too large to commit: config 1 has a 128100x768 embedding).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from frozenbilm_b200.config import CONFIG_BASE, CONFIG_BERT_TINY, CONFIG_TINY  # noqa: E402
from frozenbilm_b200.synth import synth_params, synth_batch, is_trainable  # noqa: E402
from oracle.ref_loader import build_reference_bert, build_reference_deberta  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (config, batch, n_text, n_frames, ragged, mode)
CASES = {
    "tiny_b2_f2_t32": (CONFIG_TINY.replace(max_feats=2), 2, 32, 2, True, "mlm"),
    "tiny_b3_f10_t118": (CONFIG_TINY, 3, 118, 10, True, "mlm"),
    # L = 352 touches the log-bucket zone (|i-j| > 128), buckets -221..221
    "tiny_b2_f10_t342": (CONFIG_TINY, 2, 342, 10, True, "mlm"),
    # downstream answer head (mc.py: n_ans=2, no labels)
    "tiny_ans2_b2_f10_t22": (CONFIG_TINY.replace(n_ans=2), 2, 22, 10, True, "ans"),
    # BASELINE.json configs[0]
    "base_b2_f2_t32": (CONFIG_BASE, 2, 32, 2, True, "mlm"),
    # BERT variant (reference model/bert.py): plain attention, absolute positions, no adapters
    "bert_tiny_b3_f10_t118": (CONFIG_BERT_TINY, 3, 118, 10, True, "mlm"),
    "bert_tiny_ans2_b2_f10_t22": (CONFIG_BERT_TINY.replace(n_ans=2), 2, 22, 10, True, "ans"),
}

LOGIT_COLS = 97   # stride over the vocab axis
GRAD_STRIDE = 53


def logit_slice(logits):
    V = logits.shape[-1]
    cols = np.arange(0, V, max(1, V // LOGIT_COLS))[:LOGIT_COLS]
    return logits[..., cols], cols


def run_case(name, cfg, B, Lt, F, ragged, mode, seed=1234, out_dir=HERE):
    torch.manual_seed(0)
    P = synth_params(cfg, seed)
    batch = synth_batch(cfg, B, Lt, seed + 1, n_frames=F, ragged=ragged)
    if cfg.arch == "bert":
        model = build_reference_bert(cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim,
                                     n_ans=cfg.n_ans)
    else:
        model = build_reference_deberta(
            cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim,
            ds_factor_attn=cfg.ds_factor_attn, ds_factor_ff=cfg.ds_factor_ff, dropout=cfg.adapter_dropout,
            n_ans=cfg.n_ans)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected, unexpected
    assert all(("decoder" in k or "position_ids" in k) for k in missing), missing
    model.train()  # all dropout p = 0 in these configs, so train == eval
    names_t = [n for n, p in model.named_parameters() if p.requires_grad]
    assert sorted(names_t) == sorted(n for n in P if is_trainable(n, cfg)), "freezing rule mismatch"
    kw = dict(video=batch["video"], video_mask=batch["video_mask"], input_ids=batch["input_ids"],
              attention_mask=batch["attention_mask"])
    fx = {}
    # DebertaV2ForMaskedLM.forward (deberta.py:1414-1427) has no output_hidden_states argument: it always asks the
    # encoder for them (the enhanced mask decoder needs them) and returns them; BertForMaskedLM.forward takes the flag.
    hs_kw = dict(output_hidden_states=True) if cfg.arch == "bert" else {}
    if mode == "mlm":
        out = model(labels=batch["labels"], **hs_kw, **kw)
        out.loss.backward()
        fx["loss"] = np.float64(out.loss.item())
        for n, p in model.named_parameters():
            if p.requires_grad:
                g = p.grad.detach().reshape(-1)
                fx["gnorm/" + n] = np.float64(g.double().norm().item())
                fx["gslice/" + n] = g[::GRAD_STRIDE].numpy().copy()
    else:
        with torch.no_grad():
            out = model(**hs_kw, **kw)
    lg = out.logits.detach().numpy()
    sl, cols = logit_slice(lg)
    fx["logits_slice"] = sl.copy()
    fx["logits_cols"] = cols
    fx["logits_sum"] = np.float64(lg.astype(np.float64).sum())
    fx["logits_abs_sum"] = np.float64(np.abs(lg.astype(np.float64)).sum())
    hs = out.hidden_states
    fx["hidden_abs_mean"] = np.array([h.detach().abs().double().mean().item() for h in hs])
    fx["hidden_slice"] = np.stack([h.detach()[:, :, ::37].numpy() for h in hs])
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **fx)
    print(name, "loss", fx.get("loss"), "logits", lg.shape, "abs-mean", np.abs(lg).mean())
    return fx


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(name, *spec)
