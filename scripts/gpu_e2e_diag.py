"""End-to-end diagnostics on the GPU: drop-in model vs golden vectors from the unmodified reference."""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests.golden_util import CASES, GRAD_STRIDE, load_case
from frozenbilm_b200.model import DebertaV2ForMaskedLM
from frozenbilm_b200.synth import is_trainable

dev = torch.device("cuda")
only = sys.argv[1:]
for name in CASES:
    if only and name not in only:
        continue
    cfg, P, batch, fx, mode = load_case(name)
    model = DebertaV2ForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim,
                                 ds_factor_attn=cfg.ds_factor_attn, ds_factor_ff=cfg.ds_factor_ff,
                                 dropout=cfg.adapter_dropout, n_ans=cfg.n_ans)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected, unexpected
    model.to(dev)
    kw = {k: v.to(dev) for k, v in batch.items() if k != "labels"}
    cols = fx["logits_cols"]
    # ---- eval logits
    model.eval()
    with torch.no_grad():
        out = model(**kw)
    lg = out["logits"].float().cpu().numpy()
    err = np.abs(lg[..., cols] - fx["logits_slice"]).max()
    print(f"[{name}] eval logits max-abs-err {err:.4e} (ref abs-mean {np.abs(fx['logits_slice']).mean():.3f})", flush=True)
    if mode != "mlm":
        continue
    # ---- train-mode forward/backward (all dropout p = 0 in these configs)
    model.train()
    out = model(labels=batch["labels"].to(dev), **kw)
    loss = out["loss"]
    print(f"[{name}] loss {loss.item():.6f} ref {float(fx['loss']):.6f}", flush=True)
    loss.backward()
    worst = []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        g = p.grad.detach().float().cpu().reshape(-1)
        ref_norm = float(fx["gnorm/" + n])
        ref_slice = fx["gslice/" + n]
        rel = abs(g.double().norm().item() - ref_norm) / max(ref_norm, 1e-12)
        sl_err = np.abs(g[::GRAD_STRIDE].numpy() - ref_slice).max() / max(np.abs(ref_slice).max(), 1e-12)
        worst.append((max(rel, sl_err), rel, sl_err, n, ref_norm))
    worst.sort(reverse=True)
    for w in worst[:12]:
        print(f"   grad {w[3]}: norm-rel-err {w[1]:.3e} slice-rel-err {w[2]:.3e} (ref norm {w[4]:.3e})")
    print(f"[{name}] median grad err {np.median([w[0] for w in worst]):.3e}, max {worst[0][0]:.3e}", flush=True)
    lg2 = out["logits"].float().cpu().numpy()
    print(f"[{name}] lazy train logits err {np.abs(lg2[..., cols] - fx['logits_slice']).max():.4e}")
