import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import load_case
from tests.test_model_gpu import build, to_dev, DEV
from frozenbilm_b200.synth import is_trainable
from oracle import frozenbilm_oracle as orc
for name in ["tiny_b3_f10_t118", "tiny_b2_f10_t342"]:
    cfg, P, batch, fx, mode = load_case(name)
    names_t = [n for n in P if is_trainable(n, cfg)]
    kw = {k: batch[k] for k in ("video", "video_mask", "input_ids", "attention_mask")}
    ref, rg = orc.forward_backward(P, cfg, names_t, labels=batch["labels"], bf16_gemm_inputs=True, **kw)
    ref32, rg32 = orc.forward_backward(P, cfg, names_t, labels=batch["labels"], **kw)
    model = build(cfg, P).train()
    out = model(labels=batch["labels"].to(DEV), **to_dev(batch)); out["loss"].backward()
    print(name, "loss", out["loss"].item(), ref["loss"].item(), "logit err", (out["logits"].float().cpu() - ref["logits"].detach()).abs().max().item(), "scale", ref["logits"].abs().mean().item())
    g = dict(model.named_parameters())
    res = []
    for n in names_t:
        a, b, c = g[n].grad.float().cpu().reshape(-1).double(), rg[n].reshape(-1).double(), rg32[n].reshape(-1).double()
        cos = (a @ b / (a.norm() * b.norm())).item(); cos32 = (c @ b / (c.norm() * b.norm())).item()
        res.append((cos, n, abs(a.norm().item() / b.norm().item() - 1), cos32))
    res.sort()
    for r in res[:10]:
        print(f"  cos(mine, emu)={r[0]:.5f} cos(fp32, emu)={r[3]:.5f} normrel={r[2]:.2e} {r[1]}")
