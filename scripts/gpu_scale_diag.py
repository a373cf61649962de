import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200.config import CONFIG_TINY, CONFIG_XL
from frozenbilm_b200.synth import synth_batch, synth_params
from tests.test_model_gpu import build, to_dev, DEV
cfg = CONFIG_XL.replace(num_hidden_layers=2, vocab_size=4096, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, adapter_dropout=0.0)
P = synth_params(cfg, 3); b = synth_batch(cfg, 4, 342, 4)
model = build(cfg, P).train(); kw = to_dev(b); lab = b["labels"].to(DEV)
def run(s, touch_logits=False, perm=None):
    model.zero_grad(set_to_none=True)
    k2 = kw if perm is None else {k: v[perm] for k, v in kw.items()}
    l2 = lab if perm is None else lab[perm]
    out = model(labels=l2, **k2); (s * out["loss"]).backward()
    g = {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}
    if touch_logits: out["logits"].sum().item()
    return g
def cmp(a, b, s, tag):
    rows = sorted([(((b[n] - s * a[n]).norm() / (s * a[n]).norm()).item(), n) for n in a], reverse=True)
    tot = (torch.cat([(b[n] - s * a[n]).reshape(-1) for n in a]).norm() / torch.cat([(s * a[n]).reshape(-1) for n in a]).norm()).item()
    print(tag, "total", f"{tot:.3e}", " worst:", [(f"{r[0]:.2e}", r[1].split("deberta.")[-1]) for r in rows[:3]])
perm = torch.tensor([2, 0, 3, 1], device=DEV)
g1 = run(1.0)
cmp(g1, run(4.0), 4.0, "scale4")
cmp(g1, run(1.0, touch_logits=True), 1.0, "touch (this run)")
cmp(g1, run(1.0), 1.0, "run after touch")
cmp(g1, run(4.0, perm=perm), 4.0, "scale4+perm")
cmp(g1, run(1.0, perm=perm), 1.0, "perm")
