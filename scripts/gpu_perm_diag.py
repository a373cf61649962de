import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200.config import CONFIG_XL, CONFIG_TINY
from frozenbilm_b200.synth import synth_batch, synth_params
from tests.test_model_gpu import build, to_dev, DEV
for cfgname, cfg, Lt in [("xl2", CONFIG_XL.replace(num_hidden_layers=2, vocab_size=4096, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, adapter_dropout=0.0), 342),
                         ("tiny", CONFIG_TINY, 342)]:
    P = synth_params(cfg, 3)
    b = synth_batch(cfg, 4, Lt, 4)
    model = build(cfg, P).train()
    kw = to_dev(b); lab = b["labels"].to(DEV)
    def run(perm=None):
        model.zero_grad(set_to_none=True)
        k2 = kw if perm is None else {k: v[perm] for k, v in kw.items()}
        l2 = lab if perm is None else lab[perm]
        out = model(labels=l2, **k2); out["loss"].backward()
        return {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}, out["loss"].item()
    g0, l0 = run(); g1, l1 = run()
    perm = torch.tensor([2, 0, 3, 1], device=DEV)
    g2, l2 = run(perm)
    perm2 = torch.tensor([1, 0, 2, 3], device=DEV)   # keeps the ragged sample last
    g3, l3 = run(perm2)
    print(cfgname, "loss", l0, l1, l2, l3)
    def rel(a, b): return ((a - b).norm() / (a.norm() + 1e-30)).item()
    rows = [(rel(g0[n], g2[n]), rel(g0[n], g1[n]), rel(g0[n], g3[n]), n) for n in g0]
    rows.sort(reverse=True)
    for r in rows[:14]:
        print(f"  perm {r[0]:.3e} rerun {r[1]:.3e} perm2 {r[2]:.3e} {r[3]}")
