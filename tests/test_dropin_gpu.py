"""Proof of the drop-in claim: the reference's OWN training / evaluation loops run unmodified against this build.

``main.train_one_epoch`` (main.py:24-97) and ``mc.train_one_epoch`` / ``mc.evaluate`` (mc.py:25-199) are imported from
the byte-identical copies under baseline/_ref (baseline/install_ref.py; /root/reference in the container) with
``from model import build_model, get_tokenizer`` resolved by ``frozenbilm_b200.install_as_model()``.  Only the things
the loops receive as ARGUMENTS are stand-ins: a stub tokenizer (no tokenizer files exist offline), an in-memory data
loader and an ``args`` namespace; ``datasets`` / ``hostlist`` (imported by the scripts, unused by the loops) are empty
stub modules.  The model is built by the reference-named ``build_model(args)`` from a config directory."""
import importlib
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference" if os.path.isfile("/root/reference/main.py") else os.path.join(ROOT, "baseline", "_ref")
MASK_ID, PAD_ID, CLS_ID, SEP_ID, VOCAB = 3, 0, 1, 2, 1000


class StubTokenizer:
    """Whitespace 'tokenizer' with the call signature and attributes the loops use (main.py:45-55, mc.py:50-72,
    util/misc.py:19-47)."""
    mask_token, _pad_token = "[MASK]", "[PAD]"
    pad_token_id, mask_token_id = PAD_ID, MASK_ID

    def __call__(self, text, add_special_tokens=True, max_length=64, padding="longest", truncation=True,
                 return_tensors="pt"):
        rows = []
        for t in text:
            ids = [CLS_ID] + [MASK_ID if w == "[MASK]" else 4 + (hash(w) % (VOCAB - 4)) for w in t.split()][:max_length - 2]
            rows.append(ids + [SEP_ID])
        L = max(len(r) for r in rows)
        ids = torch.tensor([r + [PAD_ID] * (L - len(r)) for r in rows])
        return {"input_ids": ids, "attention_mask": (ids != PAD_ID).long()}

    def get_special_tokens_mask(self, val, already_has_special_tokens=True):
        return [1 if t in (CLS_ID, SEP_ID) else 0 for t in val]

    def convert_tokens_to_ids(self, tok):
        return MASK_ID

    def __len__(self):
        return VOCAB


@pytest.fixture(scope="module")
def ref_scripts():
    if not os.path.isfile(os.path.join(REF, "main.py")):
        pytest.skip("reference scripts not staged (python baseline/install_ref.py)")
    import frozenbilm_b200
    saved = {k: sys.modules.get(k) for k in ("model", "datasets", "hostlist", "util", "util.dist", "util.misc",
                                             "util.metrics", "args", "main", "mc")}
    frozenbilm_b200.install_as_model()
    ds = types.ModuleType("datasets")
    for name in ("build_videotext_dataset", "videotext_collate_fn", "build_mc_dataset", "mc_collate_fn"):
        setattr(ds, name, None)                       # imported by the scripts' headers, never called by the loops
    sys.modules["datasets"] = ds
    sys.modules["hostlist"] = types.ModuleType("hostlist")   # util/dist.py:7, SLURM only
    for k in ("util", "util.dist", "util.misc", "util.metrics", "args", "main", "mc"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    try:
        yield importlib.import_module("main"), importlib.import_module("mc")
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _args(tmp_path, n_ans):
    from frozenbilm_b200.config import CONFIG_TINY
    cfg_dir = tmp_path / "deberta-v2-tiny"                      # build_model dispatches on the substring "deberta"
    CONFIG_TINY.replace(vocab_size=VOCAB).to_hf().save_pretrained(str(cfg_dir))
    return types.SimpleNamespace(
        model_name=str(cfg_dir), scratch=True, features_dim=768, use_video=True, max_feats=10, freeze_lm=True,
        freeze_mlm=True, ft_ln=True, ds_factor_attn=8, ds_factor_ff=8, dropout=0.1, n_ans=n_ans, freeze_last=True,
        epochs=3, print_freq=100, max_tokens=64, mlm_prob=0.15, lr=3e-3, schedule="linear_with_warmup",
        fraction_warmup_steps=0.1, beta1=0.9, beta2=0.95, weight_decay=0.0)


WORDS = "a man is playing the guitar on stage while people watch and clap in a large hall with lights".split()


def _sentence(g, n):
    return " ".join(WORDS[int(i)] for i in torch.randint(0, len(WORDS), (n,), generator=g))


def test_main_train_one_epoch_runs_unmodified(ref_scripts, tmp_path):
    main, _ = ref_scripts
    from model import build_model                                  # the reference's import line (main.py:17)
    args = _args(tmp_path, 0)
    torch.manual_seed(0)
    model = build_model(args)
    assert type(model).__module__.startswith("frozenbilm_b200.")
    device = torch.device("cuda")
    model.to(device)
    g = torch.Generator().manual_seed(1)
    loader = [{"video": torch.randn(4, 10, 768, generator=g), "video_len": torch.tensor([10, 7, 10, 3]),
               "text": [_sentence(g, int(n)) for n in torch.randint(8, 30, (4,), generator=g)]} for _ in range(6)]
    params = [p for n, p in model.named_parameters() if p.requires_grad]          # main.py:177-188
    opt = torch.optim.Adam(params, lr=args.lr, betas=(args.beta1, args.beta2), weight_decay=args.weight_decay)
    before = torch.cat([p.detach().reshape(-1).clone() for p in params])
    losses = []
    for epoch in range(args.epochs):
        torch.manual_seed(123)                                     # same mask_tokens draws every epoch
        stats = main.train_one_epoch(model=model, tokenizer=StubTokenizer(), data_loader=loader, optimizer=opt,
                                     device=device, epoch=epoch, args=args, max_norm=0.1)
        losses.append(stats["loss"])
    after = torch.cat([p.detach().reshape(-1) for p in params])
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses
    assert (after - before).abs().max().item() > 0
    for n, p in model.named_parameters():
        assert (p.grad is not None) == p.requires_grad, n


def test_mc_train_and_evaluate_run_unmodified(ref_scripts, tmp_path):
    _, mc = ref_scripts
    from model import build_model
    args = _args(tmp_path, 2)
    torch.manual_seed(0)
    model = build_model(args)
    device = torch.device("cuda")
    model.to(device)
    tok = StubTokenizer()
    # mc.py:281-305: answer embeddings of "yes" / "no" from the word embeddings
    model.set_answer_embeddings(torch.tensor([[5, 0], [6, 7]]).to(device), freeze_last=args.freeze_last)
    g = torch.Generator().manual_seed(2)
    C, B = 5, 3

    def batch(k):
        text = [[_sentence(g, 6) + " [MASK] " + _sentence(g, 4) for _ in range(B)] for _ in range(C)]   # text[aid][b]
        return {"video": torch.randn(B, 10, 768, generator=g), "video_len": torch.tensor([10, 5, 8]), "text": text,
                "answer_id": torch.randint(0, C, (B,), generator=g), "qid": [f"q{k}_{i}" for i in range(B)],
                "type": [0] * B}

    class Dataset:
        mc = C

        def __len__(self):
            return 4 * B

    class Loader(list):
        dataset = Dataset()

    loader = Loader(batch(k) for k in range(4))
    params = [p for n, p in model.named_parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=args.lr, betas=(args.beta1, args.beta2))
    losses = []
    for epoch in range(args.epochs):
        stats = mc.train_one_epoch(model=model, tokenizer=tok, data_loader=loader, optimizer=opt, device=device,
                                   epoch=epoch, args=args, max_norm=0.1)     # 5 forwards, ONE backward per step
        losses.append(stats["loss"])
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses
    results, acc = mc.evaluate(model=model, tokenizer=tok, data_loader=loader, device=device,
                               dataset_name="synthetic", args=args)
    assert len(results) == 4 * B and 0.0 <= acc <= 1.0
    # the batched fast path gives the same predictions as mc.evaluate's per-candidate loop
    model.eval()
    b = loader[0]
    from util.misc import get_mask
    vm = get_mask(b["video_len"], 10).to(device)
    enc = [tok(b["text"][a], max_length=args.max_tokens) for a in range(C)]
    Lt = max(e["input_ids"].shape[1] for e in enc)
    pad = lambda t: torch.nn.functional.pad(t, (0, Lt - t.shape[1]))
    ids = torch.stack([pad(e["input_ids"]) for e in enc], 1).reshape(B * C, Lt)
    am = torch.stack([pad(e["attention_mask"]) for e in enc], 1).reshape(B * C, Lt)
    sc = model.answer_scores(ids.to(device), am.to(device), video=b["video"].to(device), video_mask=vm,
                             mask_token_id=MASK_ID, candidates=C, return_probs=True)
    per = []
    with torch.no_grad():
        for a in range(C):
            lg = model(video=b["video"].to(device), video_mask=vm, input_ids=enc[a]["input_ids"].to(device),
                       attention_mask=enc[a]["attention_mask"].to(device))["logits"]
            per.append(lg[:, 10:enc[a]["input_ids"].size(1) + 10][enc[a]["input_ids"] == MASK_ID].softmax(-1)[:, 0])
    per = torch.stack(per, 1)
    assert (sc["probs"][:, 0].view(B, C) - per).abs().max().item() <= 5e-3
