"""Golden fixtures for the two host-side pieces of the path that round 1 left unpinned (container only: needs
/root/reference).

    python tests/golden/make_aux_golden.py

* answer_embeddings_golden.npz -- DebertaV2ForMaskedLM.set_answer_embeddings of the UNMODIFIED reference
  (model/deberta.py:1358-1380) on the synthetic tiny model: same-size case and the re-initialising (resize) branch.
* mlm_stats_golden.json -- selection / 80-10-10 / special-pad statistics of the UNMODIFIED reference mask_tokens
  (util/misc.py:14-56) over 1.3 M tokens with a stub tokenizer (the reference draws with torch's global generator,
  so parity with the counter-RNG kernel is distributional: tests compare within binomial error).
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from frozenbilm_b200.config import CONFIG_TINY  # noqa: E402
from frozenbilm_b200.synth import synth_params  # noqa: E402
from oracle.ref_loader import REF_ROOT, build_reference_deberta  # noqa: E402

V_MLM, MASK_ID, PAD_ID, CLS_ID, SEP_ID = 128100, 128000, 0, 1, 2


def answer_tokens(n_ans, width, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    a2tok = torch.randint(1, vocab, (n_ans, width), generator=g)
    lens = torch.randint(1, width + 1, (n_ans,), generator=g)
    for i in range(n_ans):
        a2tok[i, lens[i]:] = 0           # right-padded with the pad id, like videoqa's a2tok
    a2tok[-1] = 0                         # an all-pad answer: sum_tokens clamps to 1
    return a2tok


def make_answer_golden():
    fx = {}
    for tag, n0, n1 in (("same", 3, 3), ("resize", 2, 7)):
        cfg = CONFIG_TINY.replace(n_ans=n0)
        model = build_reference_deberta(cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim,
                                        ds_factor_attn=cfg.ds_factor_attn, ds_factor_ff=cfg.ds_factor_ff, dropout=0.0,
                                        n_ans=n0)
        model.load_state_dict(synth_params(cfg, 1234), strict=False)
        model.eval()
        a2tok = answer_tokens(n1, 5, cfg.vocab_size, 99)
        bias_before = model.answer_bias.detach().clone()
        model.set_answer_embeddings(a2tok, freeze_last=True)
        fx[tag + "/a2tok"] = a2tok.numpy()
        fx[tag + "/weight"] = model.answer_embeddings.weight.detach().numpy().copy()
        fx[tag + "/bias_head"] = model.answer_bias.detach().numpy()[:min(n0, n1)].copy()
        fx[tag + "/bias_before"] = bias_before.numpy()
        fx[tag + "/n_ans"] = np.int64(model.n_ans)
        fx[tag + "/requires_grad"] = np.array([model.answer_embeddings.weight.requires_grad,
                                               model.answer_bias.requires_grad])
    np.savez_compressed(os.path.join(HERE, "answer_embeddings_golden.npz"), **fx)
    print("answer_embeddings_golden.npz", {k: np.asarray(v).shape for k, v in fx.items()})


class StubTokenizer:
    """The four things mask_tokens touches (util/misc.py:19-47)."""
    mask_token = "[MASK]"
    _pad_token = "[PAD]"
    pad_token_id = PAD_ID

    def get_special_tokens_mask(self, val, already_has_special_tokens=True):
        return [1 if t in (CLS_ID, SEP_ID) else 0 for t in val]

    def convert_tokens_to_ids(self, tok):
        return MASK_ID

    def __len__(self):
        return V_MLM


def mlm_inputs(seed=5, rows=3840, width=342):
    rng = np.random.default_rng(seed)
    ids = rng.integers(5, V_MLM, size=(rows, width)).astype(np.int64)
    ids[ids == MASK_ID] = 5
    ids[:, 0] = CLS_ID
    ids[:, -1] = SEP_ID
    lens = rng.integers(width // 2, width + 1, size=rows)
    for r in range(rows):
        ids[r, lens[r]:] = PAD_ID
        ids[r, lens[r] - 1] = SEP_ID
    return ids


def mlm_stats(ids, out, labels):
    """Counts that characterise util/misc.py:14-56 (all integers)."""
    special = (ids == CLS_ID) | (ids == SEP_ID)
    pad = ids == PAD_ID
    sel = labels != -100
    elig = ~special & ~pad
    is_mask = sel & (out == MASK_ID)
    changed = sel & (out != ids) & (out != MASK_ID)
    kept = sel & (out == ids)
    return dict(tokens=int(ids.size), eligible=int(elig.sum()), selected=int(sel.sum()),
                selected_special_or_pad=int((sel & ~elig).sum()), to_mask=int(is_mask.sum()),
                to_random=int(changed.sum()), kept=int(kept.sum()),
                labels_equal_original=bool(np.array_equal(labels[sel], ids[sel])),
                unselected_untouched=bool(np.array_equal(out[~sel], ids[~sel])),
                random_id_mean=float(out[changed].mean()), random_id_min=int(out[changed].min()),
                random_id_max=int(out[changed].max()))


def load_reference_misc():
    spec = importlib.util.spec_from_file_location("_ref_util_misc", os.path.join(REF_ROOT, "util", "misc.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_mlm_stats(seed=0):
    misc = load_reference_misc()
    ids = mlm_inputs()
    torch.manual_seed(seed)
    out, labels = misc.mask_tokens(torch.from_numpy(ids.copy()), StubTokenizer(), 0.15)
    return mlm_stats(ids, out.numpy(), labels.numpy())


if __name__ == "__main__":
    make_answer_golden()
    st = reference_mlm_stats()
    json.dump(st, open(os.path.join(HERE, "mlm_stats_golden.json"), "w"), indent=1)
    print("mlm_stats_golden.json", st)
