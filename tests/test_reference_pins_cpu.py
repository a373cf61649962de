"""CPU pins of the host-side pieces against the UNMODIFIED reference:

* the committed golden recipe regenerates a fixture bit-for-bit (when /root/reference is present);
* set_answer_embeddings (deberta.py:1358-1380), same-size and resize branches, vs golden output of the reference
  (and vs the live reference when present);
* the MLM-corruption oracle (oracle/data_oracle.py: mlm_corrupt, which the CUDA kernel matches bit-exactly) vs
  statistics of the reference's mask_tokens (util/misc.py:14-56) over 1.3 M tokens, within binomial error.
"""
import importlib.util
import json
import math
import os

import numpy as np
import pytest
import torch

from frozenbilm_b200.config import CONFIG_TINY
from frozenbilm_b200.synth import synth_params
from oracle import data_oracle as dor
from oracle.ref_loader import reference_available

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_golden_recipe_regenerates_fixture_bit_for_bit(tmp_path):
    mg = _load("_make_golden", os.path.join(GOLD, "make_golden.py"))
    for name in ("tiny_b2_f2_t32", "tiny_ans2_b2_f10_t22"):
        fx = mg.run_case(name, *mg.CASES[name], out_dir=str(tmp_path))
        ref = np.load(os.path.join(GOLD, name + ".npz"))
        assert set(ref.keys()) == set(fx.keys())
        for k in ref.keys():
            assert np.array_equal(np.asarray(fx[k]), ref[k]), f"{name}: {k} differs from the committed fixture"


def _our_model(n_ans):
    from frozenbilm_b200.model import DebertaV2ForMaskedLM
    cfg = CONFIG_TINY.replace(n_ans=n_ans)
    m = DebertaV2ForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim,
                             ds_factor_attn=cfg.ds_factor_attn, ds_factor_ff=cfg.ds_factor_ff, dropout=0.0, n_ans=n_ans)
    m.load_state_dict(synth_params(cfg, 1234), strict=False)
    return m.eval()


@pytest.mark.parametrize("tag,n0", [("same", 3), ("resize", 2)])
def test_set_answer_embeddings_matches_reference(tag, n0):
    fx = np.load(os.path.join(GOLD, "answer_embeddings_golden.npz"))
    a2tok = torch.from_numpy(fx[tag + "/a2tok"])
    m = _our_model(n0)
    m.set_answer_embeddings(a2tok, freeze_last=True)
    assert m.n_ans == int(fx[tag + "/n_ans"]) == a2tok.shape[0]
    w = m.answer_embeddings.weight.detach().numpy()
    assert w.shape == fx[tag + "/weight"].shape
    assert np.array_equal(w, fx[tag + "/weight"]), "answer embeddings differ from the reference's"
    # the reference assigns the bias to an ATTRIBUTE (deberta.py:1375): answer_bias keeps its values
    nb = min(n0, a2tok.shape[0])
    assert np.array_equal(m.answer_bias.detach().numpy()[:nb], fx[tag + "/bias_head"])
    assert np.array_equal(fx[tag + "/bias_head"], fx[tag + "/bias_before"][:nb])
    assert m.answer_bias.shape[0] == a2tok.shape[0]
    assert [m.answer_embeddings.weight.requires_grad, m.answer_bias.requires_grad] == \
        fx[tag + "/requires_grad"].tolist() == [False, False]
    # an all-pad answer row (sum_tokens clamped to 1) is the zero vector
    assert not w[-1].any()


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_set_answer_embeddings_live_reference():
    aux = _load("_make_aux_golden", os.path.join(GOLD, "make_aux_golden.py"))
    from oracle.ref_loader import build_reference_deberta
    cfg = CONFIG_TINY.replace(n_ans=2)
    ref = build_reference_deberta(cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim,
                                  ds_factor_attn=8, ds_factor_ff=8, dropout=0.0, n_ans=2)
    ref.load_state_dict(synth_params(cfg, 1234), strict=False)
    ref.eval()
    a2tok = aux.answer_tokens(11, 4, cfg.vocab_size, 3)
    ref.set_answer_embeddings(a2tok)
    m = _our_model(2)
    m.set_answer_embeddings(a2tok)
    assert torch.equal(m.answer_embeddings.weight, ref.answer_embeddings.weight) and m.n_ans == ref.n_ans == 11


def test_unfrozen_recipes_are_refused():
    """--ft_lm / --ft_mlm / --ft_last would need weight gradients this build does not produce: fail loudly."""
    from frozenbilm_b200.model import BertForMaskedLM, DebertaV2ForMaskedLM
    from frozenbilm_b200.config import CONFIG_BERT_TINY
    cfg = CONFIG_TINY
    for kw in (dict(freeze_lm=False), dict(freeze_mlm=False), dict(n_ans=2, freeze_last=False)):
        with pytest.raises(NotImplementedError):
            DebertaV2ForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats, **kw)
    with pytest.raises(NotImplementedError):
        BertForMaskedLM(CONFIG_BERT_TINY.to_hf(), max_feats=10, freeze_lm=False)
    m = _our_model(2)
    with pytest.raises(NotImplementedError):
        m.set_answer_embeddings(torch.ones(2, 3, dtype=torch.long), freeze_last=False)


# ---------------------------------------------------------------------------------------------------------
# f3: device-side MLM corruption vs the reference's mask_tokens
# ---------------------------------------------------------------------------------------------------------
def _z(count, n, p):
    return abs(count - n * p) / math.sqrt(max(n * p * (1 - p), 1e-12))


def _check_stats(st, other=None):
    """Exact invariants + binomial bounds (|z| <= 5) for the rates util/misc.py:14-56 promises."""
    assert st["selected_special_or_pad"] == 0 and st["labels_equal_original"] and st["unselected_untouched"]
    assert st["to_mask"] + st["to_random"] + st["kept"] == st["selected"]
    assert _z(st["selected"], st["eligible"], 0.15) <= 5
    assert _z(st["to_mask"], st["selected"], 0.8) <= 5
    # random replacement: 10 % of the selected, minus the 1/V draws that hit the original or the mask id
    assert _z(st["to_random"], st["selected"], 0.1) <= 5 and _z(st["kept"], st["selected"], 0.1) <= 5
    assert 0 <= st["random_id_min"] and st["random_id_max"] < 128100
    n = st["to_random"]
    assert abs(st["random_id_mean"] - 128099 / 2) <= 5 * (128100 / math.sqrt(12)) / math.sqrt(n)
    if other is not None:
        # two-sample comparison of the rates (reference vs oracle), |z| <= 5
        for key, den in (("selected", "eligible"), ("to_mask", "selected"), ("to_random", "selected")):
            p1, p2 = st[key] / st[den], other[key] / other[den]
            pp = (st[key] + other[key]) / (st[den] + other[den])
            z = abs(p1 - p2) / math.sqrt(pp * (1 - pp) * (1 / st[den] + 1 / other[den]))
            assert z <= 5, (key, p1, p2, z)


def test_mlm_corrupt_oracle_matches_reference_statistics():
    aux = _load("_make_aux_golden", os.path.join(GOLD, "make_aux_golden.py"))
    golden = json.load(open(os.path.join(GOLD, "mlm_stats_golden.json")))
    ids = aux.mlm_inputs()
    assert ids.size == golden["tokens"] >= 1_000_000
    special = ((ids == aux.CLS_ID) | (ids == aux.SEP_ID)).astype(np.uint8)
    out, labels = dor.mlm_corrupt(ids, special, aux.PAD_ID, aux.MASK_ID, aux.V_MLM, 0.15, seed=2024)
    ours = aux.mlm_stats(ids, out, labels)
    assert ours["eligible"] == golden["eligible"]
    _check_stats(golden)           # the reference's own run obeys its stated rates
    _check_stats(ours, golden)     # ... and the oracle is statistically indistinguishable from it
    if reference_available():
        live = aux.reference_mlm_stats(seed=0)
        assert live == golden, "the committed statistics are not what the reference produces here"
