"""Load the UNMODIFIED reference model files by path: from /root/reference (container) or from the byte-identical
copies baseline/install_ref.py stages under baseline/_ref (git-ignored; travels to the GPU box).

TEST / BENCH INFRASTRUCTURE ONLY: used by tests/golden/*.py to generate golden vectors, by the CPU tests that pin the
oracle restatement against the real reference, and by bench.py's CPU legs (--impl reference, cpu_baseline) to time the
reference's own implementation.  The product path (frozenbilm_b200/) never imports this module.

The shims below are version-compat only (transformers 5.5 / torch 2.11 vs the
reference's pinned transformers 4.17 / torch 1.8.1); none changes arithmetic.
See SURVEY.md section 8(c).
"""
import dataclasses
import importlib.util
import os
import sys
import types

import numpy as np
import torch

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")   # baseline/install_ref.py


def _pick_root():
    for cand in (os.environ.get("FROZENBILM_REFERENCE"), "/root/reference", _STAGED):
        if cand and os.path.isfile(os.path.join(cand, "model", "deberta.py")):
            return cand
    return "/root/reference"


REF_ROOT = _pick_root()


def reference_available() -> bool:
    """The reference checkout itself (container) -- what the golden generators need."""
    return os.path.isfile(os.path.join("/root/reference", "model", "deberta.py")) or \
        bool(os.environ.get("FROZENBILM_REFERENCE"))


def reference_model_files_available() -> bool:
    """The reference's model files, from the checkout or staged by baseline/install_ref.py (enough to run its path)."""
    return os.path.isfile(os.path.join(REF_ROOT, "model", "deberta.py"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_reference_deberta():
    """Returns the reference's model.deberta module (with compat shims applied)."""
    if "deberta" in _cache:
        return _cache["deberta"]
    if not hasattr(np, "int"):
        np.int = int  # deberta.py:588 uses the removed alias
    pkg = types.ModuleType("model")
    pkg.__path__ = [os.path.join(REF_ROOT, "model")]
    saved = sys.modules.get("model")
    sys.modules["model"] = pkg
    try:
        _load("model.adapter", os.path.join(REF_ROOT, "model", "adapter.py"))
        deb = _load("model.deberta", os.path.join(REF_ROOT, "model", "deberta.py"))
    finally:
        # keep our synthetic package registered under a private alias only
        sys.modules["_ref_model_pkg"] = pkg
        if saved is not None:
            sys.modules["model"] = saved
        else:
            sys.modules.pop("model", None)
    # shim 1: transformers>=5 init_weights needs post_init bookkeeping
    deb.DebertaV2PreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    # shim 3: torch>=1.11 _softmax_backward_data takes a dtype, not a tensor
    orig = deb._softmax_backward_data
    deb._softmax_backward_data = lambda g, o, dim, t: orig(g, o, dim, t.dtype)
    # shim 4: ModelOutput subclasses must be dataclasses in transformers 5
    if not dataclasses.is_dataclass(deb.BaseModelOutput):
        deb.BaseModelOutput = dataclasses.dataclass(deb.BaseModelOutput)
    _cache["deberta"] = deb
    return deb


def build_reference_deberta(hf_config, *, max_feats, features_dim=768, ds_factor_attn=8,
                            ds_factor_ff=8, dropout=0.0, n_ans=0, freeze_lm=True,
                            freeze_mlm=True, ft_ln=True, freeze_last=True):
    deb = load_reference_deberta()
    model = deb.DebertaV2ForMaskedLM(
        config=hf_config, max_feats=max_feats, features_dim=features_dim,
        freeze_lm=freeze_lm, freeze_mlm=freeze_mlm, ds_factor_attn=ds_factor_attn,
        ds_factor_ff=ds_factor_ff, ft_ln=ft_ln, dropout=dropout, n_ans=n_ans,
        freeze_last=freeze_last)
    return model


def load_reference_bert():
    """Returns the reference's model.bert module (shims 5-7 of SURVEY.md section 8c)."""
    if "bert" in _cache:
        return _cache["bert"]
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    for name in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = None
    bert = _load("_ref_model_bert", os.path.join(REF_ROOT, "model", "bert.py"))
    # transformers 4.17.0 semantics of the additive key mask: (1 - m) * -10000 (requirements.txt:10)
    bert.BertModel.get_extended_attention_mask = \
        lambda self, m, shape, device=None: (1.0 - m[:, None, None, :].float()) * -10000.0
    bert.BertModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    bert.BertPreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    bert.BertPreTrainedModel.post_init = lambda self: self.apply(self._init_weights)
    _cache["bert"] = bert
    return bert


def build_reference_bert(hf_config, *, max_feats, features_dim=768, n_ans=0, freeze_lm=True, freeze_mlm=True,
                         ft_ln=True, freeze_last=True):
    bert = load_reference_bert()
    return bert.BertForMaskedLM(config=hf_config, max_feats=max_feats, features_dim=features_dim,
                                freeze_lm=freeze_lm, freeze_mlm=freeze_mlm, ft_ln=ft_ln, n_ans=n_ans,
                                freeze_last=freeze_last)
