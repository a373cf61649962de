"""Drop-in mirror of the reference's ``model`` package surface (reference model/__init__.py:14-112).

    from frozenbilm_b200.model import build_model, get_tokenizer      # or: frozenbilm_b200.install_as_model()

``build_model(args)`` reads the same ``args`` fields as the reference and returns an ``nn.Module`` with the
reference's parameter names, ``forward`` signature and outputs; the compute runs on the sm_100a kernels of
libfrozenbilm_sm100.so.  Only the BiLM hot path is provided: DeBERTa-V2 (+adapters) and BERT.  The
autoregressive GPT-Neo / GPT-J baselines are out of scope and raise NotImplementedError like an unknown name.
"""
from .bert import BertForMaskedLM
from .deberta import DebertaV2ForMaskedLM


def build_model(args):
    """Reference model/__init__.py:14-91 (same substring dispatch on args.model_name)."""
    name = args.model_name
    if "deberta" in name:
        from transformers import DebertaV2Config

        kw = dict(features_dim=args.features_dim if args.use_video else 0, max_feats=args.max_feats,
                  freeze_lm=args.freeze_lm, freeze_mlm=args.freeze_mlm, ft_ln=args.ft_ln,
                  ds_factor_attn=args.ds_factor_attn, ds_factor_ff=args.ds_factor_ff, dropout=args.dropout,
                  n_ans=args.n_ans, freeze_last=args.freeze_last)
        if args.scratch:
            config = DebertaV2Config.from_pretrained(pretrained_model_name_or_path=name, local_files_only=True)
            return DebertaV2ForMaskedLM(config=config, **kw)
        return DebertaV2ForMaskedLM.from_pretrained(pretrained_model_name_or_path=name, local_files_only=True, **kw)
    if "bert" in name:
        assert (not args.ds_factor_ff) and (not args.ds_factor_attn) and (not args.scratch)
        return BertForMaskedLM.from_pretrained(
            features_dim=args.features_dim if args.use_video else 0, max_feats=args.max_feats,
            freeze_lm=args.freeze_lm, freeze_mlm=args.freeze_mlm, ft_ln=args.ft_ln, n_ans=args.n_ans,
            freeze_last=args.freeze_last, pretrained_model_name_or_path=name, local_files_only=True)
    raise NotImplementedError


def get_tokenizer(args):
    """Reference model/__init__.py:94-112."""
    name = args.model_name
    if "deberta" in name:
        from transformers import DebertaV2Tokenizer

        return DebertaV2Tokenizer.from_pretrained(name, local_files_only=True)
    if "bert" in name:
        from transformers import BertTokenizer

        return BertTokenizer.from_pretrained(name, local_files_only=True)
    raise NotImplementedError
