"""Static description of the FrozenBiLM masked-LM hot path.

Mirrors the fields the reference reads from ``DebertaV2Config`` and from ``args``
(reference model/__init__.py:14-47, model/deberta.py:1293-1350).  Only the
switches the released DeBERTa-V2-XL checkpoint uses are supported:
relative_attention, position_buckets=256, norm_rel_ebd=layer_norm, share_att_key,
pos_att_type=p2c|c2p, conv_kernel_size=3, conv_act=gelu, position_biased_input=False,
type_vocab_size=0, hidden_act=gelu.  Anything else raises NotImplementedError.
"""
from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class FBConfig:
    hidden_size: int = 1536
    num_hidden_layers: int = 24
    num_attention_heads: int = 24
    intermediate_size: int = 6144
    vocab_size: int = 128100
    max_position_embeddings: int = 512
    position_buckets: int = 256
    layer_norm_eps: float = 1e-7
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    pad_token_id: int = 0
    # FrozenBiLM additions (reference model/deberta.py:1293-1311)
    features_dim: int = 768
    max_feats: int = 10
    ds_factor_attn: int = 8
    ds_factor_ff: int = 8
    adapter_dropout: float = 0.1
    n_ans: int = 0
    freeze_lm: bool = True
    freeze_mlm: bool = True
    ft_ln: bool = True
    freeze_last: bool = True
    # "deberta" (model/deberta.py) or "bert" (model/bert.py: absolute positions + token types, no adapters,
    # no relative attention / conv / enhanced mask decoder, additive key mask)
    arch: str = "deberta"
    type_vocab_size: int = 0

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def r_attn(self) -> int:
        return self.hidden_size // self.ds_factor_attn if self.ds_factor_attn else 0

    @property
    def r_ff(self) -> int:
        return self.hidden_size // self.ds_factor_ff if self.ds_factor_ff else 0

    def to_dict(self):
        return asdict(self)

    def replace(self, **kw) -> "FBConfig":
        d = asdict(self)
        d.update(kw)
        return FBConfig(**d)

    def to_hf(self):
        """A transformers.DebertaV2Config carrying the XL switches."""
        if self.arch == "bert":
            from transformers import BertConfig

            return BertConfig(hidden_size=self.hidden_size, num_hidden_layers=self.num_hidden_layers,
                              num_attention_heads=self.num_attention_heads,
                              intermediate_size=self.intermediate_size, vocab_size=self.vocab_size,
                              max_position_embeddings=self.max_position_embeddings,
                              type_vocab_size=self.type_vocab_size, layer_norm_eps=self.layer_norm_eps,
                              hidden_act="gelu", hidden_dropout_prob=self.hidden_dropout_prob,
                              attention_probs_dropout_prob=self.attention_probs_dropout_prob,
                              pad_token_id=self.pad_token_id)
        from transformers import DebertaV2Config

        return DebertaV2Config(
            hidden_size=self.hidden_size, num_hidden_layers=self.num_hidden_layers,
            num_attention_heads=self.num_attention_heads, intermediate_size=self.intermediate_size,
            vocab_size=self.vocab_size, max_position_embeddings=self.max_position_embeddings,
            relative_attention=True, position_buckets=self.position_buckets,
            norm_rel_ebd="layer_norm", share_att_key=True, pos_att_type="p2c|c2p",
            conv_kernel_size=3, conv_act="gelu", position_biased_input=False,
            max_relative_positions=-1, type_vocab_size=0, layer_norm_eps=self.layer_norm_eps,
            hidden_act="gelu", hidden_dropout_prob=self.hidden_dropout_prob,
            attention_probs_dropout_prob=self.attention_probs_dropout_prob,
            pad_token_id=self.pad_token_id)

    @staticmethod
    def from_hf(hf, **extra) -> "FBConfig":
        def need(name, want):
            got = getattr(hf, name, None)
            if isinstance(want, (list, tuple, set)):
                ok = got in want
            else:
                ok = got == want
            if not ok:
                raise NotImplementedError(
                    f"frozenbilm_b200 supports only the DeBERTa-V2-XL switch set; config.{name}={got!r}")

        need("relative_attention", True)
        need("share_att_key", True)
        need("position_biased_input", False)
        need("type_vocab_size", 0)
        need("hidden_act", "gelu")
        pat = getattr(hf, "pos_att_type", None)
        if isinstance(pat, str):
            pat = [x.strip() for x in pat.lower().split("|")]
        if sorted(pat or []) != ["c2p", "p2c"]:
            raise NotImplementedError(f"pos_att_type={pat!r}")
        if "layer_norm" not in str(getattr(hf, "norm_rel_ebd", "none")).lower():
            raise NotImplementedError("norm_rel_ebd must be layer_norm")
        if getattr(hf, "conv_kernel_size", 0) != 3 or getattr(hf, "conv_act", "tanh") != "gelu":
            raise NotImplementedError("conv_kernel_size=3 / conv_act=gelu required")
        if getattr(hf, "max_relative_positions", -1) >= 1:
            raise NotImplementedError("max_relative_positions must be -1")
        return FBConfig(
            hidden_size=hf.hidden_size, num_hidden_layers=hf.num_hidden_layers,
            num_attention_heads=hf.num_attention_heads, intermediate_size=hf.intermediate_size,
            vocab_size=hf.vocab_size, max_position_embeddings=hf.max_position_embeddings,
            position_buckets=hf.position_buckets, layer_norm_eps=hf.layer_norm_eps,
            hidden_dropout_prob=hf.hidden_dropout_prob,
            attention_probs_dropout_prob=hf.attention_probs_dropout_prob,
            pad_token_id=getattr(hf, "pad_token_id", 0) or 0, **extra)


def bert_from_hf(hf, **extra) -> FBConfig:
    """FBConfig of a transformers.BertConfig (reference model/bert.py path)."""
    if getattr(hf, "hidden_act", "gelu") != "gelu":
        raise NotImplementedError("hidden_act must be gelu")
    if getattr(hf, "position_embedding_type", "absolute") != "absolute" or getattr(hf, "is_decoder", False):
        raise NotImplementedError("only the absolute-position bidirectional BERT encoder is supported")
    return FBConfig(hidden_size=hf.hidden_size, num_hidden_layers=hf.num_hidden_layers,
                    num_attention_heads=hf.num_attention_heads, intermediate_size=hf.intermediate_size,
                    vocab_size=hf.vocab_size, max_position_embeddings=hf.max_position_embeddings,
                    position_buckets=0, layer_norm_eps=hf.layer_norm_eps,
                    hidden_dropout_prob=hf.hidden_dropout_prob,
                    attention_probs_dropout_prob=hf.attention_probs_dropout_prob,
                    pad_token_id=getattr(hf, "pad_token_id", 0) or 0, arch="bert",
                    type_vocab_size=hf.type_vocab_size, ds_factor_attn=0, ds_factor_ff=0, adapter_dropout=0.0,
                    **extra)


# BASELINE.json configs[0]: the CPU-runnable parity case (DeBERTa-base dims + XL switches)
CONFIG_BASE = FBConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                       intermediate_size=3072, vocab_size=128100, max_feats=2,
                       hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                       adapter_dropout=0.0)
# BASELINE.json configs[1..3]: DeBERTa-V2-XL, WebVid MLM shape
CONFIG_XL = FBConfig()
# small config used by the fast parity tests (same code paths: 2 layers -> layer 0 + conv,
# last layer + 2 EMD passes; d=64 heads)
CONFIG_TINY = FBConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                       intermediate_size=1024, vocab_size=1000, max_feats=10,
                       hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                       adapter_dropout=0.0)
# BERT variant (reference model/bert.py; BASELINE.json configs[4] uses BERT-large dims)
CONFIG_BERT_TINY = FBConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=1024,
                            vocab_size=1000, max_feats=10, position_buckets=0, layer_norm_eps=1e-12,
                            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, adapter_dropout=0.0,
                            ds_factor_attn=0, ds_factor_ff=0, arch="bert", type_vocab_size=2)
CONFIG_BERT_LARGE = CONFIG_BERT_TINY.replace(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                                             intermediate_size=4096, vocab_size=30522, hidden_dropout_prob=0.1,
                                             attention_probs_dropout_prob=0.1)
