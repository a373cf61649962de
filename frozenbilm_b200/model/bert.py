"""BertForMaskedLM drop-in (reference model/bert.py:704-872): same constructor arguments, parameter names,
``forward(video=, video_mask=, input_ids=, ...)`` signature and outputs; arithmetic on the sm_100a kernels through
frozenbilm_b200.engine.BertEngine.  The torch sub-modules are parameter holders only."""
from __future__ import annotations

import os

import torch
from torch import nn

from ..config import bert_from_hf
from ..engine import BertEngine
from .deberta import FrozenBiLMBase, _Holder


class BertForMaskedLM(FrozenBiLMBase):
    engine_cls = BertEngine
    word_embeddings_path = ("bert", "embeddings", "word_embeddings")
    head_bias_name = "cls.predictions.bias"

    def __init__(self, config, max_feats=10, features_dim=768, freeze_lm=True, freeze_mlm=True, ft_ln=True, n_ans=0,
                 freeze_last=True):
        super().__init__()
        self._require_frozen(freeze_lm, freeze_mlm, freeze_last)
        self.config = config
        self.fb = bert_from_hf(config, features_dim=features_dim, max_feats=max_feats, n_ans=n_ans,
                               freeze_lm=freeze_lm, freeze_mlm=freeze_mlm, ft_ln=ft_ln, freeze_last=freeze_last)
        fb = self.fb
        H, I, V = fb.hidden_size, fb.intermediate_size, fb.vocab_size
        eps = fb.layer_norm_eps
        bert = _Holder()
        emb = _Holder()
        emb.word_embeddings = nn.Embedding(V, H, padding_idx=fb.pad_token_id)
        emb.position_embeddings = nn.Embedding(fb.max_position_embeddings, H)
        emb.token_type_embeddings = nn.Embedding(fb.type_vocab_size, H)
        emb.LayerNorm = nn.LayerNorm(H, eps=eps)
        emb.register_buffer("position_ids", torch.arange(fb.max_position_embeddings).expand((1, -1)))
        if features_dim:
            emb.linear_video = nn.Linear(features_dim, H)
        bert.embeddings = emb
        enc = _Holder()
        layers = []
        for _ in range(fb.num_hidden_layers):
            lay = _Holder()
            att = _Holder()
            sa = _Holder()
            sa.query, sa.key, sa.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)
            att.self = sa
            ao = _Holder()
            ao.dense = nn.Linear(H, H)
            ao.LayerNorm = nn.LayerNorm(H, eps=eps)
            att.output = ao
            lay.attention = att
            inter = _Holder()
            inter.dense = nn.Linear(H, I)
            lay.intermediate = inter
            out = _Holder()
            out.dense = nn.Linear(I, H)
            out.LayerNorm = nn.LayerNorm(H, eps=eps)
            lay.output = out
            layers.append(lay)
        enc.layer = nn.ModuleList(layers)
        bert.encoder = enc
        self.bert = bert
        cls = _Holder()
        pred = _Holder()
        tr = _Holder()
        tr.dense = nn.Linear(H, H)
        tr.LayerNorm = nn.LayerNorm(H, eps=eps)
        pred.transform = tr
        pred.decoder = nn.Linear(H, V, bias=False)
        pred.bias = nn.Parameter(torch.zeros(V))
        pred.decoder.bias = pred.bias
        cls.predictions = pred
        self.cls = cls
        self.features_dim = features_dim
        self.max_feats = max_feats
        self.n_ans = n_ans
        std = getattr(config, "initializer_range", 0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(0.0, std)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.data.normal_(0.0, std)
                if m.padding_idx is not None:
                    m.weight.data[m.padding_idx].zero_()
        # freezing rules (bert.py:547-553 and :748-750 -- the whole head, LayerNorm included)
        if freeze_lm:
            for n, p in self.bert.named_parameters():
                if ("linear_video" not in n) and ("adapter" not in n):
                    if ft_ln and "LayerNorm" in n:
                        continue
                    p.requires_grad_(False)
        if freeze_mlm:
            for p in self.cls.parameters():
                p.requires_grad_(False)
        if n_ans:
            self.answer_embeddings = nn.Embedding(n_ans, H)
            self.answer_bias = nn.Parameter(torch.zeros(n_ans))
            if freeze_last:
                self.answer_embeddings.requires_grad_(False)
                self.answer_bias.requires_grad_(False)
        self._post_init_state()

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, local_files_only=True, **kw):
        from transformers import BertConfig

        path = pretrained_model_name_or_path
        config = BertConfig.from_pretrained(path, local_files_only=local_files_only)
        model = cls(config=config, **kw)
        st, pt = os.path.join(path, "model.safetensors"), os.path.join(path, "pytorch_model.bin")
        if os.path.exists(st):
            from safetensors.torch import load_file

            sd = load_file(st)
        elif os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu")
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
        sd = {k.replace("LayerNorm.gamma", "LayerNorm.weight").replace("LayerNorm.beta", "LayerNorm.bias"): v
              for k, v in sd.items()}
        if "cls.predictions.decoder.weight" not in sd and "bert.embeddings.word_embeddings.weight" in sd:
            sd["cls.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]   # HF weight tying
        model.load_state_dict(sd, strict=False)
        return model

    def forward(self, video=None, video_mask=None, input_ids=None, attention_mask=None, token_type_ids=None,
                position_ids=None, head_mask=None, inputs_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, labels=None, output_attentions=None, output_hidden_states=None,
                return_dict=None, mlm=False):
        """Argument order of the reference (bert.py:792-809); callers use keywords."""
        return super().forward(input_ids=input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids,
                               position_ids=position_ids, inputs_embeds=inputs_embeds, labels=labels, video=video,
                               video_mask=video_mask, mlm=mlm)
