"""DebertaV2ForMaskedLM drop-in (reference model/deberta.py:1288-1501).

Same constructor arguments, parameter names / shapes / requires_grad pattern, ``forward`` signature,
``set_answer_embeddings`` and outputs as the reference class; the arithmetic is executed by
frozenbilm_b200.engine.DebertaEngine on the sm_100a kernels.  The torch sub-modules below are
PARAMETER HOLDERS ONLY (so that ``state_dict()`` keys equal the reference's); they are never called.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
from torch import nn

from ..config import FBConfig
from ..engine import DebertaEngine
from .. import lib


class _Holder(nn.Module):
    pass


def _adapter(ds_factor, H):
    m = _Holder()
    m.down = nn.Linear(H, H // ds_factor)
    m.up = nn.Linear(H // ds_factor, H)
    return m


class FBMaskedLMOutput(dict):
    """Supports out["loss"], out["logits"], out.loss, out.logits like transformers' MaskedLMOutput
    (reference deberta.py:1496-1501).  In the MLM-loss path ``logits`` ([B, F+Lt, V], 5.8 GB at the
    WebVid shape) is materialised lazily on first access."""

    def __init__(self, loss=None, logits=None, logits_fn=None):
        super().__init__()
        self._logits_fn = logits_fn
        if loss is not None:
            dict.__setitem__(self, "loss", loss)
        if logits is not None:
            dict.__setitem__(self, "logits", logits)

    def __getitem__(self, k):
        if isinstance(k, int):
            return self.to_tuple()[k]
        if k == "logits" and not dict.__contains__(self, "logits") and self._logits_fn is not None:
            dict.__setitem__(self, "logits", self._logits_fn())
        return dict.__getitem__(self, k)

    def __contains__(self, k):
        return dict.__contains__(self, k) or (k == "logits" and self._logits_fn is not None)

    @property
    def loss(self):
        return self.get("loss")

    @property
    def logits(self):
        return self["logits"] if "logits" in self else None

    def to_tuple(self):
        return tuple(self[k] for k in ("loss", "logits") if k in self)


class _MLMLossFn(torch.autograd.Function):
    """Bridges loss.backward() (main.py:81) to the hand-scheduled selective backward.  The node owns the
    ForwardState of ITS forward, so any number of train-mode forwards may precede one backward."""

    @staticmethod
    def forward(ctx, anchor, module, state, loss):
        ctx.module, ctx.state = module, state
        return loss.clone()

    @staticmethod
    def backward(ctx, dloss):
        ctx.module._run_backward(ctx.state, scale=dloss)      # stays on the device: no host sync in the backward
        ctx.state = None
        return None, None, None, None


class _LogitsFn(torch.autograd.Function):
    """Downstream mode (mc.py / videoqa.py): gradients arrive through the answer-head logits."""

    @staticmethod
    def forward(ctx, anchor, module, state, logits):
        ctx.module, ctx.state = module, state
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        ctx.module._run_backward(ctx.state, dlogits=dlogits)
        ctx.state = None
        return None, None, None, None


class FrozenBiLMBase(nn.Module):
    """Everything the two architectures share: the flat trainable bucket, the engine life-cycle, the forward
    contract (loss / logits, lazy full logits) and the autograd bridge.  Sub-classes build the parameter holders
    with the reference's names and set ``engine_cls``."""

    engine_cls = DebertaEngine
    word_embeddings_path = ("deberta", "embeddings", "word_embeddings")
    head_bias_name = "lm_predictions.lm_head.bias"

    def _post_init_state(self):
        self._engine_obj = None
        self._flat = None
        self._anchor = None
        self._frozen_sig = None
        # data-parallel gradient synchronisation inside backward() (the reference has none): on by default
        # when torch.distributed is initialised with world_size > 1; FB_DP_SYNC=0 / model.dp_sync = False disables it
        self.dp_sync = os.environ.get("FB_DP_SYNC", "1") != "0"
        # FB_DP_OVERLAP=1: all-reduce every layer's slice as soon as its gradients are complete, on a communication
        # stream.  Measured on 2xB200 (profiles/r2_dp_overlap.md): 72.4 ms/step vs 71.9 ms with ONE all-reduce at the end
        # of the backward -- the NCCL kernels take SMs away from the persistent one-CTA-per-SM GEMM / attention kernels,
        # which costs more than the 0.5 ms the single collective leaves exposed.  Off by default.
        self.dp_overlap = os.environ.get("FB_DP_OVERLAP", "0") == "1"
        self._dp = None
        # "bf16" (training path) | "fp32x3" (fp32-class instantiation, the configs[0] parity gate)
        self.precision = os.environ.get("FB_PRECISION", "bf16")

    def set_precision(self, precision: str):
        """Selects the arithmetic of the contractions: "bf16" or "fp32x3" (see DebertaEngine)."""
        if precision not in ("bf16", "fp32x3"):
            raise ValueError(f"unknown precision {precision!r}")
        if precision != self.precision:
            self.precision = precision
            self._engine_obj = None
        return self

    @staticmethod
    def _require_frozen(freeze_lm, freeze_mlm, freeze_last):
        """The selective backward only produces gradients for adapters, LayerNorms and linear_video.  The reference's
        --ft_lm / --ft_mlm / --ft_last (UnFrozenBiLM recipes) would need dW for every LM matrix, the embeddings and
        the answer head: refuse loudly instead of silently training nothing."""
        bad = [n for n, v in (("freeze_lm", freeze_lm), ("freeze_mlm", freeze_mlm), ("freeze_last", freeze_last))
               if not v]
        if bad:
            raise NotImplementedError(
                f"{', '.join(bad)}=False (--ft_lm / --ft_mlm / --ft_last) is not supported: this build implements the "
                "FROZEN-LM path (gradients for adapters, LayerNorms and linear_video only)")

    @property
    def device(self):
        return next(self.parameters()).device

    def _word_embeddings(self):
        m = self
        for k in self.word_embeddings_path:
            m = getattr(m, k)
        return m

    # ------------------------------------------------------------------ flat trainable bucket
    def _param_dict(self):
        return OrderedDict((n, p) for n, p in self.named_parameters())

    def _engine_params(self, named):
        return {n: p.data for n, p in named.items()}

    def _ensure_flat(self, dev, named=None):
        """(Re)builds the flat fp32 bucket of the trainable set: parameters become views of it, so that one NCCL
        all-reduce per contiguous range / one fused Adam kernel covers everything that trains."""
        named = named if named is not None else self._param_dict()
        train = [(n, p) for n, p in named.items() if p.requires_grad]
        sig = tuple((n, p.data_ptr()) for n, p in train)
        if self._flat is None or self._flat["sig"] != sig or self._flat["flat32"].device != dev:
            total = sum(p.numel() for _, p in train)
            total_pad = (total + 7) // 8 * 8
            flat32 = torch.zeros(total_pad, device=dev, dtype=torch.float32)
            slices, off = {}, 0
            for n, p in train:
                k = p.numel()
                flat32[off:off + k].copy_(p.data.reshape(-1))
                p.data = flat32[off:off + k].view(p.shape)
                slices[n] = (off, off + k, tuple(p.shape))
                off += k
            self._flat = dict(flat32=flat32, flat16=torch.empty(total_pad, device=dev, dtype=torch.bfloat16),
                              grad32=torch.zeros(total_pad, device=dev, dtype=torch.float32), slices=slices,
                              sig=tuple((n, p.data_ptr()) for n, p in train), version=-1, total=total)
            self._engine_obj = None
        return train

    def _ensure_engine(self):
        dev = self.device
        if dev.type != "cuda":
            raise lib.FbError("frozenbilm_b200 has no CPU path: call model.to('cuda') before forward()")
        named = self._param_dict()
        train = self._ensure_flat(dev, named)
        # frozen tensors: version counters of the PARAMETERS (p.data has its own, which load_state_dict never bumps)
        fsig = tuple((n, p._version, p.data_ptr()) for n, p in named.items() if not p.requires_grad)
        if self._engine_obj is None:
            self._engine_obj = self.engine_cls(self.fb, self._engine_params(named), dev, self.precision)
            self._anchor = torch.zeros(1, device=dev, requires_grad=True)
            self._flat["version"] = -1
        elif fsig != self._frozen_sig:
            # load_state_dict / in-place edit / .to(): refresh the operand copies of the frozen matrices
            self._engine_obj.P = self._engine_params(named)
            self._engine_obj.prepare_frozen()
        self._frozen_sig = fsig
        fl = self._flat
        ver = fl["flat32"]._version + sum(p._version for n, p in train)
        if fl["version"] != ver:
            self._engine_obj.refresh_trainable(fl["flat32"], fl["flat16"], fl["slices"])
            fl["version"] = fl["flat32"]._version + sum(p._version for n, p in train)
        return self._engine_obj

    def _mark_trainable_fresh(self):
        """Called by the fused optimizer after it updated flat32 AND its bf16 copy in one kernel."""
        fl = self._flat
        train = [p for p in self.parameters() if p.requires_grad]
        fl["version"] = fl["flat32"]._version + sum(p._version for p in train)

    def flat_buffers(self):
        """(flat fp32 params, flat fp32 grads, flat bf16 copy) of the trainable set (after the first forward)."""
        self._ensure_engine()
        return self._flat["flat32"], self._flat["grad32"], self._flat["flat16"]

    # ------------------------------------------------------------------ reference API
    def set_answer_embeddings(self, a2tok, freeze_last=True):
        """Reference deberta.py:1358-1380: mean word embedding of the non-pad answer tokens.  The reference
        assigns the bias to an ATTRIBUTE (``answer_bias.weight``, :1375), so ``answer_bias`` itself keeps its
        values; that observable behaviour is reproduced."""
        we = self._word_embeddings().weight
        pad = self.fb.pad_token_id
        a2tok = a2tok.to(we.device)
        a2v = we.data[a2tok]
        nz = (a2tok != pad)
        sum_tokens = nz.sum(1, keepdim=True)
        if not freeze_last:
            self._require_frozen(True, True, freeze_last)
        if len(a2v) != self.n_ans:
            assert not self.training
            self.n_ans = len(a2v)
            self.answer_embeddings = nn.Embedding(self.n_ans, self.fb.hidden_size).to(we.device)
            self.answer_bias.requires_grad = False
            self.answer_bias.resize_(self.n_ans)
            self.fb = self.fb.replace(n_ans=self.n_ans)
            self._engine_obj = None
            self._flat = None
        self.answer_embeddings.weight.data = torch.div((a2v * nz.float()[:, :, None]).sum(1), sum_tokens.clamp(min=1))
        if freeze_last:
            self.answer_embeddings.requires_grad_(False)
            self.answer_bias.requires_grad_(False)
        if self._engine_obj is not None:
            self._engine_obj.P["answer_embeddings.weight"] = self.answer_embeddings.weight.data
            self._engine_obj.refresh_answer_embeddings()

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, inputs_embeds=None,
                labels=None, output_attentions=None, return_dict=None, video=None, video_mask=None, mlm=False):
        if input_ids is None:
            if inputs_embeds is not None:
                raise NotImplementedError("inputs_embeds is not supported by the fused path")
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        eng = self._ensure_engine()
        dev = self.device
        B, Lt = input_ids.shape
        use_video = bool(self.features_dim) and video is not None
        F = video.shape[1] if use_video else 0
        L = F + Lt
        training = self.training and torch.is_grad_enabled()
        fb = self.fb
        if labels is not None:
            labels_on_host = labels if labels.device.type == "cpu" else None
            labels = labels.to(dev)
            flat_lab = torch.cat([torch.full((B, F), -100, dtype=torch.long, device=dev), labels], 1).reshape(-1) \
                if F else labels.reshape(-1)
            # Number of labelled rows (the head's GEMM M): counted on the host when the labels live there; otherwise on
            # the device into pinned memory, and read only AFTER the encoder has been queued -- the wait then ends when
            # the GPU reaches the start of this forward, with the whole encoder still in its queue (no idle GPU).
            lab_ne = flat_lab != -100
            if labels_on_host is not None:
                n_sel, cnt_ev = int((labels_on_host != -100).sum()), None
            else:
                n_sel, cnt_ev = None, self._count_async(lab_ne)
        with torch.no_grad():
            eng.forward_encoder(video, video_mask, input_ids, attention_mask, training)
            state = eng.cur
            if labels is not None:
                if n_sel is None:
                    cnt_ev.synchronize()
                    n_sel = int(self._count_host[0])
                sel = torch.nonzero_static(lab_ne, size=n_sel).reshape(-1)     # size known: no host sync
                if n_sel == 0:
                    loss = torch.full((), float("nan"), device=dev)
                else:
                    loss = eng.head_mlm_loss(sel, flat_lab[sel].contiguous(), n_sel, want_grad=training)
            else:
                logits = self._full_logits(eng, state, B, L, mlm=mlm)
        if labels is not None:
            logits_fn = lambda: self._full_logits(eng, state, B, L, mlm=True, scratch=True)
            if training and n_sel > 0:
                loss = _MLMLossFn.apply(self._anchor, self, state, loss)
            else:
                state.release()
            return FBMaskedLMOutput(loss=loss, logits_fn=logits_fn)
        if training:
            logits = _LogitsFn.apply(self._anchor, self, state, logits)
        return FBMaskedLMOutput(logits=logits)

    @torch.no_grad()
    def answer_scores(self, input_ids, attention_mask=None, video=None, video_mask=None, *, mask_token_id, topk=1,
                      candidates=1, return_probs=False):
        """Inference fast path of the downstream scripts: what videoqa.py:157-170 and mc.py:150-172
        compute from ``forward()["logits"]`` -- the answer distribution AT THE [MASK] TOKEN -- without materialising
        the [B, L, n_ans] logits: the head GEMM runs on the [MASK] rows only and softmax + top-k are one kernel.

        candidates = C > 1 (multiple choice): ``input_ids`` / ``attention_mask`` are [B*C, Lt] with the C candidates
        of a question in consecutive rows, ``video`` / ``video_mask`` are [B, ...]; the visual prefix is projected
        once per clip and all B*C rows go through the encoder as one batch (mc.py runs C forwards of [B, Lt]).

        Returns dict(top_p [n, k], top_i [n, k], rows [n] (flat index of each [MASK] position in [B*C, Lt]),
        probs [n, n_ans] when return_probs).  mc.py's score of candidate c is ``probs[:, 0]`` (mc.py:172)."""
        if not self.fb.n_ans:
            raise ValueError("answer_scores needs the answer head (n_ans > 0)")
        eng = self._ensure_engine()
        dev = self.device
        ids_on_host = input_ids if input_ids.device.type == "cpu" else None
        input_ids = input_ids.to(dev)
        Bt, Lt = input_ids.shape
        use_video = bool(self.features_dim) and video is not None
        F = video.shape[1] if use_video else 0
        is_mask = input_ids.reshape(-1) == mask_token_id
        # [MASK] count: on the host when the ids live there, else read back only after the encoder has been queued
        n, cnt_ev = (int((ids_on_host == mask_token_id).sum()), None) if ids_on_host is not None \
            else (None, self._count_async(is_mask))
        eng.forward_encoder(video, video_mask, input_ids, attention_mask, False, video_repeat=candidates)
        if n is None:
            cnt_ev.synchronize()
            n = int(self._count_host[0])
        pos = torch.nonzero_static(is_mask, size=n).reshape(-1)      # the reference's boolean index, size known
        rows = (pos // Lt) * (F + Lt) + F + pos % Lt
        rows16 = eng.ws.rows("ans.rows16", n, self.fb.hidden_size, eng.AD)
        lib.move_rows(eng.cur.o2[1], rows16, idx_in=rows.contiguous())
        logits, V = eng.head_logits(rows16, n, mlm=False, scratch=True)
        k = min(int(topk), V)
        probs, top_p, top_i = lib.softmax_topk(logits, V, k, want_probs=return_probs)
        out = dict(top_p=top_p, top_i=top_i, rows=pos)
        if return_probs:
            out["probs"] = probs
        return out

    def _count_async(self, flags):
        """flags.sum() -> pinned host memory on the current stream; returns the event to wait for before reading
        ``self._count_host``."""
        if getattr(self, "_count_host", None) is None:
            self._count_host = torch.zeros(1, dtype=torch.long).pin_memory()
        self._count_host.copy_(flags.sum(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def _full_logits(self, eng, state, B, L, mlm, scratch=False):
        if eng.cur is not state:
            raise lib.FbError("the lazily materialised logits of an earlier forward are gone: read out['logits'] "
                              "before the next forward()")
        T = B * L
        n = self.fb.vocab_size if (mlm or not self.fb.n_ans) else self.fb.n_ans
        out = torch.empty(T, n, device=self.device, dtype=torch.float32)
        with torch.no_grad():
            eng.head_logits(state.o2[1], T, mlm=(mlm or not self.fb.n_ans), out=out, scratch=scratch)
        return out.view(B, L, n)

    # ------------------------------------------------------------------ data-parallel gradient all-reduce
    @staticmethod
    def _dp_world():
        import torch.distributed as dist
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def _layer_ranges(self):
        """flat-bucket range [a, b) of every encoder layer's trainable tensors (contiguous: named_parameters order)
        and the ranges in front of / behind the layers."""
        fl = self._flat
        if "layer_ranges" not in fl:
            pfx = self.engine_cls.LAYER_PREFIX
            rng = {}
            for n, (a, b, _) in fl["slices"].items():
                if n.startswith(pfx):
                    i = int(n[len(pfx):].split(".")[0])
                    lo, hi = rng.get(i, (a, b))
                    rng[i] = (min(lo, a), max(hi, b))
            spans = sorted(rng.values())
            ok = all(spans[k][1] == spans[k + 1][0] for k in range(len(spans) - 1)) if spans else False
            fl["layer_ranges"] = rng if ok else {}
            fl["rest_ranges"] = [(0, spans[0][0]), (spans[-1][1], fl["total"])] if ok else [(0, fl["total"])]
        return fl["layer_ranges"], fl["rest_ranges"]

    def _allreduce_range(self, a, b, async_op):
        import torch.distributed as dist
        if b <= a:
            return None
        t = self._flat["grad32"][a:b]
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        w = dist.all_reduce(t, async_op=False)          # gloo (CPU tests): SUM, then the 1/world factor
        t.mul_(1.0 / dist.get_world_size())
        return None

    def _dp_layer_ready(self, layer, events):
        """Called by the engine as a layer's gradients complete: all-reduce that layer's slice of the flat bucket on a
        communication stream, overlapped with the backward of the earlier layers."""
        dp = self._dp
        rng = dp["layers"].get(layer)
        if rng is None or layer in dp["done"]:
            return
        comm = dp["stream"]
        if comm is None:                                    # host-logic tests on CPU (gloo): reduce in place
            w = self._allreduce_range(rng[0], rng[1], async_op=False)
        else:
            for ev in events:
                comm.wait_event(ev)
            with torch.cuda.stream(comm):
                w = self._allreduce_range(rng[0], rng[1], async_op=True)
        dp["done"].add(layer)
        dp["works"].append(w)

    def _finalize_backward(self):
        """End of the autograd pass (queued like DistributedDataParallel's finalizer): reduce what the overlapped path
        did not cover -- everything, when several forwards shared this backward -- and join the communication."""
        dp, self._dp = self._dp, None
        if dp is None:
            return
        layers, rest = dp["layers"], dp["rest"]
        todo = list(rest) if dp["done"] else [(0, self._flat["total"])]
        if dp["done"]:
            todo += [layers[i] for i in layers if i not in dp["done"]]
        for w in dp["works"]:
            if w is not None:
                w.wait()
        for a, b in todo:
            self._allreduce_range(a, b, async_op=False)
        self._flat["synced"] = True

    # ------------------------------------------------------------------ backward bridge
    def _run_backward(self, state, scale=None, dlogits=None):
        eng = self._engine_obj
        fl = self._flat
        named = self._param_dict()
        train = [(n, p) for n, p in named.items() if p.requires_grad]
        views = {n: fl["grad32"][a:b].view(shape) for n, (a, b, shape) in fl["slices"].items()}
        aliased = [p.grad is not None and p.grad.data_ptr() == views[n].data_ptr() for n, p in train]
        if not all(aliased):
            # fresh step (zero_grad(set_to_none=True)) or foreign .grad tensors: start the bucket from zero
            foreign = {n: p.grad for (n, p), al in zip(train, aliased) if p.grad is not None and not al}
            fl["grad32"].zero_()
        else:
            foreign = {}
        hook = None
        if self.dp_sync and self._dp_world() > 1:
            if self._dp is None:
                # first backward node of this autograd pass: the finalizer runs when the whole pass is done
                layers, rest = self._layer_ranges()
                self._dp = dict(layers=layers, rest=rest, done=set(), works=[], stream=None)
                torch.autograd.Variable._execution_engine.queue_callback(self._finalize_backward)
                # overlap per layer only when this is the ONLY forward feeding this backward pass (several pending
                # forwards -- mc.py's candidate loop -- accumulate into the same slices until the pass ends)
                if self.dp_overlap and layers and not foreign and eng.pending_forwards() == 1 and \
                        self._flat["grad32"].is_cuda:
                    if "comm_stream" not in self._flat:
                        self._flat["comm_stream"] = torch.cuda.Stream(device=self.device)
                    self._dp["stream"] = self._flat["comm_stream"]
                    hook = self._dp_layer_ready
            self._flat["synced"] = False
        eng.backward(views, scale=scale, dlogits=dlogits, state=state, on_grads_ready=hook)
        for n, p in train:
            if n in foreign:
                views[n].add_(foreign[n])
            p.grad = views[n]



class DebertaV2ForMaskedLM(FrozenBiLMBase):
    def __init__(self, config, max_feats=10, features_dim=768, freeze_lm=True, freeze_mlm=True, ds_factor_attn=8,
                 ds_factor_ff=8, ft_ln=True, dropout=0.1, n_ans=0, freeze_last=True):
        super().__init__()
        self._require_frozen(freeze_lm, freeze_mlm, freeze_last)
        self.config = config
        self.fb = FBConfig.from_hf(config, features_dim=features_dim, max_feats=max_feats,
                                   ds_factor_attn=ds_factor_attn, ds_factor_ff=ds_factor_ff,
                                   adapter_dropout=dropout, n_ans=n_ans, freeze_lm=freeze_lm, freeze_mlm=freeze_mlm,
                                   ft_ln=ft_ln, freeze_last=freeze_last)
        fb = self.fb
        H, I, V = fb.hidden_size, fb.intermediate_size, fb.vocab_size
        eps = fb.layer_norm_eps
        # ---- parameter holders with the reference's module tree (deberta.py:950-995, 420-472, 1523-1570)
        deb = _Holder()
        emb = _Holder()
        emb.word_embeddings = nn.Embedding(V, H, padding_idx=fb.pad_token_id)
        emb.position_embeddings = nn.Embedding(fb.max_position_embeddings, H)
        emb.LayerNorm = nn.LayerNorm(H, eps)
        emb.register_buffer("position_ids", torch.arange(fb.max_position_embeddings).expand((1, -1)))
        if features_dim:
            emb.linear_video = nn.Linear(features_dim, H)
        deb.embeddings = emb
        enc = _Holder()
        layers = []
        for _ in range(fb.num_hidden_layers):
            lay = _Holder()
            att = _Holder()
            sa = _Holder()
            sa.query_proj, sa.key_proj, sa.value_proj = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)
            att.self = sa
            ao = _Holder()
            ao.dense = nn.Linear(H, H)
            ao.LayerNorm = nn.LayerNorm(H, eps)
            if ds_factor_attn:
                ao.adapter = _adapter(ds_factor_attn, H)
            att.output = ao
            lay.attention = att
            inter = _Holder()
            inter.dense = nn.Linear(H, I)
            lay.intermediate = inter
            out = _Holder()
            out.dense = nn.Linear(I, H)
            out.LayerNorm = nn.LayerNorm(H, eps)
            if ds_factor_ff:
                out.adapter = _adapter(ds_factor_ff, H)
            lay.output = out
            layers.append(lay)
        enc.layer = nn.ModuleList(layers)
        enc.rel_embeddings = nn.Embedding(2 * fb.position_buckets, H)
        enc.LayerNorm = nn.LayerNorm(H, eps, elementwise_affine=True)
        conv = _Holder()
        conv.conv = nn.Conv1d(H, H, 3, padding=1)
        conv.LayerNorm = nn.LayerNorm(H, eps)
        enc.conv = conv
        deb.encoder = enc
        self.deberta = deb
        head = _Holder()
        lm = _Holder()
        lm.dense = nn.Linear(H, H)
        lm.LayerNorm = nn.LayerNorm(H, eps=eps)
        lm.decoder = nn.Linear(H, V, bias=False)   # allocated but unused by forward, like the reference (:1535-1542)
        lm.bias = nn.Parameter(torch.zeros(V))
        lm.decoder.bias = lm.bias
        head.lm_head = lm
        self.lm_predictions = head
        self.features_dim = features_dim
        self.n_ans = n_ans
        self._init_weights()
        # ---- freezing rules (deberta.py:1152-1158, 1334-1339)
        if freeze_lm:
            for n, p in self.deberta.named_parameters():
                if ("linear_video" not in n) and ("adapter" not in n):
                    if ft_ln and "LayerNorm" in n:
                        continue
                    p.requires_grad_(False)
        if freeze_mlm:
            for n, p in self.lm_predictions.named_parameters():
                if ft_ln and "LayerNorm" in n:
                    continue
                p.requires_grad_(False)
        if n_ans:
            self.answer_embeddings = nn.Embedding(n_ans, H)
            self.answer_bias = nn.Parameter(torch.zeros(n_ans))
            if freeze_last:
                self.answer_embeddings.requires_grad_(False)
                self.answer_bias.requires_grad_(False)
        self._post_init_state()

    def _engine_params(self, named):
        return {n: p.data for n, p in named.items() if "decoder" not in n}   # lm_head.decoder is unused (:1535-1542)

    # ------------------------------------------------------------------ init / load
    def _init_weights(self):
        """N(0, initializer_range) matrices, zero biases, unit LayerNorm (deberta.py _init_weights);
        adapters N(0, 1e-3) clamped to +-2e-3, biases included (adapter.py:23-28)."""
        std = getattr(self.config, "initializer_range", 0.02)
        for name, m in self.named_modules():
            if isinstance(m, (nn.Linear, nn.Conv1d)):
                if ".adapter." in name:
                    nn.init.normal_(m.weight, std=1e-3)
                    nn.init.normal_(m.bias, std=1e-3)
                    m.weight.data.clamp_(-2e-3, 2e-3)
                    m.bias.data.clamp_(-2e-3, 2e-3)
                elif isinstance(m, nn.Linear):
                    m.weight.data.normal_(0.0, std)
                    if m.bias is not None:
                        m.bias.data.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.data.normal_(0.0, std)
                if m.padding_idx is not None:
                    m.weight.data[m.padding_idx].zero_()

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, local_files_only=True, **kw):
        """Loads an HF-format DeBERTa-V2 directory (config.json + pytorch_model.bin / model.safetensors),
        like PreTrainedModel.from_pretrained in the reference (model/__init__.py:34-47)."""
        from transformers import DebertaV2Config

        path = pretrained_model_name_or_path
        config = DebertaV2Config.from_pretrained(path, local_files_only=local_files_only)
        model = cls(config=config, **kw)
        sd = None
        st = os.path.join(path, "model.safetensors")
        pt = os.path.join(path, "pytorch_model.bin")
        if os.path.exists(st):
            from safetensors.torch import load_file

            sd = load_file(st)
        elif os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu")
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
        sd = {k.replace("lm_predictions.lm_head.LayerNorm.gamma", "lm_predictions.lm_head.LayerNorm.weight")
               .replace("lm_predictions.lm_head.LayerNorm.beta", "lm_predictions.lm_head.LayerNorm.bias"): v
              for k, v in sd.items()}
        model.load_state_dict(sd, strict=False)
        return model

