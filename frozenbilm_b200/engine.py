"""Kernel schedule of the FrozenBiLM masked-LM hot path (DeBERTa-V2 + adapters + visual prefix).

This module only SEQUENCES the C-ABI kernels of libfrozenbilm_sm100.so (tcgen05 GEMMs, fused
disentangled attention, LayerNorm/row kernels): PyTorch provides device memory, streams and the
tiny index/mask preparation.  There is no eager/CPU fallback.

Forward  = reference model/deberta.py:997-1058 (embeddings + visual prefix), :507-575 (encoder, the
           discarded last ordinary pass skipped), :395-419 (conv after layer 0),
           :1382-1412 (two enhanced-mask-decoder passes), :1544-1558 + :1481-1488 (head + CE).
Backward = hand-scheduled selective backward: dX through the frozen linears, dW only for adapters,
           LayerNorms and linear_video (freezing rules deberta.py:1152-1158, 1334-1339).  GEMM outputs
           needed by the backward are kept (180 GB of HBM makes a full recompute a net loss); cheap
           things are recomputed: dropout masks (Philox), GELU/ReLU derivatives, softmax probabilities.
"""
from __future__ import annotations

import math
import os

import torch

from . import lib
from .config import FBConfig
from .relpos import bucket_of_rel

BF16, F32 = torch.bfloat16, torch.float32

# dropout site ids inside one layer pass (offset = pass_id * 8 + site)
SITE_POS, SITE_PROBS, SITE_AD1, SITE_HID1, SITE_AD2, SITE_HID2 = 0, 1, 2, 3, 4, 5
OFF_EMB, OFF_CONV = 100000, 100001


def _ceil8(n):
    return (n + 7) // 8 * 8


_STREAMS = {}


def _priority_streams(device):
    """(high-priority, low-priority) stream pair of a device, created once per process."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _STREAMS.get(key)
    if st is None:
        lo_p, hi_p = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
        st = (torch.cuda.Stream(device=device, priority=hi_p), torch.cuda.Stream(device=device, priority=lo_p))
        _STREAMS[key] = st
    return st


class Workspace:
    """Named scratch tensors reused across steps (no per-step cudaMalloc).  A buffer is keyed by NAME and grows to the
    largest element count ever requested; ``get`` hands out a view of its head, so a run whose (B, L) changes from
    batch to batch (``padding='longest'`` in the reference's loaders) keeps ONE set of buffers sized for the largest
    batch instead of one set per distinct shape."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape, dtype, zero=False):
        key = (name, dtype)
        n = 1
        for d in shape:
            n *= int(d)
        t = self.bufs.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(max(n, 1), device=self.device, dtype=dtype)
            self.bufs[key] = t
        v = t[:n].view(*shape)
        if zero:
            v.zero_()
        return v

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())

    def rows(self, name, nrows, ncols, dtype):
        """[nrows, ncols] view of a growable buffer (row counts that vary per step, e.g. labelled positions)."""
        key = (name + "#rows", ncols, dtype)
        t = self.bufs.get(key)
        if t is None or t.shape[0] < nrows:
            cap = max(256, int(nrows * 1.25) + 8)
            t = torch.empty((cap, ncols), device=self.device, dtype=dtype)
            self.bufs[key] = t
        return t[:nrows]


class ForwardState:
    """Everything ONE forward leaves behind for its backward: the saved activations (in a workspace slot of their
    own), the head state, the dimensions, the mask and the dropout seed.  The autograd node of that forward holds it,
    so several train-mode forwards may be pending before one ``loss.backward()`` (the multiple-choice loop of the
    reference's mc.py runs one forward per answer candidate)."""

    def __init__(self, engine, slot):
        self.engine, self.slot = engine, slot
        self.sv = None
        self.head = None
        self.dims = None
        self.mask_u8 = None
        self.seed = None
        self.o2 = None
        self.released = False

    def release(self):
        if not self.released:
            self.released = True
            self.sv = self.head = None
            if self.slot is not None:
                self.engine._free_slots.append(self.slot)
                self.slot = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class DebertaEngine:
    LAYER_PREFIX = "deberta.encoder.layer."

    def __init__(self, cfg: FBConfig, params: dict, device, precision: str = "bf16"):
        """params: name -> fp32 CUDA tensor (the module's parameters; reference names).
        precision: "bf16" (training path: bf16 tensor-core operands, fp32 accumulation) or "fp32x3" (fp32-class
        instantiation: every contraction as a 3-pass split-bf16 product on the same tcgen05 GEMM, attention in fp32;
        the parity gate of BASELINE configs[0], see fb_precise.cu)."""
        lib.load()
        if device.type != "cuda":
            raise lib.FbError("frozenbilm_b200 has no CPU path: move the model to a CUDA device")
        if precision not in ("bf16", "fp32x3"):
            raise ValueError(f"unknown precision {precision!r} (bf16 | fp32x3)")
        self.cfg = cfg
        self.P = params
        self.device = device
        self.precise = precision == "fp32x3"
        self.AD = F32 if self.precise else BF16      # dtype of the GEMM-operand ("16") buffers
        self.ws = Workspace(device)
        self.sws = self.ws                           # where the current forward keeps what its backward needs
        self._free_slots = []
        self._n_slots = 0
        self.max_pending = int(os.environ.get("FB_MAX_PENDING_FORWARDS", "8"))
        self.cur = None
        self.saved = None
        self.head_saved = None
        self._on_grads_ready = None
        self.step_seed = 0x5EED0000
        self.cur_seed = self.step_seed
        if cfg.head_dim != 64:
            raise NotImplementedError("the fused attention kernel is specialised for head_dim = 64")
        self.use_rel = True                                   # disentangled (c2p + p2c) attention
        # attention backward: one fused kernel per pass (default) or the staged GEMM schedule (FB_ATTN_BWD=staged)
        self.fused_attn_bwd = self.precise or os.environ.get("FB_ATTN_BWD", "fused") != "staged"
        # adapter weight gradients on a low-priority side stream (FB_DW_STREAM=0 keeps everything on one stream)
        self.dw_side_stream = os.environ.get("FB_DW_STREAM", "1") != "0" and not self.precise
        self._side = None
        # ... and the whole positional-gradient chain (fb_attn_bwd_rel -> dR; FB_REL_STREAM=0 keeps it in the chain).
        # FB_REL_UNITS = tile pairs per CTA of the low-priority band-gradient kernel (0 = the kernel's own choice);
        # measured at config 2: off 75.2 ms, 8 units 73.9, 4 units 73.8, 2 units 76.0 (too many accumulator flushes)
        self.rel_side_stream = self.dw_side_stream and os.environ.get("FB_REL_STREAM", "1") != "0"
        self.rel_side_units = int(os.environ.get("FB_REL_UNITS", "0"))
        # FB_REL_PREFETCH=1: build the positional tables of pass p + 1 on the low-priority stream while pass p runs.
        # Measured at config 2: 71.9 ms/step vs 71.1 ms without -- a running low-priority GEMM CTA is not preempted, so
        # the [512 x 2H x H] projection delays the chain more than the three launches it takes off it.  Off by default.
        self.rel_prefetch = self.dw_side_stream and os.environ.get("FB_REL_PREFETCH", "0") == "1"
        # FB_REL_BATCHED=0: build the positional tables pass by pass inside the chain (three launches per pass) instead
        # of for all passes up front (_rel_tables_all)
        self.rel_batched = os.environ.get("FB_REL_BATCHED", "1") != "0"
        self._rel_bwd_batched = False
        self.attn_scale = 1.0 / math.sqrt(3 * 64)             # deberta.py:769-777
        self.layer_prefix = self.LAYER_PREFIX
        if cfg.hidden_size % 128 != 0:
            raise NotImplementedError("hidden_size must be a multiple of 128")
        self.bucket_of_rel = torch.from_numpy(bucket_of_rel(cfg.position_buckets, cfg.max_position_embeddings)).to(device) \
            if cfg.position_buckets > 0 else None
        self.W = {}
        # head parameter names (DeBERTa: lm_predictions.lm_head.*; BERT: cls.predictions.*)
        self.hn = dict(dense_b="lm_predictions.lm_head.dense.bias", ln_w="lm_predictions.lm_head.LayerNorm.weight",
                       ln_b="lm_predictions.lm_head.LayerNorm.bias", bias="lm_predictions.lm_head.bias")
        self.prepare_frozen()

    # ------------------------------------------------------------------ weight preparation
    def prepare_frozen(self):
        """bf16 operand copies of the LM matrices.  Re-run by the module whenever a frozen Parameter's version counter
        or storage changes (load_state_dict, in-place edits, .to()): see FrozenBiLMBase._ensure_engine."""
        cfg, P, W = self.cfg, self.P, self.W
        H = cfg.hidden_size
        BF16 = self.AD   # fp32-class mode keeps the operands unrounded (split at the GEMM)
        with torch.no_grad():
            # Q|K|V weights of ALL layers in one [NL, 3H, H] tensor (constant layer stride): the positional projections
            # of every layer pass are one batched GEMM (_rel_tables_all)
            NL = cfg.num_hidden_layers
            W["qkv.w.all"] = torch.empty(NL, 3 * H, H, device=self.device, dtype=BF16)
            W["qkv.b.all"] = torch.empty(NL, 3 * H, device=self.device, dtype=F32)
            for i in range(NL):
                p = f"deberta.encoder.layer.{i}."
                a = p + "attention.self."
                W[f"qkv.w.{i}"] = W["qkv.w.all"][i]
                W[f"qkv.b.{i}"] = W["qkv.b.all"][i]
                for j, n in enumerate(("query_proj", "key_proj", "value_proj")):
                    W[f"qkv.w.{i}"][j * H:(j + 1) * H].copy_(P[a + n + ".weight"])
                W[f"qkv.b.{i}"].copy_(torch.cat([P[a + "query_proj.bias"], P[a + "key_proj.bias"],
                                                 P[a + "value_proj.bias"]], 0))
                W[f"o.w.{i}"] = P[p + "attention.output.dense.weight"].to(BF16).contiguous()
                W[f"ffn1.w.{i}"] = P[p + "intermediate.dense.weight"].to(BF16).contiguous()
                W[f"ffn2.w.{i}"] = P[p + "output.dense.weight"].to(BF16).contiguous()
            c = "deberta.encoder."
            # Conv1d weight [out, in, 3] -> [out, 3*in] with column = k * H + in (matches fb_conv_im2col)
            W["conv.w"] = P[c + "conv.conv.weight"].permute(0, 2, 1).reshape(H, 3 * H).to(BF16).contiguous()
            W["head.w"] = P["lm_predictions.lm_head.dense.weight"].to(BF16).contiguous()
            W["emb16"] = P["deberta.embeddings.word_embeddings.weight"].to(BF16).contiguous()
            if cfg.n_ans:
                self.refresh_answer_embeddings()

    def refresh_answer_embeddings(self):
        E = self.P["answer_embeddings.weight"]
        n = E.shape[0]
        Ep = torch.zeros(_ceil8(n), self.cfg.hidden_size, device=self.device, dtype=self.AD)
        Ep[:n] = E.to(self.AD)
        self.W["ans16"] = Ep
        self.W["ans_n"] = n

    def refresh_trainable(self, flat32, flat16, slices):
        """flat16 = bf16 copy of the flat trainable bucket (one cast kernel); slices: name -> (a, b, shape)."""
        lib.cast_bf16(flat32, flat16)
        src = flat32 if self.precise else flat16
        self.T16 = {k: src[a:b].view(shape) for k, (a, b, shape) in slices.items()}

    def t16(self, name):
        return self.T16[name]

    # ------------------------------------------------------------------ helpers
    def drop(self, p, offset):
        return lib.drop(p if self.training else 0.0, self.cur_seed, offset)

    def _acquire_slot(self):
        """Workspace for the activations one train-mode forward saves (returned to the pool by its backward, or when
        its autograd node is dropped)."""
        if self._free_slots:
            return self._free_slots.pop()
        if self._n_slots >= self.max_pending:
            raise lib.FbError(f"{self._n_slots} train-mode forwards are pending without a backward "
                              "(FB_MAX_PENDING_FORWARDS); run them under torch.no_grad() or call backward()")
        self._n_slots += 1
        return Workspace(self.device)

    def begin_forward(self, training):
        self.training = bool(training)
        self.step_seed += 1
        self.cur_seed = self.step_seed
        st = ForwardState(self, self._acquire_slot() if self.training else None)
        st.seed = self.cur_seed
        self.sws = st.slot if st.slot is not None else self.ws
        self.cur = st
        self.saved = None
        self.head_saved = None
        return st

    def _video_prefix(self, video, B, F, L, e32, wname, bname, repeat=1):
        """linear_video(video) into the F visual rows of every sample of e32 [B*L, H] (deberta.py:993-1010).
        ``repeat`` = C: one clip serves C consecutive samples and is projected once."""
        cfg, ws, dev = self.cfg, self.ws, self.device
        Dv, H = cfg.features_dim, cfg.hidden_size
        Bv = B // repeat
        if repeat > 1 and self.training:
            raise NotImplementedError("video_repeat > 1 is an inference path (the candidates share one projection)")
        v16 = self.sws.get("video16", (Bv * F, _ceil8(Dv)), self.AD)
        vin = video.to(dev, F32).reshape(Bv * F, Dv).contiguous()
        if _ceil8(Dv) != Dv:
            v16.zero_()
            v16[:, :Dv] = vin.to(self.AD)
        else:
            lib.cast_bf16(vin, v16)
        ar = torch.arange(B, device=dev)[:, None] * L
        if self.precise or repeat > 1:
            pv32 = ws.get("video.proj", (Bv * F, H), F32)
            lib.gemm(v16, self.t16(wname), pv32, bias=self.P[bname], N=H, K=Dv)
            fr = torch.arange(F, device=dev)[None, :]
            idx_in = ((torch.arange(B, device=dev) // repeat)[:, None] * F + fr).reshape(-1).contiguous()
            lib.move_rows(pv32, e32, idx_in=idx_in, idx_out=(ar + fr).reshape(-1).contiguous())
        else:
            lib.gemm(v16, self.t16(wname), e32, bias=self.P[bname], M=F, N=H, K=Dv,
                     batch={"outer": (B, F * v16.stride(0), 0, L * H)})
        idx_out = (ar + F + torch.arange(L - F, device=dev)[None, :]).reshape(-1).contiguous()
        return v16, idx_out

    # ------------------------------------------------------------------ forward
    def forward_encoder(self, video, video_mask, input_ids, attention_mask, training, video_repeat=1):
        """``video_repeat`` = C > 1: ``video`` holds one clip per C consecutive rows of ``input_ids`` (the answer
        candidates of one question, mc.py:150-172); the visual prefix is projected once per clip."""
        cfg, P, W, ws = self.cfg, self.P, self.W, self.ws
        st = self.begin_forward(training)
        sws, AD = self.sws, self.AD
        dev = self.device
        H, NL, h = cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads
        B, Lt = input_ids.shape
        use_video = bool(cfg.features_dim) and video is not None
        F = video.shape[1] if use_video else 0
        L = F + Lt
        T = B * L
        if L > cfg.max_position_embeddings or L > 512:
            raise ValueError(f"sequence length {L} exceeds max_position_embeddings / kernel limit 512")
        self.dims = st.dims = (B, L, F, Lt, T)
        hid_p, att_p, ad_p = cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob, cfg.adapter_dropout
        e = "deberta.embeddings."
        sv = {"passes": []} if training else None

        # ---- masks (tiny index math stays in torch)
        if attention_mask is None:
            attention_mask = torch.ones(B, Lt, device=dev, dtype=torch.long)
        if use_video:
            if video_mask is None:
                video_mask = torch.ones(video.shape[0], F, device=dev, dtype=torch.long)
            if video_repeat > 1:
                if video.shape[0] * video_repeat != B:
                    raise ValueError(f"video_repeat={video_repeat}: {video.shape[0]} clips for {B} text rows")
                video_mask = video_mask.to(dev).repeat_interleave(video_repeat, 0)
            m = torch.cat([video_mask.to(dev), attention_mask.to(dev)], 1)
        else:
            m = attention_mask.to(dev)
        mask_u8 = (m != 0).to(torch.uint8).contiguous()
        self.mask_u8 = st.mask_u8 = mask_u8

        # ---- embeddings: [linear_video(video) ; word_emb[ids]] -> LN -> * mask -> dropout
        e32 = sws.get("e32", (T, H), F32)
        v16, idx_out = None, None
        if use_video:
            v16, idx_out = self._video_prefix(video, B, F, L, e32, e + "linear_video.weight", e + "linear_video.bias",
                                              video_repeat)
        ids = input_ids.to(dev).reshape(-1).contiguous()
        lib.move_rows(P[e + "word_embeddings.weight"], e32, idx_in=ids, idx_out=idx_out)
        x32 = ws.get("x32.0", (T, H), F32)
        x16 = ws.get("x16.0", (T, H), AD)
        emb_mean, emb_rstd = sws.get("emb.mean", (T,), F32), sws.get("emb.rstd", (T,), F32)
        lib.layernorm_fwd(e32, P[e + "LayerNorm.weight"], P[e + "LayerNorm.bias"], cfg.layer_norm_eps,
                          rowmask=mask_u8, out_drop=self.drop(hid_p, OFF_EMB), y32=x32, y16=x16,
                          mean=emb_mean, rstd=emb_rstd)
        x0_16 = x16

        # ---- relative embeddings: R = LN(rel_embeddings.weight) (trainable LN), once per forward
        c = "deberta.encoder."
        S2 = 2 * cfg.position_buckets
        R32 = ws.get("R32", (S2, H), F32)
        rel_mean, rel_rstd = sws.get("rel.mean", (S2,), F32), sws.get("rel.rstd", (S2,), F32)
        lib.layernorm_fwd(P[c + "rel_embeddings.weight"], P[c + "LayerNorm.weight"], P[c + "LayerNorm.bias"],
                          cfg.layer_norm_eps, y32=R32, mean=rel_mean, rstd=rel_rstd)
        self._R32 = R32
        self._rel_tables_cache = {}
        self._rel_side_ready = None
        self._R32_event = torch.cuda.current_stream().record_event() if (self.use_rel and self.rel_prefetch) else None
        # the positional tables of pass p + 1 are built on the side stream while pass p runs
        pass_layers = list(range(NL - 1)) + [NL - 1, NL - 1]
        if self.use_rel and self.rel_batched and not self.precise:
            self._rel_tables_all(L, pass_layers)
        self._rel_tables_prefetch(pass_layers[0], 0, L)

        # ---- encoder: NL-1 ordinary passes (the NL-th only feeds hidden_states[-1], which nothing reads)
        pass_id = 0
        for i in range(NL - 1):
            self._rel_tables_prefetch(pass_layers[pass_id + 1], pass_id + 1, L)
            qkv = sws.get(f"qkv.{pass_id}", (T, 3 * H), AD) if training else ws.get("qkv", (T, 3 * H), AD)
            lib.gemm(x16, W[f"qkv.w.{i}"], qkv, bias=W[f"qkv.b.{i}"])
            y32, y16, rec = self._layer_pass(i, pass_id, x32, qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], training)
            if training:
                rec["qkv"] = qkv
            if i == 0:
                # ConvLayer on the EMBEDDING output; LN(layer0_out + gelu(dropout(conv))) * mask
                col16 = ws.get("col16", (T, 3 * H), AD)
                lib.conv_im2col(x0_16, B, L, H, col16)
                sc32 = sws.get("conv.s", (T, H), F32)
                cg16 = sws.get("conv.pre", (T, H), AD)
                lib.gemm(col16, W["conv.w"], sc32, bias=P[c + "conv.conv.bias"], rowmask=mask_u8,
                         dropout=self.drop(hid_p, OFF_CONV), drop_pos=1, D_pre=cg16, act=lib.FB_ACT_GELU, resid=y32)
                cm, cr = sws.get("conv.mean", (T,), F32), sws.get("conv.rstd", (T,), F32)
                y32b, y16b = ws.get("x32.c", (T, H), F32), ws.get("x16.c", (T, H), AD)
                lib.layernorm_fwd(sc32, P[c + "conv.LayerNorm.weight"], P[c + "conv.LayerNorm.bias"],
                                  cfg.layer_norm_eps, rowmask=mask_u8, y32=y32b, y16=y16b, mean=cm, rstd=cr)
                y32, y16 = y32b, y16b
                if training:
                    sv.update(conv_s=sc32, conv_pre=cg16, conv_mean=cm, conv_rstd=cr)
            if training:
                sv["passes"].append(rec)
            x32, x16 = y32, y16
            pass_id += 1

        # ---- enhanced mask decoder: 2 passes of the last layer, Q from z / o1, K/V from h[-2] (computed once)
        last = NL - 1
        kv32, kv16 = x32, x16
        z32, z16 = ws.get("emd.z32", (T, H), F32), ws.get("emd.z16", (T, H), AD)
        lib.add_rowbcast(kv32, P[e + "position_embeddings.weight"][:L].contiguous(), B, L, H, z32, z16)
        kvp = sws.get("emd.kv", (T, 2 * H), AD)
        lib.gemm(kv16, W[f"qkv.w.{last}"][H:], kvp, bias=W[f"qkv.b.{last}"][H:].contiguous())
        q32, q16 = z32, z16
        for j in range(2):
            if j == 0:
                self._rel_tables_prefetch(last, pass_id + 1, L)
            qp = sws.get(f"emd.q.{j}", (T, H), AD)
            lib.gemm(q16, W[f"qkv.w.{last}"][:H], qp, bias=W[f"qkv.b.{last}"][:H].contiguous())
            y32, y16, rec = self._layer_pass(last, pass_id, q32, qp, kvp[:, :H], kvp[:, H:], training, tag=f"e{j}")
            if training:
                rec["q"] = qp
                rec["kv"] = kvp
                sv["passes"].append(rec)
            q32, q16 = y32, y16
            pass_id += 1
        if training:
            sv.update(e32=e32, emb_mean=emb_mean, emb_rstd=emb_rstd, rel_mean=rel_mean, rel_rstd=rel_rstd,
                      video16=v16, use_video=use_video)
            self.saved = st.sv = sv
        self.o2_32, self.o2_16 = q32, q16
        st.o2 = (q32, q16)
        return q32, q16

    def _rel_tables_compute(self, layer, pass_id, L):
        """PQrel / PKrel for one layer pass: Linear_{q|k}(dropout(R)) expanded to relative positions
        (fp32-class mode: the un-expanded [2S, 2H] projections, the fp32 attention indexes them by bucket)."""
        cfg, W, ws, sws = self.cfg, self.W, self.ws, self.sws
        H, h = cfg.hidden_size, cfg.num_attention_heads
        S2 = 2 * cfg.position_buckets
        tagk = f".{pass_id}" if self.training else f".L{layer}"
        tws = sws if self.training else ws
        Rd16 = ws.get("Rd16", (S2, H), self.AD)
        lib.dropout_apply(self._R32, self.drop(cfg.hidden_dropout_prob, pass_id * 8 + SITE_POS), out16=Rd16)
        if self.precise:
            pqk = tws.get("pqk32" + tagk, (S2, 2 * H), F32)
            lib.gemm(Rd16, W[f"qkv.w.{layer}"][:2 * H], pqk, bias=W[f"qkv.b.{layer}"][:2 * H].contiguous())
            return (pqk, None, None)
        pqk = ws.get("pqk16", (S2, 2 * H), BF16)
        lib.gemm(Rd16, W[f"qkv.w.{layer}"][:2 * H], pqk, bias=W[f"qkv.b.{layer}"][:2 * H].contiguous())
        rel_rows, _ = lib.rel_table_geometry(L)
        pq = tws.get("PQrel" + tagk, (h, rel_rows, 64), BF16)
        pk = tws.get("PKrel" + tagk, (h, rel_rows, 64), BF16)
        pkr = tws.get("PKrev" + tagk, (h, rel_rows, 64), BF16) if (self.training and self.fused_attn_bwd) else None
        lib.build_rel_tables(pqk, self.bucket_of_rel, L, h, out_pq=pq, out_pk=pk, out_pk_rev=pkr)
        return (pq, pk, pkr)

    def _rel_tables_all(self, L, pass_layers):
        """PQrel / PKrel of EVERY layer pass of this forward in four launches (dropout, one batched GEMM over the
        layers, the trailing passes that reuse the last layer, the table expansion) instead of three per pass inside
        the chain: the projections depend on the relative embeddings and the frozen weights only, never on the
        activations.  Same per-pass dropout sites as _rel_tables_compute (bit-identical tables)."""
        cfg, W, ws = self.cfg, self.W, self.ws
        H, h, NL = cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers
        S2 = 2 * cfg.position_buckets
        if pass_layers[:NL] != list(range(NL)):
            return                                     # unexpected pass order: the per-pass path handles it
        p = cfg.hidden_dropout_prob if self.training else 0.0
        tws = self.sws if self.training else ws
        wall, ball = W["qkv.w.all"], W["qkv.b.all"]
        if p == 0.0:
            # without dropout the tables only depend on the LAYER
            keys = [("L", l) for l in range(NL)]
            Rd16 = ws.get("Rd16", (S2, H), BF16)
            lib.dropout_apply(self._R32, self.drop(0.0, SITE_POS), out16=Rd16)
            pqk = ws.get("pqk16.all", (NL, S2, 2 * H), BF16)
            lib.gemm(Rd16, wall[0][:2 * H], pqk[0], bias=ball[0][:2 * H],
                     batch={"outer": (NL, 0, 3 * H * H, S2 * 2 * H), "bias_stride": 3 * H})
        else:
            NP = len(pass_layers)
            keys = [("P", q) for q in range(NP)]
            Rd16 = ws.get("Rd16.all", (NP, S2, H), BF16)
            lib.dropout_batched(self._R32, self.drop(p, SITE_POS), 8, out16=Rd16)     # pass q: site q * 8 + SITE_POS
            pqk = ws.get("pqk16.all", (NP, S2, 2 * H), BF16)
            lib.gemm(Rd16[0], wall[0][:2 * H], pqk[0], bias=ball[0][:2 * H],
                     batch={"outer": (NL, S2 * H, 3 * H * H, S2 * 2 * H), "bias_stride": 3 * H})
            for q in range(NL, NP):                    # the enhanced-mask-decoder passes reuse the last layer
                l = pass_layers[q]
                lib.gemm(Rd16[q], wall[l][:2 * H], pqk[q], bias=ball[l][:2 * H])
        nb = len(keys)
        rel_rows, _ = lib.rel_table_geometry(L)
        pq = tws.get("PQrel.all", (nb, h, rel_rows, 64), BF16)
        pk = tws.get("PKrel.all", (nb, h, rel_rows, 64), BF16)
        pkr = tws.get("PKrev.all", (nb, h, rel_rows, 64), BF16) if (self.training and self.fused_attn_bwd) else None
        lib.build_rel_tables_batched(pqk, self.bucket_of_rel, L, h, pq, pk, pkr)
        for j, key in enumerate(keys):
            self._rel_tables_cache[key] = ((pq[j], pk[j], pkr[j] if pkr is not None else None), None)

    def _rel_tables_prefetch(self, layer, pass_id, L):
        """Issues the table construction of a LATER pass on the low-priority side stream: it only depends on the
        relative embeddings (R), never on the activations, so it is taken off the critical path of the forward (three
        small launches per pass: dropout, a [512 x 2H x H] GEMM, the table expansion)."""
        if not self.use_rel or not self.rel_prefetch:
            return
        p = self.cfg.hidden_dropout_prob if self.training else 0.0
        key = ("L", layer) if p == 0.0 else ("P", pass_id)
        if key in self._rel_tables_cache:
            return
        _, lo = _priority_streams(self.device)
        if self._rel_side_ready is None:
            lo.wait_event(self._R32_event)          # R = LN(rel_embeddings) of THIS forward
            self._rel_side_ready = True
        with torch.cuda.stream(lo):
            out = self._rel_tables_compute(layer, pass_id, L)
            ev = lo.record_event()
        self._rel_tables_cache[key] = (out, ev)

    def _rel_tables(self, layer, pass_id, L):
        p = self.cfg.hidden_dropout_prob if self.training else 0.0
        key = ("L", layer) if p == 0.0 else ("P", pass_id)     # without dropout the tables only depend on the layer
        hit = self._rel_tables_cache.get(key)
        if hit is None:
            hit = (self._rel_tables_compute(layer, pass_id, L), None)
            self._rel_tables_cache[key] = hit
        out, ev = hit
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        return out

    def _sublayer_out(self, name, i, pass_id, u_in16, W_dense, b_dense, ad_prefix, ln_prefix, resid32, site_ad,
                      site_hid, training, tag):
        """dense -> adapter -> dropout -> + residual -> LayerNorm  (deberta.py:254-260 / 328-334)."""
        cfg, P, ws, AD = self.cfg, self.P, self.ws, self.AD
        tws = self.sws if training else ws
        T = u_in16.shape[0]
        H = cfg.hidden_size
        sfx = f".{tag}" if training else ""
        hid = self.drop(cfg.hidden_dropout_prob, pass_id * 8 + site_hid)
        s32 = tws.get(f"{name}.s{sfx}", (T, H), F32)
        rec = {}
        r = cfg.r_attn if name == "att" else cfg.r_ff
        if r:
            u16 = tws.get(f"{name}.u{sfx}", (T, H), AD)
            lib.gemm(u_in16, W_dense, u16, bias=b_dense)
            z16 = tws.get(f"{name}.z{sfx}", (T, r), AD)
            lib.gemm(u16, self.t16(ad_prefix + "down.weight"), z16, bias=P[ad_prefix + "down.bias"],
                     act=lib.FB_ACT_RELU, dropout=self.drop(cfg.adapter_dropout, pass_id * 8 + site_ad), drop_pos=2)
            lib.gemm(z16, self.t16(ad_prefix + "up.weight"), s32, bias=P[ad_prefix + "up.bias"], aux=u16,
                     aux_op=lib.FB_AUX_ADD, dropout=hid, drop_pos=2, resid=resid32)
            rec.update(u=u16, z=z16)
        else:
            lib.gemm(u_in16, W_dense, s32, bias=b_dense, dropout=hid, drop_pos=2, resid=resid32)
        y32 = ws.get(f"{name}.y32.{pass_id % 2}", (T, H), F32)
        y16 = ws.get(f"{name}.y16.{pass_id % 2}", (T, H), AD)
        mean = tws.get(f"{name}.mean{sfx}", (T,), F32)
        rstd = tws.get(f"{name}.rstd{sfx}", (T,), F32)
        lib.layernorm_fwd(s32, P[ln_prefix + "weight"], P[ln_prefix + "bias"], cfg.layer_norm_eps, y32=y32, y16=y16,
                          mean=mean, rstd=rstd)
        rec.update(s=s32, mean=mean, rstd=rstd)
        return y32, y16, rec

    def _layer_pass(self, i, pass_id, xq32, Q, K, V, training, tag=None):
        """One DebertaV2Layer pass (deberta.py:351-375) given the projected Q/K/V views."""
        cfg, P, W, ws, AD = self.cfg, self.P, self.W, self.ws, self.AD
        tws = self.sws if training else ws
        B, L, F, Lt, T = self.dims
        H, h, I = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        tag = tag if tag is not None else str(pass_id)
        sfx = f".{tag}" if training else ""
        p = f"{self.layer_prefix}{i}."
        pq, pk, pkr = self._rel_tables(i, pass_id, L) if self.use_rel else (None, None, None)
        ctx16 = tws.get("ctx" + sfx, (T, H), AD)
        lse = tws.get("lse" + sfx, (B, h, L), F32)
        att_drop = self.drop(cfg.attention_probs_dropout_prob, pass_id * 8 + SITE_PROBS)
        s2 = None
        if self.precise:
            # fp32-class mode: pq holds the [2S, 2H] projections; the backward recomputes the probabilities
            lib.attn_fwd_f32(Q, K, V, pq, self.bucket_of_rel, self.mask_u8, ctx16, lse, B=B, L=L, h=h,
                             scale=self.attn_scale, drop=att_drop, key_mask_only=not self.use_rel)
        else:
            s2 = tws.get("s2" + sfx, (B, h, L, (L + 127) // 128 * 128), lib.SCORE_DTYPE) if training else None
            lib.attn_fwd(Q, K, V, pk, pq, self.mask_u8, ctx16, lse, B=B, L=L, h=h, scale=self.attn_scale,
                         dropout_p=att_drop.p, seed=att_drop.seed, offset=att_drop.offset, s_out=s2,
                         key_mask_only=not self.use_rel)
        a32, a16, rec1 = self._sublayer_out("att", i, pass_id, ctx16, W[f"o.w.{i}"], P[p + "attention.output.dense.bias"],
                                            p + "attention.output.adapter.", p + "attention.output.LayerNorm.",
                                            xq32, SITE_AD1, SITE_HID1, training, tag)
        f16 = ws.get("ffn.f", (T, I), AD)
        h1 = tws.get("ffn.h1" + sfx, (T, I), AD) if training else None
        lib.gemm(a16, W[f"ffn1.w.{i}"], f16, bias=P[p + "intermediate.dense.bias"], act=lib.FB_ACT_GELU, D_pre=h1)
        y32, y16, rec2 = self._sublayer_out("ffn", i, pass_id, f16, W[f"ffn2.w.{i}"], P[p + "output.dense.bias"],
                                            p + "output.adapter.", p + "output.LayerNorm.", a32, SITE_AD2, SITE_HID2,
                                            training, tag)
        rec = None
        if training:
            rec = dict(layer=i, pass_id=pass_id, Q=Q, K=K, V=V, pq=pq, pk=pk, pkr=pkr, ctx=ctx16, lse=lse, att=rec1,
                       ffn=rec2, h1=h1, s2=s2)
        return y32, y16, rec

    # ------------------------------------------------------------------ head
    def head_logits(self, rows16, n_rows, mlm, out=None, scratch=False):
        """logits = LN(gelu(dense(x))) . E^T + bias for the given rows (deberta.py:1544-1558).
        ``scratch`` uses separate buffers and leaves the state saved for the backward untouched."""
        cfg, P, W, AD = self.cfg, self.P, self.W, self.AD
        ws = self.ws if scratch else self.sws
        H = cfg.hidden_size
        hn = self.hn
        pfx = "headx." if scratch else "head."
        hd32 = ws.rows(pfx + "hd32", n_rows, H, F32)
        pre16 = ws.rows(pfx + "pre16", n_rows, H, AD)
        lib.gemm(rows16, W["head.w"], hd32, bias=P[hn["dense_b"]], act=lib.FB_ACT_GELU, D_pre=pre16)
        t16 = ws.rows(pfx + "t16", n_rows, H, AD)
        mean, rstd = ws.rows(pfx + "mean", n_rows, 1, F32), ws.rows(pfx + "rstd", n_rows, 1, F32)
        lib.layernorm_fwd(hd32, P[hn["ln_w"]], P[hn["ln_b"]], cfg.layer_norm_eps, y16=t16,
                          mean=mean, rstd=rstd)
        if cfg.n_ans and not mlm:
            E, bias, V = W["ans16"], P["answer_bias"], W["ans_n"]
        else:
            E, bias, V = W["emb16"], P[hn["bias"]], cfg.vocab_size
        if out is None:
            out = ws.rows(pfx + "logits", n_rows, _ceil8(V), F32)
        lib.gemm(t16, E, out, bias=bias, N=V)
        if scratch:
            return out, V
        self.head_saved = dict(hd32=hd32, pre16=pre16, t16=t16, mean=mean, rstd=rstd, E=E, V=V, n_rows=n_rows)
        if self.cur is not None:
            self.cur.head = self.head_saved
        return out, V

    def head_mlm_loss(self, sel, labels_sel, n_sel, want_grad):
        """MLM loss over the labelled positions only (CrossEntropyLoss ignores -100, deberta.py:1481-1488): the
        head runs on n_sel rows instead of all B*L and the [B, L, V] logits are never materialised."""
        cfg, ws = self.cfg, self.sws
        H = cfg.hidden_size
        rows16 = ws.rows("head.rows16", n_sel, H, self.AD)
        lib.move_rows(self.o2_16, rows16, idx_in=sel)
        logits, V = self.head_logits(rows16, n_sel, mlm=True)
        loss_sum = ws.get("head.loss", (1,), F32, zero=True)
        dlog16 = ws.rows("head.dlog16", n_sel, _ceil8(V), self.AD) if want_grad else None
        lib.cross_entropy(logits, V, labels_sel, loss_sum, 1.0 / n_sel, dlog16, 1.0 / n_sel)
        self.head_saved.update(sel=sel, dlog16=dlog16, mode="mlm")
        return loss_sum[0].clone()   # never a view of the reused workspace (a kept loss must not change later)

    # ------------------------------------------------------------------ backward
    def pending_forwards(self):
        """Train-mode forwards whose backward has not run yet."""
        return self._n_slots - len(self._free_slots)

    def _grads_ready(self, group):
        """Every gradient of ``group`` (a layer index) has been ISSUED: hands the data-parallel hook one event per
        stream that contributes to them, so the layer's slice of the flat bucket can be all-reduced while the
        backward of the earlier layers is still running."""
        cb = self._on_grads_ready
        if cb is None:
            return
        evs = [torch.cuda.current_stream().record_event()]
        if self._side is not None:
            evs.append(self._side.record_event())
        cb(group, evs)

    def backward(self, gviews, scale=None, dlogits=None, state=None, on_grads_ready=None):
        """Selective backward of ONE forward (``state``; default: the latest): accumulates into ``gviews``
        (name -> fp32 view of the flat gradient bucket).  ``on_grads_ready(layer, events)`` is called as soon as the
        gradients of an encoder layer are complete on the device timeline described by ``events``."""
        self._on_grads_ready = on_grads_ready
        st = state if state is not None else self.cur
        if st is None or st.sv is None or st.released:
            raise lib.FbError("backward() without a training-mode forward (or its state was already consumed)")
        # the engine attributes the backward helpers read are those of THAT forward
        self.saved, self.head_saved, self.dims, self.mask_u8 = st.sv, st.head, st.dims, st.mask_u8
        self.cur_seed, self.training = st.seed, True
        self.g = gviews
        try:
            self._backward_impl(scale, dlogits)
        finally:
            self.saved = self.head_saved = None
            self._on_grads_ready = None
            st.release()

    def _backward_impl(self, scale, dlogits):
        if not self.dw_side_stream:
            self._side = None
            dy32 = self._bwd_head(scale, dlogits)
            self._bwd_encoder(dy32)
            return
        # The adapter weight gradients (dW = dY^T X, bias column sums: ~100 us of small split-K GEMMs per layer pass)
        # feed nothing downstream.  They are issued on a LOW-priority stream while the backward chain runs on a
        # high-priority one, so their CTAs are only scheduled on SMs the chain leaves idle: the partial last wave of
        # the attention kernels (768 CTAs on 148 SMs) and every kernel's tail.  Their operands get per-pass buffers,
        # so the chain never waits for them; the streams join before the optimizer.
        cur = torch.cuda.current_stream()
        hi, lo = _priority_streams(self.device)
        hi.wait_stream(cur)
        lo.wait_stream(cur)
        self._side = lo
        # no dependent launches while the side stream has work: kernels parked in griddepcontrol.wait would occupy
        # the SM slots the low-priority CTAs are meant to fill (72.3 vs 70.8 ms/step)
        pdl = lib.set_pdl(False)
        try:
            with torch.cuda.stream(hi):
                dy32 = self._bwd_head(scale, dlogits)
                self._bwd_encoder(dy32)
        finally:
            lib.set_pdl(pdl)
        self._side = None
        cur.wait_stream(hi)
        cur.wait_stream(lo)

    def _bwd_head(self, scale, dlogits):
        cfg, P, W, ws, hs = self.cfg, self.P, self.W, self.ws, self.head_saved
        B, L, F, Lt, T = self.dims
        H = cfg.hidden_size
        hn = self.hn
        n = hs["n_rows"]
        # ---- head: logits -> t -> LN -> gelu -> dense
        dt32 = ws.rows("bw.dt32", n, H, F32)
        if dlogits is None:
            # (K = V = 128100 with only 13 x 8 output tiles at config 2: 30 % of the SMs idle.  Splitting K seven ways
            # measured 731 -> 590 us, but its fp32 atomics sit at the HEAD of the backward chain: the 1e-7 summation-order
            # noise is amplified by every bf16 operand rounding downstream and the step stops being reproducible run to
            # run at the 1e-3 level -- not worth 0.2 % of the step.  All other split-K / atomic reductions of the
            # backward end in terminal gradient buffers.)
            lib.gemm(hs["dlog16"], hs["E"], dt32, b_mn_major=True, K=hs["V"], N=H)
            if torch.is_tensor(scale):
                dt32.mul_(scale.to(dt32.dtype))         # d(loss) from autograd: a device scalar, never read on the host
            elif scale is not None and scale != 1.0:
                dt32.mul_(scale)
        else:
            nans = hs["V"]
            d16 = ws.get("bw.dlog_ans", (n, _ceil8(nans)), self.AD, zero=True)
            d16[:, :nans] = dlogits.reshape(n, nans).to(self.AD)
            lib.gemm(d16, hs["E"], dt32, b_mn_major=True, N=H)
        dhd32 = ws.rows("bw.dhd32", n, H, F32)
        lib.layernorm_bwd(dt32, hs["hd32"], hs["mean"], hs["rstd"], P[hn["ln_w"]], dx32=dhd32,
                          dgamma=self._g(hn["ln_w"]), dbeta=self._g(hn["ln_b"]))
        dpre16 = ws.rows("bw.dpre16", n, H, self.AD)
        lib.gelu_bwd(dhd32, hs["pre16"], dpre16)
        dy32 = ws.get("bw.dy32.a", (T, H), F32)
        if dlogits is None:
            drows32 = ws.rows("bw.drows32", n, H, F32)
            lib.gemm(dpre16, W["head.w"], drows32, b_mn_major=True)
            dy32.zero_()
            lib.move_rows(drows32, dy32, idx_out=hs["sel"])
        else:
            lib.gemm(dpre16, W["head.w"], dy32, b_mn_major=True)
        return dy32

    def _bwd_encoder(self, dy32):
        cfg, P, W, ws, sv = self.cfg, self.P, self.W, self.ws, self.saved
        B, L, F, Lt, T = self.dims
        H = cfg.hidden_size
        self._dR32 = ws.get("bw.dR32", (2 * cfg.position_buckets, H), F32, zero=True)
        passes = sv["passes"]
        # batched positional-gradient chain: the passes are layers 0 .. NL-1 in order, then repeats of the last layer
        self._rel_bwd_batched = (self.use_rel and self.rel_batched and not self.precise and self.fused_attn_bwd and
                                 [r["layer"] for r in passes[:cfg.num_hidden_layers]] ==
                                 list(range(min(cfg.num_hidden_layers, len(passes)))) and
                                 all(r["pass_id"] == k for k, r in enumerate(passes)))
        NL = cfg.num_hidden_layers
        last = NL - 1
        # ---- enhanced mask decoder passes (reverse): shared K/V projection of h[-2]
        dq32, dkv_b = self._bwd_pass(passes[-1], dy32, emd=True, slot=0)
        dz32, dkv_a = self._bwd_pass(passes[-2], dq32, emd=True, slot=1)
        self._grads_ready(last)
        dx32 = ws.get("bw.dx32.e", (T, H), F32)
        wkv = W[f"qkv.w.{last}"][H:]
        lib.gemm(dkv_b, wkv, dx32, b_mn_major=True, resid=dz32)
        lib.gemm(dkv_a, wkv, dx32, b_mn_major=True, accumulate=True)
        # ---- ordinary passes, last to first
        c = "deberta.encoder."
        dcol16 = None
        for rec in reversed(passes[:-2]):
            if rec["layer"] == 0:
                # y0' = LN_conv(y0 + gelu(dropout(conv(x0) * m))) * m
                ds_c = ws.get("bw.ds_c", (T, H), F32)
                lib.layernorm_bwd(dx32, sv["conv_s"], sv["conv_mean"], sv["conv_rstd"], P[c + "conv.LayerNorm.weight"],
                                  rowmask=self.mask_u8, dx32=ds_c, dgamma=self._g(c + "conv.LayerNorm.weight"),
                                  dbeta=self._g(c + "conv.LayerNorm.bias"))
                dconv16 = ws.get("bw.dconv16", (T, H), self.AD)
                lib.gelu_bwd(ds_c, sv["conv_pre"], dconv16, rowmask=self.mask_u8,
                             dropout=self.drop(cfg.hidden_dropout_prob, OFF_CONV))
                dcol16 = ws.get("bw.dcol16", (T, 3 * H), self.AD)
                lib.gemm(dconv16, W["conv.w"], dcol16, b_mn_major=True)
                dx32 = ds_c
            dx32 = self._bwd_pass(rec, dx32, emd=False, slot=rec["pass_id"] % 2)
            self._grads_ready(rec["layer"])
        if dcol16 is not None:
            lib.conv_col2im(dcol16, B, L, H, dx32, accumulate=True)
        self._bwd_embeddings(dx32, "deberta.embeddings.", rowmask=self.mask_u8)
        # ---- relative-embedding LayerNorm (trainable under ft_ln): gamma/beta only, the table is frozen
        if self._rel_bwd_batched:
            if self._side is not None and self.rel_side_stream:
                with torch.cuda.stream(self._side):
                    self._bwd_rel_all()
            else:
                self._bwd_rel_all()
        if self._side is not None and self.rel_side_stream:
            torch.cuda.current_stream().wait_stream(self._side)   # dR has been accumulated on the side stream
        lib.layernorm_bwd(self._dR32, P[c + "rel_embeddings.weight"], sv["rel_mean"], sv["rel_rstd"],
                          P[c + "LayerNorm.weight"], dgamma=self._g(c + "LayerNorm.weight"),
                          dbeta=self._g(c + "LayerNorm.bias"))

    def _bwd_embeddings(self, dx32, e, rowmask):
        """dropout -> (* mask) -> LayerNorm -> [linear_video ; word_emb (frozen)]"""
        cfg, P, ws, sv = self.cfg, self.P, self.ws, self.saved
        B, L, F, Lt, T = self.dims
        H = cfg.hidden_size
        de16 = ws.get("bw.de16", (T, H), self.AD)
        lib.layernorm_bwd(dx32, sv["e32"], sv["emb_mean"], sv["emb_rstd"], P[e + "LayerNorm.weight"],
                          rowmask=rowmask, out_drop=self.drop(cfg.hidden_dropout_prob, OFF_EMB), dx16=de16,
                          dgamma=self._g(e + "LayerNorm.weight"), dbeta=self._g(e + "LayerNorm.bias"))
        if sv["use_video"]:
            dv16 = ws.get("bw.dv16", (B * F, H), self.AD)
            idx = (torch.arange(B, device=self.device)[:, None] * L + torch.arange(F, device=self.device)[None, :])
            lib.move_rows(de16, dv16, idx_in=idx.reshape(-1).contiguous())
            v16 = sv["video16"]
            lib.gemm(dv16, v16, self._g(e + "linear_video.weight"), a_mn_major=True, b_mn_major=True,
                     N=cfg.features_dim, accumulate=True)
            lib.colsum_bf16(dv16, self._g(e + "linear_video.bias"))

    def _g(self, name):
        """fp32 gradient view of a trainable tensor, or None when it is frozen (kernels skip null outputs)."""
        return self.g.get(name)

    def _bwd_sublayer(self, name, rec, dy32, ad_prefix, ln_prefix, W_dense, pass_id, site_ad, site_hid, slot):
        """Backward of LN(dropout(adapter(dense(x))) + resid): returns (ds32 = grad of the residual branch,
        dxin16 = grad of the dense input)."""
        cfg, P, ws = self.cfg, self.P, self.ws
        T, H = dy32.shape
        r = cfg.r_attn if name == "att" else cfg.r_ff
        ds32 = ws.get(f"bw.{name}.ds32.{slot}", (T, H), F32)
        r_ = cfg.r_attn if name == "att" else cfg.r_ff
        # per-pass operand buffers when the weight gradients run asynchronously (never overwritten before they are read)
        gw16 = ws.get(f"bw.{name}.gw16" + (f".{pass_id}" if (self._side is not None and r_) else ""), (T, H), self.AD)
        lib.layernorm_bwd(dy32, rec["s"], rec["mean"], rec["rstd"], P[ln_prefix + "weight"],
                          in_drop=self.drop(cfg.hidden_dropout_prob, pass_id * 8 + site_hid), dx32=ds32, dx16=gw16,
                          dgamma=self._g(ln_prefix + "weight"), dbeta=self._g(ln_prefix + "bias"),
                          dx16_colsum=self._g(ad_prefix + "up.bias") if r else None)
        if r:
            u16, z16 = rec["u"], rec["z"]
            wu, wd = self.t16(ad_prefix + "up.weight"), self.t16(ad_prefix + "down.weight")
            ad = self.drop(cfg.adapter_dropout, pass_id * 8 + site_ad)
            inv_keep = 1.0 / (1.0 - ad.p) if ad.p > 0 else 1.0
            side = self._side
            dz16 = ws.get(f"bw.{name}.dz16" + (f".{pass_id}" if side is not None else ""), (T, r), self.AD)
            # d(zd) = gw . Wu ; through dropout and ReLU: zd > 0 <=> kept and pre-activation > 0
            lib.gemm(gw16, wu, dz16, b_mn_major=True, alpha=inv_keep, aux=z16, aux_op=lib.FB_AUX_MUL_RELU_MASK)
            sk = max(1, min(32, T // 512))

            def weight_grads():
                lib.gemm(gw16, z16, self._g(ad_prefix + "up.weight"), a_mn_major=True, b_mn_major=True, split_k=sk,
                         accumulate=True)
                lib.gemm(dz16, u16, self._g(ad_prefix + "down.weight"), a_mn_major=True, b_mn_major=True, split_k=sk,
                         accumulate=True)
                lib.colsum_bf16(dz16, self._g(ad_prefix + "down.bias"))

            if side is None:
                weight_grads()
            else:
                ev = torch.cuda.current_stream().record_event()
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    weight_grads()
            du16 = ws.get(f"bw.{name}.du16", (T, H), self.AD)
            lib.gemm(dz16, wd, du16, b_mn_major=True, aux=gw16, aux_op=lib.FB_AUX_ADD)
        else:
            du16 = gw16
        return ds32, du16

    def _bwd_pass(self, rec, dy32, emd, slot):
        cfg, P, W, ws = self.cfg, self.P, self.W, self.ws
        B, L, F, Lt, T = self.dims
        H, h, I = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        i, pass_id = rec["layer"], rec["pass_id"]
        p = f"{self.layer_prefix}{i}."
        # ---- FFN sub-layer
        ds2, du2 = self._bwd_sublayer("ffn", rec["ffn"], dy32, p + "output.adapter.", p + "output.LayerNorm.",
                                      W[f"ffn2.w.{i}"], pass_id, SITE_AD2, SITE_HID2, slot)
        dh1 = ws.get("bw.dh1", (T, I), self.AD)
        lib.gemm(du2, W[f"ffn2.w.{i}"], dh1, b_mn_major=True, aux=rec["h1"], aux_op=lib.FB_AUX_MUL_GELU_GRAD)
        da32 = ws.get("bw.da32", (T, H), F32)
        lib.gemm(dh1, W[f"ffn1.w.{i}"], da32, b_mn_major=True, resid=ds2)
        # ---- attention sub-layer
        ds1, du1 = self._bwd_sublayer("att", rec["att"], da32, p + "attention.output.adapter.",
                                      p + "attention.output.LayerNorm.", W[f"o.w.{i}"], pass_id, SITE_AD1,
                                      SITE_HID1, slot)
        dctx16 = ws.get("bw.dctx16", (T, H), self.AD)
        lib.gemm(du1, W[f"o.w.{i}"], dctx16, b_mn_major=True)
        dqkv16 = ws.get(f"bw.dqkv16.{slot if emd else 0}", (T, 3 * H), self.AD)
        if self.fused_attn_bwd:
            self._attn_bwd_fused(rec, dctx16, dqkv16)
        else:
            dqkv32 = self._attn_bwd(rec, dctx16)
            lib.cast_bf16(dqkv32, dqkv16)
        dxq = ws.get(f"bw.dxq.{slot}", (T, H), F32)
        if not emd:
            lib.gemm(dqkv16, W[f"qkv.w.{i}"], dxq, b_mn_major=True, resid=ds1)
            return dxq
        lib.gemm(dqkv16[:, :H], W[f"qkv.w.{i}"][:H], dxq, b_mn_major=True, resid=ds1)
        return dxq, dqkv16[:, H:]

    def _attn_bwd_fused(self, rec, dctx16, dqkv16):
        """Attention backward as one fused tcgen05 kernel per pass (fb_attn_bwd) + the band-gradient kernel of the
        positional projections (fb_attn_bwd_rel); dQ | dK | dV land in ``dqkv16`` as bf16."""
        cfg, W, ws = self.cfg, self.W, self.ws
        B, L, F, Lt, T = self.dims
        H, h = cfg.hidden_size, cfg.num_attention_heads
        i, pass_id = rec["layer"], rec["pass_id"]
        Q, K, V = rec["Q"], rec["K"], rec["V"]
        lp = (L + 127) // 128 * 128
        rel = self.use_rel
        ad = self.drop(cfg.attention_probs_dropout_prob, pass_id * 8 + SITE_PROBS)
        if self.precise:
            # fp32-class mode: fp32 CUDA-core backward straight into the bucket rows of d(Linear_{q|k}(dropout(R)))
            dpqk32 = ws.get("bw.dpqk32", (2 * cfg.position_buckets, 2 * H), F32, zero=True) if rel else None
            lib.attn_bwd_f32(Q, K, V, rec["pq"], self.bucket_of_rel, self.mask_u8, rec["ctx"], rec["lse"], dctx16,
                             dqkv16[:, :H], dqkv16[:, H:2 * H], dqkv16[:, 2 * H:], dpqk32, B=B, L=L, h=h,
                             scale=self.attn_scale, drop=ad, key_mask_only=not rel)
            if rel:
                self._bwd_rel_from_buckets(dpqk32, i, pass_id)
            return
        # The gradients of the positional projections (fb_attn_bwd_rel and the chain down to dR) feed nothing on the
        # token path: with the side stream they run at low priority from a per-pass copy of dS'.
        side = self._side if self.rel_side_stream else None
        ds = ws.get("bw.ds" + (f".{pass_id}" if side is not None else ""), (B, h, L, lp), BF16) if rel else None
        lib.attn_bwd(Q, K, V, rec["ctx"], dctx16, rec["s2"], rec["lse"], rec["pkr"], rec["pq"],
                     dqkv16[:, :H], dqkv16[:, H:2 * H], dqkv16[:, 2 * H:], ds_out=ds, B=B, L=L, h=h,
                     scale=self.attn_scale, dropout_p=ad.p, seed=ad.seed, offset=ad.offset)
        if not rel:
            return
        rel_rows, _ = lib.rel_table_geometry(L)
        batched = self._rel_bwd_batched
        if batched:
            # every pass keeps its band gradients; the chain down to dR runs ONCE for all passes (_bwd_rel_all)
            NP = len(self.saved["passes"])
            gpk = ws.get("bw.gpk.all", (NP, h, rel_rows, 64), F32)[pass_id]
            gpq = ws.get("bw.gpq.all", (NP, h, rel_rows, 64), F32)[pass_id]
        else:
            gpk = ws.get("bw.gpk", (h, rel_rows, 64), F32)
            gpq = ws.get("bw.gpq", (h, rel_rows, 64), F32)

        def positional_grads():
            lib.attn_bwd_rel(ds, Q, K, gpk, gpq, B=B, L=L, h=h, scale=self.attn_scale,
                             units_per_cta=self.rel_side_units if side is not None else 0)
            if not batched:
                self._bwd_rel_tables(gpq, gpk, i, pass_id, L, h)

        if side is None:
            positional_grads()
        else:
            ev = torch.cuda.current_stream().record_event()
            side.wait_event(ev)
            with torch.cuda.stream(side):
                positional_grads()

    def _bwd_rel_all(self):
        """d(per-relative-position tables) of ALL passes -> d(Linear_{q|k}(dropout(R))) -> dR in five launches (fold
        onto the bucket rows, cast, one batched GEMM over the layers + the trailing passes, dropout backward summed
        over the passes) instead of five per pass; the mirror image of _rel_tables_all."""
        cfg, W, ws = self.cfg, self.W, self.ws
        B, L, F, Lt, T = self.dims
        H, h, NL = cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers
        S2 = 2 * cfg.position_buckets
        passes = self.saved["passes"]
        NP = len(passes)
        rel_rows, _ = lib.rel_table_geometry(L)
        gpk = ws.get("bw.gpk.all", (NP, h, rel_rows, 64), F32)
        gpq = ws.get("bw.gpq.all", (NP, h, rel_rows, 64), F32)
        dpqk32 = ws.get("bw.dpqk32.all", (NP, S2, 2 * H), F32, zero=True)
        lib.rel_to_bucket_batched(gpq, gpk, self.bucket_of_rel, L, h, dpqk32)
        dpqk16 = ws.get("bw.dpqk16.all", (NP, S2, 2 * H), BF16)
        lib.cast_bf16(dpqk32, dpqk16)
        dRd32 = ws.get("bw.dRd32.all", (NP, S2, H), F32)
        wall = W["qkv.w.all"]
        nlead = min(NL, NP)
        lib.gemm(dpqk16[0], wall[0][:2 * H], dRd32[0], b_mn_major=True,
                 batch={"outer": (nlead, S2 * 2 * H, 3 * H * H, S2 * H)})
        for q in range(nlead, NP):
            lib.gemm(dpqk16[q], wall[passes[q]["layer"]][:2 * H], dRd32[q], b_mn_major=True)
        lib.dropout_batched(dRd32, self.drop(cfg.hidden_dropout_prob, SITE_POS), 8, out32_accum=self._dR32)

    def _bwd_rel_tables(self, gpq, gpk, i, pass_id, L, h):
        """d(per-relative-position tables) -> d(Linear_{q|k}(dropout(R))) -> accumulate dR (deberta.py:847-853)."""
        cfg, ws = self.cfg, self.ws
        dpqk32 = ws.get("bw.dpqk32", (2 * cfg.position_buckets, 2 * cfg.hidden_size), F32, zero=True)
        lib.rel_to_bucket(gpq, gpk, self.bucket_of_rel, L, h, dpqk32)
        self._bwd_rel_from_buckets(dpqk32, i, pass_id)

    def _bwd_rel_from_buckets(self, dpqk32, i, pass_id):
        cfg, W, ws = self.cfg, self.W, self.ws
        H = cfg.hidden_size
        S2 = 2 * cfg.position_buckets
        dpqk16 = ws.get("bw.dpqk16", (S2, 2 * H), self.AD)
        lib.cast_bf16(dpqk32, dpqk16)
        dRd32 = ws.get("bw.dRd32", (S2, H), F32)
        lib.gemm(dpqk16, W[f"qkv.w.{i}"][:2 * H], dRd32, b_mn_major=True)
        lib.dropout_apply(dRd32, self.drop(cfg.hidden_dropout_prob, pass_id * 8 + SITE_POS), out32=self._dR32,
                          accumulate=True)

    def _attn_bwd(self, rec, dctx16):
        """Attention backward from GEMMs + the staging kernels of fb_attention.cu (see fb_attn_bwd_prep)."""
        cfg, W, ws = self.cfg, self.W, self.ws
        B, L, F, Lt, T = self.dims
        H, h = cfg.hidden_size, cfg.num_attention_heads
        i, pass_id = rec["layer"], rec["pass_id"]
        Q, K, V = rec["Q"], rec["K"], rec["V"]
        lp = (L + 127) // 128 * 128
        rel_rows, rel_off = lib.rel_table_geometry(L)
        scale = self.attn_scale
        rel = self.use_rel
        # dP = dO . V^T
        dp = ws.get("bw.dp", (B, h, L, lp), BF16)
        lib.gemm(dctx16, V, dp.view(-1, lp), M=L, N=L, K=64,
                 batch={"outer": (B, L * H, L * V.stride(0), h * L * lp), "inner": (h, 64, 64, L * lp)})
        ds = ws.get("bw.ds", (B, h, L, lp), BF16)
        pd = ws.get("bw.pd", (B, h, L, lp), BF16)
        z1 = ws.get("bw.z1", (h, B * L, rel_rows), BF16) if rel else None
        z2 = ws.get("bw.z2", (h, B * L, rel_rows), BF16) if rel else None
        ad = self.drop(cfg.attention_probs_dropout_prob, pass_id * 8 + SITE_PROBS)
        lib.attn_bwd_prep(rec["s2"], rec["lse"], dp, rec["ctx"], dctx16, ds, pd, z1, z2, B=B, L=L, h=h, scale=scale,
                          dropout_p=ad.p, seed=ad.seed, offset=ad.offset)
        dqkv32 = ws.get("bw.dqkv32", (T, 3 * H), F32)
        ds2d, pd2d = ds.view(-1, lp), pd.view(-1, lp)
        if rel:
            z1_2d, z2_2d = z1.view(-1, rel_rows), z2.view(-1, rel_rows)
            pk2d, pq2d = rec["pk"].view(-1, 64), rec["pq"].view(-1, 64)
        dq32, dk32, dv32 = dqkv32[:, :H], dqkv32[:, H:2 * H], dqkv32[:, 2 * H:]
        sq, sk_, svv = Q.stride(0), K.stride(0), V.stride(0)
        o3 = 3 * H
        # dV = Pd^T . dO
        lib.gemm(pd2d, dctx16, dv32, a_mn_major=True, b_mn_major=True, M=L, N=64, K=L,
                 batch={"outer": (B, h * L * lp, L * H, L * o3), "inner": (h, L * lp, 64, 64)})
        # dQ = dS . K + Z1 . PKrel
        lib.gemm(ds2d, K, dq32, b_mn_major=True, M=L, N=64, K=L,
                 batch={"outer": (B, h * L * lp, L * sk_, L * o3), "inner": (h, L * lp, 64, 64)})
        if rel:
            lib.gemm(z1_2d, pk2d, dq32, b_mn_major=True, M=L, N=64, K=rel_rows, accumulate=True,
                     batch={"outer": (B, L * rel_rows, 0, L * o3), "inner": (h, B * L * rel_rows, rel_rows * 64, 64)})
        # dK = dS^T . Q + Z2 . PQrel
        lib.gemm(ds2d, Q, dk32, a_mn_major=True, b_mn_major=True, M=L, N=64, K=L,
                 batch={"outer": (B, h * L * lp, L * sq, L * o3), "inner": (h, L * lp, 64, 64)})
        if not rel:
            return dqkv32
        lib.gemm(z2_2d, pq2d, dk32, b_mn_major=True, M=L, N=64, K=rel_rows, accumulate=True,
                 batch={"outer": (B, L * rel_rows, 0, L * o3), "inner": (h, B * L * rel_rows, rel_rows * 64, 64)})
        # positional projections: dPKrel = Z1^T . Q, dPQrel = Z2^T . K (sum over batch inside the K dimension)
        gpk = ws.get("bw.gpk", (h, rel_rows, 64), F32)
        gpq = ws.get("bw.gpq", (h, rel_rows, 64), F32)
        lib.gemm(z1_2d, Q, gpk.view(-1, 64), a_mn_major=True, b_mn_major=True, M=rel_rows, N=64, K=B * L,
                 batch={"outer": (h, B * L * rel_rows, 64, rel_rows * 64)})
        lib.gemm(z2_2d, K, gpq.view(-1, 64), a_mn_major=True, b_mn_major=True, M=rel_rows, N=64, K=B * L,
                 batch={"outer": (h, B * L * rel_rows, 64, rel_rows * 64)})
        self._bwd_rel_tables(gpq, gpk, i, pass_id, L, h)
        return dqkv32


class BertEngine(DebertaEngine):
    """BERT variant of the path (reference model/bert.py, BASELINE.json configs[4]): visual prefix + absolute
    position / token-type embeddings (:246-278), plain attention with an additive key mask (:141-191), post-LN
    blocks without adapters (:281-292, 349-360), MLM head on ``cls.predictions.decoder`` (:74-96).  Same kernels
    as the DeBERTa path: fb_attn_fwd in plain mode (no relative tables), the same GEMM / LayerNorm schedule."""

    LAYER_PREFIX = "bert.encoder.layer."

    def __init__(self, cfg, params, device, precision="bf16"):
        super().__init__(cfg, params, device, precision)
        self.use_rel = False
        self.attn_scale = 1.0 / math.sqrt(64)
        self.layer_prefix = self.LAYER_PREFIX
        self.hn = dict(dense_b="cls.predictions.transform.dense.bias", ln_w="cls.predictions.transform.LayerNorm.weight",
                       ln_b="cls.predictions.transform.LayerNorm.bias", bias="cls.predictions.bias")

    def prepare_frozen(self):
        cfg, P, W = self.cfg, self.P, self.W
        BF16 = self.AD
        with torch.no_grad():
            for i in range(cfg.num_hidden_layers):
                p = f"bert.encoder.layer.{i}."
                a = p + "attention.self."
                W[f"qkv.w.{i}"] = torch.cat([P[a + "query.weight"], P[a + "key.weight"], P[a + "value.weight"]],
                                            0).to(BF16).contiguous()
                W[f"qkv.b.{i}"] = torch.cat([P[a + "query.bias"], P[a + "key.bias"], P[a + "value.bias"]],
                                            0).float().contiguous()
                W[f"o.w.{i}"] = P[p + "attention.output.dense.weight"].to(BF16).contiguous()
                W[f"ffn1.w.{i}"] = P[p + "intermediate.dense.weight"].to(BF16).contiguous()
                W[f"ffn2.w.{i}"] = P[p + "output.dense.weight"].to(BF16).contiguous()
            W["head.w"] = P["cls.predictions.transform.dense.weight"].to(BF16).contiguous()
            W["emb16"] = P["cls.predictions.decoder.weight"].to(BF16).contiguous()
            if cfg.n_ans:
                self.refresh_answer_embeddings()

    def forward_encoder(self, video, video_mask, input_ids, attention_mask, training, video_repeat=1):
        cfg, P, W, ws = self.cfg, self.P, self.W, self.ws
        st = self.begin_forward(training)
        sws, AD = self.sws, self.AD
        dev = self.device
        H, NL = cfg.hidden_size, cfg.num_hidden_layers
        B, Lt = input_ids.shape
        use_video = bool(cfg.features_dim) and video is not None
        F = video.shape[1] if use_video else 0
        L = F + Lt
        T = B * L
        if L > cfg.max_position_embeddings or L > 512:
            raise ValueError(f"sequence length {L} exceeds max_position_embeddings / kernel limit 512")
        self.dims = st.dims = (B, L, F, Lt, T)
        e = "bert.embeddings."
        sv = {"passes": []} if training else None
        if attention_mask is None:
            attention_mask = torch.ones(B, Lt, device=dev, dtype=torch.long)
        if use_video:
            if video_mask is None:
                video_mask = torch.ones(video.shape[0], F, device=dev, dtype=torch.long)
            if video_repeat > 1:
                if video.shape[0] * video_repeat != B:
                    raise ValueError(f"video_repeat={video_repeat}: {video.shape[0]} clips for {B} text rows")
                video_mask = video_mask.to(dev).repeat_interleave(video_repeat, 0)
            m = torch.cat([video_mask.to(dev), attention_mask.to(dev)], 1)
        else:
            m = attention_mask.to(dev)
        self.mask_u8 = st.mask_u8 = (m != 0).to(torch.uint8).contiguous()
        # ---- embeddings: [linear_video(video) ; word_emb[ids]] + pos + type(0) -> LN -> dropout (no mask multiply)
        g32 = ws.get("g32", (T, H), F32)
        v16, idx_out = None, None
        if use_video:
            v16, idx_out = self._video_prefix(video, B, F, L, g32, e + "linear_video.weight", e + "linear_video.bias",
                                              video_repeat)
        lib.move_rows(P[e + "word_embeddings.weight"], g32, idx_in=input_ids.to(dev).reshape(-1).contiguous(),
                      idx_out=idx_out)
        pos_type = (P[e + "position_embeddings.weight"][:L] + P[e + "token_type_embeddings.weight"][0]).contiguous()
        e32 = sws.get("e32", (T, H), F32)
        lib.add_rowbcast(g32, pos_type, B, L, H, e32)
        x32, x16 = ws.get("x32.0", (T, H), F32), ws.get("x16.0", (T, H), AD)
        emb_mean, emb_rstd = sws.get("emb.mean", (T,), F32), sws.get("emb.rstd", (T,), F32)
        lib.layernorm_fwd(e32, P[e + "LayerNorm.weight"], P[e + "LayerNorm.bias"], cfg.layer_norm_eps,
                          out_drop=self.drop(cfg.hidden_dropout_prob, OFF_EMB), y32=x32, y16=x16, mean=emb_mean,
                          rstd=emb_rstd)
        for i in range(NL):
            qkv = sws.get(f"qkv.{i}", (T, 3 * H), AD) if training else ws.get("qkv", (T, 3 * H), AD)
            lib.gemm(x16, W[f"qkv.w.{i}"], qkv, bias=W[f"qkv.b.{i}"])
            x32, x16, rec = self._layer_pass(i, i, x32, qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], training)
            if training:
                rec["qkv"] = qkv
                sv["passes"].append(rec)
        if training:
            sv.update(e32=e32, emb_mean=emb_mean, emb_rstd=emb_rstd, video16=v16, use_video=use_video)
            self.saved = st.sv = sv
        self.o2_32, self.o2_16 = x32, x16
        st.o2 = (x32, x16)
        return x32, x16

    def _bwd_encoder(self, dy32):
        sv = self.saved
        dx32 = dy32
        for rec in reversed(sv["passes"]):
            dx32 = self._bwd_pass(rec, dx32, emd=False, slot=rec["pass_id"] % 2)
            self._grads_ready(rec["layer"])
        self._bwd_embeddings(dx32, "bert.embeddings.", rowmask=None)
