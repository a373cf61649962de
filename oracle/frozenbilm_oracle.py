"""CPU oracle for the FrozenBiLM masked-LM hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

A plain torch-fp32 functional restatement of the reference's algorithm
(/root/reference/model/deberta.py + model/adapter.py), every function citing the
reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path
(frozenbilm_b200/) never does and fails loudly without its CUDA extension.

Parity status: PINNED.  tests/golden/make_golden.py runs the UNMODIFIED reference
(loaded by path with version-compat shims, oracle/ref_loader.py) on the synthetic
weights/inputs of frozenbilm_b200/synth.py and commits its logits / loss / trainable
gradients as fixtures; tests/test_oracle_golden.py holds this restatement to them
(fp32, max-abs-err ~1e-6).  The reference itself ships no tests or golden vectors
(SURVEY.md section 4).

Dropout: the reference draws masks with bernoulli_ (deberta.py:161) which cannot be
reproduced; like the reference in .eval() or with p=0, the oracle is dropout-free.
``bf16_gemm_inputs=True`` rounds every contraction operand to bf16 (fp32 accumulate)
at the points the CUDA path does, which gives a tight comparison target in addition
to the exact-fp32 one.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as Fnn


# --------------------------------------------------------------------------------------
# relative-position buckets  (reference model/deberta.py:578-618)
# --------------------------------------------------------------------------------------
def make_log_bucket_position(relative_pos: np.ndarray, bucket_size: int, max_position: int) -> np.ndarray:
    """deberta.py:578-589.  |rel| <= mid keeps rel; beyond, log-spaced buckets up to bucket_size-1."""
    sign = np.sign(relative_pos)
    mid = bucket_size // 2
    abs_pos = np.where((relative_pos < mid) & (relative_pos > -mid), mid - 1, np.abs(relative_pos))
    log_pos = np.ceil(np.log(abs_pos / mid) / np.log((max_position - 1) / mid) * (mid - 1)) + mid
    return np.where(abs_pos <= mid, relative_pos, log_pos * sign).astype(np.int64)


def build_relative_position(query_size: int, key_size: int, bucket_size: int, max_position: int) -> np.ndarray:
    """deberta.py:592-618.  rel[i, j] = bucket(i - j), shape [q, k]."""
    q_ids = np.arange(0, query_size)
    k_ids = np.arange(0, key_size)
    rel = q_ids[:, None] - np.tile(k_ids, (q_ids.shape[0], 1))
    if bucket_size > 0 and max_position > 0:
        rel = make_log_bucket_position(rel, bucket_size, max_position)
    return rel[:query_size, :]


def delta_table(L: int, bucket_size: int, max_position: int) -> np.ndarray:
    """delta[i, j] = clamp(bucket(i - j) + S, 0, 2S-1): the single gather index shared by the
    c2p and p2c terms (deberta.py:872, 898 with the odd symmetry of the bucket function)."""
    S = bucket_size
    return np.clip(build_relative_position(L, L, bucket_size, max_position) + S, 0, 2 * S - 1)


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
class _Ctx:
    def __init__(self, bf16: bool):
        self.bf16 = bf16

    def r(self, t: torch.Tensor) -> torch.Tensor:
        """Round a contraction operand to bf16 (straight-through for autograd) if emulating."""
        if not self.bf16:
            return t
        return t + (t.detach().bfloat16().float() - t.detach())

    def lin(self, x, w, b=None):
        y = self.r(x) @ self.r(w).t()
        return y if b is None else y + b


def gelu(x):
    """erf GELU = transformers ACT2FN["gelu"] (deberta.py:306, 403, 1528)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps):
    return Fnn.layer_norm(x, (x.shape[-1],), w, b, eps)


def adapter(ctx: _Ctx, u, P, prefix):
    """model/adapter.py:33-45 with ln_before=ln_after=False: u + up(relu(down(u)))."""
    z = torch.relu(ctx.lin(u, P[prefix + "down.weight"], P[prefix + "down.bias"]))
    return u + ctx.lin(z, P[prefix + "up.weight"], P[prefix + "up.bias"])


def heads(x, h):
    """transpose_for_scores, deberta.py:712-715: [B, L, H] -> [B, h, L, d]."""
    B, L, H = x.shape
    return x.view(B, L, h, H // h).permute(0, 2, 1, 3)


# --------------------------------------------------------------------------------------
# one encoder layer pass  (deberta.py:254-375, 717-947)
# --------------------------------------------------------------------------------------
def disentangled_attention(ctx: _Ctx, xq, xkv, M, R, delta, P, prefix, cfg, return_probs=False):
    """deberta.py:717-818 + 820-947 (share_att_key, pos_att_type = c2p|p2c).

    s[b,h,i,j] = (Q_i.K_j + Q_i.PK[h,delta[i,j]] + K_j.PQ[h,delta[i,j]]) / sqrt(3 d)
    p = softmax_j over M==1, exactly 0 where M==0 (XSoftmax, deberta.py:100-138).
    """
    h = cfg.num_attention_heads
    d = cfg.head_dim
    Wq, bq = P[prefix + "query_proj.weight"], P[prefix + "query_proj.bias"]
    Wk, bk = P[prefix + "key_proj.weight"], P[prefix + "key_proj.bias"]
    Wv, bv = P[prefix + "value_proj.weight"], P[prefix + "value_proj.bias"]
    Q = heads(ctx.lin(xq, Wq, bq), h)
    K = heads(ctx.lin(xkv, Wk, bk), h)
    V = heads(ctx.lin(xkv, Wv, bv), h)
    # positional projections share the content weights, bias included (deberta.py:847-853)
    PK = ctx.lin(R, Wk, bk).view(-1, h, d).permute(1, 0, 2)  # [h, 2S, d]
    PQ = ctx.lin(R, Wq, bq).view(-1, h, d).permute(1, 0, 2)
    scale = math.sqrt(d * 3)
    Qr, Kr, Vr, PKr, PQr = ctx.r(Q), ctx.r(K), ctx.r(V), ctx.r(PK), ctx.r(PQ)
    qk = Qr @ Kr.transpose(-1, -2)                              # [B,h,L,L]
    c2p_full = Qr @ PKr.transpose(-1, -2).unsqueeze(0)          # [B,h,L,2S]
    p2c_full = Kr @ PQr.transpose(-1, -2).unsqueeze(0)          # [B,h,L(j),2S]
    B, _, L, _ = qk.shape
    idx = delta.view(1, 1, L, L).expand(B, h, L, L)
    c2p = torch.gather(c2p_full, -1, idx)                       # [.., i, j] = Q_i.PK[delta[i,j]]
    # p2c[i,j] = K_j . PQ[delta[i,j]]: gather row j at column delta[i,j]  -> index [j, i] = delta[i, j]
    p2c = torch.gather(p2c_full, -1, idx.transpose(-1, -2)).transpose(-1, -2)
    s = (qk + c2p + p2c) / scale
    Mb = M.bool()
    # XSoftmax fills with -inf then zero-fills the output; a fully masked row is 0, never NaN.
    # -1e30 gives bit-identical probabilities for live rows (exp underflows to exactly 0) and a
    # finite (then zero-filled) value for dead rows, so autograd stays NaN-free like the
    # reference's hand-written backward (deberta.py:134-138).
    s = s.masked_fill(~Mb, -1e30)
    p = torch.softmax(s, -1)
    p = p.masked_fill(~Mb, 0.0)
    c = ctx.r(p) @ Vr                                           # [B,h,L,d]
    c = c.permute(0, 2, 1, 3).reshape(B, L, h * d)
    if return_probs:
        return c, p
    return c


def layer_pass(ctx: _Ctx, xq, xkv, M, R, delta, P, i, cfg):
    """DebertaV2Layer.forward with query_states=xq, hidden_states=xkv (deberta.py:351-375);
    the attention-output residual is the QUERY source (deberta.py:290-292)."""
    p = f"deberta.encoder.layer.{i}."
    eps = cfg.layer_norm_eps
    c = disentangled_attention(ctx, xq, xkv, M, R, delta, P, p + "attention.self.", cfg)
    u = ctx.lin(c, P[p + "attention.output.dense.weight"], P[p + "attention.output.dense.bias"])
    if cfg.ds_factor_attn:
        u = adapter(ctx, u, P, p + "attention.output.adapter.")
    a = layer_norm(u + xq, P[p + "attention.output.LayerNorm.weight"], P[p + "attention.output.LayerNorm.bias"], eps)
    f = gelu(ctx.lin(a, P[p + "intermediate.dense.weight"], P[p + "intermediate.dense.bias"]))
    y = ctx.lin(f, P[p + "output.dense.weight"], P[p + "output.dense.bias"])
    if cfg.ds_factor_ff:
        y = adapter(ctx, y, P, p + "output.adapter.")
    return layer_norm(y + a, P[p + "output.LayerNorm.weight"], P[p + "output.LayerNorm.bias"], eps)


# --------------------------------------------------------------------------------------
# whole forward (deberta.py:997-1058, 507-575, 395-419, 1382-1412, 1544-1558, 1452-1488)
# --------------------------------------------------------------------------------------
def forward(P, cfg, *, video, video_mask, input_ids, attention_mask, labels=None, mlm=False,
            bf16_gemm_inputs=False, return_hidden=False):
    """Returns dict(logits [B, F+Lt, V or n_ans], loss or None[, hidden=list])."""
    ctx = _Ctx(bf16_gemm_inputs)
    eps = cfg.layer_norm_eps
    e = "deberta.embeddings."
    B, Lt = input_ids.shape
    # --- embeddings (deberta.py:997-1058): video first, no abs-pos add, LN, x mask
    emb = P[e + "word_embeddings.weight"][input_ids]
    if cfg.features_dim and video is not None:
        v = ctx.lin(video, P[e + "linear_video.weight"], P[e + "linear_video.bias"])
        emb = torch.cat([v, emb], 1)
        if video_mask is None:
            video_mask = torch.ones(video.shape[:2], dtype=attention_mask.dtype)
        m = torch.cat([video_mask, attention_mask], 1)
    else:
        m = attention_mask
    L = emb.shape[1]
    mf = m.to(emb.dtype)
    pe = P[e + "position_embeddings.weight"][:L]                     # kept for the EMD only
    x0 = layer_norm(emb, P[e + "LayerNorm.weight"], P[e + "LayerNorm.bias"], eps) * mf[:, :, None]
    # --- encoder tables (deberta.py:474-505)
    c = "deberta.encoder."
    R = layer_norm(P[c + "rel_embeddings.weight"], P[c + "LayerNorm.weight"], P[c + "LayerNorm.bias"], eps)
    delta = torch.from_numpy(delta_table(L, cfg.position_buckets, cfg.max_position_embeddings))
    M = (mf[:, None, :, None] * mf[:, None, None, :])                # [B,1,L,L] (deberta.py:480-490)
    hidden = [x0]
    x = x0
    NL = cfg.num_hidden_layers
    # the 24th ordinary pass only feeds hidden_states[-1], which nothing reads:
    # the oracle skips it; tests against the real reference show logits are unchanged.
    for i in range(NL - 1):
        y = layer_pass(ctx, x, x, M, R, delta, P, i, cfg)
        if i == 0:
            # ConvLayer (deberta.py:395-419, 549-550): conv over the EMBEDDING output x0
            w = P[c + "conv.conv.weight"]
            g = Fnn.conv1d(ctx.r(x0).permute(0, 2, 1), ctx.r(w), P[c + "conv.conv.bias"], padding=1).permute(0, 2, 1)
            g = g * mf[:, :, None]
            g = gelu(g)
            y = layer_norm(y + g, P[c + "conv.LayerNorm.weight"], P[c + "conv.LayerNorm.bias"], eps) * mf[:, :, None]
        hidden.append(y)
        x = y
    if NL == 1:
        raise NotImplementedError("need >= 2 layers (EMD reads hidden_states[-2])")
    # --- enhanced mask decoder (deberta.py:1382-1412): 2 passes of layer[-1], KV = h[-2]
    kv = x
    z = pe.unsqueeze(0).expand(B, L, -1) + kv
    o1 = layer_pass(ctx, z, kv, M, R, delta, P, NL - 1, cfg)
    o2 = layer_pass(ctx, o1, kv, M, R, delta, P, NL - 1, cfg)
    # --- head (deberta.py:1544-1558, 1473-1479)
    hp = "lm_predictions.lm_head."
    t = layer_norm(gelu(ctx.lin(o2, P[hp + "dense.weight"], P[hp + "dense.bias"])),
                   P[hp + "LayerNorm.weight"], P[hp + "LayerNorm.bias"], eps)
    if cfg.n_ans and not mlm:
        E, bvec = P["answer_embeddings.weight"], P["answer_bias"]
    else:
        E, bvec = P[e + "word_embeddings.weight"], P[hp + "bias"]
    logits = ctx.lin(t, E) + bvec
    loss = None
    if labels is not None:
        # deberta.py:1452-1462: -100 labels for the F visual slots; CE mean over labels != -100
        if cfg.features_dim and video is not None:
            labels = torch.cat([torch.full((B, L - Lt), -100, dtype=labels.dtype), labels], 1)
        loss = Fnn.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)
    out = dict(logits=logits, loss=loss)
    if return_hidden:
        out["hidden"] = hidden + [o1, o2]
        out["head_t"] = t
    return out


def bert_forward(P, cfg, *, video, video_mask, input_ids, attention_mask, labels=None, mlm=False,
                 bf16_gemm_inputs=False, return_hidden=False):
    """BERT variant, reference model/bert.py: embeddings :246-278 (visual prefix + absolute positions + token type 0,
    LayerNorm, NO mask multiply), BertSelfAttention :141-191 (scores / sqrt(d) + (1 - m_key) * -10000, every query
    row computed), BertSelfOutput / BertOutput :281-292, 349-360 (post-LN, no adapters), BertLMPredictionHead
    :74-96 (decoder weight in MLM mode, answer embeddings downstream), loss :846-860."""
    ctx = _Ctx(bf16_gemm_inputs)
    eps = cfg.layer_norm_eps
    e = "bert.embeddings."
    B, Lt = input_ids.shape
    emb = P[e + "word_embeddings.weight"][input_ids]
    if cfg.features_dim and video is not None:
        v = ctx.lin(video, P[e + "linear_video.weight"], P[e + "linear_video.bias"])
        emb = torch.cat([v, emb], 1)
        if video_mask is None:
            video_mask = torch.ones(video.shape[:2], dtype=attention_mask.dtype)
        m = torch.cat([video_mask, attention_mask], 1)
    else:
        m = attention_mask
    L = emb.shape[1]
    emb = emb + P[e + "position_embeddings.weight"][:L] + P[e + "token_type_embeddings.weight"][0]
    x = layer_norm(emb, P[e + "LayerNorm.weight"], P[e + "LayerNorm.bias"], eps)
    ext = (1.0 - m.to(emb.dtype))[:, None, None, :] * -10000.0
    h, d = cfg.num_attention_heads, cfg.head_dim
    hidden = [x]
    for i in range(cfg.num_hidden_layers):
        p = f"bert.encoder.layer.{i}."
        a = p + "attention.self."
        Q = heads(ctx.lin(x, P[a + "query.weight"], P[a + "query.bias"]), h)
        K = heads(ctx.lin(x, P[a + "key.weight"], P[a + "key.bias"]), h)
        V = heads(ctx.lin(x, P[a + "value.weight"], P[a + "value.bias"]), h)
        s = ctx.r(Q) @ ctx.r(K).transpose(-1, -2) / math.sqrt(d) + ext
        pr = torch.softmax(s, -1)
        c = (ctx.r(pr) @ ctx.r(V)).permute(0, 2, 1, 3).reshape(B, L, h * d)
        u = ctx.lin(c, P[p + "attention.output.dense.weight"], P[p + "attention.output.dense.bias"])
        a1 = layer_norm(u + x, P[p + "attention.output.LayerNorm.weight"], P[p + "attention.output.LayerNorm.bias"], eps)
        f = gelu(ctx.lin(a1, P[p + "intermediate.dense.weight"], P[p + "intermediate.dense.bias"]))
        y = ctx.lin(f, P[p + "output.dense.weight"], P[p + "output.dense.bias"])
        x = layer_norm(y + a1, P[p + "output.LayerNorm.weight"], P[p + "output.LayerNorm.bias"], eps)
        hidden.append(x)
    hp = "cls.predictions."
    t = layer_norm(gelu(ctx.lin(x, P[hp + "transform.dense.weight"], P[hp + "transform.dense.bias"])),
                   P[hp + "transform.LayerNorm.weight"], P[hp + "transform.LayerNorm.bias"], eps)
    if cfg.n_ans and not mlm:
        logits = ctx.lin(t, P["answer_embeddings.weight"]) + P["answer_bias"]
    else:
        logits = ctx.lin(t, P[hp + "decoder.weight"]) + P[hp + "bias"]
    loss = None
    if labels is not None:
        if cfg.features_dim and video is not None:
            labels = torch.cat([torch.full((B, L - Lt), -100, dtype=labels.dtype), labels], 1)
        loss = Fnn.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)
    out = dict(logits=logits, loss=loss)
    if return_hidden:
        out["hidden"] = hidden
    return out


def forward_backward(P, cfg, trainable_names, **kw):
    """Runs forward + autograd backward; returns (out, {name: grad}) for the trainable set."""
    Pg = {k: (v.clone().requires_grad_(True) if k in trainable_names else v) for k, v in P.items()}
    out = (bert_forward if cfg.arch == "bert" else forward)(Pg, cfg, **kw)
    out["loss"].backward()
    grads = {k: Pg[k].grad for k in trainable_names if Pg[k].grad is not None}
    return out, grads
