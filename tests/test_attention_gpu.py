"""Fused disentangled attention forward vs an fp32 torch restatement of deberta.py:717-947 on the
same bf16 operands.  Tolerance: the kernel rounds P to bf16 before P.V and stages the positional
terms in fp16, so ctx is compared with atol 2e-2*max|V| (measured errors are ~3e-3)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_attention(Q, K, V, PQb, PKb, delta, mask, h, scale):
    """Q,K,V: [B, L, H] fp32; PQb/PKb: [2S, H]; delta: [L, L] long; mask [B, L]."""
    B, L, H = Q.shape
    d = H // h
    def hd(x):
        return x.view(B, L, h, d).permute(0, 2, 1, 3)
    Qh, Kh, Vh = hd(Q), hd(K), hd(V)
    PK = PKb.view(-1, h, d).permute(1, 0, 2)
    PQ = PQb.view(-1, h, d).permute(1, 0, 2)
    qk = Qh @ Kh.transpose(-1, -2)
    c2p_full = Qh @ PK.transpose(-1, -2).unsqueeze(0)
    p2c_full = Kh @ PQ.transpose(-1, -2).unsqueeze(0)
    idx = delta.view(1, 1, L, L).expand(B, h, L, L)
    c2p = torch.gather(c2p_full, -1, idx)
    p2c = torch.gather(p2c_full, -1, idx.transpose(-1, -2)).transpose(-1, -2)
    s = (qk + c2p + p2c) * scale
    m = mask.float()
    M = (m[:, None, :, None] * m[:, None, None, :]).bool()
    s = s.masked_fill(~M, -1e30)
    p = torch.softmax(s, -1).masked_fill(~M, 0.0)
    lse = torch.logsumexp(s, -1)
    ctx = (p @ Vh).permute(0, 2, 1, 3).reshape(B, L, H)
    return ctx, p, lse


@pytest.mark.parametrize("B,L,h", [(2, 34, 4), (3, 128, 4), (2, 352, 4), (1, 512, 2), (2, 129, 3)])
def test_attention_forward(B, L, h):
    from frozenbilm_b200 import lib
    from frozenbilm_b200.relpos import bucket_of_rel
    from oracle.frozenbilm_oracle import delta_table
    dev = "cuda"
    H = h * 64
    g = torch.Generator(device=dev).manual_seed(L * 13 + h)
    qkv = (torch.randn(B * L, 3 * H, device=dev, generator=g)).bfloat16()
    pqk = (torch.randn(512, 2 * H, device=dev, generator=g)).bfloat16()
    mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
    if B > 1:
        mask[-1, L - max(1, L // 5):] = 0
        mask[-1, 1] = 0
    bor = torch.from_numpy(bucket_of_rel(256, 512)).to(dev)
    PQrel, PKrel = lib.build_rel_tables(pqk, bor, L, h)
    ctx = torch.full((B * L, H), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, h, L, device=dev)
    scale = 1.0 / math.sqrt(64 * 3)
    lib.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h, scale=scale)
    torch.cuda.synchronize()
    delta = torch.from_numpy(delta_table(L, 256, 512)).to(dev)
    qf = qkv.float().view(B, L, 3 * H)
    ref_ctx, ref_p, ref_lse = ref_attention(qf[..., :H].contiguous(), qf[..., H:2 * H].contiguous(),
                                            qf[..., 2 * H:].contiguous(), pqk.float()[:, :H], pqk.float()[:, H:],
                                            delta, mask, h, scale)
    got = ctx.float().view(B, L, H)
    err = (got - ref_ctx).abs().max().item()
    assert math.isfinite(err) and err < 3e-2, f"ctx max abs err {err}"
    live = mask.bool()[:, None, :].expand(B, h, L)
    lse_nat = lse * math.log(2.0)
    lerr = (lse_nat[live] - ref_lse[live]).abs().max().item()
    assert lerr < 2e-2, f"lse err {lerr}"
    assert torch.isinf(lse[~live]).all() if (~live).any() else True
    # masked query rows produce exactly zero context (XSoftmax zero-fill)
    dead = ~mask.bool()
    if dead.any():
        assert got[dead].abs().max().item() == 0.0


def test_attention_dropout_statistics():
    """Dropout keeps ~(1-p) of the probability mass and is reproducible for a fixed (seed, offset)."""
    from frozenbilm_b200 import lib
    from frozenbilm_b200.relpos import bucket_of_rel
    dev = "cuda"
    B, L, h = 2, 256, 2
    H = h * 64
    g = torch.Generator(device=dev).manual_seed(1)
    qkv = torch.randn(B * L, 3 * H, device=dev, generator=g).bfloat16()
    qkv[:, 2 * H:] = 1.0  # V = 1 -> ctx = sum of kept probs / (1-p)
    pqk = torch.randn(512, 2 * H, device=dev, generator=g).bfloat16()
    mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
    bor = torch.from_numpy(bucket_of_rel(256, 512)).to(dev)
    PQrel, PKrel = lib.build_rel_tables(pqk, bor, L, h)
    outs = []
    for seed in (7, 7, 8):
        ctx = torch.empty(B * L, H, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, h, L, device=dev)
        lib.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h,
                     scale=0.02, dropout_p=0.1, seed=seed, offset=3)
        outs.append(ctx.float())
    assert torch.equal(outs[0], outs[1])
    assert not torch.equal(outs[0], outs[2])
    assert abs(outs[0].mean().item() - 1.0) < 0.02
