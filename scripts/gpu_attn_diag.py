import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib
from frozenbilm_b200.relpos import bucket_of_rel
from oracle.frozenbilm_oracle import delta_table
from tests.test_attention_gpu import ref_attention
dev = "cuda"
for (B, L, h) in [(1, 128, 1), (2, 34, 4), (2, 352, 4), (1, 512, 2)]:
    H = h * 64
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B * L, 3 * H, device=dev, generator=g).bfloat16()
    pqk = torch.randn(512, 2 * H, device=dev, generator=g).bfloat16()
    mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
    bor = torch.from_numpy(bucket_of_rel(256, 512)).to(dev)
    PQrel, PKrel = lib.build_rel_tables(pqk, bor, L, h)
    ctx = torch.full((B * L, H), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, h, L, device=dev)
    scale = 1.0 / math.sqrt(192)
    delta = torch.from_numpy(delta_table(L, 256, 512)).to(dev)
    qf = qkv.float().view(B, L, 3 * H)
    for variant in ("full", "qk_only", "c2p_only", "p2c_only"):
        pq, pk, q = PQrel, PKrel, qkv
        pqk_v = pqk.clone()
        if variant == "qk_only": pqk_v.zero_()
        if variant == "c2p_only": pqk_v[:, :H] = 0
        if variant == "p2c_only": pqk_v[:, H:] = 0
        pq, pk = lib.build_rel_tables(pqk_v, bor, L, h)
        lib.attn_fwd(qkv[:, :H], qkv[:, H:2*H], qkv[:, 2*H:], pk, pq, mask, ctx, lse, B=B, L=L, h=h, scale=scale)
        torch.cuda.synchronize()
        rc, rp, rl = ref_attention(qf[..., :H].contiguous(), qf[..., H:2*H].contiguous(), qf[..., 2*H:].contiguous(),
                                   pqk_v.float()[:, :H], pqk_v.float()[:, H:], delta, mask, h, scale)
        err = (ctx.float().view(B, L, H) - rc).abs().max().item()
        lerr = (lse * math.log(2) - rl).abs().max().item()
        print(f"B={B} L={L} h={h} {variant}: ctx err {err:.3e} lse err {lerr:.3e}", flush=True)
# timing at config-2 shape
B, L, h = 32, 352, 24; H = h * 64
qkv = torch.randn(B * L, 3 * H, device=dev).bfloat16()
pqk = torch.randn(512, 2 * H, device=dev).bfloat16()
mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
bor = torch.from_numpy(bucket_of_rel(256, 512)).to(dev)
PQrel, PKrel = lib.build_rel_tables(pqk, bor, L, h)
ctx = torch.empty(B * L, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, h, L, device=dev)
for p in (0.0, 0.1):
    for _ in range(3): lib.attn_fwd(qkv[:, :H], qkv[:, H:2*H], qkv[:, 2*H:], PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h, scale=0.07, dropout_p=p)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): lib.attn_fwd(qkv[:, :H], qkv[:, H:2*H], qkv[:, 2*H:], PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h, scale=0.07, dropout_p=p)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    fl = B * L * H * (4 * L + 8 * 256) 
    print(f"attn fwd B=32 L=352 h=24 p={p}: {ms*1e3:.1f} us  ({fl/ms/1e9:.1f} TFLOP/s algorithmic)")
