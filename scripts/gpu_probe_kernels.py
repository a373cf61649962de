"""Launches the four kernels bench.py reports a roofline for, at the BASELINE configs[1] shapes (B = 32, L = 352,
H = 1536): the FFN up-projection GEMM (+bias+GELU, two outputs) and the three attention kernels.  Under
`ncu --set full -k regex:... ` this gives one capture per kernel with an unambiguous name (the launch list of a whole
step has several GEMM shapes per template instantiation); on its own it prints CUDA-event timings."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib
from frozenbilm_b200.relpos import bucket_of_rel

dev = "cuda"
B, L, h, I = 32, 352, 24, 6144
H, T = h * 64, B * L
lp = (L + 127) // 128 * 128
torch.manual_seed(0)
qkv = (torch.randn(T, 3 * H, device=dev) * 0.5).bfloat16()
pqk = (torch.randn(512, 2 * H, device=dev) * 0.5).bfloat16()
mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
bor = torch.from_numpy(bucket_of_rel(256, 512)).to(dev)
rel_rows, _ = lib.rel_table_geometry(L)
PKrev = torch.empty(h, rel_rows, 64, device=dev, dtype=torch.bfloat16)
PQrel, PKrel = lib.build_rel_tables(pqk, bor, L, h, out_pk_rev=PKrev)
Q, K, V = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
ctx = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
dctx = (torch.randn(T, H, device=dev) * 0.1).bfloat16()
dqkv = torch.empty(T, 3 * H, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, h, L, device=dev)
s2 = torch.empty(B, h, L, lp, device=dev, dtype=lib.SCORE_DTYPE)
ds = torch.empty(B, h, L, lp, device=dev, dtype=torch.bfloat16)
gpk = torch.empty(h, rel_rows, 64, device=dev)
gpq = torch.empty(h, rel_rows, 64, device=dev)
x16 = (torch.randn(T, H, device=dev) * 0.5).bfloat16()
w1 = (torch.randn(I, H, device=dev) * 0.02).bfloat16()
b1 = torch.randn(I, device=dev) * 0.02
f16 = torch.empty(T, I, device=dev, dtype=torch.bfloat16)
h1 = torch.empty(T, I, device=dev, dtype=torch.bfloat16)
scale = 1.0 / math.sqrt(192)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

KERNELS = {
    "ffn1 gemm": lambda: lib.gemm(x16, w1, f16, bias=b1, act=lib.FB_ACT_GELU, D_pre=h1),
    "attn_fwd": lambda: lib.attn_fwd(Q, K, V, PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h, scale=scale, dropout_p=0.1,
                                     seed=3, offset=5, s_out=s2),
    "attn_bwd": lambda: lib.attn_bwd(Q, K, V, ctx, dctx, s2, lse, PKrev, PQrel, dqkv[:, :H], dqkv[:, H:2 * H],
                                     dqkv[:, 2 * H:], ds_out=ds, B=B, L=L, h=h, scale=scale, dropout_p=0.1, seed=3, offset=5),
    "attn_bwd_rel": lambda: lib.attn_bwd_rel(ds, Q, K, gpk, gpq, B=B, L=L, h=h, scale=scale),
}
if os.environ.get("PROBE_ONCE"):
    # under `ncu -k regex:"gemm_kernel|attn_fwd|attn_bwd" --launch-count 4`: exactly one capture per kernel
    lib.attn_fwd(Q, K, V, PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h, scale=scale, dropout_p=0.1, seed=3, offset=5,
                 s_out=s2)                      # warm launch that produces s2 / lse (ncu: --launch-skip 1)
    torch.cuda.synchronize()
    for fn in KERNELS.values():
        fn()
    torch.cuda.synchronize()
    sys.exit(0)
for name, fn in KERNELS.items():
    for _ in range(3):
        fn()
    tot, n = 0.0, 10
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    print(f"{name:14s}: {1e3 * tot / n:8.1f} us")
