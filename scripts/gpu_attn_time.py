"""Isolated timing of the attention kernels at the BASELINE configs[1] shape (B=32, L=352, h=24), CUDA events."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib
from frozenbilm_b200.relpos import bucket_of_rel

dev = "cuda"
B, L, h = 32, 352, 24
H = h * 64
lp = (L + 127) // 128 * 128
torch.manual_seed(0)
qkv = (torch.randn(B * L, 3 * H, device=dev) * 0.5).bfloat16()
pqk = (torch.randn(512, 2 * H, device=dev) * 0.5).bfloat16()
mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
bor = torch.from_numpy(bucket_of_rel(256, 512)).to(dev)
rel_rows, _ = lib.rel_table_geometry(L)
PKrev = torch.empty(h, rel_rows, 64, device=dev, dtype=torch.bfloat16)
PQrel, PKrel = lib.build_rel_tables(pqk, bor, L, h, out_pk_rev=PKrev)
Q, K, V = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
ctx = torch.empty(B * L, H, device=dev, dtype=torch.bfloat16)
dctx = (torch.randn(B * L, H, device=dev) * 0.1).bfloat16()
dqkv = torch.empty(B * L, 3 * H, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, h, L, device=dev)
s2 = torch.empty(B, h, L, lp, device=dev, dtype=lib.SCORE_DTYPE)
ds = torch.empty(B, h, L, lp, device=dev, dtype=torch.bfloat16)
gpk = torch.empty(h, rel_rows, 64, device=dev)
gpq = torch.empty(h, rel_rows, 64, device=dev)
scale = 1.0 / math.sqrt(192)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def fwd(train=True):
    lib.attn_fwd(Q, K, V, PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h, scale=scale, dropout_p=0.1 if train else 0.0,
                 seed=3, offset=5, s_out=s2 if train else None)


def bwd():
    lib.attn_bwd(Q, K, V, ctx, dctx, s2, lse, PKrev, PQrel, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], ds_out=ds,
                 B=B, L=L, h=h, scale=scale, dropout_p=0.1, seed=3, offset=5)


def rel():
    lib.attn_bwd_rel(ds, Q, K, gpk, gpq, B=B, L=L, h=h, scale=scale)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return 1e3 * tot / n


flops_f = B * L * H * (4 * L + 8 * 256)
print(f"attn_fwd train : {timeit(lambda: fwd(True)):8.1f} us   ({flops_f / timeit(lambda: fwd(True)) / 1e6:.0f} TFLOP/s algorithmic)")
print(f"attn_fwd eval  : {timeit(lambda: fwd(False)):8.1f} us")
fwd(True)
print(f"attn_bwd       : {timeit(bwd):8.1f} us")
print(f"attn_bwd_rel   : {timeit(rel):8.1f} us")
