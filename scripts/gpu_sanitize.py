"""Small invocations of the hand-rolled mbarrier / TMEM / TMA protocols for compute-sanitizer (memcheck, racecheck,
synccheck): the three attention kernels (forward generation 2, fused backward, band gradients), the 2-CTA cluster GEMM
with multicast B tiles and a TMA reduce-add (split-K) GEMM, at shapes small enough for the tools' 10-100x slowdown."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib
from frozenbilm_b200.relpos import bucket_of_rel

dev = "cuda"
B, L, h = 2, 200, 2
H = h * 64
lp = (L + 127) // 128 * 128
torch.manual_seed(0)
qkv = (torch.randn(B * L, 3 * H, device=dev) * 0.5).bfloat16()
pqk = (torch.randn(512, 2 * H, device=dev) * 0.5).bfloat16()
mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
mask[-1, L - 30:] = 0
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
ds = torch.zeros(B, h, L, lp, device=dev, dtype=torch.bfloat16)
gpk = torch.empty(h, rel_rows, 64, device=dev)
gpq = torch.empty(h, rel_rows, 64, device=dev)
scale = 1.0 / math.sqrt(192)
for rel in (True, False):
    pk, pq, pkr = (PKrel, PQrel, PKrev) if rel else (None, None, None)
    lib.attn_fwd(Q, K, V, pk, pq, mask, ctx, lse, B=B, L=L, h=h, scale=scale, dropout_p=0.1, seed=3, offset=5, s_out=s2,
                 key_mask_only=not rel)
    lib.attn_bwd(Q, K, V, ctx, dctx, s2, lse, pkr, pq, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:],
                 ds_out=ds if rel else None, B=B, L=L, h=h, scale=scale, dropout_p=0.1, seed=3, offset=5)
    if rel:
        lib.attn_bwd_rel(ds, Q, K, gpk, gpq, B=B, L=L, h=h, scale=scale)
torch.cuda.synchronize()
print("attention kernels ok", float(ctx.float().abs().sum()), float(dqkv.float().abs().sum()), float(gpk.abs().sum()))
# cluster GEMM (needs >= 2 * num_sms tiles and an even number of M blocks) + epilogue operands + split-K reduce-add
M, N, K = 148 * 2 * 128 // 2, 512, 256
A = (torch.randn(M, K, device=dev) * 0.1).bfloat16()
Bw = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
bias = torch.randn(N, device=dev)
resid = torch.randn(M, N, device=dev)
D = torch.empty(M, N, device=dev)
lib.gemm(A, Bw, D, bias=bias, resid=resid, block_n=256)
D2 = torch.zeros(256, N, device=dev)
lib.gemm(A[:256], Bw, D2, split_k=4)
torch.cuda.synchronize()
ref = A.float() @ Bw.float().t()
print("gemm ok", float((D - (ref + bias + resid)).abs().max()), float((D2 - ref[:256]).abs().max()))
