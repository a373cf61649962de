import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib
from frozenbilm_b200.relpos import bucket_of_rel
dev = "cuda"
B, L, h = 32, 352, 24; H = h * 64
qkv = torch.randn(B * L, 3 * H, device=dev).bfloat16()
pqk = torch.randn(512, 2 * H, device=dev).bfloat16()
mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
bor = torch.from_numpy(bucket_of_rel(256, 512)).to(dev)
PQrel, PKrel = lib.build_rel_tables(pqk, bor, L, h)
ctx = torch.empty(B * L, H, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, h, L, device=dev)
for _ in range(3):
    lib.attn_fwd(qkv[:, :H], qkv[:, H:2*H], qkv[:, 2*H:], PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h, scale=0.07, dropout_p=0.1)
torch.cuda.synchronize()
