"""Phase timeline of fb_attn_bwd's CTA (0,0) at the BASELINE shape (B=32, h=24, L=352): clock64 stamps recorded
by the kernel itself (fb_attn_bwd_set_debug), printed as per-pair deltas in SM cycles."""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from frozenbilm_b200 import lib

lib.use_debug_library()   # test-only build with the probe / phase-stamp hooks
from frozenbilm_b200.relpos import bucket_of_rel

B, L, h = 32, 352, 24
H = h * 64
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B * L, 3 * H, device=dev, generator=g).bfloat16()
pqk = torch.randn(512, 2 * H, device=dev, generator=g).bfloat16()
dO = torch.randn(B * L, H, device=dev, generator=g).bfloat16()
mask = torch.ones(B, L, dtype=torch.uint8, device=dev)
bor = torch.from_numpy(bucket_of_rel(256, 512)).to(dev)
lp = 384
scale = 1.0 / math.sqrt(192)
PKrev = torch.empty(h, 2 * 384 + 8, 64, device=dev, dtype=torch.bfloat16)
PQrel, PKrel = lib.build_rel_tables(pqk, bor, L, h, out_pk_rev=PKrev)
Q, K, V = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
ctx = torch.zeros(B * L, H, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, h, L, device=dev)
s2 = torch.empty(B, h, L, lp, device=dev, dtype=lib.SCORE_DTYPE)
lib.attn_fwd(Q, K, V, PKrel, PQrel, mask, ctx, lse, B=B, L=L, h=h, scale=scale, dropout_p=0.1, seed=3, offset=5, s_out=s2)
dqkv = torch.empty(B * L, 3 * H, device=dev, dtype=torch.bfloat16)
ds = torch.empty(B, h, L, lp, device=dev, dtype=torch.bfloat16)


def bwd():
    lib.attn_bwd(Q, K, V, ctx, dO, s2, lse, PKrev, PQrel, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], ds_out=ds,
                 B=B, L=L, h=h, scale=scale, dropout_p=0.1, seed=3, offset=5)


for _ in range(3):
    bwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    bwd()
e1.record()
torch.cuda.synchronize()
print(f"fb_attn_bwd: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per launch (no stamps)")
stamps = torch.zeros(16 * 32, dtype=torch.int64, device=dev)
fn = lib.use_debug_library().fb_attn_bwd_set_debug
fn.argtypes = [C.c_void_p]
fn(C.c_void_p(stamps.data_ptr()))
bwd()
torch.cuda.synchronize()
fn(C.c_void_p(0))
st = stamps.cpu().view(16, 32)
names_c = ["start", "dP issued", "rdyA seen", "A issued", "rdyB seen", "B issued", "mmaA done", "mmaB done", "pq landed",
           "rdyC seen", "C issued", "mmaC done"]
names_w = ["start", "dP ready", "elementwise done", "S0/Z free", "staged + Z1 built", "arrived (rdyA, rdyB)", "ds stored",
           "mmaB seen", "Z2 built (rdyC)", "epilogue done", "scores loaded + parked until mmaC"]
tot = 0
for t in range(9):
    c = [int(x) for x in st[t, :12]]
    w = [int(x) for x in st[t, 16:27]]
    print(f"pair {t}: total {c[11] - c[0]} cycles")
    print("  ctrl: " + "  ".join(f"{n} +{v - pv}" for n, v, pv in zip(names_c[1:], c[1:], c[:-1])))
    print("  warp: " + "  ".join(f"{n} +{v - pv}" for n, v, pv in zip(names_w[1:], w[1:], w[:-1])))
    tot += c[11] - c[0]
print("sum of pairs", tot, "cycles")
