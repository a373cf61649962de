"""FFN up-projection GEMM (11264 x 6144 x 1536, +bias +GELU +pre-activation output) in isolation:
timing of epilogue variants, and the target of the `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib
dev = "cuda"
T, H, I = 11264, 1536, 6144
A = torch.randn(T, H, device=dev).bfloat16(); W = (0.02 * torch.randn(I, H, device=dev)).bfloat16()
bias = torch.randn(I, device=dev); D = torch.empty(T, I, device=dev, dtype=torch.bfloat16); Dp = torch.empty_like(D)
D32 = torch.empty(T, H, device=dev); W2 = (0.02 * torch.randn(H, I, device=dev)).bfloat16(); resid = torch.randn(T, H, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
quick = "--quick" in sys.argv
variants = [("bias+gelu+pre", lambda: lib.gemm(A, W, D, bias=bias, act=lib.FB_ACT_GELU, D_pre=Dp))]
if not quick:
    variants = [("plain", lambda: lib.gemm(A, W, D)), ("bias", lambda: lib.gemm(A, W, D, bias=bias)),
                ("bias+gelu", lambda: lib.gemm(A, W, D, bias=bias, act=lib.FB_ACT_GELU)),
                ("bias+pre (two outputs, no GELU)", lambda: lib.gemm(A, W, D, bias=bias, D_pre=Dp))] + variants + \
               [("ffn2 f32+resid", lambda: lib.gemm(D, W2, D32, bias=resid[0].contiguous(), resid=resid))]
for name, fn in variants:
    us = t(fn, 3 if quick else 10)
    print(f"{name}: {us:.1f} us  {2*T*I*H/us/1e6:.1f} TFLOP/s", flush=True)
