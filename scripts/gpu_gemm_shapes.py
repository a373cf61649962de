"""Timing of the step's big GEMM shapes (T = 11264 tokens), CUDA events over back-to-back launches.
FB_GEMM_CLUSTER=0 disables the 2-CTA multicast kernel (separate process: the switch is read once)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib

dev = "cuda"
T, H, I = 11264, 1536, 6144
torch.manual_seed(0)


def t(r, c):
    return (torch.randn(r, c, device=dev) * 0.05).bfloat16()


x_h, x_i, x_3h = t(T, H), t(T, I), t(T, 3 * H)
w_qkv, w_o, w_f1, w_f2 = t(3 * H, H), t(H, H), t(I, H), t(H, I)
b_i, b_h, b_3h = torch.randn(I, device=dev), torch.randn(H, device=dev), torch.randn(3 * H, device=dev)
o_h16, o_i16, o_3h16 = torch.empty(T, H, device=dev, dtype=torch.bfloat16), torch.empty(T, I, device=dev, dtype=torch.bfloat16), torch.empty(T, 3 * H, device=dev, dtype=torch.bfloat16)
pre_i = torch.empty(T, I, device=dev, dtype=torch.bfloat16)
o_h32 = torch.empty(T, H, device=dev)
res = torch.randn(T, H, device=dev)

cases = [
    ("qkv   T x 4608 x 1536 (+bias)", 2.0 * T * 3 * H * H, lambda: lib.gemm(x_h, w_qkv, o_3h16, bias=b_3h)),
    ("ffn1  T x 6144 x 1536 (+bias+GELU, pre)", 2.0 * T * I * H, lambda: lib.gemm(x_h, w_f1, o_i16, bias=b_i, act=lib.FB_ACT_GELU, D_pre=pre_i)),
    ("ffn2  T x 1536 x 6144 (+bias)", 2.0 * T * I * H, lambda: lib.gemm(x_i, w_f2, o_h16, bias=b_h)),
    ("o     T x 1536 x 1536 (+bias)", 2.0 * T * H * H, lambda: lib.gemm(x_h, w_o, o_h16, bias=b_h)),
    ("dX qkv  T x 1536 x 4608 (B MN-major, +resid f32)", 2.0 * T * 3 * H * H, lambda: lib.gemm(x_3h, w_qkv, o_h32, b_mn_major=True, resid=res)),
    ("dX ffn2 T x 6144 x 1536 (B MN-major, gelu-grad aux)", 2.0 * T * I * H, lambda: lib.gemm(x_h, w_f2, o_i16, b_mn_major=True, aux=pre_i, aux_op=lib.FB_AUX_MUL_GELU_GRAD)),
    ("dX ffn1 T x 1536 x 6144 (B MN-major, +resid f32)", 2.0 * T * I * H, lambda: lib.gemm(x_i, w_f1, o_h32, b_mn_major=True, resid=res)),
    ("dX o    T x 1536 x 1536 (B MN-major)", 2.0 * T * H * H, lambda: lib.gemm(x_h, w_o, o_h16, b_mn_major=True)),
]
print("FB_GEMM_CLUSTER =", os.environ.get("FB_GEMM_CLUSTER", "(default on)"))
for name, fl, fn in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record()
    torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / 20
    print(f"{name:52s} {us:7.1f} us  {fl / us / 1e6:7.0f} TFLOP/s")
