"""Timing of the adapter-size GEMMs of a layer pass (config 2: T = 11264, H = 1536, r = 192), CUDA events, four
operand sets cycled (each launch's working set comes from HBM).  FB_GEMM_EG=2|3 selects the epilogue warpgroup count
of the 128 x 192 cluster kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib

dev = "cuda"
T, H, r = 11264, 1536, 192
torch.manual_seed(0)
S = []
for i in range(4):
    S.append(dict(u16=(torch.randn(T, H, device=dev) * 0.5).bfloat16(), z16=(torch.randn(T, r, device=dev) * 0.5).bfloat16(),
                  res=torch.randn(T, H, device=dev), s32=torch.empty(T, H, device=dev),
                  o16=torch.empty(T, H, device=dev, dtype=torch.bfloat16), zo=torch.empty(T, r, device=dev, dtype=torch.bfloat16)))
wu = (torch.randn(H, r, device=dev) * 0.02).bfloat16()
wd = (torch.randn(r, H, device=dev) * 0.02).bfloat16()
bu, bd = torch.randn(H, device=dev) * 0.01, torch.randn(r, device=dev) * 0.01
drop = lib.drop(0.1, 5, 7)

cases = {
    "down  T x 192 x 1536 (+bias+ReLU+dropout)": (lambda s: lib.gemm(s["u16"], wd, s["zo"], bias=bd, act=lib.FB_ACT_RELU, dropout=drop, drop_pos=2), 2 * T * H + 2 * T * r),
    "up    T x 1536 x 192 (+bias+skip+dropout+resid, fp32 out)": (lambda s: lib.gemm(s["z16"], wu, s["s32"], bias=bu, aux=s["u16"], aux_op=lib.FB_AUX_ADD, dropout=drop, drop_pos=2, resid=s["res"]), 2 * T * r + 2 * T * H + 8 * T * H),
    "du    T x 1536 x 192 (B MN-major, +aux, bf16 out)": (lambda s: lib.gemm(s["z16"], wd, s["o16"], b_mn_major=True, aux=s["u16"], aux_op=lib.FB_AUX_ADD), 2 * T * r + 4 * T * H),
    "dz    T x 192 x 1536 (B MN-major, ReLU mask)": (lambda s: lib.gemm(s["u16"], wu, s["zo"], b_mn_major=True, aux=s["z16"], aux_op=lib.FB_AUX_MUL_RELU_MASK), 2 * T * H + 4 * T * r),
}
print("FB_GEMM_EG =", os.environ.get("FB_GEMM_EG", "(default)"))
for name, (fn, nbytes) in cases.items():
    for i in range(4):
        fn(S[i])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(40):
        fn(S[i % 4])
    e.record()
    torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / 40
    print(f"{name:60s} {us:7.1f} us  {nbytes / us / 1e6:5.2f} TB/s algorithmic")
