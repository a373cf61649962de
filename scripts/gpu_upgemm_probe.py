"""Decomposition of the adapter up-projection GEMM epilogue cost (T x 1536 x 192), CUDA-event timed, cold L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib

dev = "cuda"
T, H, r, I = 11264, 1536, 192, 6144
torch.manual_seed(0)
z = torch.randn(T, r, device=dev).bfloat16()
wu = (torch.randn(H, r, device=dev) * 0.05).bfloat16()
bias = torch.randn(H, device=dev)
u16 = torch.randn(T, H, device=dev).bfloat16()
res = torch.randn(T, H, device=dev)
o32 = torch.empty(T, H, device=dev)
o16 = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
drop = lib.Dropout(); drop.p = 0.1; drop.seed = 5; drop.offset = 7


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return 1e3 * tot / n


cases = {
    "full (bias+aux+drop+resid, f32)": dict(D=o32, bias=bias, aux=u16, aux_op=lib.FB_AUX_ADD, dropout=drop, drop_pos=2, resid=res),
    "no dropout": dict(D=o32, bias=bias, aux=u16, aux_op=lib.FB_AUX_ADD, resid=res),
    "no resid": dict(D=o32, bias=bias, aux=u16, aux_op=lib.FB_AUX_ADD, dropout=drop, drop_pos=2),
    "no aux": dict(D=o32, bias=bias, dropout=drop, drop_pos=2, resid=res),
    "bias only f32": dict(D=o32, bias=bias),
    "bias only bf16": dict(D=o16, bias=bias),
    "full, bf16 out": dict(D=o16, bias=bias, aux=u16, aux_op=lib.FB_AUX_ADD, dropout=drop, drop_pos=2, resid=res),
}
for bn in (0, 256, 192, 128, 64):
    for name, kw in cases.items():
        kw = dict(kw); D = kw.pop("D")
        t = timeit(lambda: lib.gemm(z, wu, D, block_n=bn, **kw))
        print(f"bn={bn:3d} {name:34s} {t:7.1f} us")
# reference points: down projection and a plain H x H GEMM
zz = torch.empty(T, r, device=dev, dtype=torch.bfloat16)
wd = (torch.randn(r, H, device=dev) * 0.05).bfloat16()
for bn in (0, 64, 128, 192):
    print(f"down T x 192 x 1536 bn={bn}: {timeit(lambda: lib.gemm(u16, wd, zz, bias=torch.zeros(r, device=dev), act=lib.FB_ACT_RELU, block_n=bn)):7.1f} us")
    for sk in (2, 3):
        pass
w = (torch.randn(H, H, device=dev) * 0.02).bfloat16()
print(f"dense T x 1536 x 1536 bf16 out: {timeit(lambda: lib.gemm(u16, w, o16, bias=bias)):7.1f} us")
print(f"dense T x 1536 x 1536 f32 out + resid: {timeit(lambda: lib.gemm(u16, w, o32, bias=bias, resid=res)):7.1f} us")
# pure elementwise reference for the same traffic as the up-projection epilogue
print(f"torch elementwise o32 = res + u16 (same bytes): {timeit(lambda: torch.add(res, u16, out=o32)):7.1f} us")
