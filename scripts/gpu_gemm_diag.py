"""GPU bring-up diagnostics for fb_gemm: correctness over layouts + TFLOP/s on the hot shapes."""
import math, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib

dev = "cuda"
print(torch.cuda.get_device_name(0))
g = torch.Generator(device=dev).manual_seed(0)

def run(M, N, K, a_mn, b_mn, bn):
    def mk(r, c, mn):
        rows, cols = (c, r) if mn else (r, c)
        ld = (cols + 7) // 8 * 8
        return torch.randn(rows, ld, device=dev, generator=g).bfloat16()[:, :cols]
    A, B = mk(M, K, a_mn), mk(N, K, b_mn)
    D = torch.full((M, (N + 7) // 8 * 8), float("nan"), device=dev)[:, :N]
    lib.gemm(A, B, D, a_mn_major=a_mn, b_mn_major=b_mn, block_n=bn)
    torch.cuda.synchronize()
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    ref = Af @ Bf.t()
    err = (D - ref).abs().max().item()
    print(f"M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} bn={bn}: max_abs_err={err:.3e} ref_absmax={ref.abs().max().item():.2f}", flush=True)
    return err

bad = 0
for (M, N, K) in [(128, 64, 64), (128, 256, 64), (128, 128, 256), (256, 512, 192), (300, 200, 136), (1000, 1536, 1536)]:
    for a_mn, b_mn in [(False, False), (False, True), (True, False), (True, True)]:
        for bn in ([64, 128, 192, 256] if M <= 300 else [0]):
            try:
                e = run(M, N, K, a_mn, b_mn, bn)
                bad += (not (e < 1e-2))
            except Exception as ex:
                print("EXC", M, N, K, a_mn, b_mn, bn, ex); bad += 1
print("bad cases:", bad)

def bench(M, N, K, a_mn=False, b_mn=False, bn=0, out=torch.bfloat16, iters=20, **kw):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    D = torch.empty(M, N, device=dev, dtype=out)
    for _ in range(3): lib.gemm(A, B, D, a_mn_major=a_mn, b_mn_major=b_mn, block_n=bn, **kw)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): lib.gemm(A, B, D, a_mn_major=a_mn, b_mn_major=b_mn, block_n=bn, **kw)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"bench M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)

if bad == 0 or "--bench" in sys.argv:
    T = 11264
    for bn in (128, 192, 256):
        bench(T, 1536, 1536, bn=bn)
        bench(T, 6144, 1536, bn=bn)
        bench(T, 1536, 6144, bn=bn)
    bench(T, 1536, 1536, b_mn=True)
    bench(T, 192, 1536)
    bench(T, 1536, 192)
    bench(8192, 8192, 8192, bn=256)
    # cuBLAS reference for context
    A = torch.randn(T, 1536, device=dev).bfloat16(); B = torch.randn(6144, 1536, device=dev).bfloat16()
    for _ in range(3): A @ B.t()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): A @ B.t()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f"cublas T x 6144 x 1536: {ms*1e3:.1f} us {2*T*6144*1536/ms/1e9:.1f} TFLOP/s")
