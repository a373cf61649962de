"""Isolated timing of the LayerNorm kernels at config 2 (T = 11264 rows, H = 1536), CUDA events, L2 flushed by cycling
through four input sets (4 x 242 MB > 126 MB L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frozenbilm_b200 import lib

dev = "cuda"
T, H = 11264, 1536
sets = []
for i in range(4):
    x, dy = torch.randn(T, H, device=dev), torch.randn(T, H, device=dev)
    sets.append((x, dy, torch.empty(T, H, device=dev), torch.empty(T, H, device=dev, dtype=torch.bfloat16)))
gamma, beta = torch.randn(H, device=dev), torch.randn(H, device=dev)
mean, rstd = torch.zeros(T, device=dev), torch.ones(T, device=dev)
dg, db, cs = (torch.zeros(H, device=dev) for _ in range(3))
od = lib.drop(0.1, 3, 5)


def bwd(i):
    x, dy, o32, o16 = sets[i % 4]
    lib.layernorm_bwd(dy, x, mean, rstd, gamma, in_drop=od, dx32=o32, dx16=o16, dgamma=dg, dbeta=db, dx16_colsum=cs)


def fwd(i):
    x, dy, o32, o16 = sets[i % 4]
    lib.layernorm_fwd(x, gamma, beta, 1e-7, y32=o32, y16=o16, mean=mean, rstd=rstd)


for name, fn, nbytes in (("ln_fwd", fwd, 10), ("ln_bwd", bwd, 14)):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(40):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / 40
    print(f"{name}: {us:.1f} us  {nbytes * T * H / us / 1e6:.2f} TB/s algorithmic")
