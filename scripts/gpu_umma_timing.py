"""tcgen05.mma issue/complete cycles per instruction for the operand-major / shape combinations the attention
kernels use (one CTA, 64 back-to-back MMAs on resident shared-memory operands), via fb_umma_probe."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from frozenbilm_b200 import lib

lib.use_debug_library()   # test-only build with the probe / phase-stamp hooks


def desc(lbo, sbo, layout):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | layout << 61


def idesc(m, n, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


a = torch.zeros(131072, dtype=torch.uint8, device="cuda")
b = torch.zeros(65536, dtype=torch.uint8, device="cuda")
cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
NM = 64
cases = [
    # name, M, N, a_mn, b_mn, a_desc, b_desc, a_off(ks), b_off(ks)
    ("A K-major sw128, B K-major sw128, 128x128 (S=QK^T, dP)", 128, 128, 0, 0, desc(16, 1024, 2), desc(16, 1024, 2),
     lambda k: (k % 4) * 32, lambda k: (k % 4) * 32),
    ("A K-major sw128, B K-major sw128, 128x256 (band scores)", 128, 256, 0, 0, desc(16, 1024, 2), desc(16, 1024, 2),
     lambda k: (k % 4) * 32, lambda k: (k % 4) * 32),
    ("A K-major sw128, B MN-major sw128, 128x64 (P.V, dS.K, Z.band)", 128, 64, 0, 1, desc(16, 1024, 2), desc(16, 1024, 2),
     lambda k: (k % 4) * 32, lambda k: (k % 8) * 2048),
    ("A MN-major sw128 (LBO 16K), B MN-major, 128x64 (Pd^T.dO, dS^T.Q)", 128, 64, 1, 1, desc(16384, 1024, 2), desc(16, 1024, 2),
     lambda k: (k % 8) * 2048, lambda k: (k % 8) * 2048),
    ("A MN-major sw128 (LBO 8K), B MN-major, 128x64", 128, 64, 1, 1, desc(8192, 1024, 2), desc(16, 1024, 2),
     lambda k: (k % 4) * 2048, lambda k: (k % 8) * 2048),
    ("A MN-major sw128 (LBO 16K), B K-major, 128x64", 128, 64, 1, 0, desc(16384, 1024, 2), desc(16, 1024, 2),
     lambda k: (k % 8) * 2048, lambda k: (k % 4) * 32),
    ("A MN-major sw128 (LBO 16K), B MN-major, 128x128", 128, 128, 1, 1, desc(16384, 1024, 2), desc(16384, 1024, 2),
     lambda k: (k % 8) * 2048, lambda k: (k % 8) * 2048),
    ("A K-major none (LBO 128, SBO 4096), B MN-major sw128, 128x64", 128, 64, 0, 1, desc(128, 4096, 0), desc(16, 1024, 2),
     lambda k: (k % 16) * 256, lambda k: (k % 8) * 2048),
    ("A MN-major none (LBO 4224, SBO 128), B MN-major sw128, 128x64", 128, 64, 1, 1, desc(4224, 128, 0), desc(16, 1024, 2),
     lambda k: (k % 8) * 8448, lambda k: (k % 8) * 2048),
    ("A K-major sw128, B MN-major, 128x16 (transpose by identity)", 128, 16, 0, 1, desc(16, 1024, 2), desc(16, 1024, 2),
     lambda k: (k % 4) * 32, lambda k: 0),
    ("A MN-major sw128 (LBO 16K), B K-major none, 128x16 (transpose by identity)", 128, 16, 1, 0, desc(16384, 1024, 2), desc(128, 256, 0),
     lambda k: (k % 8) * 2048, lambda k: 0),
    ("A K-major sw128, B K-major, 64x128 (M = 64)", 64, 128, 0, 0, desc(16, 1024, 2), desc(16, 1024, 2),
     lambda k: (k % 4) * 32, lambda k: (k % 4) * 32),
]
for name, M, N, amn, bmn, ad, bd, ao, bo in cases:
    try:
        for rep in range(2):
            lib.umma_probe(a, b, [ao(k) for k in range(NM)], [bo(k) for k in range(NM)], ad, bd, idesc(M, N, amn, bmn),
                           max(32, N), cycles=cyc)
            torch.cuda.synchronize()
        c = cyc.cpu().tolist()
        print(f"{name:78s} issue {c[0] / NM:6.1f}  complete {c[1] / NM:6.1f} cycles/MMA  "
              f"({2 * M * N * 16 / (c[1] / NM):7.0f} flop/cycle)", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(name, "EXCEPTION", repr(ex)[:160])
        break
