"""Pins the TMEM accumulator layout of tcgen05.mma.cta_group::1 with M = 64 (which lane / column holds D[m, n]) on a
B200, with fb_umma_probe: D[m, n] = (m + 1) + (n + 1) / 256 from one K = 16 MMA, read back as the full
[128 lanes x N columns] image.  Needed before a 64-row-tile attention kernel relies on it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from frozenbilm_b200 import lib

lib.use_debug_library()   # test-only build with the probe / phase-stamp hooks
from gpu_umma_probe import b_image_mn_sw128, bf16_bytes, desc, idesc

for Mm, Nn in ((64, 64),):
    A = np.zeros((Mm, 16), np.float32)
    A[:, 0] = np.arange(1, Mm + 1)
    A[:, 1] = 1.0
    Bm = np.zeros((64, 16), np.float32)       # the helper builds 64-column MN-major tiles
    Bm[:, 0] = 1.0
    Bm[:, 1] = np.arange(1, 65) / 256.0
    a_img = np.zeros(128 * 64, np.uint16)      # K-major, 128B-swizzled rows of 64 bf16 (validated layout E0)
    bits = bf16_bytes(A)
    for m in range(Mm):
        for k in range(16):
            off = m * 128 + ((((k >> 3)) ^ (m & 7)) << 4) + (k & 7) * 2
            a_img[off // 2] = bits[m, k]
    b_img, b_offs = b_image_mn_sw128(Bm)
    if Nn == 128:                              # second 64-column chunk of the MN-major B tile: n + 64
        Bm2 = Bm.copy()
        Bm2[:, 1] = np.arange(65, 129) / 256.0
        b2, _ = b_image_mn_sw128(Bm2)
        b_img = np.concatenate([b_img, np.zeros(4096 - b_img.size % 4096 if b_img.size % 4096 else 0, np.uint16)])
        pad = 8192 // 2 - b_img.size
        b_img = np.concatenate([b_img, np.zeros(max(pad, 0), np.uint16), b2])
    a = torch.from_numpy(a_img.view(np.uint8).copy()).cuda()
    b = torch.from_numpy(b_img.view(np.uint8).copy()).cuda()
    bdesc = desc(8192 if Nn == 128 else 16, 1024, 2)      # LBO = distance between 64-column chunks
    out = lib.umma_probe(a, b, [0], [0], desc(16, 1024, 2), bdesc, idesc(Mm, Nn, 0, 1), Nn)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    m_of = np.floor(o + 1e-6).astype(np.int64) - 1
    n_of = np.rint((o - np.floor(o + 1e-6)) * 256).astype(np.int64) - 1
    print(f"== M={Mm} N={Nn}: lanes holding data (value = (m+1) + (n+1)/256)")
    for lane in range(128):
        ms, ns = m_of[lane], n_of[lane]
        ok = (ms >= 0) & (ms < Mm) & (ns >= 0) & (ns < Nn) & (np.abs(o[lane] - (ms + 1) - (ns + 1) / 256.0) < 1e-6)
        if ok.any():
            cols = np.nonzero(ok)[0]
            same_m = len(set(ms[cols].tolist())) == 1
            lin = bool((ns[cols] == cols).all())
            print(f"lane {lane:3d}: {len(cols):3d} cols [{cols[0]}..{cols[-1]}]  m={'%d' % ms[cols[0]] if same_m else sorted(set(ms[cols].tolist()))[:6]}"
                  f"  n(col)={'col' if lin else ns[cols][:8].tolist()}")
print("probe done")
