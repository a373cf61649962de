"""Pins the shared-memory operand layout semantics of tcgen05.mma on a B200 with fb_umma_probe:
SWIZZLE_NONE K-major / MN-major A operands, which of LBO / SBO strides which direction, and whether a 'skewed'
SBO (group stride +- one core-matrix column) reads the linear address the formula predicts.  Prints one line per
(experiment, hypothesis): MATCH / mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from frozenbilm_b200 import lib

lib.use_debug_library()   # test-only build with the probe / phase-stamp hooks

M, N = 128, 64


def bf16_bytes(x):
    """float array of bf16-exact values -> uint16 bit patterns"""
    return (np.asarray(x, np.float32).view(np.uint32) >> 16).astype(np.uint16)


def bf16_val(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def desc(lbo, sbo, layout):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | layout << 61


def idesc(m, n, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def b_image_mn_sw128(Bm):
    """B[n, k] as the validated MN-major 128B-swizzled tile [k rows][64 n]: returns (uint16 image, per-kstep offsets)"""
    Kt = Bm.shape[1]
    img = np.zeros(Kt * 64, np.uint16)
    bits = bf16_bytes(Bm)
    for k in range(Kt):
        for n in range(N):
            off = k * 128 + (((n >> 3) ^ (k & 7)) << 4) + (n & 7) * 2
            img[off // 2] = bits[n, k]
    return img, [ks * 2048 for ks in range(Kt // 16)]


def run(a_img16, a_offs, a_desc, a_mn, Bm):
    b_img, b_offs = b_image_mn_sw128(Bm)
    a = torch.from_numpy(a_img16.view(np.uint8).copy()).cuda()
    b = torch.from_numpy(b_img.view(np.uint8).copy()).cuda()
    out = lib.umma_probe(a, b, a_offs, b_offs, a_desc, desc(16, 1024, 2), idesc(M, N, a_mn, 1), N)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def expect(a_img16, addr, Kt, Bm):
    """D[m, n] = sum_k A[m, k] B[n, k] with A[m, k] read from the image at byte address addr(m, k)"""
    A = np.zeros((M, Kt), np.float32)
    vals = bf16_val(a_img16)
    for m in range(M):
        for k in range(Kt):
            A[m, k] = vals[addr(m, k) // 2]
    return A @ Bm.T


def report(name, got, hyps):
    for hname, exp in hyps:
        ok = np.array_equal(got, exp)
        print(f"{name:34s} {hname:46s} {'MATCH' if ok else 'mismatch (max diff %.3g)' % np.abs(got - exp).max()}", flush=True)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    want = lambda tag: (not only) or only == tag
    rng = np.random.default_rng(0)
    Kt = 256
    Bm = rng.integers(-3, 4, size=(N, Kt)).astype(np.float32)
    nbytes = 131072
    rnd = bf16_bytes(rng.integers(-3, 4, size=nbytes // 2).astype(np.float32))

    # E0: the validated SWIZZLE_128B K-major layout (sanity check of the bench itself)
    sw = lambda m, k: (k >> 6) * 16384 + m * 128 + ((((k & 63) >> 3) ^ (m & 7)) << 4) + (k & 7) * 2
    if want("E0"):
        got = run(rnd, [(ks >> 2) * 16384 + (ks & 3) * 32 for ks in range(Kt // 16)], desc(16, 1024, 2), 0, Bm)
        report("E0 sw128 K-major", got, [("validated layout", expect(rnd, sw, Kt, Bm))])

    # E1: SWIZZLE_NONE K-major, plain strides: Sk (K-direction core-matrix stride) = 128, Sm (8-row group stride) = 4096
    Sk, Sm = 128, 4096
    plain = lambda m, k: (m >> 3) * Sm + (m & 7) * 16 + (k >> 3) * Sk + (k & 7) * 2
    offs = [ks * 2 * Sk for ks in range(Kt // 16)]
    for nm, d in (("LBO=K-stride SBO=M-stride", desc(Sk, Sm, 0)), ("LBO=M-stride SBO=K-stride", desc(Sm, Sk, 0))):
        if not want("E1"):
            break
        try:
            got = run(rnd, offs, d, 0, Bm)
            report("E1 none K-major " + nm, got, [("addr=(m>>3)Sm+(m&7)16+(k>>3)Sk+(k&7)2", expect(rnd, plain, Kt, Bm))])
        except Exception as ex:  # noqa: BLE001
            print("E1", nm, "EXCEPTION", repr(ex)[:200])

    # E2/E3: skewed group stride: the MMA sees row group G shifted by +-G core-matrix columns
    for sign in (+1, -1):
        if not want("E2" if sign > 0 else "E3"):
            continue
        base = 16 * Sk if sign < 0 else 0
        skew = lambda m, k, s=sign, b=base: b + (m >> 3) * (Sm + s * Sk) + (m & 7) * 16 + (k >> 3) * Sk + (k & 7) * 2
        offs = [base + ks * 2 * Sk for ks in range(Kt // 16)]
        K2 = 192   # keep every address inside the image
        for nm, d in (("LBO=K SBO=M", desc(Sk, Sm + sign * Sk, 0)), ("LBO=M SBO=K", desc(Sm + sign * Sk, Sk, 0))):
            try:
                got = run(rnd, offs[:K2 // 16], d, 0, Bm[:, :K2].copy())
                report(f"E{2 if sign > 0 else 3} none K-major skew {sign:+d} " + nm, got,
                       [("linear address formula", expect(rnd, skew, K2, Bm[:, :K2]))])
            except Exception as ex:  # noqa: BLE001
                print("E2/3", nm, "EXCEPTION", repr(ex)[:200])

    # E4: SWIZZLE_NONE MN-major A: element (m, k) at (m>>3)*X + (m&7)*2 + (k>>3)*Y + (k&7)*16
    X, Y = 128, 4096      # X: stride between 8-element M groups, Y: stride between 8-k-row groups
    K4 = 128
    mn = lambda m, k: (m >> 3) * X + (m & 7) * 2 + (k >> 3) * Y + (k & 7) * 16
    offs = [ks * 2 * Y for ks in range(K4 // 16)]
    for nm, d in (("LBO=M-stride SBO=K-stride", desc(X, Y, 0)), ("LBO=K-stride SBO=M-stride", desc(Y, X, 0))):
        if not want("E4"):
            break
        try:
            got = run(rnd, offs, d, 1, Bm[:, :K4].copy())
            report("E4 none MN-major " + nm, got, [("addr=(m>>3)X+(m&7)2+(k>>3)Y+(k&7)16", expect(rnd, mn, K4, Bm[:, :K4]))])
        except Exception as ex:  # noqa: BLE001
            print("E4", nm, "EXCEPTION", repr(ex)[:200])

    # E5: MN-major, skewed: storage written K-major-style (row group G = k>>3 of the MN view is the 'a' group of Z):
    # element (m = c, k = a) at (a>>3)*(Sm - Sk) + (a&7)*16 + (c>>3)*Sk + (c&7)*2   [the Z1 buffer read as Z1^T]
    K5 = 128
    zt = lambda m, k: (k >> 3) * (Sm + Sk) + (k & 7) * 16 + (m >> 3) * Sk + (m & 7) * 2
    offs = [ks * 2 * (Sm + Sk) for ks in range(K5 // 16)]
    for nm, d in (("LBO=M-stride SBO=K-stride", desc(Sk, Sm + Sk, 0)), ("LBO=K-stride SBO=M-stride", desc(Sm + Sk, Sk, 0))):
        if not want("E5"):
            break
        try:
            got = run(rnd, offs, d, 1, Bm[:, :K5].copy())
            report("E5 none MN-major skewed " + nm, got, [("Z^T view of the skewed K-major buffer", expect(rnd, zt, K5, Bm[:, :K5]))])
        except Exception as ex:  # noqa: BLE001
            print("E5", nm, "EXCEPTION", repr(ex)[:200])
    print("probe done")


if __name__ == "__main__":
    main()
