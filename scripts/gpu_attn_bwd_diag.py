"""Diagnostics for the fused attention backward: per-output, per-128-row-tile error tables against torch autograd
(never asserts; prints everything so that one GPU call localises a bug)."""
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.test_attention_bwd_gpu import reference_grads, run_case


def tile_errs(a, b, L, B):
    """a, b: [B*L, H]; returns list of (tile, max err) over row tiles of 128 inside each sample."""
    H = a.shape[1]
    a, b = a.view(B, L, H), b.view(B, L, H)
    out = []
    for t in range((L + 127) // 128):
        sl = slice(t * 128, min(L, t * 128 + 128))
        out.append(((a[:, sl] - b[:, sl]).abs().max().item(), b[:, sl].abs().max().item()))
    return out


def main():
    cases = [(2, 34, 2, True), (2, 128, 2, True), (2, 352, 3, True), (3, 129, 2, True), (2, 34, 2, False),
             (2, 266, 2, False)]
    for B, L, h, rel in cases:
        try:
            out, ins = run_case(B, L, h, rel)
            ref = reference_grads(ins, B, L, h, rel)
        except Exception as ex:  # noqa: BLE001
            print(f"case B={B} L={L} h={h} rel={rel}: EXCEPTION {ex!r}")
            try:
                torch.cuda.synchronize()
            except Exception as ex2:  # noqa: BLE001
                print("  device error persists:", repr(ex2)[:200])
                return
            continue
        print(f"case B={B} L={L} h={h} rel={rel}")
        for k in ("dq", "dk", "dv"):
            te = tile_errs(out[k], ref[k], L, B)
            print(f"  {k}: " + "  ".join(f"tile{t}: err {e:.4g} / max {m:.4g}" for t, (e, m) in enumerate(te)),
                  " nan:", int(torch.isnan(out[k]).sum()))
        if rel:
            H = h * 64
            a, b = out["dpqk"], ref["dpqk"]
            for name, sl in (("dPQ", slice(0, H)), ("dPK", slice(H, 2 * H))):
                err = (a[:, sl] - b[:, sl]).abs()
                rows = err.max(1).values
                worst = int(rows.argmax())
                print(f"  {name}: max err {err.max().item():.4g} / max {b[:, sl].abs().max().item():.4g} (worst bucket row "
                      f"{worst}); rows with err > 2%: {int((rows > 0.02 * b[:, sl].abs().max()).sum())} nan: {int(torch.isnan(a[:, sl]).sum())}")
    print("done")


if __name__ == "__main__":
    main()
