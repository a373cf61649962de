"""tcgen05 GEMM (fb_gemm through the C-ABI) against a torch fp32 reference of the same op.

Operands are bf16 on both sides, so the only difference is fp32 accumulation order:
tolerance 2e-3 relative to the row scale (sqrt(K) * |a| * |b|) -- see the test body."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, a_mn, b_mn, bias=None, act=0, aux=None, aux_op=0, resid=None, alpha=1.0):
    Af = A.float().transpose(-1, -2) if a_mn else A.float()
    Bf = B.float().transpose(-1, -2) if b_mn else B.float()
    v = alpha * (Af @ Bf.transpose(-1, -2))
    if bias is not None:
        v = v + bias
    pre = v.clone()
    if act == 1:
        v = torch.nn.functional.gelu(v)
    elif act == 2:
        v = torch.relu(v)
    if aux is not None:
        a = aux.float()
        if aux_op == 1:
            x = a
            cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
            pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
            v = v * (cdf + x * pdf)
        elif aux_op == 2:
            v = v * (a > 0)
        else:
            v = v + a
    if resid is not None:
        v = v + resid
    return v, pre


SHAPES = [
    # M, N, K
    (128, 256, 64), (128, 128, 128), (256, 512, 256), (300, 200, 136), (1000, 1536, 1536),
    (77, 64, 72), (2816, 192, 1536), (2816, 1536, 192), (512, 1000, 256), (64, 1000, 776),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("block_n", [0, 64, 128, 192, 256])
def test_gemm_layouts(M, N, K, a_mn, b_mn, block_n):
    from frozenbilm_b200 import lib
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    dev = "cuda"

    def mk(r, c, mn):
        # leading dims padded to a multiple of 8 elements (TMA needs 16-byte row pitch)
        rows, cols = (c, r) if mn else (r, c)
        ld = (cols + 7) // 8 * 8
        t = torch.randn(rows, ld, device=dev, generator=g).bfloat16()
        return t[:, :cols]

    A = mk(M, K, a_mn)
    B = mk(N, K, b_mn)
    ldd = (N + 7) // 8 * 8
    D = torch.full((M, ldd), float("nan"), device=dev, dtype=torch.bfloat16)[:, :N]
    lib.gemm(A, B, D, a_mn_major=a_mn, b_mn_major=b_mn, block_n=block_n)
    ref, _ = _ref(A, B, a_mn, b_mn)
    err = (D.float() - ref).abs().max().item()
    tol = 2e-2 * math.sqrt(K)  # bf16 output rounding of values ~ sqrt(K)
    assert math.isfinite(err) and err < tol, f"max abs err {err} (tol {tol})"
    rel = ((D.float() - ref).norm() / ref.norm()).item()
    assert rel < 4e-3, rel


def test_gemm_epilogues():
    from frozenbilm_b200 import lib
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    M, N, K = 777, 1160, 520
    A = (0.2 * torch.randn(M, K, device=dev, generator=g)).bfloat16()
    B = (0.2 * torch.randn(N, K, device=dev, generator=g)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    aux = torch.randn(M, N, device=dev, generator=g).bfloat16()
    resid = torch.randn(M, N, device=dev, generator=g)
    # bias + gelu, with the pre-activation side output, fp32 out
    D = torch.empty(M, N, device=dev)
    Dp = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    lib.gemm(A, B, D, bias=bias, act=lib.FB_ACT_GELU, D_pre=Dp)
    ref, pre = _ref(A, B, False, False, bias=bias, act=1)
    assert (D - ref).abs().max().item() < 2e-3
    assert (Dp.float() - pre).abs().max().item() < 3e-2
    # relu + residual, bf16 out, alpha
    D = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    lib.gemm(A, B, D, bias=bias, act=lib.FB_ACT_RELU, resid=resid, alpha=0.5)
    ref, _ = _ref(A, B, False, False, bias=bias, act=2, resid=resid, alpha=0.5)
    assert (D.float() - ref).abs().max().item() < 5e-2
    for op in (1, 2, 3):
        D = torch.empty(M, N, device=dev)
        lib.gemm(A, B, D, aux=aux, aux_op=op)
        ref, _ = _ref(A, B, False, False, aux=aux, aux_op=op)
        assert (D - ref).abs().max().item() < 2e-3, op
    # accumulate and split-K (atomics) into fp32
    D0 = torch.randn(M, N, device=dev, generator=g)
    D = D0.clone()
    lib.gemm(A, B, D, accumulate=True)
    ref, _ = _ref(A, B, False, False)
    assert (D - (D0 + ref)).abs().max().item() < 2e-3
    D = D0.clone()
    lib.gemm(A, B, D, split_k=4, bias=bias)
    assert (D - (D0 + ref + bias)).abs().max().item() < 2e-3


def test_gemm_dw_shape_splitk():
    """dW = dY^T X with both operands MN-major and a long K (= tokens), as in the adapter backward."""
    from frozenbilm_b200 import lib
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(11)
    T, H, r = 5000, 1536, 192
    dY = (0.1 * torch.randn(T, r, device=dev, generator=g)).bfloat16()
    X = (0.1 * torch.randn(T, H, device=dev, generator=g)).bfloat16()
    D = torch.zeros(r, H, device=dev)
    lib.gemm(dY, X, D, a_mn_major=True, b_mn_major=True, split_k=16)
    ref = dY.float().t() @ X.float()
    assert (D - ref).abs().max().item() < 5e-3 * math.sqrt(T) * 0.01 + 2e-3


def test_gemm_batched():
    from frozenbilm_b200 import lib
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    b, M, N, K = 6, 352, 352, 64
    A = torch.randn(b, M, K, device=dev, generator=g).bfloat16()
    B = torch.randn(b, N, K, device=dev, generator=g).bfloat16()
    D = torch.empty(b, M, N, device=dev)
    lib.gemm(A, B, D)
    ref = A.float() @ B.float().transpose(1, 2)
    assert (D - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K,block_n", [(5632, 2048, 320, 256), (5632, 1536, 384, 192), (5600, 1500, 200, 256),
                                           (11264, 1536, 192, 0), (5632, 3072, 256, 0)])
@pytest.mark.parametrize("b_mn", [False, True])
def test_gemm_cluster_multicast_path(M, N, K, block_n, b_mn):
    """Shapes with >= 2 x 148 tiles and an even number of 128-row blocks take the 2-CTA cluster kernel (each CTA
    loads half of the shared B tile, multicast to its peer).  Plain, epilogue-operand and ragged-edge cases."""
    from frozenbilm_b200 import lib
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(M + N + K + int(b_mn))
    A = (0.3 * torch.randn(M, K, device=dev, generator=g)).bfloat16()
    ldb = (N + 7) // 8 * 8
    B = (0.3 * torch.randn(K, ldb, device=dev, generator=g)).bfloat16()[:, :N] if b_mn else \
        (0.3 * torch.randn(N, K, device=dev, generator=g)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    ldn = (N + 7) // 8 * 8
    resid = torch.randn(M, ldn, device=dev, generator=g)[:, :N]
    aux = torch.randn(M, ldn, device=dev, generator=g).bfloat16()[:, :N]
    ref, _ = _ref(A, B, False, b_mn, bias=bias)
    tol = 3e-3 * math.sqrt(K) * 0.09 + 1e-3
    D = torch.full((M, ldn), float("nan"), device=dev)[:, :N]
    lib.gemm(A, B, D, b_mn_major=b_mn, bias=bias, block_n=block_n)
    assert (D - ref).abs().max().item() < tol
    # coalesced epilogue operands on top of the cluster kernel
    D = torch.full((M, ldn), float("nan"), device=dev)[:, :N]
    lib.gemm(A, B, D, b_mn_major=b_mn, bias=bias, aux=aux, aux_op=lib.FB_AUX_ADD, resid=resid, block_n=block_n)
    assert (D - (ref + aux.float() + resid)).abs().max().item() < tol
    # bf16 output + GELU with the pre-activation side output
    Db = torch.empty(M, ldn, device=dev, dtype=torch.bfloat16)[:, :N]
    Dp = torch.empty(M, ldn, device=dev, dtype=torch.bfloat16)[:, :N]
    lib.gemm(A, B, Db, b_mn_major=b_mn, bias=bias, act=lib.FB_ACT_GELU, D_pre=Dp, block_n=block_n)
    assert (Dp.float() - ref).abs().max().item() < 3e-2 + tol
    assert (Db.float() - torch.nn.functional.gelu(ref)).abs().max().item() < 3e-2 + tol
    # repeated launches reuse the pipeline barriers across many tiles: results must be identical
    D2 = torch.empty_like(Db)
    lib.gemm(A, B, D2, b_mn_major=b_mn, bias=bias, act=lib.FB_ACT_GELU, D_pre=Dp, block_n=block_n)
    assert torch.equal(D2, Db)
    # GELU' x aux epilogue (the dX of the FFN down-projection; 128 x 192 tiles run it with three epilogue warpgroups)
    refg, _ = _ref(A, B, False, b_mn, aux=aux, aux_op=1)
    lib.gemm(A, B, Db, b_mn_major=b_mn, aux=aux, aux_op=lib.FB_AUX_MUL_GELU_GRAD, block_n=block_n)
    assert (Db.float() - refg).abs().max().item() < 3e-2 + tol
