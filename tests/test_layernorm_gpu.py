"""fb_layernorm_fwd / fb_layernorm_bwd (through the C-ABI) against torch autograd of the same op:
y = LayerNorm(x) * rowmask (deberta.py:1043-1054, 259, 333), dx / dgamma / dbeta from y.backward(dy); the bf16 copy of
dx and its column sums (the adapter up-projection bias gradient) against the fp32 dx.  Ragged row counts (not a
multiple of the 4 rows a CTA walks per step, fewer rows than SMs, many rows per persistent CTA)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("M,H", [(3, 256), (130, 768), (1642, 1536), (11264, 1536), (777, 2048), (5000, 1024)])
@pytest.mark.parametrize("masked", [False, True])
def test_layernorm_fwd_bwd_match_autograd(M, H, masked):
    from frozenbilm_b200 import lib
    g = torch.Generator(device=DEV).manual_seed(M + H)
    x = torch.randn(M, H, device=DEV, generator=g) * 1.7 + 0.3
    gamma = torch.randn(H, device=DEV, generator=g)
    beta = torch.randn(H, device=DEV, generator=g)
    dy = torch.randn(M, H, device=DEV, generator=g)
    rowmask = (torch.rand(M, device=DEV, generator=g) > 0.2).to(torch.uint8) if masked else None
    eps = 1e-7
    y32 = torch.empty(M, H, device=DEV)
    y16 = torch.empty(M, H, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    lib.layernorm_fwd(x, gamma, beta, eps, rowmask=rowmask, y32=y32, y16=y16, mean=mean, rstd=rstd)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (H,), gr, br, eps)
    if masked:
        ref = ref * rowmask[:, None].float()
    torch.testing.assert_close(y32, ref.detach(), rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(y16.float(), ref.detach(), rtol=1e-2, atol=1e-2)
    ref.backward(dy)
    dx32 = torch.empty(M, H, device=DEV)
    dx16 = torch.empty(M, H, device=DEV, dtype=torch.bfloat16)
    dgamma, dbeta, csum = (torch.zeros(H, device=DEV) for _ in range(3))
    lib.layernorm_bwd(dy, x, mean, rstd, gamma, rowmask=rowmask, dx32=dx32, dx16=dx16, dgamma=dgamma, dbeta=dbeta,
                      dx16_colsum=csum)
    scale = xr.grad.abs().max().item()
    assert (dx32 - xr.grad).abs().max().item() <= 2e-5 * max(scale, 1.0) * 10
    torch.testing.assert_close(dgamma, gr.grad, rtol=1e-4, atol=1e-4 * max(1.0, M ** 0.5))
    torch.testing.assert_close(dbeta, br.grad, rtol=1e-4, atol=1e-4 * max(1.0, M ** 0.5))
    assert torch.equal(dx16, dx32.to(torch.bfloat16))
    torch.testing.assert_close(csum, dx16.float().sum(0), rtol=1e-4, atol=1e-4 * max(1.0, M ** 0.5))
    # gamma / beta gradients only (the relative-embedding LayerNorm: no dx requested)
    dg2, db2 = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    lib.layernorm_bwd(dy, x, mean, rstd, gamma, rowmask=rowmask, dgamma=dg2, dbeta=db2)
    torch.testing.assert_close(dg2, dgamma, rtol=1e-5, atol=1e-5 * max(1.0, M ** 0.5))
    torch.testing.assert_close(db2, dbeta, rtol=1e-5, atol=1e-5 * max(1.0, M ** 0.5))


def test_layernorm_bwd_dropout_sites_regenerate_forward_masks():
    """out_drop (the dropout applied to the LayerNorm OUTPUT in the forward) and in_drop (the hidden dropout in front of
    the residual sum, applied to the bf16 copy of dx) use the same counter RNG as the forward kernels."""
    from frozenbilm_b200 import lib
    M, H = 300, 1536
    g = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn(M, H, device=DEV, generator=g)
    gamma, beta = torch.randn(H, device=DEV, generator=g), torch.randn(H, device=DEV, generator=g)
    dy = torch.randn(M, H, device=DEV, generator=g)
    od, idr = lib.drop(0.1, 77, 5), lib.drop(0.1, 77, 6)
    y_plain, y_drop = torch.empty(M, H, device=DEV), torch.empty(M, H, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    lib.layernorm_fwd(x, gamma, beta, 1e-7, y32=y_plain, mean=mean, rstd=rstd)
    lib.layernorm_fwd(x, gamma, beta, 1e-7, out_drop=od, y32=y_drop, mean=mean, rstd=rstd)
    fmask = torch.where(y_plain != 0, y_drop / y_plain, torch.ones_like(y_plain))   # 0 or 1 / keep
    dx_ref, dx_got = torch.empty(M, H, device=DEV), torch.empty(M, H, device=DEV)
    lib.layernorm_bwd((dy * fmask).contiguous(), x, mean, rstd, gamma, dx32=dx_ref)
    dx16 = torch.empty(M, H, device=DEV, dtype=torch.bfloat16)
    lib.layernorm_bwd(dy, x, mean, rstd, gamma, out_drop=od, in_drop=idr, dx32=dx_got, dx16=dx16)
    torch.testing.assert_close(dx_got, dx_ref, rtol=1e-5, atol=1e-5)
    imask = torch.empty(M, H, device=DEV)
    lib.dropout_apply(torch.ones(M, H, device=DEV), idr, out32=imask)
    assert torch.equal(dx16, (dx_got * imask).to(torch.bfloat16))
