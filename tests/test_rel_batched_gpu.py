"""The positional tables of ALL layer passes built up front (engine._rel_tables_all: batched dropout, one batched
GEMM over the layers, batched table expansion) against the per-pass path (three launches per pass inside the chain):
the kernels must be bit-identical per pass, and a train step must give the same loss and gradients either way.
Reference math: deberta.py:779 (dropout of the relative embeddings), :847-853 (Linear_{q|k}), :578-618 (buckets)."""
import pytest
import torch

from frozenbilm_b200.config import CONFIG_TINY
from frozenbilm_b200.synth import synth_batch, synth_params

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_dropout_batched_matches_per_pass():
    from frozenbilm_b200 import lib
    torch.manual_seed(0)
    M, N, nb = 64, 256, 5
    x = torch.randn(M, N, device=DEV)
    out16 = torch.empty(nb, M, N, device=DEV, dtype=torch.bfloat16)
    lib.dropout_batched(x, lib.drop(0.1, 1234, 3), 8, out16=out16)
    for b in range(nb):
        ref = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        lib.dropout_apply(x, lib.drop(0.1, 1234, 3 + 8 * b), out16=ref)
        assert torch.equal(ref, out16[b]), b
    assert not torch.equal(out16[0], out16[1])
    # reduce direction: out += sum_b dropout_b(g[b])
    g = torch.randn(nb, M, N, device=DEV)
    acc = torch.full((M, N), 0.5, device=DEV)
    lib.dropout_batched(g, lib.drop(0.1, 1234, 3), 8, out32_accum=acc)
    ref = torch.full((M, N), 0.5, device=DEV)
    for b in range(nb):
        lib.dropout_apply(g[b].contiguous(), lib.drop(0.1, 1234, 3 + 8 * b), out32=ref, accumulate=True)
    torch.testing.assert_close(acc, ref, rtol=1e-6, atol=1e-6)
    # p = 0: plain cast / sum
    lib.dropout_batched(x, lib.drop(0.0, 1, 0), 8, out16=out16)
    assert torch.equal(out16[2], x.to(torch.bfloat16))


def test_gemm_batched_bias_stride():
    from frozenbilm_b200 import lib
    torch.manual_seed(1)
    nb, M, N, K = 3, 128, 192, 256
    A = torch.randn(nb, M, K, device=DEV).to(torch.bfloat16)
    Wall = torch.randn(nb, N + 64, K, device=DEV).to(torch.bfloat16)       # rows [0, N) of every slab are used
    ball = torch.randn(nb, N + 64, device=DEV)
    D = torch.empty(nb, M, N, device=DEV, dtype=torch.bfloat16)
    lib.gemm(A[0], Wall[0][:N], D[0], bias=ball[0][:N],
             batch={"outer": (nb, M * K, (N + 64) * K, M * N), "bias_stride": N + 64})
    for b in range(nb):
        ref = A[b].float() @ Wall[b, :N].float().t() + ball[b, :N]
        torch.testing.assert_close(D[b].float(), ref, rtol=2e-2, atol=2e-2)


def test_rel_tables_batched_match_per_pass():
    from frozenbilm_b200 import lib
    from frozenbilm_b200.relpos import bucket_of_rel
    torch.manual_seed(2)
    h, S2, L, nb = 4, 64, 200, 3
    H = h * 64
    bor = torch.from_numpy(bucket_of_rel(S2 // 2, 512)).to(DEV)
    pqk = torch.randn(nb, S2, 2 * H, device=DEV).to(torch.bfloat16)
    rel_rows, _ = lib.rel_table_geometry(L)
    pq, pk, pkr = (torch.empty(nb, h, rel_rows, 64, device=DEV, dtype=torch.bfloat16) for _ in range(3))
    lib.build_rel_tables_batched(pqk, bor, L, h, pq, pk, pkr)
    for b in range(nb):
        r_pkr = torch.empty(h, rel_rows, 64, device=DEV, dtype=torch.bfloat16)
        r_pq, r_pk = lib.build_rel_tables(pqk[b], bor, L, h, out_pk_rev=r_pkr)
        assert torch.equal(r_pq, pq[b]) and torch.equal(r_pk, pk[b]) and torch.equal(r_pkr, pkr[b])
    # and the fold back onto the bucket rows
    gpq = torch.randn(nb, h, rel_rows, 64, device=DEV)
    gpk = torch.randn(nb, h, rel_rows, 64, device=DEV)
    out = torch.zeros(nb, S2, 2 * H, device=DEV)
    lib.rel_to_bucket_batched(gpq, gpk, bor, L, h, out)
    for b in range(nb):
        ref = torch.zeros(S2, 2 * H, device=DEV)
        lib.rel_to_bucket(gpq[b], gpk[b], bor, L, h, ref)
        torch.testing.assert_close(out[b], ref, rtol=1e-5, atol=1e-5)     # atomics: summation order differs


@pytest.mark.parametrize("train", [True, False])
def test_step_identical_with_and_without_batched_tables(train):
    from tests.test_model_gpu import build, to_dev
    cfg = CONFIG_TINY
    P = synth_params(cfg, 7)
    batch = synth_batch(cfg, 2, 118, 11, n_frames=10)
    res = []
    for batched in (True, False):
        m = build(cfg, P)
        m.train(train)
        eng = m._ensure_engine()
        eng.rel_batched = batched
        out = m(**to_dev(batch), labels=batch["labels"].to(DEV))
        loss = out["loss"]
        grads = None
        if train:
            loss.backward()
            grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        res.append((float(loss.detach()), grads))
    # the cross-entropy sums its rows with fp32 atomics: equal up to summation order
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[1][0])
    if train:
        assert res[0][1].keys() == res[1][1].keys() and len(res[0][1]) > 10
        for n in res[0][1]:
            # dW / positional gradients are reduced with fp32 atomics: same values up to summation order
            torch.testing.assert_close(res[0][1][n], res[1][1][n], rtol=1e-4, atol=1e-7, msg=n)
