"""Pins the oracle restatement (oracle/frozenbilm_oracle.py) to the golden vectors produced by
the UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from frozenbilm_b200.synth import is_trainable
from oracle import frozenbilm_oracle as orc
from tests.golden_util import CASES, GRAD_STRIDE, load_case, max_abs


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    if name.startswith("base") and not torch.get_num_threads() >= 1:
        pytest.skip("")
    cfg, P, batch, fx, mode = load_case(name)
    torch.set_grad_enabled(True)
    kw = dict(video=batch["video"], video_mask=batch["video_mask"], input_ids=batch["input_ids"],
              attention_mask=batch["attention_mask"])
    if mode == "mlm":
        names_t = [n for n in P if is_trainable(n, cfg)]
        out, grads = orc.forward_backward(P, cfg, names_t, labels=batch["labels"], return_hidden=True, **kw)
        assert abs(out["loss"].item() - float(fx["loss"])) < 2e-5 * max(1.0, abs(float(fx["loss"])))
    else:
        with torch.no_grad():
            out = (orc.bert_forward if cfg.arch == "bert" else orc.forward)(P, cfg, return_hidden=True, **kw)
        assert out["loss"] is None
    lg = out["logits"].detach().numpy()
    cols = fx["logits_cols"]
    err = max_abs(lg[..., cols], fx["logits_slice"])
    assert err < 2e-4, f"logits max-abs-err {err}"
    np.testing.assert_allclose(lg[..., cols], fx["logits_slice"], rtol=1e-3, atol=1e-4)
    assert abs(lg.astype(np.float64).sum() - float(fx["logits_sum"])) < 1e-6 * float(fx["logits_abs_sum"]) + 1e-3
    # hidden states x0 .. x_{NL-1} (the reference's x_NL is the discarded pass)
    NL = cfg.num_hidden_layers
    for i in range(NL):
        h = out["hidden"][i].detach().numpy()[:, :, ::37]
        assert max_abs(h, fx["hidden_slice"][i]) < 1e-4, f"hidden[{i}]"
    if mode == "mlm":
        for n in names_t:
            g = grads[n].reshape(-1)
            ref_norm = float(fx["gnorm/" + n])
            assert abs(g.double().norm().item() - ref_norm) <= 1e-3 * ref_norm + 1e-7, n
            np.testing.assert_allclose(g[::GRAD_STRIDE].numpy(), fx["gslice/" + n],
                                       rtol=2e-3, atol=2e-4 * max(ref_norm, 1e-6), err_msg=n)


def test_bucket_function_properties():
    """Odd symmetry and identity zone of the log-bucket map (deberta.py:578-589); L=352 touches
    only buckets -221..221."""
    rel = orc.build_relative_position(352, 352, 256, 512)
    assert (rel == -rel.T).all()
    i, j = np.indices(rel.shape)
    near = np.abs(i - j) <= 128
    assert (rel[near] == (i - j)[near]).all()
    assert rel.min() == -221 and rel.max() == 221
    # monotone in i-j
    d = orc.delta_table(512, 256, 512)
    assert d.min() >= 0 and d.max() <= 511
    first_col = d[:, 0]
    assert (np.diff(first_col) >= 0).all()
