"""Host-side logic of the path that runs without a GPU: the relative-position bucket table against the oracle's
restatement of build_relative_position (reference deberta.py:578-618), geometry helpers, the freezing rules and the
flat trainable bucket layout (reference deberta.py:1152-1158, 1334-1339)."""
import numpy as np
import pytest

from frozenbilm_b200.config import CONFIG_TINY, CONFIG_XL
from frozenbilm_b200.relpos import bucket_of_rel
from frozenbilm_b200.synth import is_trainable, synth_params
from oracle import frozenbilm_oracle as orc


@pytest.mark.parametrize("L", [34, 138, 352, 512])
def test_bucket_table_matches_relative_position_of_the_reference(L):
    S, MAXP = 256, 512
    tab = bucket_of_rel(S, MAXP)
    assert tab.shape == (1023,) and tab.dtype == np.int32
    rel = orc.build_relative_position(L, L, S, MAXP)          # [1, L, L] bucketed i - j (deberta.py:592-618)
    rel = np.asarray(rel).reshape(L, L)
    i, j = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    c2p_index = np.clip(rel + S, 0, 2 * S - 1)                 # deberta.py:872
    assert np.array_equal(tab[(i - j) + 511], c2p_index)
    # odd symmetry: the p2c index (deberta.py:898, built from -rel) is the mirrored table
    assert np.array_equal(np.clip(-rel + S, 0, 2 * S - 1), tab[(j - i) + 511])


def test_rel_table_geometry_covers_every_tile_band():
    from frozenbilm_b200.lib import rel_table_geometry
    for L in (1, 34, 128, 129, 352, 512):
        rows, off = rel_table_geometry(L)
        lp = (L + 127) // 128 * 128
        assert off >= lp and rows >= off + lp + 1 and rows % 8 == 0
        # the 256-row band of any (query tile, key block) pair stays inside the table
        for i0 in range(0, lp, 128):
            for j0 in range(0, lp, 128):
                band0 = i0 - j0 - 127 + off
                assert 0 <= band0 and band0 + 256 <= rows


def test_freezing_rules_and_trainable_count_xl():
    # reference: trainable <=> adapter / linear_video / LayerNorm (ft_ln); head: only lm_head.LayerNorm
    names = {
        "deberta.embeddings.linear_video.weight": True,
        "deberta.embeddings.word_embeddings.weight": False,
        "deberta.embeddings.LayerNorm.weight": True,
        "deberta.encoder.layer.3.attention.output.adapter.down.weight": True,
        "deberta.encoder.layer.3.attention.self.query_proj.weight": False,
        "deberta.encoder.layer.3.output.LayerNorm.bias": True,
        "deberta.encoder.rel_embeddings.weight": False,
        "deberta.encoder.LayerNorm.weight": True,
        "lm_predictions.lm_head.LayerNorm.weight": True,
        "lm_predictions.lm_head.dense.weight": False,
        "lm_predictions.lm_head.bias": False,
        "answer_embeddings.weight": False,
    }
    for n, want in names.items():
        assert is_trainable(n, CONFIG_XL) == want, n
    P = synth_params(CONFIG_TINY, 1)
    n_train = sum(v.numel() for k, v in P.items() if is_trainable(k, CONFIG_TINY))
    assert 0 < n_train < sum(v.numel() for v in P.values()) // 4
