"""Shared helpers for tests that compare against tests/golden/*.npz."""
import os

import numpy as np

from frozenbilm_b200.config import CONFIG_BASE, CONFIG_BERT_TINY, CONFIG_TINY
from frozenbilm_b200.synth import synth_batch, synth_params

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror tests/golden/make_golden.py::CASES
CASES = {
    "tiny_b2_f2_t32": (CONFIG_TINY.replace(max_feats=2), 2, 32, 2, True, "mlm"),
    "tiny_b3_f10_t118": (CONFIG_TINY, 3, 118, 10, True, "mlm"),
    "tiny_b2_f10_t342": (CONFIG_TINY, 2, 342, 10, True, "mlm"),
    "tiny_ans2_b2_f10_t22": (CONFIG_TINY.replace(n_ans=2), 2, 22, 10, True, "ans"),
    "base_b2_f2_t32": (CONFIG_BASE, 2, 32, 2, True, "mlm"),
    # BERT variant (reference model/bert.py): plain attention, absolute positions, no adapters
    "bert_tiny_b3_f10_t118": (CONFIG_BERT_TINY, 3, 118, 10, True, "mlm"),
    "bert_tiny_ans2_b2_f10_t22": (CONFIG_BERT_TINY.replace(n_ans=2), 2, 22, 10, True, "ans"),
}
GRAD_STRIDE = 53


def load_case(name, seed=1234):
    cfg, B, Lt, F, ragged, mode = CASES[name]
    P = synth_params(cfg, seed)
    batch = synth_batch(cfg, B, Lt, seed + 1, n_frames=F, ragged=ragged)
    fx = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    return cfg, P, batch, fx, mode


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))
