"""Oracle of the input-side kernels (oracle/data_oracle.py) against the reference's golden output and the
reference's stated semantics (util/misc.py:14-56, datasets/videotext_dataset.py:23-44)."""
import os

import numpy as np

from oracle import data_oracle as dor

GOLD = os.path.join(os.path.dirname(__file__), "golden", "frames_golden.npz")


def test_gather_frames_oracle_matches_reference_golden():
    g = np.load(GOLD)
    n = g["nframes"]
    off = np.concatenate([[0], np.cumsum(n)[:-1]])
    video, vlen = dor.gather_frames(g["shard"], off, n, 10)
    assert np.array_equal(vlen, g["video_len"])
    assert np.array_equal(video, g["video"].astype(np.float32))      # fp16 -> fp32 is exact: bit-exact parity


def test_gather_frames_edge_cases():
    shard = np.arange(5 * 4, dtype=np.float16).reshape(5, 4)
    v, l = dor.gather_frames(shard, [0, 5, 0], [5, 0, 2], 3)
    assert l.tolist() == [3, 0, 2]
    assert np.array_equal(v[0], shard[[0, 1, 3]].astype(np.float32))   # (j * 5) // 3 = 0, 1, 3
    assert not v[1].any()
    assert np.array_equal(v[2, :2], shard[:2].astype(np.float32)) and not v[2, 2].any()


def test_mlm_corrupt_semantics():
    rng = np.random.default_rng(0)
    V, MASK, PAD = 128100, 128000, 0
    ids = rng.integers(5, V, size=(64, 342)).astype(np.int64)
    ids[:, 0] = 1
    ids[:, -1] = 2
    ids[10:20, 300:] = PAD
    special = (ids == 1) | (ids == 2)
    out, labels = dor.mlm_corrupt(ids, special.astype(np.uint8), PAD, MASK, V, 0.15, seed=77)
    sel = labels != -100
    assert not sel[special].any() and not sel[ids == PAD].any()          # never special / pad
    assert np.array_equal(labels[sel], ids[sel]) and np.array_equal(out[~sel], ids[~sel])
    elig = ~special & (ids != PAD)
    rate = sel.sum() / elig.sum()
    assert abs(rate - 0.15) < 0.01
    frac_mask = (out[sel] == MASK).mean()
    changed = (out[sel] != ids[sel]) & (out[sel] != MASK)
    assert abs(frac_mask - 0.8) < 0.02 and abs(changed.mean() - 0.1) < 0.015
    assert out.min() >= 0 and out.max() < V
    # deterministic in (seed, position); different seeds decorrelate
    out2, labels2 = dor.mlm_corrupt(ids, special.astype(np.uint8), PAD, MASK, V, 0.15, seed=77)
    assert np.array_equal(out, out2) and np.array_equal(labels, labels2)
    _, labels3 = dor.mlm_corrupt(ids, special.astype(np.uint8), PAD, MASK, V, 0.15, seed=78)
    both = ((labels3 != -100) & sel).sum() / sel.sum()
    assert abs(both - 0.15) < 0.02
    # probability 0 / 1
    _, l0 = dor.mlm_corrupt(ids, None, None, MASK, V, 0.0, seed=1)
    assert (l0 == -100).all()
    _, l1 = dor.mlm_corrupt(ids, None, None, MASK, V, 1.0, seed=1)
    assert (l1 != -100).mean() > 0.999999


def test_feature_shard_packing_and_roundtrip(tmp_path):
    """Host side of the feature reader: packing, offsets, missing videos, save / load (no GPU involved)."""
    import torch
    from frozenbilm_b200.data import FeatureShard
    rng = np.random.default_rng(1)
    arrays = [rng.standard_normal((3, 8)).astype(np.float16), None, rng.standard_normal((5, 8)).astype(np.float16),
              np.zeros((0, 8), np.float16)]
    fs = FeatureShard.from_arrays(arrays, features_dim=8, device="cpu")
    assert fs.nframes.tolist() == [3, 0, 5, 0] and fs.offsets.tolist() == [0, 3, 3, 8]
    assert fs.shard.dtype == torch.float16 and fs.shard.shape == (8, 8)
    assert np.array_equal(fs.shard[3:8].numpy(), arrays[2])
    p = tmp_path / "shard.npz"
    fs.save(p)
    fs2 = FeatureShard.load(p, device="cpu")
    assert torch.equal(fs2.shard, fs.shard) and torch.equal(fs2.offsets, fs.offsets) and torch.equal(fs2.nframes, fs.nframes)
    # a suffix-less path round-trips too (np.savez appends ".npz")
    fs.save(tmp_path / "shard2")
    assert torch.equal(FeatureShard.load(tmp_path / "shard2", device="cpu").shard, fs.shard)
    # the oracle on the packed representation reproduces the per-video rule
    v, l = dor.gather_frames(fs2.shard.numpy(), fs2.offsets.numpy(), fs2.nframes.numpy(), 4)
    assert l.tolist() == [3, 0, 4, 0] and np.array_equal(v[2], arrays[2][[0, 1, 2, 3]].astype(np.float32))
    # files: a missing path becomes an empty video (reference videotext_dataset.py:39-41)
    np.save(tmp_path / "7.mp4.npy", arrays[0])
    fs3 = FeatureShard.from_files([str(tmp_path / "7.mp4.npy"), str(tmp_path / "missing.mp4.npy")], features_dim=8, device="cpu")
    assert fs3.nframes.tolist() == [3, 0]
