"""GPU parity of the input-side kernels (fb_mlm_corrupt, fb_gather_frames) through the C-ABI against the CPU oracle
(bit-exact: integer / fp16->fp32 work) and the reference's golden output."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as dor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "frames_golden.npz")


@pytest.mark.parametrize("shape,p,use_special,pad", [((32, 342), 0.15, True, 0), ((1, 7), 0.5, False, None),
                                                     ((5, 128), 1.0, True, 3), ((3, 33), 0.0, False, 0)])
def test_mlm_corrupt_bit_exact_vs_oracle(shape, p, use_special, pad):
    from frozenbilm_b200 import lib
    rng = np.random.default_rng(5)
    V, MASK = 128100, 128000
    ids = rng.integers(0, V, size=shape).astype(np.int64)
    ids[:, -1] = 2
    special = (ids == 2).astype(np.uint8) if use_special else None
    for seed in (0, 12345678901234567):
        ref_out, ref_lab = dor.mlm_corrupt(ids, special, pad, MASK, V, p, seed)
        t = torch.from_numpy(ids).cuda()
        sp = torch.from_numpy(special).cuda() if special is not None else None
        out, lab = lib.mlm_corrupt(t, sp, pad, MASK, V, p, seed)
        assert np.array_equal(out.cpu().numpy(), ref_out)
        assert np.array_equal(lab.cpu().numpy(), ref_lab)


def test_mask_tokens_drop_in_surface():
    from frozenbilm_b200.data import mask_tokens

    class Tok:
        mask_token, pad_token, _pad_token, pad_token_id = "[MASK]", "[PAD]", "[PAD]", 0
        all_special_ids = [0, 1, 2, 3, 128000]

        def __len__(self):
            return 128100

        def convert_tokens_to_ids(self, t):
            return 128000

    g = torch.Generator().manual_seed(0)
    ids = torch.randint(5, 128000, (16, 64), generator=g)
    ids[:, 0], ids[:, -1] = 1, 2
    ids[3, 40:] = 0
    orig = ids.clone()
    dev_ids = ids.cuda()
    out, labels = mask_tokens(dev_ids, Tok(), 0.15, seed=9)
    assert out.data_ptr() == dev_ids.data_ptr()                       # in place, like the reference
    out, labels = out.cpu(), labels.cpu()
    sel = labels != -100
    special = (orig <= 3)
    assert not sel[special].any()
    assert torch.equal(labels[sel], orig[sel]) and torch.equal(out[~sel], orig[~sel])
    assert 0.08 < sel.float().mean() < 0.22

    class NoMask(Tok):
        mask_token = None
    with pytest.raises(ValueError):
        mask_tokens(dev_ids, NoMask(), 0.15)


def test_gather_frames_matches_reference_golden_and_oracle():
    from frozenbilm_b200 import lib
    from frozenbilm_b200.data import FeatureShard
    g = np.load(GOLD)
    n = g["nframes"]
    off = np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64)
    shard = torch.from_numpy(g["shard"]).cuda()
    video, vlen = lib.gather_frames(shard, torch.from_numpy(off).cuda(), torch.from_numpy(n).cuda(), 10)
    assert np.array_equal(vlen.cpu().numpy(), g["video_len"])
    assert np.array_equal(video.cpu().numpy(), g["video"].astype(np.float32))
    # bf16 output = rounding of the same values
    vb = torch.empty(len(n), 10, 768, device="cuda", dtype=torch.bfloat16)
    lib.gather_frames(shard, torch.from_numpy(off).cuda(), torch.from_numpy(n).cuda(), 10, out_bf16=vb)
    assert torch.equal(vb.cpu(), torch.from_numpy(g["video"].astype(np.float32)).to(torch.bfloat16))
    # ragged random case (odd feature width -> scalar path) against the oracle, through the FeatureShard surface
    rng = np.random.default_rng(3)
    arrays = [None if k == 4 else rng.standard_normal((int(c), 100)).astype(np.float16)
              for k, c in enumerate(rng.integers(1, 60, size=12))]
    fs = FeatureShard.from_arrays(arrays, features_dim=100)
    idx = [11, 4, 0, 7, 7]
    b = fs.batch(idx, max_feats=6)
    cnt = [0 if arrays[i] is None else len(arrays[i]) for i in idx]
    rows = [arrays[i] if arrays[i] is not None else np.zeros((0, 100), np.float16) for i in idx]
    ref_v, ref_l = dor.gather_frames(np.concatenate(rows, 0), np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt, 6)
    assert np.array_equal(b["video"].cpu().numpy(), ref_v) and np.array_equal(b["video_len"].cpu().numpy(), ref_l)
