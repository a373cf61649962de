"""Golden vectors for the frame gather, produced by the UNMODIFIED reference
datasets/videotext_dataset.py (VideoText_Dataset.__getitem__ + videotext_collate_fn) on .npy files written to a
temp directory.  Run in the build container (needs /root/reference):  python tests/golden/make_data_golden.py"""
import importlib.util
import os
import sys
import tempfile

import numpy as np
import pandas as pd

REF = "/root/reference/datasets/videotext_dataset.py"
spec = importlib.util.spec_from_file_location("ref_videotext_dataset", REF)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

rng = np.random.default_rng(1234)
D, F = 768, 10
counts = [37, 10, 3, 0, 11, 1, 250, 9]          # more / equal / fewer frames than max_feats, a missing video
with tempfile.TemporaryDirectory() as tmp:
    feats = []
    for vid, n in enumerate(counts):
        if n == 0:
            feats.append(np.zeros((0, D), np.float16))
            continue                               # no file: the reference's except branch
        a = rng.standard_normal((n, D)).astype(np.float16)
        feats.append(a)
        np.save(os.path.join(tmp, f"{vid}.mp4.npy"), a)
    csv = os.path.join(tmp, "d.csv")
    pd.DataFrame({"video_id": list(range(len(counts))), "text": ["t"] * len(counts)}).to_csv(csv, index=False)
    ds = mod.VideoText_Dataset(csv, tmp, max_feats=F, features_dim=D)
    batch = mod.videotext_collate_fn([ds[i] for i in range(len(counts))])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frames_golden.npz")
np.savez_compressed(out, shard=np.concatenate(feats, 0), nframes=np.array(counts, np.int32),
                    video=batch["video"].numpy().astype(np.float16), video_len=batch["video_len"].numpy())
print("wrote", out, os.path.getsize(out), "bytes")
