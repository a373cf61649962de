"""CPU oracle (TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product path) for the input-side kernels
of frozenbilm_b200/csrc/fb_data.cu.

* ``gather_frames``  restates reference datasets/videotext_dataset.py:23-44 (+ collate :47-57) on a packed shard.
  Pinned: tests/golden/frames_golden.npz was produced by the UNMODIFIED reference ``VideoText_Dataset.__getitem__``
  / ``videotext_collate_fn`` on .npy files written to a temp directory (tests/golden/make_data_golden.py).
* ``mlm_corrupt``    restates reference util/misc.py:14-56 (mask_tokens) with a counter-based RNG instead of torch's
  global generator.  Parity with the reference is DISTRIBUTIONAL (selection rate, 80/10/10 split, special/pad tokens
  never selected, labels = original ids / -100): the reference's draws are not reproducible outside torch.  The CUDA
  kernel is held bit-exactly to THIS function (same integer arithmetic).
"""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def _mix32(x):
    x = x & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x21F0AAAD)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x735A2D97)) & _M32
    x ^= x >> np.uint64(15)
    return x


def _data_hash(idx, seed, stream):
    """fb_data.cu: data_hash (uint32 arithmetic carried out in uint64 and masked)."""
    idx = idx.astype(np.uint64)
    lo, hi = idx & _M32, idx >> np.uint64(32)
    seed_lo, seed_hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    st = np.uint64(stream)
    x = (lo * np.uint64(0x9E3779B1) + seed_lo) & _M32
    x ^= (((hi + st) & _M32) * np.uint64(0x85EBCA77) + seed_hi) & _M32
    x = _mix32(x)
    x = _mix32((x + st * np.uint64(0xC2B2AE3D) + np.uint64(0x27D4EB2F)) & _M32)
    return x


def mlm_corrupt(ids, special_mask, pad_id, mask_id, vocab, mlm_probability, seed):
    """ids: int64 array (any shape).  Returns (corrupted ids, labels) like util/misc.py:14-56."""
    ids = np.asarray(ids, dtype=np.int64)
    flat = ids.reshape(-1)
    idx = np.arange(flat.size, dtype=np.uint64)
    eligible = np.ones(flat.size, dtype=bool)
    if special_mask is not None:
        eligible &= np.asarray(special_mask).reshape(-1) == 0
    if pad_id is not None and pad_id >= 0:
        eligible &= flat != pad_id
    t = float(np.float32(mlm_probability)) * 4294967296.0
    thresh = np.uint64(0xFFFFFFFF if t >= 4294967295.0 else int(t))
    u0 = _data_hash(idx, seed, 0)
    masked = eligible & (u0 < thresh)
    u1 = _data_hash(idx, seed, 1)
    u2 = _data_hash(idx, seed, 2)
    out = flat.copy()
    to_mask = masked & (u1 < np.uint64(3435973837))
    to_rand = masked & (u1 >= np.uint64(3865470566))
    out[to_mask] = mask_id
    out[to_rand] = ((u2[to_rand] * np.uint64(vocab)) >> np.uint64(32)).astype(np.int64)
    labels = np.where(masked, flat, -100).astype(np.int64)
    return out.reshape(ids.shape), labels.reshape(ids.shape)


def gather_frames(shard, offsets, nframes, max_feats):
    """shard: float16 [rows, D]; returns (video float32 [B, max_feats, D], video_len int64 [B])."""
    shard = np.asarray(shard)
    B, D = len(nframes), shard.shape[1]
    out = np.zeros((B, max_feats, D), dtype=np.float32)
    vlen = np.zeros(B, dtype=np.int64)
    for b in range(B):
        n = int(nframes[b])
        video = shard[int(offsets[b]):int(offsets[b]) + n].astype(np.float32)
        if n > max_feats:                                    # videotext_dataset.py:26-31
            for j in range(max_feats):
                out[b, j] = video[(j * n) // max_feats]
            vlen[b] = max_feats
        elif n > 0:                                          # :32-38 (n < max_feats: zero padded) and n == max_feats
            out[b, :n] = video
            vlen[b] = n
        # n == 0: missing / corrupted file -> zeros, video_len 0  (:39-41)
    return out, vlen
