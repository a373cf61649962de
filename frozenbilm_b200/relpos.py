"""Relative-position bucket table for the fused attention (host side, computed once).

bucket_of_rel[rel + 511] = clamp(bucket(rel) + S, 0, 2S - 1) for rel = i - j in [-511, 511], where
bucket() is DeBERTa-V2's log-bucket map (reference model/deberta.py:578-589, make_log_bucket_position
with bucket_size = position_buckets, max_position = max_position_embeddings) and the +S / clamp is
deberta.py:872 (c2p_pos) -- the p2c index (deberta.py:898) is the same table by odd symmetry.
The reference rebuilds this [L, L] table with numpy on the CPU every forward; it only depends on
i - j, so one 1023-entry table serves every sequence length.
"""
import numpy as np


def bucket_of_rel(position_buckets: int, max_position: int) -> np.ndarray:
    S = position_buckets
    rel = np.arange(-511, 512, dtype=np.int64)
    mid = S // 2
    sign = np.sign(rel)
    abs_pos = np.where((rel < mid) & (rel > -mid), mid - 1, np.abs(rel)).astype(np.float64)
    log_pos = np.ceil(np.log(abs_pos / mid) / np.log((max_position - 1) / mid) * (mid - 1)) + mid
    bucket = np.where(abs_pos <= mid, rel, log_pos * sign).astype(np.int64)
    return np.clip(bucket + S, 0, 2 * S - 1).astype(np.int32)
