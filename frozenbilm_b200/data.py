"""Device-side input pipeline of the masked-LM hot path.

Mirrors the two reference functions that prepare a training batch (same names and argument meaning), but the work
runs on the GPU through the C-ABI (fb_mlm_corrupt / fb_gather_frames); there is no CPU fallback.

    mask_tokens(inputs, tokenizer, mlm_probability)        reference util/misc.py:14-56 (called at main.py:52-58)
    FeatureShard(...).batch(video_indices, max_feats)      reference datasets/videotext_dataset.py:18-44 + :47-57
"""
import numpy as np
import torch

from . import lib

_step_seed = [0x5EED0001]


def mask_tokens(inputs: torch.Tensor, tokenizer, mlm_probability: float, seed=None):
    """Same contract as the reference's ``mask_tokens``: returns ``(inputs, labels)`` with 80 % [MASK] / 10 % random /
    10 % unchanged at the selected positions and ``labels = -100`` elsewhere; special and pad tokens are never
    selected.  ``inputs`` must live on the GPU (int64); it is corrupted in place like the reference does.
    ``seed`` (optional) makes the draw reproducible; by default a process-wide counter advances every call."""
    if tokenizer.mask_token is None:
        raise ValueError(
            "This tokenizer does not have a mask token which is necessary for masked language modeling. Remove the "
            "--mlm flag if you want to use this tokenizer.")
    if not inputs.is_cuda:
        raise lib.FbError("mask_tokens: inputs must be a CUDA tensor (there is no CPU path)")
    if seed is None:
        _step_seed[0] = (_step_seed[0] * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        seed = _step_seed[0]
    special_ids = torch.tensor(sorted(set(tokenizer.all_special_ids)), device=inputs.device, dtype=torch.int64)
    special = torch.isin(inputs, special_ids).to(torch.uint8).contiguous()
    pad_id = tokenizer.pad_token_id if getattr(tokenizer, "_pad_token", tokenizer.pad_token) is not None else None
    mask_id = tokenizer.convert_tokens_to_ids(tokenizer.mask_token)
    ids = inputs.contiguous()
    out, labels = lib.mlm_corrupt(ids, special, pad_id, mask_id, len(tokenizer), mlm_probability, seed, ids_out=ids)
    if out.data_ptr() != inputs.data_ptr():
        inputs.copy_(out)
    return inputs, labels


class FeatureShard:
    """All videos' CLIP features ([n_i, D] fp16 each) packed back to back in one device tensor.

    ``FeatureShard.from_arrays(list_of_arrays)`` packs in-memory arrays (``None`` = missing video);
    ``FeatureShard.from_files(paths)`` loads the reference's ``<video_id>.mp4.npy`` files (missing / unreadable files
    become empty videos, as the reference's bare ``except`` does)."""

    def __init__(self, shard: torch.Tensor, offsets: torch.Tensor, nframes: torch.Tensor):
        self.shard, self.offsets, self.nframes = shard, offsets, nframes

    @classmethod
    def from_arrays(cls, arrays, features_dim=768, device="cuda"):
        n = [0 if a is None else int(len(a)) for a in arrays]
        off = np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        rows = [np.asarray(a, dtype=np.float16).reshape(-1, features_dim) for a in arrays if a is not None and len(a)]
        packed = np.concatenate(rows, 0) if rows else np.zeros((1, features_dim), np.float16)
        return cls(torch.from_numpy(packed).to(device), torch.from_numpy(off).to(device),
                   torch.tensor(n, dtype=torch.int32, device=device))

    @classmethod
    def from_files(cls, paths, features_dim=768, device="cuda"):
        arrays = []
        for p in paths:
            try:
                arrays.append(np.load(p))
            except Exception:  # noqa: BLE001  (reference videotext_dataset.py:39: missing video or corrupted file)
                arrays.append(None)
        return cls.from_arrays(arrays, features_dim, device)

    def save(self, path):
        """One file for the whole feature set (what a dataset preparation step writes once): packed fp16 rows +
        per-video offsets / frame counts, readable with ``FeatureShard.load``."""
        np.savez(self._npz_path(path), shard=self.shard.cpu().numpy(), offsets=self.offsets.cpu().numpy(),
                 nframes=self.nframes.cpu().numpy())

    @staticmethod
    def _npz_path(path):
        """np.savez appends ".npz" to a suffix-less path; save and load agree on the final name."""
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"

    @classmethod
    def load(cls, path, device="cuda"):
        z = np.load(cls._npz_path(path))
        return cls(torch.from_numpy(z["shard"]).to(device), torch.from_numpy(z["offsets"]).to(device),
                   torch.from_numpy(z["nframes"]).to(device))

    def batch(self, indices, max_feats=10, bf16=False):
        """-> {"video": [B, max_feats, D] (fp32, or bf16 for the projection GEMM), "video_len": int64 [B]}."""
        idx = torch.as_tensor(indices, device=self.shard.device, dtype=torch.int64)
        off, n = self.offsets[idx].contiguous(), self.nframes[idx].contiguous()
        B, D = idx.numel(), self.shard.shape[1]
        out = torch.empty(B, max_feats, D, device=self.shard.device, dtype=torch.bfloat16 if bf16 else torch.float32)
        video, vlen = lib.gather_frames(self.shard, off, n, max_feats, out_f32=None if bf16 else out,
                                        out_bf16=out if bf16 else None)
        return {"video": video, "video_len": vlen}
