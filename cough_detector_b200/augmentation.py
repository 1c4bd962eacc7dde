"""SpecAugment on the GPU feature cache.

Mirror of the reference's ``SpecAugment`` (src/augmentation.py:271-331): with probability ``p`` apply
``n_freq_masks`` x ``T.FrequencyMasking(freq_mask_param)`` then ``n_time_masks`` x ``T.TimeMasking(time_mask_param)``,
mask value 0.  The mask intervals are drawn on the host with the same calls in the same order as the reference
(``random.random()`` for ``p``; per mask ``torch.rand(1) * param`` then ``torch.rand(1) * (size - value)``,
torchaudio ``functional.mask_along_axis``), so with equal seeds the masked cells are the reference's.  Only the
fill runs on the device (``cd_spec_augment``: stores, no reads).  The waveform augmentations of the reference
(``AudioAugmentor``: CPU-side random effects on raw audio) are out of scope.
"""
from __future__ import annotations

import random
from typing import List, Optional

import torch

from . import _lib


class SpecAugment:
    """Drop-in for reference ``SpecAugment`` (src/augmentation.py:271-331)."""

    def __init__(self, freq_mask_param: int = 10, time_mask_param: int = 20, n_freq_masks: int = 2,
                 n_time_masks: int = 2, p: float = 0.5):
        if n_freq_masks > 8 or n_time_masks > 8:
            raise NotImplementedError("cd_spec_augment takes at most 8 masks per axis")
        self.freq_mask_param, self.time_mask_param = int(freq_mask_param), int(time_mask_param)
        self.n_freq_masks, self.n_time_masks = int(n_freq_masks), int(n_time_masks)
        self.p = p

    # ---- host: the reference's random draws -----------------------------------------------------
    @staticmethod
    def _interval(mask_param: int, size: int) -> List[int]:
        """One mask_along_axis draw (p = 1.0): [start, end).  Consumes two torch.rand(1), or none when mask_param < 1
        (the reference returns before drawing)."""
        if mask_param < 1:
            return [0, 0]
        value = torch.rand(1) * mask_param
        min_value = torch.rand(1) * (size - value)
        start = int(min_value.long())
        end = start + int(value.long())
        return [start, end]

    def sample(self, n_rows: int, n_cols: int = _lib.N_FRAMES) -> Optional[List[int]]:
        """The intervals of ONE __call__: None when the reference would return its input untouched (probability 1 - p),
        else 2 (n_freq_masks + n_time_masks) ints {start, end} -- rows first, then columns."""
        if random.random() > self.p:
            return None
        out: List[int] = []
        for _ in range(self.n_freq_masks):
            out += self._interval(self.freq_mask_param, n_rows)
        for _ in range(self.n_time_masks):
            out += self._interval(self.time_mask_param, n_cols)
        return out

    def sample_batch(self, batch: int, n_rows: int, n_cols: int = _lib.N_FRAMES) -> torch.Tensor:
        """``batch`` independent __call__s in order (what a DataLoader worker does sample by sample): (batch, 2 (nf + nt)) int32,
        all-zero rows for the clips left untouched."""
        width = 2 * (self.n_freq_masks + self.n_time_masks)
        rows = []
        for _ in range(batch):
            m = self.sample(n_rows, n_cols)
            rows.append([0] * width if m is None else m)
        return torch.tensor(rows, dtype=torch.int32).reshape(batch, width)

    # ---- device ------------------------------------------------------------------------------------
    def apply_masks(self, spectrogram: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        """Out of place like the reference (masked_fill): (B, 1, rows, 101) + (B, width) -> masked copy on the input's device."""
        h = _lib.get_handle(spectrogram.device if spectrogram.is_cuda else None)
        x = spectrogram.to(device=f"cuda:{h.device}", dtype=torch.float32).contiguous()
        if x.data_ptr() == spectrogram.data_ptr():
            x = x.clone()
        h.spec_augment_(x, masks, self.n_freq_masks, self.n_time_masks)
        return x if spectrogram.is_cuda else x.to(spectrogram.device)

    def __call__(self, spectrogram: torch.Tensor) -> torch.Tensor:
        """(..., freq, time) -- usually (channels, freq, time) or (batch, channels, freq, time); every leading index shares ONE
        set of masks, as in the reference (mask_along_axis: "All examples will have the same mask interval")."""
        if spectrogram.dim() < 2:
            raise ValueError(f"Spectrogram must have at least two dimensions (time and frequency) ({spectrogram.dim()} given).")
        n_rows, n_cols = int(spectrogram.shape[-2]), int(spectrogram.shape[-1])
        m = self.sample(n_rows, n_cols)
        if m is None:
            return spectrogram
        x = spectrogram.reshape(-1, 1, n_rows, n_cols)
        masks = torch.tensor(m, dtype=torch.int32).repeat(x.shape[0], 1)
        return self.apply_masks(x, masks).reshape(spectrogram.shape)

    def augment_batch_(self, features: torch.Tensor) -> torch.Tensor:
        """Additive, batched: independent masks per clip (one __call__ each, in order), one launch, in place on a cuda batch."""
        masks = self.sample_batch(features.shape[0], int(features.shape[-2]), int(features.shape[-1]))
        h = _lib.get_handle(features.device)
        return h.spec_augment_(features, masks, self.n_freq_masks, self.n_time_masks)
