"""Training-side use of the feature path: the reference's ``CoughDataset`` with its feature
cache filled in batches on the GPU.

The reference extracts features per sample inside ``__getitem__`` (src/dataset.py:121-173: load -> resample -> mono ->
[waveform augmentation] -> normalize -> pad/trim -> extract_features, 1.33 ms per clip on a CPU core) and keeps them in
a ``{path: tensor}`` cache when ``cache_spectrograms`` is set.  Here the same cache is filled by ``precompute_cache``:
files are decoded on the host, packed per sample rate, and pushed through ``AudioPreprocessor.process_batch`` (one
upload, one resample + mono launch, one feature launch per chunk); 16-bit PCM wav files stay int16 until they are on
the GPU (half the host->device bytes; the scale cancels in the peak normalisation, so the features equal the float path
bit for bit).  ``__getitem__`` keeps the reference's behaviour sample by sample; ``feature_batches`` is the additive,
batched way to feed ``train_epoch``: device-resident ``(B, 1, rows, 101)`` batches with SpecAugment applied by
``cd_spec_augment``.

File decode itself is host I/O outside the hot path: stdlib ``wave`` for PCM wav, otherwise ``torchaudio.load`` like the
reference (src/dataset.py:141-148, same error text).  There is no CPU feature path.
"""
from __future__ import annotations

import wave
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset

from .augmentation import SpecAugment
from .preprocessing import AudioPreprocessor


def decode_audio(path: str) -> Tuple[torch.Tensor, int]:
    """(channels, samples) waveform + sample rate.  16-bit PCM wav -> int16 tensor (what the file holds); anything else
    through torchaudio.load (float32), failing as loudly as the reference does (src/dataset.py:141-148)."""
    if str(path).lower().endswith(".wav"):
        try:
            with wave.open(str(path), "rb") as w:
                if w.getsampwidth() == 2 and w.getcomptype() == "NONE":
                    n, c, sr = w.getnframes(), w.getnchannels(), w.getframerate()
                    pcm = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, c).T
                    return torch.from_numpy(np.array(pcm, order="C")), sr            # a writable copy (frombuffer views are read-only)
        except (wave.Error, EOFError):
            pass                                            # not plain PCM: let torchaudio try
    try:
        import torchaudio
        return torchaudio.load(path)
    except Exception as e:
        raise RuntimeError(
            f"Failed to load audio file: {path}\n"
            f"Error: {e}\n"
            f"If you see 'torchcodec' errors, run: pip install torchcodec"
        ) from e


class CoughDataset(Dataset):
    """Drop-in for reference ``CoughDataset`` (src/dataset.py:23-173): ``data_dir/{non_cough,cough}/*.{wav,...}``."""

    AUDIO_EXTENSIONS = {".wav", ".mp3", ".flac", ".ogg", ".webm"}

    def __init__(self, data_dir: str, preprocessor: AudioPreprocessor, audio_augmentor=None,
                 spec_augmentor: Optional[SpecAugment] = None, is_training: bool = True,
                 cache_spectrograms: bool = False):
        self.data_dir = Path(data_dir)
        self.preprocessor = preprocessor
        self.audio_augmentor = audio_augmentor
        self.spec_augmentor = spec_augmentor
        self.is_training = is_training
        self.cache_spectrograms = cache_spectrograms
        self.classes = ["non_cough", "cough"]
        self.class_to_idx = {c: i for i, c in enumerate(self.classes)}
        self.samples = self._collect_samples()
        self.cache: Optional[Dict[str, torch.Tensor]] = {} if cache_spectrograms else None
        self.class_counts = self._count_classes()
        self.sample_weights = self._compute_sample_weights()

    # ---- bookkeeping, as the reference (:82-116) ---------------------------------------------------
    def _collect_samples(self) -> List[Tuple[str, int]]:
        samples = []
        for class_name in self.classes:
            class_dir = self.data_dir / class_name
            if not class_dir.exists():
                print(f"Warning: Class directory {class_dir} not found")
                continue
            label = self.class_to_idx[class_name]
            for audio_file in sorted(class_dir.iterdir()):
                if audio_file.suffix.lower() in self.AUDIO_EXTENSIONS:
                    samples.append((str(audio_file), label))
        return samples

    def _count_classes(self) -> Dict[int, int]:
        counts = {i: 0 for i in range(len(self.classes))}
        for _, label in self.samples:
            counts[label] += 1
        return counts

    def _compute_sample_weights(self) -> torch.Tensor:
        total = len(self.samples)
        return torch.tensor([total / (len(self.classes) * self.class_counts[label]) for _, label in self.samples])

    def __len__(self) -> int:
        return len(self.samples)

    # ---- the batched cache fill ----------------------------------------------------------------------
    def precompute_cache(self, batch_size: int = 4096, device: Optional[str] = None) -> torch.Tensor:
        """Fill the feature cache for every file without waveform augmentation (what the reference's cache holds after a
        pass with ``audio_augmentor=None``), ``batch_size`` files per GPU call, and return the features as ONE tensor
        ``(N, 1, rows, 101)`` in sample order on ``device`` (default: the preprocessor's device if it is a GPU, else CPU).
        The per-path ``cache`` dict holds views of that tensor."""
        if device is None:
            device = self.preprocessor.device if str(self.preprocessor.device).startswith("cuda") else "cpu"
        rows = self.preprocessor.get_num_features()
        out = torch.empty(len(self.samples), 1, rows, self.preprocessor.get_expected_time_frames(), device=device)
        for lo in range(0, len(self.samples), batch_size):
            by_rate: Dict[Tuple[int, bool], List[Tuple[int, torch.Tensor]]] = {}
            for i in range(lo, min(lo + batch_size, len(self.samples))):
                w, sr = decode_audio(self.samples[i][0])
                by_rate.setdefault((int(sr), w.dtype == torch.int16), []).append((i, w))
            for (sr, _), items in by_rate.items():
                feats = self.preprocessor.process_batch([w for _, w in items], sr, out_device=out.device)
                out[torch.tensor([i for i, _ in items], device=out.device)] = feats
        self.features = out
        if self.cache is None:
            self.cache = {}
        for i, (path, _) in enumerate(self.samples):
            self.cache[path] = out[i]
        return out

    # ---- per sample, as the reference (:121-173) -----------------------------------------------------
    def __getitem__(self, idx: int) -> Tuple[torch.Tensor, int]:
        audio_path, label = self.samples[idx]
        if self.cache is not None and audio_path in self.cache:
            spectrogram = self.cache[audio_path].clone()
            if self.is_training and self.spec_augmentor:
                spectrogram = self.spec_augmentor(spectrogram)
            return spectrogram, label
        waveform, sr = decode_audio(audio_path)
        if self.is_training and self.audio_augmentor:
            # the waveform augmentations are CPU-side random effects on 16 kHz mono audio (out of scope here): condition
            # first like the reference (:151-156), then hand the augmented waveform to the GPU for the rest
            if waveform.dtype == torch.int16:
                waveform = waveform.to(torch.float32) / 32768.0
            waveform = self.preprocessor.to_mono(self.preprocessor.resample(waveform, sr))
            waveform = self.audio_augmentor.augment(waveform)
            sr = self.preprocessor.sample_rate
        features = self.preprocessor.process(waveform, sr)            # normalize -> pad/trim -> extract_features, one kernel
        if self.cache is not None:
            self.cache[audio_path] = features.clone()
        if self.is_training and self.spec_augmentor:
            features = self.spec_augmentor(features)
        return features, label

    # ---- batched feed for train_epoch (additive) -------------------------------------------------------
    def feature_batches(self, batch_size: int, shuffle: bool = True, generator: Optional[torch.Generator] = None,
                        device: Optional[str] = None) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """(inputs (B, 1, rows, 101), labels (B,)) batches from the precomputed cache, gathered and SpecAugment-ed on the
        device (independent masks per clip, drawn in batch order).  The tensors a ``train_epoch`` loop moves to the device
        per batch (src/train.py:84-85) are already there."""
        if not hasattr(self, "features"):
            self.precompute_cache(device=device)
        feats = self.features if device is None else self.features.to(device)
        labels = torch.tensor([l for _, l in self.samples], device=feats.device)
        n = len(self.samples)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        for lo in range(0, n, batch_size):
            idx = order[lo:lo + batch_size].to(feats.device)
            x = feats[idx]                                           # a fresh tensor: the cache is never written
            if self.is_training and self.spec_augmentor and x.is_cuda:
                self.spec_augmentor.augment_batch_(x)
            elif self.is_training and self.spec_augmentor:
                m = self.spec_augmentor.sample_batch(x.shape[0], int(x.shape[2]), int(x.shape[3]))
                x = self.spec_augmentor.apply_masks(x, m)
            yield x, labels[idx]
