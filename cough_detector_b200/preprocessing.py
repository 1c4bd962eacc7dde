"""AudioPreprocessor / RealtimePreprocessor on the B200 feature kernel.

Host-side mirror of the reference's ``src/preprocessing.py`` (class and method names, argument
meaning, return shapes, error behaviour) for the 90-feature configurations: the one the reference
trains with (``src/train.py:264-287``: ``use_pcen=False, use_pre_emphasis=False,
use_delta_delta=False, use_spectral_contrast=False``) and its ``use_pre_emphasis`` / ``use_pcen``
variants (same 90 rows), and ``use_delta_delta`` (103 rows; the CNN kernels are instantiated for both
heights).  All feature
arithmetic runs in ``libcough_b200.so`` (csrc/feat_kernel.cu); there is no torchaudio/CPU path here.
``use_spectral_contrast`` raises ``NotImplementedError`` instead of falling back; its rows are NaN in
the reference itself.

Additive, batched entry points (the reference API is strictly per clip):
``extract_features_batch``, ``process_batch`` and ``RealtimePreprocessor.stream_features``.
Batch semantics are per clip: every clip gets its own peak, top_db floor and MFCC mean/std.
Tensors come back on the device of the input (CPU in -> CPU out), like the reference.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

_FIXED = dict(sample_rate=16000, n_mels=64, n_fft=512, hop_length=160, win_length=400,
              f_min=100.0, f_max=4000.0, n_mfcc=13)


class AudioPreprocessor:
    """Drop-in for reference ``AudioPreprocessor`` (src/preprocessing.py:13-550), 90-feature config."""

    def __init__(self, sample_rate: int = 16000, n_mels: int = 64, n_fft: int = 512, hop_length: int = 160,
                 win_length: int = 400, f_min: float = 100.0, f_max: float = 4000.0, segment_duration: float = 1.0,
                 n_mfcc: int = 13, use_mfcc: bool = True, use_pcen: bool = True, use_pre_emphasis: bool = True,
                 pre_emphasis_coef: float = 0.97, use_delta_delta: bool = True, use_spectral_contrast: bool = True,
                 n_contrast_bands: int = 6, device: str = "cpu"):
        got = dict(sample_rate=sample_rate, n_mels=n_mels, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                   f_min=float(f_min), f_max=float(f_max), n_mfcc=n_mfcc)
        bad = {k: v for k, v in got.items() if v != _FIXED[k]}
        if bad or float(segment_duration) != 1.0 or not use_mfcc:
            raise NotImplementedError(
                f"the B200 feature kernel is built for {_FIXED}, segment_duration=1.0, use_mfcc=True; got {bad or ''} "
                f"segment_duration={segment_duration} use_mfcc={use_mfcc}")
        if use_spectral_contrast:
            raise NotImplementedError(
                "use_spectral_contrast is not implemented on the GPU path (and its rows are NaN in the reference itself: "
                "DESIGN.md §6); construct with use_spectral_contrast=False as the reference training config does "
                "(src/train.py:264-287).  There is deliberately no CPU fallback.")
        if not 0.0 <= float(pre_emphasis_coef) <= 1.0:
            raise ValueError(f"pre_emphasis_coef must be in [0, 1], got {pre_emphasis_coef}")
        self.sample_rate, self.n_mels, self.n_fft = sample_rate, n_mels, n_fft
        self.hop_length, self.win_length, self.f_min, self.f_max = hop_length, win_length, f_min, f_max
        self.segment_duration = segment_duration
        self.segment_samples = int(sample_rate * segment_duration)
        self.n_mfcc, self.use_mfcc = n_mfcc, use_mfcc
        self.use_pcen, self.use_pre_emphasis, self.pre_emphasis_coef = use_pcen, use_pre_emphasis, pre_emphasis_coef
        self.use_delta_delta, self.use_spectral_contrast = use_delta_delta, use_spectral_contrast
        self.n_contrast_bands = n_contrast_bands
        self.device = device

    # ---- plumbing ------------------------------------------------------------------------------
    def _gpu(self) -> "_lib.Handle":
        dev = torch.device(self.device)
        return _lib.get_handle(dev if dev.type == "cuda" else None)

    @property
    def feat_options(self) -> "_lib.FeatOptions":
        """This object's optional stages, as passed with every library call.  The flags belong to the preprocessor
        (reference :32-51), never to the shared per-device handle."""
        return _lib.FeatOptions(bool(self.use_pre_emphasis), float(self.pre_emphasis_coef), bool(self.use_pcen),
                                bool(self.use_delta_delta))

    @staticmethod
    def _back(out: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
        return out if like.is_cuda else out.to(like.device)

    # ---- waveform conditioning (host-side mirror, same semantics) -----------------------------------
    def load_audio(self, path: str) -> Tuple[torch.Tensor, int]:
        """File decode is outside the hot path; defers to torchaudio like the reference (:155-166)."""
        import torchaudio
        return torchaudio.load(path)

    def resample(self, waveform: torch.Tensor, orig_sr: int) -> torch.Tensor:
        """T.Resample(orig_sr, 16000) with torchaudio's defaults on the GPU (:168-183); identity at 16 kHz."""
        if orig_sr == self.sample_rate:
            return waveform
        if int(orig_sr) != orig_sr or orig_sr <= 0:
            raise ValueError("sample rates must be positive integers")
        h = self._gpu()
        x = waveform.to(device=f"cuda:{h.device}", dtype=torch.float32)
        return self._back(h.resample(x, int(orig_sr), self.sample_rate), waveform)

    def to_mono(self, waveform: torch.Tensor) -> torch.Tensor:
        """(C, n) -> (1, n) channel mean (:185-197)."""
        return waveform if waveform.shape[0] == 1 else waveform.mean(dim=0, keepdim=True)

    def normalize(self, waveform: torch.Tensor) -> torch.Tensor:
        """x / max|x|, unchanged when silent (:199-212).  Standalone use only: ``process`` and the batched
        entry points fuse this into the kernel."""
        peak = waveform.abs().max() if waveform.numel() else waveform.new_zeros(())
        return waveform / peak if peak > 0 else waveform

    def pad_or_trim(self, waveform: torch.Tensor, length: Optional[int] = None) -> torch.Tensor:
        """Centre trim or zero pad (left = floor(pad/2)) to ``length`` (:358-385)."""
        length = self.segment_samples if length is None else length
        n = waveform.shape[1]
        if n == length:
            return waveform
        if n > length:
            s = (n - length) // 2
            return waveform[:, s:s + length]
        p = length - n
        return torch.nn.functional.pad(waveform, (p // 2, p - p // 2))

    def compute_deltas(self, features: torch.Tensor) -> torch.Tensor:
        """(x[t+1] - x[t-1]) / 2 with replicated edges (:342-356); the kernel fuses this for the MFCC rows."""
        p = torch.nn.functional.pad(features, (1, 1), mode="replicate")
        return (p[:, :, 2:] - p[:, :, :-2]) / 2

    # ---- features ------------------------------------------------------------------------------
    def extract_features(self, waveform: torch.Tensor) -> torch.Tensor:
        """(1, 16000) conditioned waveform -> (1, 90, 101), NO normalisation (reference :432-489)."""
        if waveform.dim() != 2 or waveform.shape[0] != 1:
            raise ValueError(f"expected waveform of shape (1, samples), got {tuple(waveform.shape)}")
        if waveform.shape[1] != self.segment_samples:
            raise ValueError(f"the B200 kernel extracts features of {self.segment_samples}-sample segments; got "
                             f"{waveform.shape[1]} samples (use process() / pad_or_trim first)")
        return self.extract_features_batch(waveform, normalize=False)[0]

    def extract_mel_spectrogram(self, waveform: torch.Tensor) -> torch.Tensor:
        """Rows 0..63 of extract_features (:387-412)."""
        return self.extract_features(waveform)[:, : self.n_mels]

    def extract_mfcc(self, waveform: torch.Tensor) -> torch.Tensor:
        """Rows 64..76 of extract_features (:414-430)."""
        return self.extract_features(waveform)[:, self.n_mels: self.n_mels + self.n_mfcc]

    def extract_features_batch(self, waveforms: torch.Tensor, normalize: bool = True) -> torch.Tensor:
        """(B, 16000) -> (B, 1, 90, 101); ``normalize`` applies per-clip peak normalisation first.  int16 PCM tensors
        are taken as they are (half the bytes to the GPU; with ``normalize`` the result equals the float path bit for bit)."""
        if waveforms.dim() != 2 or waveforms.shape[1] != self.segment_samples:
            raise ValueError(f"expected (B, {self.segment_samples}), got {tuple(waveforms.shape)}")
        h = self._gpu()
        if waveforms.dtype == torch.int16:
            x = waveforms.to(device=f"cuda:{h.device}", non_blocking=True)
            return self._back(h.feat_forward_i16(x, normalize=normalize, opts=self.feat_options), waveforms)
        x = waveforms.to(device=f"cuda:{h.device}", dtype=torch.float32, non_blocking=True)
        return self._back(h.feat_forward(x, normalize=normalize, opts=self.feat_options), waveforms)

    def process(self, waveform: torch.Tensor, orig_sr: int) -> torch.Tensor:
        """resample -> mono -> normalize -> pad/trim -> features, one kernel (:491-517).  (C, n) -> (1, 90, 101)."""
        return self.process_batch([waveform], orig_sr)[0]

    def process_batch(self, waveforms: Sequence[torch.Tensor], orig_sr: int = 16000, out_device=None) -> torch.Tensor:
        """Variable-length clips [(C_i, n_i) or (n_i,)], float or int16 PCM, all at ``orig_sr`` -> (B, 1, 90, 101); per clip
        identical to ``process`` (the per-file pipeline of dataset.__getitem__, src/dataset.py:151-163).
        Batched end to end: ONE packed host buffer and host->device copy for all clips, ONE resample + to_mono launch
        over (offset, channels, length) descriptors when any clip needs it, ONE feature launch.  The result lands on
        ``out_device`` (default: the device of the first waveform, like every other method here)."""
        if not len(waveforms):
            return torch.zeros(0, 1, self.get_num_features(), _lib.N_FRAMES)
        if int(orig_sr) != orig_sr or orig_sr <= 0:
            raise ValueError("sample rates must be positive integers")
        h = self._gpu()
        dev = torch.device(f"cuda:{h.device}")
        i16 = all(w.dtype == torch.int16 for w in waveforms)
        dtype = torch.int16 if i16 else torch.float32
        rows = [w.reshape(1, -1) if w.dim() == 1 else w for w in waveforms]
        chans = [int(r.shape[0]) for r in rows]
        n_in = [int(r.shape[1]) for r in rows]
        # every clip starts 16-byte aligned in the packed buffer
        pad = 8 if i16 else 4
        sizes = [((c * n + pad - 1) // pad) * pad for c, n in zip(chans, n_in)]
        offs = [0]
        for sz in sizes[:-1]:
            offs.append(offs[-1] + sz)
        total = offs[-1] + sizes[-1] + pad
        if all(r.is_cuda for r in rows):
            packed = torch.zeros(total, dtype=dtype, device=dev)
            for r, o in zip(rows, offs):
                packed[o:o + r.numel()] = r.reshape(-1).to(dtype)
        else:
            host = torch.zeros(total, dtype=dtype).pin_memory()
            for r, o in zip(rows, offs):
                host[o:o + r.numel()] = r.reshape(-1).to(device="cpu", dtype=dtype)
            packed = host.to(dev, non_blocking=True)
        if orig_sr != self.sample_rate or i16 or any(c != 1 for c in chans):
            flat, out_off, out_len = h.condition_var(packed, offs, chans, n_in, int(orig_sr), self.sample_rate)
        else:
            flat, out_off, out_len = packed, torch.tensor(offs, dtype=torch.int64), torch.tensor(n_in, dtype=torch.int32)
        out = h.feat_forward_var(flat, out_off, out_len, opts=self.feat_options)
        if out_device is not None:
            return out.to(out_device)
        return self._back(out, waveforms[0])

    def process_file(self, path: str) -> torch.Tensor:
        waveform, sr = self.load_audio(path)
        return self.process(waveform, sr)

    def get_expected_time_frames(self) -> int:
        return (self.segment_samples // self.hop_length) + 1

    def get_num_features(self) -> int:
        """64 mel + 13 MFCC + 13 delta = 90 in the supported configuration (:536-550)."""
        n = self.n_mels
        if self.use_mfcc:
            n += 2 * self.n_mfcc + (self.n_mfcc if self.use_delta_delta else 0)
        if self.use_spectral_contrast:
            n += self.n_contrast_bands + 1
        return n


class RealtimePreprocessor(AudioPreprocessor):
    """Sliding-window streaming front end (reference :553-616): per-window peak normalisation, no pad/trim."""

    def __init__(self, window_duration: float = 1.0, hop_duration: float = 0.5, **kwargs):
        super().__init__(segment_duration=window_duration, **kwargs)
        self.window_duration = window_duration
        self.hop_duration = hop_duration
        self.window_samples = int(self.sample_rate * window_duration)
        self.hop_samples = int(self.sample_rate * hop_duration)
        if self.hop_samples <= 0:
            raise ValueError("hop_duration must be at least one sample")
        self.buffer = torch.zeros(1, 0)

    def stream_features(self, audio: torch.Tensor) -> torch.Tensor:
        """All complete windows of a 1-D stream at once: (n,) -> (windows, 1, 90, 101)."""
        h = self._gpu()
        x = audio.reshape(-1).to(device=f"cuda:{h.device}", dtype=torch.float32)
        return self._back(h.feat_forward_stream(x, self.hop_samples, opts=self.feat_options), audio)

    def add_audio(self, audio_chunk: torch.Tensor) -> List[torch.Tensor]:
        """Append a chunk; return the features of every window that became complete (:582-612)."""
        if audio_chunk.dim() == 1:
            audio_chunk = audio_chunk.unsqueeze(0)
        self.buffer = torch.cat([self.buffer.to(audio_chunk.device), audio_chunk.to(torch.float32)], dim=1)
        n = self.buffer.shape[1]
        if n < self.window_samples:
            return []
        nw = (n - self.window_samples) // self.hop_samples + 1
        feats = self.stream_features(self.buffer[0, : (nw - 1) * self.hop_samples + self.window_samples])
        self.buffer = self.buffer[:, nw * self.hop_samples:]
        return [feats[i] for i in range(nw)]

    def reset(self):
        self.buffer = torch.zeros(1, 0)


def create_preprocessor(realtime: bool = False, **kwargs) -> AudioPreprocessor:
    """Factory with the reference's signature (:619-632)."""
    return RealtimePreprocessor(**kwargs) if realtime else AudioPreprocessor(**kwargs)
