"""Detection engine on the B200 path.

Mirror of the reference's ``CoughDetectorInference`` (src/inference.py:39-247): checkpoint +
config loading, ``predict``, the per-chunk smoothing / threshold / debounce loop -- with the
feature and model arithmetic in libcough_b200.so.  Microphone capture, device listing and the
CLI (src/inference.py:250-507) are out of scope (no GPU work).

Additive: ``detect_stream`` runs a whole recording in one fused call (shared per-GPU kernels,
every window of the stream at once) and applies the same decision rule on a *stream-time*
clock, which makes it deterministic; ``process_audio_chunk`` keeps the reference's wall clock
unless a ``clock`` callable is supplied.
"""
from __future__ import annotations

from collections import deque
from datetime import datetime
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .model import create_model, _PREC
from .preprocessing import RealtimePreprocessor


class CoughDetectorInference:
    """Same constructor and methods as the reference engine (src/inference.py:49-117)."""

    def __init__(self, model_path: str, device: str = "auto", confidence_threshold: float = 0.5,
                 smoothing_window: int = 3, debounce_seconds: float = 0.5, verbose: bool = True,
                 clock: Optional[Callable[[], float]] = None):
        self.verbose = verbose
        self.confidence_threshold = confidence_threshold
        self.debounce_seconds = debounce_seconds
        if device in ("auto", "cuda"):
            if not torch.cuda.is_available():
                raise _lib.CoughB200Error("no CUDA device: this engine runs on B200 only (no CPU path)")
            device = "cuda"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CoughB200Error(f"device={device!r}: this engine runs on CUDA (B200) only, there is no CPU path")
        if self.verbose:
            print(f"Using device: {self.device}")
        self._load_model(model_path)
        cfg = self.config
        # same .get() defaults as the reference (src/inference.py:89-108): a checkpoint without the
        # feature flags selects the 110-feature config, which raises NotImplementedError here.
        self.preprocessor = RealtimePreprocessor(
            sample_rate=cfg.get("sample_rate", 16000), n_mels=cfg.get("n_mels", 64), n_fft=cfg.get("n_fft", 512),
            hop_length=cfg.get("hop_length", 160), win_length=cfg.get("win_length", 400),
            f_min=cfg.get("f_min", 100.0), f_max=cfg.get("f_max", 4000.0),
            window_duration=cfg.get("segment_duration", 1.0), hop_duration=0.25,
            n_mfcc=cfg.get("n_mfcc", 13), use_mfcc=cfg.get("use_mfcc", True), use_pcen=cfg.get("use_pcen", True),
            use_pre_emphasis=cfg.get("use_pre_emphasis", True), pre_emphasis_coef=cfg.get("pre_emphasis_coef", 0.97),
            use_delta_delta=cfg.get("use_delta_delta", True),
            use_spectral_contrast=cfg.get("use_spectral_contrast", True),
            n_contrast_bands=cfg.get("n_contrast_bands", 6), device=str(self.device))
        self.prediction_history = deque(maxlen=smoothing_window)
        self.last_detection_time = 0
        self.on_cough_detected: Optional[Callable[[datetime, float], None]] = None
        self._clock = clock

    def _load_model(self, model_path: str):
        """Checkpoint dict of src/train.py:183-199: {'model_state_dict', 'config', ...} (reference :119-163)."""
        if self.verbose:
            print(f"Loading model from {model_path}")
        ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
        self.config = ckpt.get("config", {})
        model_type = self.config.get("model_type", "small")
        self.model = create_model(model_type=model_type, n_mels=90, num_classes=2, in_channels=1)
        self.model.load_state_dict(ckpt["model_state_dict"])
        self.model.to(self.device)
        self.model.eval()
        if self.verbose:
            print(f"Model loaded: {model_type}")

    @torch.no_grad()
    def predict(self, spectrogram: torch.Tensor) -> tuple:
        """(1, 90, 101) or (B, 1, 90, 101) -> (is_cough, P(cough)) of batch item 0, like the reference (:165-189)."""
        if spectrogram.dim() == 3:
            spectrogram = spectrogram.unsqueeze(0)
        _, probs = self.model.predict(spectrogram.to(self.device))
        p = probs[0, 1].item()
        return p > 0.5, p

    def process_audio_chunk(self, audio_chunk) -> Optional[tuple]:
        """Reference chunk loop (:191-241) including its early return after a detection."""
        if isinstance(audio_chunk, np.ndarray):
            audio_chunk = torch.from_numpy(audio_chunk.astype(np.float32))
        if audio_chunk.dim() == 1:
            audio_chunk = audio_chunk.unsqueeze(0)
        if audio_chunk.shape[0] > 1:
            audio_chunk = audio_chunk.mean(dim=0, keepdim=True)
        for spec in self.preprocessor.add_audio(audio_chunk):
            _, confidence = self.predict(spec)
            self.prediction_history.append(confidence)
            smoothed = float(np.mean(self.prediction_history))
            now = self._clock() if self._clock else datetime.now().timestamp()
            if smoothed >= self.confidence_threshold and (now - self.last_detection_time) >= self.debounce_seconds:
                self.last_detection_time = now
                stamp = datetime.now()
                if self.on_cough_detected:
                    self.on_cough_detected(stamp, smoothed)
                return stamp, smoothed
        return None

    def reset(self):
        self.preprocessor.reset()
        self.prediction_history.clear()
        self.last_detection_time = 0

    # ---- additive: whole-recording detection --------------------------------------------------------
    @torch.no_grad()
    def stream_probabilities(self, audio: torch.Tensor, hop_duration: float = 0.25) -> torch.Tensor:
        """P(cough) of every 1-s window of a 16 kHz mono stream at the given hop, one fused call."""
        hop = int(16000 * hop_duration)
        h = self.model._handle(torch.empty(0, device=self.device))
        x = audio.reshape(-1).to(device=self.device, dtype=torch.float32)
        if not hasattr(self.model, "precision"):            # 'small' / 'standard': features, then the fp32 network
            logits = self.model(h.feat_forward_stream(x, hop, opts=self.preprocessor.feat_options))
        else:
            logits = h.e2e_forward_stream(x, hop, _PREC[self.model.precision], opts=self.preprocessor.feat_options)
        _, probs, _ = h.predict(logits)
        return probs[:, 1]

    def detect_stream(self, audio: torch.Tensor, hop_duration: float = 0.25) -> List[Tuple[float, float]]:
        """[(time_s of the window end, smoothed confidence)] for every fired detection."""
        probs = self.stream_probabilities(audio, hop_duration).cpu().numpy()
        return [(j * hop_duration + 1.0, c) for j, c in smooth_and_debounce(
            probs, hop_duration, self.confidence_threshold, self.prediction_history.maxlen, self.debounce_seconds)]


def smooth_and_debounce(probs, hop_s: float, threshold: float, smoothing_window: int,
                        debounce_s: float) -> List[Tuple[int, float]]:
    """The decision rule of src/inference.py:220-239 over ordered per-window probabilities, on stream time
    t_j = j*hop + 1 s: mean of the last ``smoothing_window`` probabilities >= threshold and at least
    ``debounce_s`` since the previous fire.  Sequential host post-pass, O(windows)."""
    hist: deque = deque(maxlen=smoothing_window)
    last_t = float("-inf")
    fired = []
    for j, p in enumerate(probs):
        hist.append(float(p))
        sm = float(np.mean(hist))
        t = j * hop_s + 1.0
        if sm >= threshold and (t - last_t) >= debounce_s:
            last_t = t
            fired.append((j, sm))
    return fired
