"""CPU oracle for the cough_detector feature + CoughDetectorResidual hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cough_detector_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker (or as the timed CPU baseline), never as the product path.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md §4, §8c), so this oracle is pinned by *executing the reference* in
the authoring container: ``oracle/make_golden.py`` imports ``/root/reference``
unmodified, runs it on seeded synthetic clips and seeded weights, and commits
its outputs under ``tests/golden/``.  ``tests/test_oracle.py`` checks every
function below against those vectors.  The arithmetic itself lives in an
un-vendored third-party dependency of the reference: ``torchaudio>=2.0.0`` /
``torch>=2.0.0`` (reference ``requirements.txt:2-3``, lower bounds only, no lock
file); the versions the goldens were made with are torchaudio 2.11.0+cu128 and
torch 2.11.0+cu128.

Three restatements live here:

* ``Explicit32``  -- op-for-op fp32 restatement in plain torch CPU ops (no
  torchaudio): reflect-pad, frame, Hann-in-512, rFFT, power, mel matmul, dB,
  top_db floor, affine+clamp, DCT, mean/std, deltas.
* ``features_fp64`` -- the same pipeline in numpy float64; the ground truth the
  1e-4 feature tolerance is measured against (reference fp32 noise ~4e-6).
* ``cnn_forward`` / ``cnn_forward_bf16`` -- functional restatement of
  CoughDetectorResidual.forward from a reference-format ``state_dict`` (fp32, and
  a bf16-operand / fp32-accumulate emulation of what the GPU kernels compute).

File:line citations are into ``/root/reference`` unless prefixed ``torchaudio/``.
"""
from __future__ import annotations

import math
from collections import deque
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE = 16000
N_FFT = 512
WIN_LENGTH = 400
HOP_LENGTH = 160
N_MELS = 64
N_MFCC = 13
F_MIN = 100.0
F_MAX = 4000.0
SEGMENT = 16000
N_FRAMES = SEGMENT // HOP_LENGTH + 1  # 101, src/preprocessing.py:532-534
N_FEATURES = N_MELS + 2 * N_MFCC      # 90,  src/preprocessing.py:536-550
TOP_DB = 80.0
AMIN = 1e-10
BN_EPS = 1e-5


# --------------------------------------------------------------------------
# constants (torchaudio/functional/functional.py:425-587, 636-667)
# --------------------------------------------------------------------------
def hann_window(dtype=torch.float32) -> torch.Tensor:
    """Periodic Hann(400), torchaudio/transforms/_transforms.py:96 (torch.hann_window)."""
    if dtype == torch.float32:
        # ATen evaluates arange(N)*(2*pi/N) -> cos -> *(-0.5) -> +0.5 in fp32; its rounding
        # (up to 2.4e-7 abs near the window edges) is part of the reference's behaviour.
        return torch.hann_window(WIN_LENGTH, periodic=True, dtype=torch.float32)
    n = torch.arange(WIN_LENGTH, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / WIN_LENGTH)).to(dtype)


def mel_filterbank(dtype=torch.float32) -> torch.Tensor:
    """HTK triangular filterbank (257, 64), norm=None.

    torchaudio/functional/functional.py:518-587 with _hz_to_mel :425, _mel_to_hz :459,
    _create_triangular_filterbank :492.  The fp32 linspace / pow are kept in fp32 so
    band edges match torchaudio to the ulp when dtype is float32.
    """
    n_freqs = N_FFT // 2 + 1
    all_freqs = torch.linspace(0, SAMPLE_RATE // 2, n_freqs, dtype=dtype)
    m_min = 2595.0 * math.log10(1.0 + F_MIN / 700.0)
    m_max = 2595.0 * math.log10(1.0 + F_MAX / 700.0)
    m_pts = torch.linspace(m_min, m_max, N_MELS + 2, dtype=dtype)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1, dtype=dtype), torch.min(down, up))


def dct_matrix(dtype=torch.float32) -> torch.Tensor:
    """Ortho DCT-II (64, 13), torchaudio/functional/functional.py:636-667."""
    n = torch.arange(float(N_MELS), dtype=dtype)
    k = torch.arange(float(N_MFCC), dtype=dtype).unsqueeze(1)
    dct = torch.cos(math.pi / float(N_MELS) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(N_MELS))
    return dct.t().contiguous()


# --------------------------------------------------------------------------
# waveform conditioning (src/preprocessing.py:185-212, 358-385)
# --------------------------------------------------------------------------
def to_mono(w: torch.Tensor) -> torch.Tensor:
    """src/preprocessing.py:185-197."""
    return w if w.shape[0] == 1 else w.mean(dim=0, keepdim=True)


def normalize(w: torch.Tensor) -> torch.Tensor:
    """Peak normalise, skipped for an all-zero clip. src/preprocessing.py:199-212."""
    peak = w.abs().max() if w.numel() else torch.zeros(())
    return w / peak if peak > 0 else w


def pad_or_trim(w: torch.Tensor, length: int = SEGMENT) -> torch.Tensor:
    """Centre trim / symmetric zero pad (left = floor(p/2)). src/preprocessing.py:358-385."""
    n = w.shape[1]
    if n == length:
        return w
    if n > length:
        s = (n - length) // 2
        return w[:, s:s + length]
    p = length - n
    return F.pad(w, (p // 2, p - p // 2))


# --------------------------------------------------------------------------
# features, fp32 explicit restatement (src/preprocessing.py:387-489)
# --------------------------------------------------------------------------
def pre_emphasis(w: torch.Tensor, coef: float = 0.97) -> torch.Tensor:
    """apply_pre_emphasis (src/preprocessing.py:213-239): y[0] = x[0], y[n] = x[n] - coef * x[n-1].  (1, n)."""
    return torch.cat([w[:, :1], w[:, 1:] - coef * w[:, :-1]], dim=1)


def pcen01(mel: torch.Tensor, alpha: float = 0.98, delta: float = 2.0, r: float = 0.5, eps: float = 1e-6) -> torch.Tensor:
    """apply_pcen (src/preprocessing.py:305-340) + the per-clip min/max scaling of extract_mel_spectrogram (:400-404).
    (64, frames) mel POWER -> (64, frames) in [0, 1].  The smoother is avg_pool2d(kernel (1, 10), stride 1, padding
    (0, 5)) cut back to `frames` columns: the zero-padded mean of columns t-5 .. t+4 (count_include_pad)."""
    sm = F.avg_pool2d(mel[None, None], kernel_size=(1, 10), stride=(1, 1), padding=(0, 5))[0, 0][:, :mel.shape[1]]
    pc = (mel / (eps + sm).pow(alpha) + delta).pow(r) - delta ** r
    return (pc - pc.min()) / (pc.max() - pc.min() + 1e-8)


class Explicit32:
    """fp32 restatement of extract_features for the 90-feature configs: the training config (train.py:264-287) and,
    additionally, its use_pre_emphasis / use_pcen variants (same 90 rows; SURVEY.md 8f rank 1, partial)."""

    def __init__(self, use_pcen: bool = False, use_pre_emphasis: bool = False, pre_emphasis_coef: float = 0.97,
                 use_delta_delta: bool = False):
        self.use_pcen, self.use_pre_emphasis, self.coef = use_pcen, use_pre_emphasis, pre_emphasis_coef
        self.use_delta_delta = use_delta_delta
        self.win512 = torch.zeros(N_FFT)
        off = (N_FFT - WIN_LENGTH) // 2  # torch.stft centres the 400-window in 512 (offset 56)
        self.win512[off:off + WIN_LENGTH] = hann_window()
        self.fb = mel_filterbank()
        self.dct = dct_matrix()

    def mel_power(self, w: torch.Tensor) -> torch.Tensor:
        """(1, n) -> (64, frames).  T.MelSpectrogram: torchaudio/transforms/_transforms.py:621-631,
        functional.py:121-145 (reflect pad n_fft//2, power=2)."""
        xp = F.pad(w.unsqueeze(0), (N_FFT // 2, N_FFT // 2), mode="reflect").squeeze(0)
        frames = xp.unfold(-1, N_FFT, HOP_LENGTH).squeeze(0)          # (frames, 512)
        spec = torch.fft.rfft(frames * self.win512, dim=-1)            # (frames, 257)
        power = spec.real ** 2 + spec.imag ** 2
        return (power @ self.fb).t().contiguous()                      # (64, frames)

    @staticmethod
    def to_db(mel: torch.Tensor) -> torch.Tensor:
        """AmplitudeToDB('power', top_db=80): torchaudio/functional/functional.py:390-404."""
        db = 10.0 * torch.log10(torch.clamp(mel, min=AMIN))
        return torch.max(db, db.amax() - TOP_DB)

    def features(self, w: torch.Tensor) -> torch.Tensor:
        """(1, 16000) normalised clip -> (1, 90, frames)."""
        if self.use_pre_emphasis:
            w = pre_emphasis(w, self.coef)                             # :452, both blocks use the emphasised signal
        mel = self.mel_power(w)
        db = self.to_db(mel)                                           # :398,407 and again :425
        if self.use_pcen:
            mel01 = pcen01(mel)                                        # :400-404
        else:
            mel01 = ((db + 80.0) / 80.0).clamp(0.0, 1.0)               # :409-410
        mf = (db.t() @ self.dct).t()                                   # torchaudio _transforms.py:716
        mf = (mf - mf.mean()) / (mf.std() + 1e-8)                      # :428 (unbiased std)
        pad = F.pad(mf.unsqueeze(0), (1, 1), mode="replicate").squeeze(0)
        delta = (pad[:, 2:] - pad[:, :-2]) / 2                         # :353-355
        rows = [mel01, mf, delta]
        if self.use_delta_delta:                                       # :463-466: deltas of the deltas
            pad2 = F.pad(delta.unsqueeze(0), (1, 1), mode="replicate").squeeze(0)
            rows.append((pad2[:, 2:] - pad2[:, :-2]) / 2)
        return torch.cat(rows, dim=0).unsqueeze(0)                     # :487

    def process(self, w: torch.Tensor) -> torch.Tensor:
        """normalize -> pad_or_trim -> extract_features (src/preprocessing.py:508-515, 16 kHz mono)."""
        return self.features(pad_or_trim(normalize(to_mono(w))))


def features_fp64(w: np.ndarray, use_pcen: bool = False, use_pre_emphasis: bool = False, coef: float = 0.97,
                  use_delta_delta: bool = False) -> np.ndarray:
    """float64 ground truth: (n,) raw clip -> (90, 101).  Same steps as Explicit32.process."""
    x = np.asarray(w, dtype=np.float64).reshape(-1)
    peak = np.abs(x).max() if x.size else 0.0
    if peak > 0:
        x = x / peak
    n = x.size
    if n > SEGMENT:
        s = (n - SEGMENT) // 2
        x = x[s:s + SEGMENT]
    elif n < SEGMENT:
        p = SEGMENT - n
        x = np.pad(x, (p // 2, p - p // 2))
    return features_fp64_window(x, use_pcen, use_pre_emphasis, coef, use_delta_delta)


def features_fp64_window(x: np.ndarray, use_pcen: bool = False, use_pre_emphasis: bool = False, coef: float = 0.97,
                         use_delta_delta: bool = False) -> np.ndarray:
    """float64 features of an already conditioned 16000-sample window."""
    if use_pre_emphasis:
        x = np.concatenate([x[:1], x[1:] - coef * x[:-1]])
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    nfr = (xp.size - N_FFT) // HOP_LENGTH + 1
    idx = np.arange(N_FFT)[None, :] + HOP_LENGTH * np.arange(nfr)[:, None]
    win = np.zeros(N_FFT)
    off = (N_FFT - WIN_LENGTH) // 2
    # the reference's window IS the fp32 table (its rounding leaks ~1e-7 of a DC / tonal peak into
    # far bins); the ground truth keeps that table and does everything downstream in float64
    win[off:off + WIN_LENGTH] = hann_window(torch.float32).double().numpy()
    spec = np.fft.rfft(xp[idx] * win, axis=-1)
    power = spec.real ** 2 + spec.imag ** 2
    mel = power @ mel_filterbank(torch.float32).double().numpy()               # (frames, 64)
    db = 10.0 * np.log10(np.maximum(mel, AMIN))
    db = np.maximum(db, db.max() - TOP_DB)
    if use_pcen:
        mp = np.pad(mel, ((5, 5), (0, 0)))                                     # (frames + 10, 64), zero padded
        cs = np.concatenate([np.zeros((1, mel.shape[1])), np.cumsum(mp, axis=0)])
        sm = (cs[10:10 + mel.shape[0]] - cs[:mel.shape[0]]) / 10.0             # mean of frames t-5 .. t+4
        pc = (mel / (1e-6 + sm) ** 0.98 + 2.0) ** 0.5 - 2.0 ** 0.5
        mel01 = ((pc - pc.min()) / (pc.max() - pc.min() + 1e-8)).T
    else:
        mel01 = np.clip((db + 80.0) / 80.0, 0.0, 1.0).T
    mf = (db @ dct_matrix(torch.float32).double().numpy()).T                   # (13, frames)
    mf = (mf - mf.mean()) / (mf.std(ddof=1) + 1e-8)
    pad = np.pad(mf, ((0, 0), (1, 1)), mode="edge")
    delta = (pad[:, 2:] - pad[:, :-2]) / 2
    rows = [mel01, mf, delta]
    if use_delta_delta:
        pad2 = np.pad(delta, ((0, 0), (1, 1)), mode="edge")
        rows.append((pad2[:, 2:] - pad2[:, :-2]) / 2)
    return np.concatenate(rows, axis=0)


def ill_conditioned_fraction(x: np.ndarray, rel: float = 1e-9) -> float:
    """Fraction of above-floor mel cells whose power is < rel * (that frame's largest FFT-bin power).

    An fp32 FFT leaves rounding noise of ~1e-15 of the frame's peak bin power in every bin, so such
    cells (more than ~90 dB below their frame's peak: DC and noiseless pure tones) differ
    by more than 1e-4 between ANY two fp32 implementations -- the reference itself is 7.7e-4 away
    from the float64 truth on a DC clip.  Parity tests hold these clips to a looser, stated bound."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    idx = np.arange(N_FFT)[None, :] + HOP_LENGTH * np.arange((xp.size - N_FFT) // HOP_LENGTH + 1)[:, None]
    win = np.zeros(N_FFT)
    off = (N_FFT - WIN_LENGTH) // 2
    win[off:off + WIN_LENGTH] = hann_window(torch.float32).double().numpy()
    spec = np.fft.rfft(xp[idx] * win, axis=-1)
    power = spec.real ** 2 + spec.imag ** 2
    mel = power @ mel_filterbank(torch.float32).double().numpy()
    db = 10.0 * np.log10(np.maximum(mel, AMIN))
    above = db > db.max() - TOP_DB
    bad = above & (mel < rel * power.max(axis=1, keepdims=True))
    return float(bad.sum()) / max(1, int(above.sum()))


def conditioned_window(w: np.ndarray) -> np.ndarray:
    """normalize + pad_or_trim in float64 (the input features_fp64_window expects)."""
    x = np.asarray(w, dtype=np.float64).reshape(-1)
    peak = np.abs(x).max() if x.size else 0.0
    if peak > 0:
        x = x / peak
    n = x.size
    if n > SEGMENT:
        s = (n - SEGMENT) // 2
        x = x[s:s + SEGMENT]
    elif n < SEGMENT:
        p = SEGMENT - n
        x = np.pad(x, (p // 2, p - p // 2))
    return x


def stream_windows(n: int, hop: int, window: int = SEGMENT) -> int:
    """Number of windows RealtimePreprocessor.add_audio emits for n buffered samples
    (src/preprocessing.py:601-610): while len >= window: emit, drop hop."""
    return 0 if n < window else (n - window) // hop + 1


def stream_features(x: torch.Tensor, hop: int, ex: Optional[Explicit32] = None) -> torch.Tensor:
    """RealtimePreprocessor semantics: per-window peak normalise, no pad/trim.
    (n,) -> (windows, 1, 90, 101).  src/preprocessing.py:582-612."""
    ex = ex or Explicit32()
    x = x.reshape(1, -1)
    out = []
    for j in range(stream_windows(x.shape[1], hop)):
        out.append(ex.features(normalize(x[:, j * hop:j * hop + SEGMENT])))
    return torch.stack(out) if out else torch.zeros(0, 1, N_FEATURES, N_FRAMES)


def spectral_contrast_band_edges(n_freq: int = N_FFT // 2 + 1, n_bands: int = 6) -> list:
    """The band edges of extract_spectral_contrast (src/preprocessing.py:268-269): logspace(0, log10(n_freq), n_bands + 2)
    truncated to int -- [1, 2, 4, 10, 23, 52, 116, 256] (10^log10(257) rounds just below 257) for the default 257 bins / 6 bands."""
    return torch.clamp(torch.logspace(0, float(np.log10(n_freq)), n_bands + 2).int(), 0, n_freq).tolist()


def spectral_contrast(w: torch.Tensor, n_bands: int = 6) -> torch.Tensor:
    """Restatement of extract_spectral_contrast (src/preprocessing.py:242-303) WITHOUT the spectral centroid row -- kept
    only to pin down why the reference's 110-feature constructor default is not a parity target: the first band
    [1, 2) holds ONE bin, so top_idx = max(1, int(0.8)) = 1 and sorted_band[:, 1:, :] is EMPTY, its mean is NaN (:287-291),
    and the block-wide normalisation (:301) spreads that NaN over every row.  Returns (1, n_bands + 1, frames)."""
    xp = F.pad(w.unsqueeze(0), (N_FFT // 2, N_FFT // 2), mode="reflect").squeeze(0)
    frames = xp.unfold(-1, N_FFT, HOP_LENGTH).squeeze(0)
    win = torch.zeros(N_FFT)
    off = (N_FFT - WIN_LENGTH) // 2
    win[off:off + WIN_LENGTH] = hann_window()
    spec = torch.fft.rfft(frames * win, dim=-1)
    power = (spec.real ** 2 + spec.imag ** 2).t().unsqueeze(0)           # (1, 257, frames)
    edges = spectral_contrast_band_edges(power.shape[1], n_bands)
    out = torch.zeros(1, n_bands + 1, power.shape[2])
    for i in range(n_bands):
        lo, hi = edges[i], max(edges[i + 1], edges[i] + 1)
        band, _ = torch.sort(power[:, lo:min(hi, power.shape[1]), :], dim=1)
        n = band.shape[1]
        top, bot = max(1, int(n * 0.8)), max(1, int(n * 0.2))
        out[:, i, :] = torch.log1p(band[:, top:, :].mean(dim=1)) - torch.log1p(band[:, :bot, :].mean(dim=1))
    return (out - out.mean()) / (out.std() + 1e-8)


# --------------------------------------------------------------------------
# CoughDetectorResidual forward (src/model.py:210-293), functional form
# --------------------------------------------------------------------------
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def cnn_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """fp32 eval-mode forward, (B,1,90,101) -> logits (B,2).  src/model.py:253-259, 285-293."""
    x = F.conv2d(x, sd["conv1.0.weight"], sd["conv1.0.bias"], stride=2, padding=3)
    x = F.max_pool2d(F.relu(_bn(x, sd, "conv1.1")), 2)                # :227-232
    for i in (0, 1):
        p = f"res_blocks.{i}."
        idt = _bn(F.conv2d(x, sd[p + "skip.0.weight"], sd[p + "skip.0.bias"], stride=2), sd, p + "skip.1")
        o = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=2, padding=1), sd, p + "bn1"))
        o = _bn(F.conv2d(o, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1), sd, p + "bn2")
        x = F.relu(o + idt)
    x = x.mean(dim=(2, 3))                                            # AdaptiveAvgPool2d(1) + Flatten
    return F.linear(x, sd["fc.2.weight"], sd["fc.2.bias"])           # Dropout is identity in eval


def cnn_forward_small(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """CoughDetectorSmall in eval mode, functional form (src/model.py:160-201): conv + BN + ReLU + pool, three
    depthwise-separable blocks (depthwise 3x3, pointwise 1x1, BN, ReLU, pool / global pool), Linear 128 -> 64 -> 2."""
    f = "features."
    x = F.max_pool2d(F.relu(_bn(F.conv2d(x, sd[f + "0.weight"], sd[f + "0.bias"], padding=1), sd, f + "1")), 2)
    for a, last in ((4, False), (9, False), (14, True)):
        c = sd[f"{f}{a}.weight"].shape[0]
        x = F.conv2d(x, sd[f"{f}{a}.weight"], sd[f"{f}{a}.bias"], padding=1, groups=c)
        x = F.relu(_bn(F.conv2d(x, sd[f"{f}{a + 1}.weight"], sd[f"{f}{a + 1}.bias"]), sd, f"{f}{a + 2}"))
        x = x.mean(dim=(2, 3)) if last else F.max_pool2d(x, 2)
    x = F.relu(F.linear(x, sd["classifier.1.weight"], sd["classifier.1.bias"]))
    return F.linear(x, sd["classifier.4.weight"], sd["classifier.4.bias"])


def cnn_forward_standard(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """CoughDetector in eval mode, functional form (src/model.py:86-128): four ConvBlocks (:31-40), GAP, fc."""
    for i in range(4):
        p = f"conv_layers.{i}."
        x = F.max_pool2d(F.relu(_bn(F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1), sd, p + "bn")), 2)
    x = x.mean(dim=(2, 3))
    x = F.relu(F.linear(x, sd["fc.0.weight"], sd["fc.0.bias"]))
    return F.linear(x, sd["fc.3.weight"], sd["fc.3.bias"])


def fold_bn(w, b, sd, p) -> Tuple[torch.Tensor, torch.Tensor]:
    """w' = w*g/sqrt(v+eps), b' = (b-mu)*g/sqrt(v+eps)+beta."""
    s = sd[p + ".weight"].double() / torch.sqrt(sd[p + ".running_var"].double() + BN_EPS)
    wf = w.double() * s.view(-1, 1, 1, 1)
    bf = (b.double() - sd[p + ".running_mean"].double()) * s + sd[p + ".bias"].double()
    return wf.float(), bf.float()


def _q(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def cnn_forward_bf16(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """Emulates the GPU numerics: BN folded into bf16 weights, bf16 activations between
    layers, fp32 accumulation and fp32 bias/residual/pool math.  Used to state the bf16
    tolerance band, not as a parity target by itself."""
    def conv(x, wkey, bkey, bn, **kw):
        w, b = fold_bn(sd[wkey], sd[bkey], sd, bn)
        return F.conv2d(_q(x).double(), _q(w).double(), None, **kw).float() + b.view(1, -1, 1, 1)
    x = F.max_pool2d(F.relu(conv(x, "conv1.0.weight", "conv1.0.bias", "conv1.1", stride=2, padding=3)), 2)
    for i in (0, 1):
        p = f"res_blocks.{i}."
        idt = conv(x, p + "skip.0.weight", p + "skip.0.bias", p + "skip.1", stride=2)
        o = F.relu(conv(x, p + "conv1.weight", p + "conv1.bias", p + "bn1", stride=2, padding=1))
        o = conv(o, p + "conv2.weight", p + "conv2.bias", p + "bn2", padding=1)
        x = F.relu(o + idt)
    x = x.mean(dim=(2, 3))
    return F.linear(x, sd["fc.2.weight"], sd["fc.2.bias"])


def predict(logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """softmax + argmax, src/model.py:261-265."""
    probs = F.softmax(logits, dim=1)
    return probs.argmax(dim=1), probs


# --------------------------------------------------------------------------
# decision post-pass (src/inference.py:216-241) on a stream-time clock
# --------------------------------------------------------------------------
def decisions_stream(probs: Sequence[float], hop_s: float, threshold: float = 0.5,
                     smoothing_window: int = 3, debounce_s: float = 0.5,
                     chunk_windows: Optional[Sequence[int]] = None) -> List[Tuple[int, float]]:
    """Sequential smoothing / threshold / debounce over per-window P(cough).

    Mirrors process_audio_chunk: deque(maxlen) mean (src/inference.py:111,220-223), fire when
    smoothed >= threshold and time since last fire >= debounce (:229-230).  ``datetime.now()``
    (:226) is replaced by stream time t_j = j*hop_s + window end; last_detection_time starts at
    0 (:114) which in wall-clock terms is "infinitely long ago", so the first fire is always
    allowed.  ``chunk_windows`` (number of windows produced by each add_audio call) reproduces
    the early ``return`` at :239 that drops the remaining windows of a chunk after a fire.
    Returns [(window_index, smoothed_confidence)].
    """
    hist: deque = deque(maxlen=smoothing_window)
    fired: List[Tuple[int, float]] = []
    last_t = -math.inf
    bounds = None
    if chunk_windows is not None:
        bounds = np.cumsum([0] + list(chunk_windows))
    j, n = 0, len(probs)
    while j < n:
        hist.append(float(probs[j]))
        sm = float(np.mean(hist))
        t = j * hop_s + 1.0
        if sm >= threshold and (t - last_t) >= debounce_s:
            last_t = t
            fired.append((j, sm))
            if bounds is not None:  # early return: skip to the next chunk's first window
                k = int(np.searchsorted(bounds, j, side="right"))
                j = int(bounds[k]) if k < len(bounds) else n
                continue
        j += 1
    return fired


# --------------------------------------------------------------------------
# resampling (src/preprocessing.py:146-183 -> torchaudio T.Resample defaults)
# --------------------------------------------------------------------------
def resample_kernel(orig_sr: int, new_sr: int):
    """(kernel (new, 2*width+orig) fp32, width, orig, new): torchaudio/functional/functional.py:1305-1404 with
    sinc_interp_hann, lowpass_filter_width=6, rolloff=0.99, dtype=None (float64 maths, the phase term through
    float32 as torch's int64-arange / int division produces it)."""
    g = math.gcd(int(orig_sr), int(new_sr))
    orig, new = int(orig_sr) // g, int(new_sr) // g
    lpw = 6
    base = min(orig, new) * 0.99
    width = math.ceil(lpw * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, :] / orig
    t = (torch.arange(0, -new, -1)[:, None] / new).double() + idx
    t = (t * base).clamp_(-lpw, lpw)
    window = torch.cos(t * math.pi / lpw / 2) ** 2
    t = t * math.pi
    k = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t) * window * (base / orig)
    return k.float(), width, orig, new


def resample(w: torch.Tensor, orig_sr: int, new_sr: int = SAMPLE_RATE) -> torch.Tensor:
    """(rows, n) -> (rows, ceil(new*n/orig)); torchaudio/functional/functional.py:1406-1432 (pad, strided conv1d, trim)."""
    if orig_sr == new_sr:
        return w
    k, width, orig, new = resample_kernel(orig_sr, new_sr)
    rows, n = w.shape
    x = F.pad(w, (width, width + orig))
    y = F.conv1d(x[:, None], k[:, None], stride=orig).transpose(1, 2).reshape(rows, -1)
    return y[:, : math.ceil(new * n / orig)]


# --------------------------------------------------------------------------
# SpecAugment (src/augmentation.py:271-331; torchaudio/functional/functional.py mask_along_axis)
# --------------------------------------------------------------------------
def spec_augment_intervals(n_rows: int, n_cols: int = N_FRAMES, freq_mask_param: int = 10, time_mask_param: int = 20,
                           n_freq_masks: int = 2, n_time_masks: int = 2, p: float = 0.5) -> Optional[List[Tuple[int, int]]]:
    """One SpecAugment.__call__ as its random draws: ``random.random() > p`` -> None (input returned untouched,
    src/augmentation.py:314-315); otherwise n_freq_masks row intervals then n_time_masks column intervals [start, end),
    each from ``value = rand(1) * param; min_value = rand(1) * (size - value); start = long(min_value);
    end = long(min_value) + long(value)`` -- and NO draw when param < 1 (mask_along_axis returns early).  Uses the global
    ``random`` / ``torch`` generators exactly like the reference."""
    import random
    if random.random() > p:
        return None
    out = []
    for param, size, n in ((freq_mask_param, n_rows, n_freq_masks), (time_mask_param, n_cols, n_time_masks)):
        for _ in range(n):
            if param < 1:
                out.append((0, 0))
                continue
            value = torch.rand(1) * param
            min_value = torch.rand(1) * (size - value)
            out.append((int(min_value.long()), int(min_value.long() + value.long())))
    return out


def spec_augment_apply(x: torch.Tensor, intervals: Optional[Sequence[Tuple[int, int]]], n_freq_masks: int = 2) -> torch.Tensor:
    """masked_fill(mask, 0.0) of the row intervals (the first n_freq_masks) and the column intervals (the rest) on (..., F, T)."""
    if intervals is None:
        return x
    y = x.clone()
    for k, (s, e) in enumerate(intervals):
        if k < n_freq_masks:
            y[..., s:e, :] = 0.0
        else:
            y[..., :, s:e] = 0.0
    return y
