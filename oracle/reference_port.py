"""CPU port of the reference's per-clip path, built on the same third-party calls.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/cough_oracle.py header).  This is the
``cpu_baseline`` (kind "port") that ``bench.py`` times beside the GPU path: it makes the
same torchaudio / torch calls, in the same order and with the same redundancy, as the
reference's per-clip loop -- including the mel spectrogram being computed twice (once by
the mel block, once inside T.MFCC) -- so its clips/s is a faithful stand-in
for ``/root/reference`` on a box where ``/root/reference`` does not exist.

Call sites followed: src/preprocessing.py:94-127 (transform construction, 90-feature flags
of src/train.py:264-287), :199-212 normalize, :358-385 pad_or_trim, :387-430 mel + MFCC,
:342-356 deltas, :432-489 concat; src/model.py:210-293 via torch.nn modules.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchaudio.transforms as T

from . import cough_oracle as co


class PortPreprocessor:
    def __init__(self, use_pcen: bool = False, use_pre_emphasis: bool = False, pre_emphasis_coef: float = 0.97,
                 use_delta_delta: bool = False):
        self.use_pcen, self.use_pre_emphasis, self.coef = use_pcen, use_pre_emphasis, pre_emphasis_coef
        self.use_delta_delta = use_delta_delta
        mk = dict(n_fft=co.N_FFT, win_length=co.WIN_LENGTH, hop_length=co.HOP_LENGTH,
                  n_mels=co.N_MELS, f_min=co.F_MIN, f_max=co.F_MAX)
        self.mel = T.MelSpectrogram(sample_rate=co.SAMPLE_RATE, power=2.0, normalized=False,
                                    center=True, pad_mode="reflect", **mk)
        self.to_db = T.AmplitudeToDB(stype="power", top_db=80)
        self.mfcc = T.MFCC(sample_rate=co.SAMPLE_RATE, n_mfcc=co.N_MFCC, melkwargs=mk)

    def features(self, w: torch.Tensor) -> torch.Tensor:
        if self.use_pre_emphasis:
            w = co.pre_emphasis(w, self.coef)
        if self.use_pcen:
            mel01 = co.pcen01(self.mel(w)[0]).unsqueeze(0)
        else:
            mel01 = ((self.to_db(self.mel(w)) + 80) / 80).clamp(0, 1)
        mf = self.mfcc(w)
        mf = (mf - mf.mean()) / (mf.std() + 1e-8)
        pad = F.pad(mf, (1, 1), mode="replicate")
        delta = (pad[:, :, 2:] - pad[:, :, :-2]) / 2
        rows = [mel01, mf, delta]
        if self.use_delta_delta:
            pad2 = F.pad(delta, (1, 1), mode="replicate")
            rows.append((pad2[:, :, 2:] - pad2[:, :, :-2]) / 2)
        return torch.cat(rows, dim=1)

    def process(self, w: torch.Tensor) -> torch.Tensor:
        return self.features(co.pad_or_trim(co.normalize(co.to_mono(w))))


class _Block(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=2, padding=1)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(cout)
        self.skip = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=2), nn.BatchNorm2d(cout))

    def forward(self, x):
        return F.relu(self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x))))) + self.skip(x))


class PortResidual(nn.Module):
    """Same module tree / state_dict keys as CoughDetectorResidual (src/model.py:210-247)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(1, 32, 7, stride=2, padding=3), nn.BatchNorm2d(32),
                                   nn.ReLU(), nn.MaxPool2d(2))
        self.res_blocks = nn.ModuleList([_Block(32, 64), _Block(64, 128)])
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Sequential(nn.Flatten(), nn.Dropout(0.5), nn.Linear(128, 2))

    def forward(self, x):
        x = self.conv1(x)
        for b in self.res_blocks:
            x = b(x)
        return self.fc(self.global_pool(x))


def build(sd: Dict[str, torch.Tensor]):
    m = PortResidual()
    m.load_state_dict(sd)
    return PortPreprocessor(), m.eval()


@torch.no_grad()
def run_clips(pp: PortPreprocessor, model: nn.Module, clips) -> torch.Tensor:
    """Per-clip loop normalize -> pad_or_trim -> extract_features -> model(x[None])."""
    out = []
    for w in clips:
        out.append(model(pp.process(w.reshape(1, -1)).unsqueeze(0))[0])
    return torch.stack(out)
