"""CoughDetectorResidual whose forward runs in libcough_b200.so.

Host-side mirror of the reference's ``src/model.py`` for the one architecture on the hot path
(``CoughDetectorResidual``, src/model.py:210-293).  The module tree, parameter names and shapes
are the reference's, so ``state_dict`` / ``load_state_dict`` / checkpoints (src/train.py:183-199)
interchange unchanged; ``forward`` hands the parameters to the C ABI (``cd_cnn_load`` folds
BatchNorm and repacks, once per weight version) and calls ``cd_cnn_forward``.

Inference only: ``forward`` in training mode raises (batch-statistics BatchNorm, Dropout and the
backward pass are outside the north-star path).  There is no PyTorch fallback.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from . import _lib

#: precision used when a model does not say otherwise: "bf16" (tcgen05 tensor cores) or "fp32" (CUDA cores)
DEFAULT_PRECISION = "bf16"
_PREC = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16}


class ResidualBlock(nn.Module):
    """Parameter container with the reference's names (src/model.py:268-283); computed by the fused kernels."""

    def __init__(self, in_channels: int, out_channels: int, stride: int = 2):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=stride, padding=1)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.skip = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, stride=stride), nn.BatchNorm2d(out_channels))

    def forward(self, x):  # pragma: no cover - blocks are never run on their own
        raise RuntimeError("ResidualBlock is executed inside CoughDetectorResidual.forward (fused CUDA kernels)")


class CoughDetectorResidual(nn.Module):
    """7x7/s2 stem + BN + ReLU + maxpool2, two residual blocks, GAP, Linear(128->2) (src/model.py:210-265)."""

    def __init__(self, n_mels: int = 64, num_classes: int = 2, in_channels: int = 1,
                 channels: Tuple[int, ...] = (32, 64, 128), dropout: float = 0.5, precision: str = None):
        super().__init__()
        if num_classes != 2 or in_channels != 1 or tuple(channels) != (32, 64, 128):
            raise NotImplementedError("the B200 kernels are built for in_channels=1, channels=(32,64,128), num_classes=2")
        self.conv1 = nn.Sequential(nn.Conv2d(in_channels, channels[0], 7, stride=2, padding=3),
                                   nn.BatchNorm2d(channels[0]), nn.ReLU(), nn.MaxPool2d(2))
        self.res_blocks = nn.ModuleList()
        cin = channels[0]
        for cout in channels[1:]:
            self.res_blocks.append(ResidualBlock(cin, cout))
            cin = cout
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Sequential(nn.Flatten(), nn.Dropout(dropout), nn.Linear(channels[-1], num_classes))
        self.precision = precision or DEFAULT_PRECISION

    # ---- weight upload, once per (device, parameter version) -----------------------------------------
    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def _handle(self, x: torch.Tensor) -> "_lib.Handle":
        h = _lib.get_handle(x.device if x.is_cuda else None)
        key = (id(self), self._weights_key())
        if h._weights_key != key:
            h.load_state_dict(self.state_dict(), key=key)
        return h

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, 1, 90, 101) -- or (B, 1, 103, 101) for use_delta_delta features -- -> logits (B, 2) (reference :253-259)."""
        if self.training:
            raise RuntimeError("cough_detector_b200 is inference-only: call model.eval() before forward()")
        if self.precision not in _PREC:
            raise ValueError(f"precision must be one of {list(_PREC)}, got {self.precision!r}")
        if x.dim() == 4 and x.shape[2] not in (_lib.N_FEATURES, _lib.N_FEATURES_DD):
            raise NotImplementedError(f"the B200 CNN kernels are built for {_lib.N_FEATURES} or {_lib.N_FEATURES_DD} feature rows, "
                                      f"got {x.shape[2]} (spectral-contrast inputs have no kernels; no fallback)")
        h = self._handle(x)
        xd = x.detach().to(device=f"cuda:{h.device}", dtype=torch.float32, non_blocking=True)
        logits = h.cnn_forward(xd, _PREC[self.precision])
        return logits if x.is_cuda else logits.to(x.device)

    def predict(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """softmax + argmax (reference :261-265): (LongTensor[B], Tensor[B, 2])."""
        logits = self.forward(x)
        h = _lib.get_handle(x.device if x.is_cuda else None)
        preds, probs, _ = h.predict(logits.to(f"cuda:{h.device}"))
        return (preds, probs) if x.is_cuda else (preds.to(x.device), probs.to(x.device))

    def forward_audio(self, audio: torch.Tensor, return_features: bool = False, preprocessor=None):
        """Additive fused entry: (B, 16000) waveforms -> logits, features never leave the GPU.  `preprocessor`
        (an AudioPreprocessor) supplies the use_pre_emphasis / use_pcen / use_delta_delta options for this call;
        without one the training configuration (all off) is used."""
        if self.training:
            raise RuntimeError("cough_detector_b200 is inference-only: call model.eval() before forward_audio()")
        h = self._handle(audio)
        opts = preprocessor.feat_options if preprocessor is not None else _lib.TRAINING_CONFIG
        if audio.dtype not in (torch.float32, torch.int16):
            audio = audio.to(torch.float32)
        if not audio.is_cuda and not return_features:
            # host audio (float32 or int16 PCM): the library's pipelined host path -- chunked copies overlapped with the
            # kernels, int16 at half the PCIe bytes -- instead of one blocking upload
            out = h.e2e_forward_host(audio.detach(), _PREC[self.precision], opts=opts)
            return out.clone()
        xd = audio.detach().to(device=f"cuda:{h.device}", non_blocking=True)
        out = h.e2e_forward(xd, _PREC[self.precision], return_features, opts=opts)
        if audio.is_cuda:
            return out
        return tuple(o.to(audio.device) for o in out) if return_features else out.to(audio.device)


class _PlainModel(nn.Module):
    """Shared forward / predict of the two model types that run on the fp32 CUDA-core kernels (csrc/cnn_plain.cu)."""
    _kind = ""

    def _handle(self, x: torch.Tensor) -> "_lib.Handle":
        h = _lib.get_handle(x.device if x.is_cuda else None)
        key = (id(self), tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values()))
        if getattr(h, "_plain_keys", {}).get(self._kind) != key:
            h.load_plain_state_dict(self._kind, self.state_dict(), key=key)
        return h

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("cough_detector_b200 is inference-only: call model.eval() before forward()")
        if x.dim() == 3:
            x = x.unsqueeze(0)
        if x.dim() != 4 or x.shape[1] != 1:
            raise NotImplementedError("the B200 kernels take (B, 1, rows, 101) feature tensors (in_channels = 1)")
        h = self._handle(x)
        logits = h.plain_forward(self._kind, x.detach().to(device=f"cuda:{h.device}", dtype=torch.float32, non_blocking=True))
        return logits if x.is_cuda else logits.to(x.device)

    def predict(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        logits = self.forward(x)
        h = _lib.get_handle(x.device if x.is_cuda else None)
        preds, probs, _ = h.predict(logits.to(f"cuda:{h.device}"))
        return (preds, probs) if x.is_cuda else (preds.to(x.device), probs.to(x.device))


class ConvBlock(nn.Module):
    """Parameter container with the reference's names (src/model.py:11-40): conv, bn (+ ReLU, MaxPool2d, Dropout2d)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1, padding: int = 1, pool_size: int = 2):
        super().__init__()
        if (kernel_size, stride, padding, pool_size) != (3, 1, 1, 2):
            raise NotImplementedError("the B200 kernel is the reference's default ConvBlock: 3x3, stride 1, padding 1, pool 2")
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.bn = nn.BatchNorm2d(out_channels)
        self.pool = nn.MaxPool2d(2)
        self.dropout = nn.Dropout2d(0.1)


class CoughDetector(_PlainModel):
    """model_type 'standard' (src/model.py:43-143): four ConvBlocks 1 -> 32 -> 64 -> 128 -> 256, GAP, Linear 256 -> 128 -> 2.
    fp32 on CUDA cores (off the north-star path; the tensor-core kernels are CoughDetectorResidual's)."""
    _kind = "standard"

    def __init__(self, n_mels: int = 64, num_classes: int = 2, in_channels: int = 1,
                 channels: Tuple[int, ...] = (32, 64, 128, 256), fc_hidden: int = 128, dropout: float = 0.5):
        super().__init__()
        if num_classes != 2 or in_channels != 1 or tuple(channels) != (32, 64, 128, 256) or fc_hidden != 128:
            raise NotImplementedError("the B200 kernels are built for in_channels=1, channels=(32,64,128,256), fc_hidden=128, num_classes=2")
        self.n_mels, self.num_classes = n_mels, num_classes
        layers, cur = [], in_channels
        for c in channels:
            layers.append(ConvBlock(cur, c))
            cur = c
        self.conv_layers = nn.Sequential(*layers)
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Sequential(nn.Linear(channels[-1], fc_hidden), nn.ReLU(), nn.Dropout(dropout), nn.Linear(fc_hidden, num_classes))


class CoughDetectorSmall(_PlainModel):
    """model_type 'small' (src/model.py:146-207), the reference's default: conv 1 -> 16, three depthwise-separable blocks
    16 -> 32 -> 64 -> 128, GAP, Linear 128 -> 64 -> 2.  fp32 on CUDA cores."""
    _kind = "small"

    def __init__(self, n_mels: int = 64, num_classes: int = 2, in_channels: int = 1):
        super().__init__()
        if num_classes != 2 or in_channels != 1:
            raise NotImplementedError("the B200 kernels are built for in_channels=1, num_classes=2")
        self.features = nn.Sequential(
            nn.Conv2d(in_channels, 16, 3, padding=1), nn.BatchNorm2d(16), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(16, 16, 3, padding=1, groups=16), nn.Conv2d(16, 32, 1), nn.BatchNorm2d(32), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(32, 32, 3, padding=1, groups=32), nn.Conv2d(32, 64, 1), nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(64, 64, 3, padding=1, groups=64), nn.Conv2d(64, 128, 1), nn.BatchNorm2d(128), nn.ReLU(),
            nn.AdaptiveAvgPool2d((1, 1)))
        self.classifier = nn.Sequential(nn.Flatten(), nn.Linear(128, 64), nn.ReLU(), nn.Dropout(0.3), nn.Linear(64, num_classes))


def create_model(model_type: str = "standard", **kwargs) -> nn.Module:
    """Reference factory (src/model.py:296-316): 'residual' is the hot path (tcgen05 kernels, bf16 or fp32); 'small' (the
    reference's default model_type) and 'standard' run on fp32 CUDA-core kernels; an unknown type raises ValueError."""
    models = {"standard": CoughDetector, "small": CoughDetectorSmall, "residual": CoughDetectorResidual}
    if model_type not in models:
        raise ValueError(f"Unknown model type: {model_type}. Choose from {list(models.keys())}")
    return models[model_type](**kwargs)


def count_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
