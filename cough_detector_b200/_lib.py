"""ctypes binding of libcough_b200.so (C ABI in include/cough_b200.h).

The library is the product; this file only marshals pointers.  There is NO fallback: if the
shared object is missing or no sm_100 device is usable, the first call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass
from typing import Dict, Optional

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libcough_b200.so")

SEGMENT = 16000
N_FEATURES = 90
N_FRAMES = 101
PRECISION_FP32 = 0
PRECISION_BF16 = 1
FEAT_NORMALIZE = 1
FEAT_PRE_EMPHASIS = 2
FEAT_PCEN = 4
FEAT_DELTA_DELTA = 8
N_FEATURES_DD = 103

_f32p = C.POINTER(C.c_float)


@dataclass(frozen=True)
class FeatOptions:
    """The optional feature stages of one AudioPreprocessor (src/preprocessing.py:32-51).  Passed with EVERY call
    (``flags`` / ``pre_emphasis_coef`` of the C ABI): the handle remembers nothing, so objects with different
    flags can share a handle.  The default is the training configuration (src/train.py:264-287)."""
    use_pre_emphasis: bool = False
    pre_emphasis_coef: float = 0.97
    use_pcen: bool = False
    use_delta_delta: bool = False

    def flags(self, normalize: bool = True) -> int:
        return ((FEAT_NORMALIZE if normalize else 0) | (FEAT_PRE_EMPHASIS if self.use_pre_emphasis else 0)
                | (FEAT_PCEN if self.use_pcen else 0) | (FEAT_DELTA_DELTA if self.use_delta_delta else 0))

    @property
    def n_features(self) -> int:
        return N_FEATURES_DD if self.use_delta_delta else N_FEATURES


TRAINING_CONFIG = FeatOptions()


class _ConvBn(C.Structure):
    _fields_ = [(n, _f32p) for n in ("weight", "bias", "bn_weight", "bn_bias", "bn_running_mean", "bn_running_var")]


class _ResidualWeights(C.Structure):
    _fields_ = [(n, _ConvBn) for n in ("stem", "b0_conv1", "b0_conv2", "b0_skip", "b1_conv1", "b1_conv2", "b1_skip")] \
        + [("fc_weight", _f32p), ("fc_bias", _f32p)]


class _Conv(C.Structure):
    _fields_ = [("weight", _f32p), ("bias", _f32p)]


class _Bn(C.Structure):
    _fields_ = [(n, _f32p) for n in ("weight", "bias", "running_mean", "running_var")]


class _DwPwBlock(C.Structure):
    _fields_ = [("dw", _Conv), ("pw", _Conv), ("bn", _Bn)]


class _SmallWeights(C.Structure):
    _fields_ = [("conv1", _Conv), ("bn1", _Bn), ("block", _DwPwBlock * 3)] + [(n, _f32p) for n in ("fc1_weight", "fc1_bias", "fc2_weight", "fc2_bias")]


class _StandardWeights(C.Structure):
    _fields_ = [("conv", _Conv * 4), ("bn", _Bn * 4)] + [(n, _f32p) for n in ("fc1_weight", "fc1_bias", "fc2_weight", "fc2_bias")]


# every symbol include/cough_b200.h declares: name -> (restype, argtypes)
_VP, _I64, _I, _F = C.c_void_p, C.c_int64, C.c_int, C.c_float
SYMBOLS = {
    "cd_create": (_I, [_I, C.POINTER(_VP)]),
    "cd_destroy": (_I, [_VP]),
    "cd_last_error": (C.c_char_p, [_VP]),
    "cd_version": (C.c_char_p, []),
    "cd_get_tables": (_I, [_VP, _VP, _VP]),
    "cd_feat_forward": (_I, [_VP, _VP, _I64, _I64, _VP, _I, _F, _VP]),
    "cd_feat_forward_i16": (_I, [_VP, _VP, _I64, _I64, _VP, _I, _F, _VP]),
    "cd_feat_forward_var": (_I, [_VP, _VP, _VP, _VP, _I64, _VP, _I, _F, _VP]),
    "cd_feat_forward_stream": (_I, [_VP, _VP, _I64, _I64, _I64, _VP, _I, _F, _VP]),
    "cd_stream_windows": (_I64, [_I64, _I64]),
    "cd_resample_length": (_I64, [_I64, _I, _I]),
    "cd_resample": (_I, [_VP, _VP, _I64, _I64, _I, _I, _VP, _VP]),
    "cd_condition_var": (_I, [_VP, _VP, _I, _VP, _VP, _VP, _VP, _I64, _I64, _I, _I, _VP, _VP]),
    "cd_host_alloc": (_I, [C.POINTER(_VP), C.c_size_t, _I]),
    "cd_host_free": (_I, [_VP]),
    "cd_cnn_load": (_I, [_VP, C.POINTER(_ResidualWeights)]),
    "cd_cnn_forward": (_I, [_VP, _VP, _I64, _VP, _I, _VP]),
    "cd_cnn_forward_rows": (_I, [_VP, _VP, _I, _I64, _VP, _I, _VP]),
    "cd_small_load": (_I, [_VP, C.POINTER(_SmallWeights)]),
    "cd_small_forward": (_I, [_VP, _VP, _I, _I64, _VP, _VP]),
    "cd_standard_load": (_I, [_VP, C.POINTER(_StandardWeights)]),
    "cd_standard_forward": (_I, [_VP, _VP, _I, _I64, _VP, _VP]),
    "cd_spec_augment": (_I, [_VP, _VP, _I, _I64, _VP, _I, _I, _VP]),
    "cd_predict": (_I, [_VP, _VP, _I64, C.c_float, _VP, _VP, _VP, _VP]),
    "cd_e2e_forward": (_I, [_VP, _VP, _I64, _I64, _VP, _VP, _I, _F, _I, _VP]),
    "cd_e2e_forward_i16": (_I, [_VP, _VP, _I64, _I64, _VP, _VP, _I, _F, _I, _VP]),
    "cd_e2e_forward_stream": (_I, [_VP, _VP, _I64, _I64, _I64, _VP, _I, _F, _I, _VP]),
    "cd_e2e_forward_host": (_I, [_VP, _VP, _I64, _VP, _I, _F, _I]),
    "cd_e2e_forward_host_i16": (_I, [_VP, _VP, _I64, _VP, _I, _F, _I]),
    "cd_selftest_umma": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _VP]),
    "cd_launch_count": (_I64, [_VP]),
}

_lib = None
_lock = threading.Lock()


def load_library() -> C.CDLL:
    """dlopen the in-tree library and set the prototypes.  No GPU needed for this step."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                    "cough_detector_b200 has no CPU or PyTorch fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


class CoughB200Error(RuntimeError):
    pass


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Handle:
    """One cd_handle on one CUDA device."""

    def __init__(self, device: int):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise CoughB200Error("no CUDA device: cough_detector_b200 runs on B200 (sm_100a) only, there is no CPU path")
        self.device = int(device)
        h = _VP()
        rc = self.lib.cd_create(self.device, C.byref(h))
        if rc != 0:
            raise CoughB200Error(f"cd_create({device}) failed ({rc}): {self.lib.cd_last_error(None).decode()}")
        self._h = h
        self._weights_key = None
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cd_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise CoughB200Error(f"{what} failed ({rc}): {self.lib.cd_last_error(self._h).decode()}")

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev_f32(self, t: torch.Tensor, what: str) -> torch.Tensor:
        if not (t.is_cuda and t.device.index == self.device):
            raise CoughB200Error(f"{what} must live on cuda:{self.device}")
        return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()

    @property
    def launches(self) -> int:
        return int(self.lib.cd_launch_count(self._h))

    # ---- features ---------------------------------------------------------------------------
    # Every call takes the feature options explicitly (``opts``: FeatOptions); the handle keeps none.
    def feat_forward(self, audio: torch.Tensor, normalize: bool = True, opts: FeatOptions = TRAINING_CONFIG) -> torch.Tensor:
        """(B, 16000) cuda fp32 -> (B, 1, 90 | 103, 101)."""
        audio = self._dev_f32(audio, "audio")
        if audio.dim() != 2 or audio.shape[1] != SEGMENT:
            raise CoughB200Error(f"audio must be (B, {SEGMENT}), got {tuple(audio.shape)}")
        B = audio.shape[0]
        out = torch.empty(B, 1, opts.n_features, N_FRAMES, device=audio.device, dtype=torch.float32)
        self._check(self.lib.cd_feat_forward(self._h, _ptr(audio), SEGMENT, B, _ptr(out), opts.flags(normalize),
                                             opts.pre_emphasis_coef, self._stream()), "cd_feat_forward")
        return out

    def feat_forward_i16(self, audio: torch.Tensor, normalize: bool = True, opts: FeatOptions = TRAINING_CONFIG) -> torch.Tensor:
        """(B, 16000) cuda int16 PCM -> (B, 1, 90 | 103, 101); equals feat_forward(audio / 32768) bit for bit."""
        if not (audio.is_cuda and audio.dtype == torch.int16 and audio.dim() == 2 and audio.shape[1] == SEGMENT):
            raise CoughB200Error("feat_forward_i16 wants a cuda int16 (B, 16000) tensor")
        audio = audio.contiguous()
        B = audio.shape[0]
        out = torch.empty(B, 1, opts.n_features, N_FRAMES, device=audio.device, dtype=torch.float32)
        self._check(self.lib.cd_feat_forward_i16(self._h, _ptr(audio), SEGMENT, B, _ptr(out), opts.flags(normalize),
                                                 opts.pre_emphasis_coef, self._stream()), "cd_feat_forward_i16")
        return out

    def feat_forward_var(self, flat: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor,
                         opts: FeatOptions = TRAINING_CONFIG) -> torch.Tensor:
        flat = self._dev_f32(flat, "audio")
        offsets = offsets.to(device=flat.device, dtype=torch.int64).contiguous()
        lengths = lengths.to(device=flat.device, dtype=torch.int32).contiguous()
        B = offsets.numel()
        out = torch.empty(B, 1, opts.n_features, N_FRAMES, device=flat.device, dtype=torch.float32)
        self._check(self.lib.cd_feat_forward_var(self._h, _ptr(flat), _ptr(offsets), _ptr(lengths), B, _ptr(out),
                                                 opts.flags(), opts.pre_emphasis_coef, self._stream()), "cd_feat_forward_var")
        return out

    def stream_windows(self, n: int, hop: int) -> int:
        return int(self.lib.cd_stream_windows(n, hop))

    def feat_forward_stream(self, audio: torch.Tensor, hop: int, opts: FeatOptions = TRAINING_CONFIG) -> torch.Tensor:
        audio = self._dev_f32(audio.reshape(-1), "audio")
        n = audio.numel()
        nw = self.stream_windows(n, hop)
        out = torch.empty(nw, 1, opts.n_features, N_FRAMES, device=audio.device, dtype=torch.float32)
        self._check(self.lib.cd_feat_forward_stream(self._h, _ptr(audio), n, hop, nw, _ptr(out), opts.flags(),
                                                    opts.pre_emphasis_coef, self._stream()), "cd_feat_forward_stream")
        return out

    # ---- resampling -------------------------------------------------------------------------
    def resample(self, wave: torch.Tensor, orig_sr: int, new_sr: int = 16000) -> torch.Tensor:
        """(rows, n) cuda fp32 -> (rows, ceil(new_sr * n / orig_sr)), torchaudio T.Resample defaults."""
        wave = self._dev_f32(wave, "waveform")
        if wave.dim() == 1:
            wave = wave.unsqueeze(0)
        rows, n = wave.shape
        n_out = int(self.lib.cd_resample_length(n, int(orig_sr), int(new_sr)))
        out = torch.empty(rows, n_out, device=wave.device, dtype=torch.float32)
        self._check(self.lib.cd_resample(self._h, _ptr(wave), rows, n, int(orig_sr), int(new_sr), _ptr(out), self._stream()),
                    "cd_resample")
        return out

    def condition_var(self, packed: torch.Tensor, in_offsets, channels, n_in, orig_sr: int, new_sr: int = 16000):
        """Batched resample + to_mono of clips packed in one device buffer (cd_condition_var).  packed: 1-D cuda fp32 or
        int16; in_offsets / channels / n_in: per-clip python ints.  Returns (mono fp32 buffer, out_offsets, out_lengths)
        -- the arguments of feat_forward_var."""
        if not (packed.is_cuda and packed.device.index == self.device and packed.dtype in (torch.float32, torch.int16)):
            raise CoughB200Error(f"packed audio must be a float32 or int16 tensor on cuda:{self.device}")
        packed = packed.contiguous()
        B = len(n_in)
        n_out = [int(self.lib.cd_resample_length(int(n), int(orig_sr), int(new_sr))) for n in n_in]
        out_off, acc = [], 0
        for n in n_out:
            out_off.append(acc)
            acc += (n + 3) & ~3
        desc = torch.tensor([list(in_offsets), out_off, list(channels), list(n_in)], dtype=torch.int64)
        d64 = desc[:2].contiguous().to(packed.device, non_blocking=True)
        d32 = desc[2:].to(torch.int32).contiguous().to(packed.device, non_blocking=True)
        out = torch.zeros(acc + 4, device=packed.device, dtype=torch.float32)
        self._check(self.lib.cd_condition_var(self._h, _ptr(packed), 1 if packed.dtype == torch.int16 else 0, _ptr(d64[0]),
                                              _ptr(d32[0]), _ptr(d32[1]), _ptr(d64[1]), B, max(n_in) if B else 0,
                                              int(orig_sr), int(new_sr), _ptr(out), self._stream()), "cd_condition_var")
        return out, d64[1], torch.tensor(n_out, dtype=torch.int32)

    # ---- model ------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], key=None):
        """Upload a reference-format CoughDetectorResidual state_dict (BN folded by the library)."""
        keep = {}

        def f(name):
            t = sd[name].detach().to("cpu", torch.float32).contiguous()
            keep[name] = t
            return C.cast(t.data_ptr(), _f32p)

        def cb(conv, bn):
            return _ConvBn(f(conv + ".weight"), f(conv + ".bias"), f(bn + ".weight"), f(bn + ".bias"),
                           f(bn + ".running_mean"), f(bn + ".running_var"))

        w = _ResidualWeights(
            cb("conv1.0", "conv1.1"),
            cb("res_blocks.0.conv1", "res_blocks.0.bn1"), cb("res_blocks.0.conv2", "res_blocks.0.bn2"),
            cb("res_blocks.0.skip.0", "res_blocks.0.skip.1"),
            cb("res_blocks.1.conv1", "res_blocks.1.bn1"), cb("res_blocks.1.conv2", "res_blocks.1.bn2"),
            cb("res_blocks.1.skip.0", "res_blocks.1.skip.1"),
            f("fc.2.weight"), f("fc.2.bias"))
        expect = {"conv1.0.weight": (32, 1, 7, 7), "res_blocks.0.conv1.weight": (64, 32, 3, 3),
                  "res_blocks.0.conv2.weight": (64, 64, 3, 3), "res_blocks.0.skip.0.weight": (64, 32, 1, 1),
                  "res_blocks.1.conv1.weight": (128, 64, 3, 3), "res_blocks.1.conv2.weight": (128, 128, 3, 3),
                  "res_blocks.1.skip.0.weight": (128, 64, 1, 1), "fc.2.weight": (2, 128)}
        for k, shp in expect.items():
            if tuple(keep[k].shape) != shp:
                raise CoughB200Error(f"{k}: expected shape {shp}, got {tuple(keep[k].shape)}")
        self._check(self.lib.cd_cnn_load(self._h, C.byref(w)), "cd_cnn_load")
        self._weights_key = key

    def load_plain_state_dict(self, kind: str, sd: Dict[str, torch.Tensor], key=None):
        """Upload a reference-format CoughDetectorSmall ('small') or CoughDetector ('standard') state_dict."""
        keep = []

        def f(name):
            t = sd[name].detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            return C.cast(t.data_ptr(), _f32p)

        def conv(p):
            return _Conv(f(p + ".weight"), f(p + ".bias"))

        def bn(p):
            return _Bn(f(p + ".weight"), f(p + ".bias"), f(p + ".running_mean"), f(p + ".running_var"))

        if kind == "small":
            blocks = (_DwPwBlock * 3)(*[_DwPwBlock(conv(f"features.{a}"), conv(f"features.{a + 1}"), bn(f"features.{a + 2}")) for a in (4, 9, 14)])
            w = _SmallWeights(conv("features.0"), bn("features.1"), blocks, f("classifier.1.weight"), f("classifier.1.bias"),
                              f("classifier.4.weight"), f("classifier.4.bias"))
            self._check(self.lib.cd_small_load(self._h, C.byref(w)), "cd_small_load")
        elif kind == "standard":
            w = _StandardWeights((_Conv * 4)(*[conv(f"conv_layers.{i}.conv") for i in range(4)]),
                                 (_Bn * 4)(*[bn(f"conv_layers.{i}.bn") for i in range(4)]),
                                 f("fc.0.weight"), f("fc.0.bias"), f("fc.3.weight"), f("fc.3.bias"))
            self._check(self.lib.cd_standard_load(self._h, C.byref(w)), "cd_standard_load")
        else:
            raise CoughB200Error(f"unknown plain model kind {kind!r}")
        self._plain_keys = dict(getattr(self, "_plain_keys", {}), **{kind: key})

    def plain_forward(self, kind: str, feat: torch.Tensor) -> torch.Tensor:
        feat = self._dev_f32(feat, "features")
        if feat.dim() == 3:
            feat = feat.unsqueeze(0)
        if feat.dim() != 4 or feat.shape[1] != 1 or feat.shape[3] != N_FRAMES:
            raise CoughB200Error(f"features must be (B, 1, rows, {N_FRAMES}), got {tuple(feat.shape)}")
        B, rows = feat.shape[0], int(feat.shape[2])
        logits = torch.empty(B, 2, device=feat.device, dtype=torch.float32)
        fn = self.lib.cd_small_forward if kind == "small" else self.lib.cd_standard_forward
        self._check(fn(self._h, _ptr(feat), rows, B, _ptr(logits), self._stream()), f"cd_{kind}_forward")
        return logits

    def cnn_forward(self, feat: torch.Tensor, precision: int = PRECISION_BF16) -> torch.Tensor:
        feat = self._dev_f32(feat, "features")
        if feat.dim() == 3:
            feat = feat.unsqueeze(0)
        if feat.dim() != 4 or feat.shape[1] != 1 or feat.shape[2] not in (N_FEATURES, N_FEATURES_DD) or feat.shape[3] != N_FRAMES:
            raise CoughB200Error(f"features must be (B, 1, {N_FEATURES} | {N_FEATURES_DD}, {N_FRAMES}), got {tuple(feat.shape)}")
        B, rows = feat.shape[0], int(feat.shape[2])
        logits = torch.empty(B, 2, device=feat.device, dtype=torch.float32)
        self._check(self.lib.cd_cnn_forward_rows(self._h, _ptr(feat), rows, B, _ptr(logits), precision, self._stream()),
                    "cd_cnn_forward_rows")
        return logits

    def predict(self, logits: torch.Tensor, threshold: float = 0.5):
        logits = self._dev_f32(logits, "logits")
        B = logits.shape[0]
        probs = torch.empty(B, 2, device=logits.device, dtype=torch.float32)
        preds = torch.empty(B, device=logits.device, dtype=torch.int64)
        dec = torch.empty(B, device=logits.device, dtype=torch.uint8)
        self._check(self.lib.cd_predict(self._h, _ptr(logits), B, threshold, _ptr(probs), _ptr(preds), _ptr(dec),
                                        self._stream()), "cd_predict")
        return preds, probs, dec.bool()

    def spec_augment_(self, feat: torch.Tensor, masks: torch.Tensor, n_freq_masks: int, n_time_masks: int) -> torch.Tensor:
        """In place: zero the row / column intervals of ``masks`` (B, 2 (n_freq_masks + n_time_masks)) int32 in
        ``feat`` (B, 1, rows, 101) cuda fp32 (cd_spec_augment)."""
        if not (feat.is_cuda and feat.device.index == self.device and feat.dtype == torch.float32 and feat.is_contiguous()):
            raise CoughB200Error(f"features must be a contiguous fp32 tensor on cuda:{self.device}")
        if feat.dim() != 4 or feat.shape[1] != 1 or feat.shape[3] != N_FRAMES:
            raise CoughB200Error(f"features must be (B, 1, rows, {N_FRAMES}), got {tuple(feat.shape)}")
        B = feat.shape[0]
        masks = masks.to(device=feat.device, dtype=torch.int32).contiguous()
        if tuple(masks.shape) != (B, 2 * (n_freq_masks + n_time_masks)):
            raise CoughB200Error(f"masks must be ({B}, {2 * (n_freq_masks + n_time_masks)}), got {tuple(masks.shape)}")
        self._check(self.lib.cd_spec_augment(self._h, _ptr(feat), int(feat.shape[2]), B, _ptr(masks), n_freq_masks,
                                             n_time_masks, self._stream()), "cd_spec_augment")
        return feat

    def selftest_umma(self, a: torch.Tensor, b: torch.Tensor, shift: int = 0, mode: int = 0) -> torch.Tensor:
        """a[shift:shift+128] (K) x b (N, K)^T on one tcgen05 tile; CPU fp32 tensors in, CPU fp32 (128, N) out."""
        a = a.to("cpu", torch.float32).contiguous()
        b = b.to("cpu", torch.float32).contiguous()
        rows = 256 if mode == 3 else 128                    # mode 3: one cta_group::2 MMA stream over a CTA pair
        assert a.shape[0] == rows + shift
        d = torch.empty(rows, b.shape[0], dtype=torch.float32)
        self._check(self.lib.cd_selftest_umma(self._h, b.shape[0], a.shape[1], shift, mode, _ptr(a), _ptr(b), _ptr(d)),
                    "cd_selftest_umma")
        return d

    # ---- end to end -------------------------------------------------------------------------
    def e2e_forward(self, audio: torch.Tensor, precision: int = PRECISION_BF16, return_features: bool = False,
                    opts: FeatOptions = TRAINING_CONFIG, normalize: bool = True):
        i16 = audio.dtype == torch.int16                    # int16 PCM stays int16: cd_e2e_forward_i16
        if i16:
            if not (audio.is_cuda and audio.device.index == self.device):
                raise CoughB200Error(f"audio must live on cuda:{self.device}")
            audio = audio.contiguous()
        else:
            audio = self._dev_f32(audio, "audio")
        if audio.dim() != 2 or audio.shape[1] != SEGMENT:
            raise CoughB200Error(f"audio must be (B, {SEGMENT}), got {tuple(audio.shape)}")
        B = audio.shape[0]
        logits = torch.empty(B, 2, device=audio.device, dtype=torch.float32)
        feat = torch.empty(B, 1, opts.n_features, N_FRAMES, device=audio.device, dtype=torch.float32) if return_features else None
        fn = self.lib.cd_e2e_forward_i16 if i16 else self.lib.cd_e2e_forward
        self._check(fn(self._h, _ptr(audio), SEGMENT, B, _ptr(logits), _ptr(feat), opts.flags(normalize),
                       opts.pre_emphasis_coef, precision, self._stream()), "cd_e2e_forward")
        return (logits, feat) if return_features else logits

    def e2e_forward_stream(self, audio: torch.Tensor, hop: int, precision: int = PRECISION_BF16,
                           opts: FeatOptions = TRAINING_CONFIG) -> torch.Tensor:
        audio = self._dev_f32(audio.reshape(-1), "audio")
        n = audio.numel()
        nw = self.stream_windows(n, hop)
        logits = torch.empty(nw, 2, device=audio.device, dtype=torch.float32)
        self._check(self.lib.cd_e2e_forward_stream(self._h, _ptr(audio), n, hop, nw, _ptr(logits), opts.flags(),
                                                   opts.pre_emphasis_coef, precision, self._stream()), "cd_e2e_forward_stream")
        return logits

    def e2e_forward_host(self, audio: torch.Tensor, precision: int = PRECISION_BF16, out: Optional[torch.Tensor] = None,
                         opts: FeatOptions = TRAINING_CONFIG, normalize: bool = True) -> torch.Tensor:
        """Host (B, 16000) fp32 or int16 tensor (ideally pinned) -> host (B, 2) logits; copies are inside the call."""
        if audio.is_cuda or audio.dtype not in (torch.float32, torch.int16) or audio.dim() != 2 or audio.shape[1] != SEGMENT:
            raise CoughB200Error("e2e_forward_host wants a CPU float32 or int16 (B, 16000) tensor")
        audio = audio.contiguous()
        B = audio.shape[0]
        if out is None:
            out = torch.empty(B, 2, dtype=torch.float32).pin_memory()
        fn = self.lib.cd_e2e_forward_host_i16 if audio.dtype == torch.int16 else self.lib.cd_e2e_forward_host
        self._check(fn(self._h, _ptr(audio), B, _ptr(out), opts.flags(normalize), opts.pre_emphasis_coef, precision),
                    "cd_e2e_forward_host")
        return out


_handles: Dict[int, Handle] = {}
_handles_lock = threading.Lock()              # not _lock: Handle() takes that one in load_library()


def get_handle(device=None) -> Handle:
    """Process-wide handle for a CUDA device (default: torch's current device)."""
    if device is None:
        if not torch.cuda.is_available():
            load_library()
            raise CoughB200Error("no CUDA device: cough_detector_b200 runs on B200 (sm_100a) only, there is no CPU path")
        device = torch.cuda.current_device()
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    device = int(device)
    with _handles_lock:                       # check-and-create under one lock: never two shared handles per device
        h = _handles.get(device)
        if h is None:
            h = _handles[device] = Handle(device)
    return h


def get_tables():
    """(hann[400], fb[257,64], dct[64,13]) as the library holds them; needs no GPU."""
    lib = load_library()
    hann = torch.empty(400)
    fb = torch.empty(257, 64)
    dct = torch.empty(64, 13)
    lib.cd_get_tables(hann.data_ptr(), fb.data_ptr(), dct.data_ptr())
    return hann, fb, dct


def host_alloc(n_elems: int, dtype=torch.float32, write_combined: bool = False) -> torch.Tensor:
    """A page-locked (optionally write-combined) host tensor from cd_host_alloc, freed when the tensor dies."""
    import numpy as np
    lib = load_library()
    itemsize = torch.empty(0, dtype=dtype).element_size()
    p = _VP()
    rc = lib.cd_host_alloc(C.byref(p), n_elems * itemsize, 1 if write_combined else 0)
    if rc != 0:
        raise CoughB200Error(f"cd_host_alloc failed ({rc}): {lib.cd_last_error(None).decode()}")
    buf = (C.c_char * (n_elems * itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8)
    t = torch.from_numpy(arr).view(dtype)

    class _Owner:
        def __del__(self, lib=lib, p=p):
            lib.cd_host_free(p)
    t._cd_owner = _Owner()
    return t
