"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run once in the authoring container (``python -m oracle.make_golden``); the GPU box has no
/root/reference, so only the committed vectors travel.  Versions used: see ``meta`` in each
file.  Inputs are stored as fp16 mantissas times a per-clip fp32 scale so the fixture stays
small while the fp32 input fed to the reference and to the CUDA path is bit-identical
(``load_clips`` in tests/golden_io.py re-creates it with the same fp32 multiply).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cough_detector_b200 import synth  # noqa: E402  (pure-python helper, no CUDA)

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
FLAGS = dict(use_pcen=False, use_pre_emphasis=False, use_delta_delta=False, use_spectral_contrast=False)


def quantize_clip(x: torch.Tensor):
    """x -> (fp16 mantissa array, fp32 scale) with x_q = float32(m) * scale exactly reproducible."""
    peak = float(x.abs().max()) if x.numel() else 0.0
    scale = np.float32(peak if peak > 0 else 1.0)
    m = (x.numpy().astype(np.float32) / scale).astype(np.float16)
    return m, scale


def dequantize_clip(m: np.ndarray, scale) -> torch.Tensor:
    return torch.from_numpy(m.astype(np.float32) * np.float32(scale))


FLAG_COMBOS = {"pre": dict(use_pcen=False, use_pre_emphasis=True), "pcen": dict(use_pcen=True, use_pre_emphasis=False),
               "pre_pcen": dict(use_pcen=True, use_pre_emphasis=True)}


def flag_goldens():
    """tests/golden/features_flags.npz: the reference's process() with use_pre_emphasis / use_pcen (90 rows each) on the
    clips of features.npz, and its RealtimePreprocessor with both flags on the stream of stream.npz (hop 0.1 s)."""
    sys.path.insert(0, REF)
    import torchaudio
    from src.preprocessing import AudioPreprocessor, RealtimePreprocessor  # the reference

    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLD, "features.npz"))
    clips, o = [], 0
    for n, sc in zip(z["lens"], z["scales"]):
        clips.append(dequantize_clip(z["mant"][o:o + n], sc))
        o += int(n)
    out = {"meta": np.array([torch.__version__, torchaudio.__version__])}
    for name, fl in FLAG_COMBOS.items():
        pp = AudioPreprocessor(use_delta_delta=False, use_spectral_contrast=False, **fl)
        assert pp.get_num_features() == 90
        out[f"features_{name}"] = np.stack([pp.process(x.reshape(1, -1), 16000)[0].numpy() for x in clips]).astype(np.float32)
    # use_delta_delta: rows 0..89 are the 90-row features again (checked here), so only rows 90..102 are stored
    for name, fl in (("dd", dict(use_pcen=False, use_pre_emphasis=False)), ("pre_pcen_dd", FLAG_COMBOS["pre_pcen"])):
        pp = AudioPreprocessor(use_delta_delta=True, use_spectral_contrast=False, **fl)
        assert pp.get_num_features() == 103
        full = np.stack([pp.process(x.reshape(1, -1), 16000)[0].numpy() for x in clips]).astype(np.float32)
        base = z["features"] if name == "dd" else out["features_pre_pcen"]
        assert np.array_equal(full[:, :90], base)
        out[f"ddrows_{name}"] = full[:, 90:]
        if name == "dd":
            # the reference network on the 103-row input (it is fully convolutional up to its global pool), both weight sets
            from src.model import create_model
            zm = np.load(os.path.join(GOLD, "model.npz"))
            sd0 = {k[4:]: torch.from_numpy(zm[k]) for k in zm.files if k.startswith("sd0/")}
            sd1 = dict(sd0)
            sd1.update({k[4:]: torch.from_numpy(zm[k]) for k in zm.files if k.startswith("sd1/")})
            model = create_model("residual", n_mels=103, num_classes=2, in_channels=1).eval()
            with torch.no_grad():
                for wname, sd in (("default", sd0), ("random", sd1)):
                    model.load_state_dict(sd)
                    out[f"logits_dd_{wname}"] = model(torch.from_numpy(full).unsqueeze(1)).numpy()
    zs = np.load(os.path.join(GOLD, "stream.npz"))
    stream = dequantize_clip(zs["mant"], zs["scale"])
    rp = RealtimePreprocessor(window_duration=1.0, hop_duration=0.1, use_delta_delta=False, use_spectral_contrast=False,
                              **FLAG_COMBOS["pre_pcen"])
    got = []
    for c in range(0, stream.numel(), 1600):
        got += rp.add_audio(stream[c:c + 1600])
    out["stream_pre_pcen_hop100"] = torch.stack(got).numpy()
    np.savez_compressed(os.path.join(GOLD, "features_flags.npz"), **out)
    print("features_flags.npz", os.path.getsize(os.path.join(GOLD, "features_flags.npz")))


def other_model_goldens():
    """tests/golden/models_other.npz: the reference's 'small' (its default model_type) and 'standard' networks with
    BatchNorm-randomised weights (seeds 11 / 12 on top of torch.manual_seed(0) init), their logits on the golden
    features (90 rows) and on the 103-row delta-delta features."""
    sys.path.insert(0, REF)
    from src.model import create_model
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLD, "features.npz"))
    zf = np.load(os.path.join(GOLD, "features_flags.npz"))
    x90 = torch.from_numpy(z["features"]).unsqueeze(1)
    x103 = torch.cat([torch.from_numpy(z["features"]), torch.from_numpy(zf["ddrows_dd"])], dim=1).unsqueeze(1)
    out = {"meta": np.array([torch.__version__])}
    for kind, seed in (("small", 11), ("standard", 12)):
        torch.manual_seed(0)
        model = create_model(kind, n_mels=90, num_classes=2, in_channels=1).eval()
        sd = synth.randomized_state_dict(model, seed=seed, fc_scale=1.0)
        last = "classifier.4" if kind == "small" else "fc.3"
        if kind == "standard":      # big tensors are stored as fp16: everything below uses the SAME rounded weights
            sd = {k: (v.half().float() if v.numel() > 100000 else v) for k, v in sd.items()}
        model.load_state_dict(sd)
        with torch.no_grad():       # random-init networks barely react to the input: spread P(cough) around 0.7
            sd = synth.spread_fc(sd, model(x90), weight_key=last + ".weight", bias_key=last + ".bias")
        model.load_state_dict(sd)
        with torch.no_grad():
            out[f"{kind}/logits90"] = model(x90).numpy()
            out[f"{kind}/logits103"] = model(x103).numpy()
        for k, v in sd.items():
            if not k.endswith("num_batches_tracked"):
                out[f"{kind}/sd/{k}"] = v.numpy().astype(np.float16 if v.numel() > 100000 else np.float32)
    np.savez_compressed(os.path.join(GOLD, "models_other.npz"), **out)
    print("models_other.npz", os.path.getsize(os.path.join(GOLD, "models_other.npz")))


def contrast_goldens():
    """tests/golden/contrast_nan.npz: what the UNMODIFIED reference returns with its constructor defaults (110 rows) on
    golden clips 0..3 -- the NaN pattern that makes use_spectral_contrast a documented non-target."""
    sys.path.insert(0, REF)
    from src.preprocessing import AudioPreprocessor
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLD, "features.npz"))
    pp = AudioPreprocessor()                               # constructor defaults: PCEN, pre-emphasis, delta-delta, contrast
    assert pp.get_num_features() == 110
    nan_rows, finite_rows, o = [], [], 0
    for n, sc in list(zip(z["lens"], z["scales"]))[:4]:
        x = dequantize_clip(z["mant"][o:o + n], sc)
        o += int(n)
        f = pp.process(x.reshape(1, -1), 16000)[0]
        nan_rows.append(torch.isnan(f).all(dim=1).numpy())
        finite_rows.append(torch.isfinite(f).all(dim=1).numpy())
    np.savez_compressed(os.path.join(GOLD, "contrast_nan.npz"), nan_rows=np.stack(nan_rows), finite_rows=np.stack(finite_rows),
                        meta=np.array([torch.__version__]))
    print("contrast_nan.npz", np.stack(nan_rows).sum(axis=1))


AUGMENT_CASES = [dict(rows=90, kw=dict(p=0.8)), dict(rows=103, kw=dict(p=0.8)),
                 dict(rows=90, kw=dict(freq_mask_param=15, time_mask_param=30, n_freq_masks=1, n_time_masks=3, p=1.0)),
                 dict(rows=90, kw=dict(freq_mask_param=0, time_mask_param=20, n_freq_masks=2, n_time_masks=2, p=1.0))]


def augment_goldens():
    """tests/golden/spec_augment.npz: which cells the UNMODIFIED reference SpecAugment (src/augmentation.py:271-331) zeroes,
    for 8 consecutive __call__s per case after random.seed(s) / torch.manual_seed(s) -- the inputs are strictly positive
    (0.5 + rand), so a zero cell is a masked cell.  Stored as packed bits, (case, call, rows * 101)."""
    import random
    sys.path.insert(0, REF)
    from src.augmentation import SpecAugment
    torch.set_num_threads(1)
    out = {"meta": np.array([torch.__version__])}
    for ci, case in enumerate(AUGMENT_CASES):
        random.seed(40 + ci)
        torch.manual_seed(40 + ci)
        sa = SpecAugment(**case["kw"])
        zs = []
        for call in range(8):
            x = 0.5 + torch.rand(1, case["rows"], 101, generator=torch.Generator().manual_seed(1000 + call))
            y = sa(x)
            zs.append(np.packbits((y == 0).numpy().reshape(-1)))
        out[f"zero_{ci}"] = np.stack(zs)
    # a 4-D batch shares one set of masks (mask_along_axis is not iid)
    random.seed(77)
    torch.manual_seed(77)
    sa = SpecAugment(p=1.0)
    xb = 0.5 + torch.rand(3, 1, 90, 101, generator=torch.Generator().manual_seed(5))
    out["zero_batch"] = np.packbits((sa(xb) == 0).numpy().reshape(-1))
    np.savez_compressed(os.path.join(GOLD, "spec_augment.npz"), **out)
    print("spec_augment.npz", os.path.getsize(os.path.join(GOLD, "spec_augment.npz")))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "augment":
        return augment_goldens()
    if len(sys.argv) > 1 and sys.argv[1] == "contrast":
        return contrast_goldens()
    if len(sys.argv) > 1 and sys.argv[1] == "flags":
        return flag_goldens()
    if len(sys.argv) > 1 and sys.argv[1] == "models":
        return other_model_goldens()
    sys.path.insert(0, REF)
    import torchaudio
    from src.preprocessing import AudioPreprocessor, RealtimePreprocessor  # the reference
    from src.model import create_model, count_parameters
    from src.inference import CoughDetectorInference

    torch.set_num_threads(1)
    meta = np.array([torch.__version__, torchaudio.__version__])
    os.makedirs(GOLD, exist_ok=True)

    # ---- features: reference process() (normalize -> pad_or_trim -> extract_features) ----
    pp = AudioPreprocessor(**FLAGS)
    assert pp.get_num_features() == 90 and pp.get_expected_time_frames() == 101
    raw = synth.synth_clips(20, seed=0) + synth.edge_clips()
    mant, scales, feats = [], [], []
    for x in raw:
        m, s = quantize_clip(x)
        xq = dequantize_clip(m, s)
        mant.append(m)
        scales.append(s)
        feats.append(pp.process(xq.reshape(1, -1), 16000)[0].numpy())
    feats = np.stack(feats).astype(np.float32)
    lens = np.array([m.size for m in mant], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "features.npz"), mant=np.concatenate(mant), lens=lens,
                        scales=np.array(scales, dtype=np.float32), features=feats, meta=meta,
                        hann=pp.mel_spectrogram.spectrogram.window.numpy(),
                        fb=pp.mel_spectrogram.mel_scale.fb.numpy(),
                        dct=pp.mfcc_transform.dct_mat.numpy())

    # ---- model: default init (seed 0) and BN-randomised weights, logits on golden features ----
    torch.manual_seed(0)
    model = create_model("residual", n_mels=90, num_classes=2, in_channels=1).eval()
    assert count_parameters(model) == 290370
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    sd1 = synth.randomized_state_dict(model, seed=1)
    x = torch.from_numpy(feats).unsqueeze(1)
    with torch.no_grad():
        model.load_state_dict(sd1)
        sd1 = synth.spread_fc(sd1, model(x))
    out = {}
    with torch.no_grad():
        for name, sd in (("default", sd0), ("random", sd1)):
            model.load_state_dict(sd)
            logits = model(x)
            preds, probs = model.predict(x)
            out[f"logits_{name}"] = logits.numpy()
            out[f"probs_{name}"] = probs.numpy()
            out[f"preds_{name}"] = preds.numpy()
    w = {f"sd0/{k}": v.numpy() for k, v in sd0.items()}
    w.update({f"sd1/{k}": v.numpy() for k, v in sd1.items()
              if not torch.equal(v, sd0[k])})
    np.savez_compressed(os.path.join(GOLD, "model.npz"), meta=meta, **w, **out)

    # ---- streaming: RealtimePreprocessor.add_audio, hop 0.25 s (reference default) and 0.1 s ----
    g = torch.Generator().manual_seed(7)
    stream = torch.cat([synth._burst(g, 16000), synth._sines(g, 12000), 0.02 * torch.randn(8000, generator=g)])
    m, s = quantize_clip(stream)
    stream = dequantize_clip(m, s)
    st = {"mant": m, "scale": np.float32(s), "meta": meta}
    for hop in (0.25, 0.1):
        rp = RealtimePreprocessor(window_duration=1.0, hop_duration=hop, **FLAGS)
        got = []
        for c in range(0, stream.numel(), 1600):          # 100 ms chunks like the mic loop
            got += rp.add_audio(stream[c:c + 1600])
        st[f"feat_hop{int(hop * 1000)}"] = torch.stack(got).numpy()
    np.savez_compressed(os.path.join(GOLD, "stream.npz"), **st)

    # ---- resampling: AudioPreprocessor.resample (T.Resample defaults) from the rates the datasets use ----
    g = torch.Generator().manual_seed(11)
    rs = {"meta": meta}
    for sr in (44100, 48000, 8000, 22050, 32000):
        n = int(0.37 * sr) + 13
        t = torch.arange(n) / sr
        xw = torch.stack([0.4 * torch.sin(2 * 3.14159265 * 431.0 * t) + 0.05 * torch.randn(n, generator=g),
                          0.3 * torch.randn(n, generator=g)])
        mq, sq = quantize_clip(xw.reshape(-1))
        xw = dequantize_clip(mq, sq).reshape(2, n)
        rs[f"in_{sr}_mant"] = mq
        rs[f"in_{sr}_scale"] = np.float32(sq)
        rs[f"out_{sr}"] = pp.resample(xw, sr).numpy()
    np.savez_compressed(os.path.join(GOLD, "resample.npz"), **rs)

    # ---- inference engine: checkpoint round trip + process_audio_chunk (debounce 0 -> deterministic) ----
    cfg = dict(model_type="residual", sample_rate=16000, n_mels=64, n_fft=512, hop_length=160, win_length=400,
               f_min=100.0, f_max=4000.0, segment_duration=1.0, n_mfcc=13, use_mfcc=True, n_contrast_bands=6,
               pre_emphasis_coef=0.97, **FLAGS)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "best_model.pt")
        torch.save({"epoch": 0, "model_state_dict": sd1, "optimizer_state_dict": {}, "metrics": {},
                    "config": cfg}, path)                  # src/train.py:183-199 format
        eng = CoughDetectorInference(path, device="cpu", confidence_threshold=0.5, smoothing_window=3,
                                     debounce_seconds=0.0, verbose=False)
        fired, confs = [], []
        for c in range(0, stream.numel(), 4000):           # 250 ms chunks: one window per call
            r = eng.process_audio_chunk(stream[c:c + 4000].numpy())
            fired.append(r is not None)
            confs.append(-1.0 if r is None else float(r[1]))
        probs = []
        for f in st["feat_hop250"]:
            probs.append(eng.predict(torch.from_numpy(f))[1])
    np.savez_compressed(os.path.join(GOLD, "engine.npz"), fired=np.array(fired), confs=np.array(confs, dtype=np.float64),
                        probs=np.array(probs, dtype=np.float64), meta=meta)
    flag_goldens()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
