"""SpecAugment and the cached-feature dataset (training-side use of the feature path)."""
import os
import random
import wave

import numpy as np
import pytest
import torch

from cough_detector_b200 import AudioPreprocessor, synth
from cough_detector_b200.augmentation import SpecAugment
from cough_detector_b200.dataset import CoughDataset, decode_audio
from oracle import cough_oracle as co
from tests import golden_io as gio

FLAGS = dict(use_pcen=False, use_pre_emphasis=False, use_delta_delta=False, use_spectral_contrast=False)


def _inputs(call, rows):
    return 0.5 + torch.rand(1, rows, 101, generator=torch.Generator().manual_seed(1000 + call))


def test_oracle_spec_augment_matches_reference():
    cases, batch = gio.load_spec_augment()
    for ci, (rows, kw, zero) in enumerate(cases):
        random.seed(40 + ci)
        torch.manual_seed(40 + ci)
        full = dict(freq_mask_param=10, time_mask_param=20, n_freq_masks=2, n_time_masks=2, p=0.5)
        full.update(kw)
        touched = 0
        for call in range(8):
            iv = co.spec_augment_intervals(rows, 101, **full)
            y = co.spec_augment_apply(_inputs(call, rows), iv, full["n_freq_masks"])
            assert np.array_equal((y[0] == 0).numpy(), zero[call]), (ci, call)
            touched += iv is not None
        assert 0 < touched <= 8
    random.seed(77)
    torch.manual_seed(77)
    iv = co.spec_augment_intervals(90, 101, p=1.0)
    xb = 0.5 + torch.rand(3, 1, 90, 101, generator=torch.Generator().manual_seed(5))
    assert np.array_equal((co.spec_augment_apply(xb, iv) == 0).numpy(), batch)


def test_host_sampler_draws_like_the_reference():
    """SpecAugment.sample / sample_batch consume random.random() and torch.rand in the reference's order."""
    cases, _ = gio.load_spec_augment()
    for ci, (rows, kw, zero) in enumerate(cases):
        sa = SpecAugment(**kw)
        random.seed(40 + ci)
        torch.manual_seed(40 + ci)
        masks = sa.sample_batch(8, rows, 101)
        assert masks.shape == (8, 2 * (sa.n_freq_masks + sa.n_time_masks)) and masks.dtype == torch.int32
        for call in range(8):
            iv = [(int(masks[call, 2 * k]), int(masks[call, 2 * k + 1])) for k in range(masks.shape[1] // 2)]
            y = co.spec_augment_apply(_inputs(call, rows), iv, sa.n_freq_masks)
            assert np.array_equal((y[0] == 0).numpy(), zero[call]), (ci, call)


def test_host_sampler_edge_cases():
    """p = 0 never touches the generators' torch side; a zero mask parameter draws nothing (the reference returns before
    torch.rand); intervals stay inside the axis; more than 8 masks per axis is refused (the kernel's limit)."""
    sa = SpecAugment(p=0.0)
    torch.manual_seed(5)
    before = torch.rand(1)
    torch.manual_seed(5)
    assert sa.sample(90, 101) is None and torch.equal(torch.rand(1), before)
    x = torch.rand(1, 90, 101)
    assert sa(x) is x                                              # untouched input comes back as is, like the reference
    sa0 = SpecAugment(freq_mask_param=0, time_mask_param=0, n_freq_masks=2, n_time_masks=2, p=1.0)
    torch.manual_seed(5)
    assert sa0.sample(90, 101) == [0, 0] * 4 and torch.equal(torch.rand(1), before)
    random.seed(1)
    torch.manual_seed(1)
    m = SpecAugment(freq_mask_param=10, time_mask_param=20, n_freq_masks=3, n_time_masks=1, p=1.0).sample_batch(200, 90, 101)
    f, t = m[:, :6].reshape(-1, 2), m[:, 6:].reshape(-1, 2)
    assert (f[:, 0] >= 0).all() and (f[:, 1] <= 90).all() and ((f[:, 1] - f[:, 0]) < 10).all() and ((f[:, 1] - f[:, 0]) >= 0).all()
    assert (t[:, 0] >= 0).all() and (t[:, 1] <= 101).all() and ((t[:, 1] - t[:, 0]) < 20).all()
    with pytest.raises(NotImplementedError):
        SpecAugment(n_freq_masks=9)
    with pytest.raises(ValueError):
        SpecAugment(p=1.0)(torch.rand(101))                        # torchaudio: "at least two dimensions"


@pytest.mark.gpu
def test_spec_augment_kernel_matches_reference():
    cases, batch = gio.load_spec_augment()
    for ci, (rows, kw, zero) in enumerate(cases):
        sa = SpecAugment(**kw)
        random.seed(40 + ci)
        torch.manual_seed(40 + ci)
        for call in range(8):                                 # per-sample __call__, CPU tensor in -> CPU tensor out
            x = _inputs(call, rows)
            y = sa(x)
            assert y.device == x.device and y.shape == x.shape
            assert np.array_equal((y[0] == 0).numpy(), zero[call]), (ci, call)
            assert torch.equal(y[y != 0], x[y != 0]) and (x > 0).all()      # untouched cells are bit-identical, input not written
        random.seed(40 + ci)
        torch.manual_seed(40 + ci)
        xb = torch.stack([_inputs(call, rows) for call in range(8)]).cuda()  # the batched form: one launch, per-clip masks
        yb = sa.augment_batch_(xb.clone())
        assert np.array_equal((yb[:, 0] == 0).cpu().numpy(), zero)
    random.seed(77)
    torch.manual_seed(77)
    xb = 0.5 + torch.rand(3, 1, 90, 101, generator=torch.Generator().manual_seed(5))
    assert np.array_equal((SpecAugment(p=1.0)(xb.cuda()) == 0).cpu().numpy(), batch)   # a 4-D batch shares one set of masks


def _write_wav(path, x, sr, channels=1):
    pcm = (x.clamp(-1, 1) * 32767.0).round().to(torch.int16).reshape(channels, -1)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(pcm.numpy().T).astype("<i2").tobytes())
    return pcm


def _make_tree(root):
    """A small data_dir/{non_cough,cough} tree of 16-bit wav files: 16 kHz mono of several lengths, one 44.1 kHz stereo, one
    8 kHz.  Returns {path: (int16 (C, n), sr)}."""
    files = {}
    clips = synth.synth_clips(10, seed=21)
    for i, x in enumerate(clips):
        d = root / ("cough" if i % 3 == 0 else "non_cough")
        d.mkdir(parents=True, exist_ok=True)
        x = x / max(float(x.abs().max()), 1e-9) * 0.8
        files[str(d / f"c{i:02d}.wav")] = (_write_wav(d / f"c{i:02d}.wav", x, 16000), 16000)
    g = torch.Generator().manual_seed(3)
    st = 0.3 * torch.randn(2, 30000, generator=g)
    files[str(root / "cough" / "s44.wav")] = (_write_wav(root / "cough" / "s44.wav", st, 44100, channels=2), 44100)
    lo = 0.3 * torch.randn(1, 9000, generator=g)
    files[str(root / "non_cough" / "l8.wav")] = (_write_wav(root / "non_cough" / "l8.wav", lo, 8000), 8000)
    (root / "non_cough" / "notes.txt").write_text("not audio")
    return files


def test_dataset_bookkeeping_matches_reference_layout(tmp_path):
    files = _make_tree(tmp_path)
    ds = CoughDataset(str(tmp_path), AudioPreprocessor(device="cpu", **FLAGS), is_training=False)
    assert len(ds) == len(files) == 12
    assert ds.classes == ["non_cough", "cough"] and ds.class_counts == {0: 7, 1: 5}
    assert all(p in files for p, _ in ds.samples)
    w = ds.sample_weights                                         # total / (n_classes * class_count) (src/dataset.py:109-116)
    assert torch.allclose(w[torch.tensor([l for _, l in ds.samples]) == 1], torch.tensor(12 / (2 * 5.0)).to(w.dtype))
    x, sr = decode_audio(str(tmp_path / "cough" / "s44.wav"))
    assert x.dtype == torch.int16 and x.shape == (2, 30000) and sr == 44100
    assert torch.equal(x, files[str(tmp_path / "cough" / "s44.wav")][0])
    with pytest.raises(RuntimeError, match="Failed to load audio file"):
        decode_audio(str(tmp_path / "non_cough" / "notes.txt"))


@pytest.mark.gpu
def test_cached_dataset_matches_per_file_pipeline(tmp_path):
    """precompute_cache (batched, int16 upload, grouped by sample rate) == the reference's per-file __getitem__ pipeline
    restated by the oracle (resample -> mono -> normalize -> pad/trim -> extract_features) within the feature tolerance."""
    files = _make_tree(tmp_path)
    pp = AudioPreprocessor(device="cuda", **FLAGS)
    ds = CoughDataset(str(tmp_path), pp, spec_augmentor=None, is_training=False, cache_spectrograms=True)
    feats = ds.precompute_cache(batch_size=5)                     # several chunks, mixed rates inside a chunk
    assert feats.shape == (12, 1, 90, 101) and feats.is_cuda
    ex = co.Explicit32()
    for i, (path, label) in enumerate(ds.samples):
        pcm, sr = files[path]
        w = co.to_mono(co.resample(pcm.to(torch.float32) / 32768.0, sr))
        ref = ex.process(w)
        tol = 1e-4 if sr == 16000 else 2e-4                       # off-rate clips add the resampler's 2e-6 ahead of the dB
        assert (feats[i].cpu() - ref).abs().max().item() <= tol, (path, sr)
        item, lab = ds[i]                                         # cache hit: a clone of the cached features
        assert lab == label and torch.equal(item, feats[i]) and item.data_ptr() != feats[i].data_ptr()
    # without a cache the per-sample path computes the same thing one file at a time
    ds2 = CoughDataset(str(tmp_path), pp, is_training=False)
    for i in (0, 5, 10, 11):
        assert torch.equal(ds2[i][0].to(feats.device), feats[i]), i


@pytest.mark.gpu
def test_feature_batches_feed_the_model(tmp_path):
    from cough_detector_b200 import create_model
    _make_tree(tmp_path)
    pp = AudioPreprocessor(device="cuda", **FLAGS)
    ds = CoughDataset(str(tmp_path), pp, spec_augmentor=SpecAugment(p=1.0), is_training=True, cache_spectrograms=True)
    cached = ds.precompute_cache().clone()
    model = create_model("residual").eval()
    random.seed(9)
    torch.manual_seed(9)
    seen, masked = 0, 0
    for x, y in ds.feature_batches(5, shuffle=True, generator=torch.Generator().manual_seed(1)):
        assert x.is_cuda and x.shape[1:] == (1, 90, 101) and y.shape == (x.shape[0],)
        logits = model(x)
        assert logits.shape == (x.shape[0], 2) and torch.isfinite(logits).all()
        seen += x.shape[0]
        masked += int((x[:, 0, :, :].abs().sum(dim=2) == 0).any(dim=1).sum())      # a fully zeroed row = a frequency mask
    assert seen == 12 and masked >= 6
    assert torch.equal(ds.features, cached)                       # augmentation never writes the cache
