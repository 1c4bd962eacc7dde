"""The oracle (oracle/) against the golden vectors produced by the unmodified reference."""
import numpy as np
import torch

from oracle import cough_oracle as co
from oracle import ref_build
from oracle import reference_port as rp
from tests import golden_io as gio


def test_constants_match_torchaudio():
    _, _, z = gio.load_features()
    assert np.array_equal(co.hann_window().numpy(), z["hann"])
    assert np.abs(co.hann_window(torch.float64).numpy() - z["hann"]).max() <= 3e-7
    assert np.array_equal(co.mel_filterbank().numpy(), z["fb"])
    assert np.array_equal(co.dct_matrix().numpy(), z["dct"])
    fb = z["fb"]
    assert (fb != 0).sum() == 244 and fb[:4].max() == 0 and fb[128:].max() == 0


def test_explicit32_matches_reference_features():
    clips, feats, _ = gio.load_features()
    ex = co.Explicit32()
    worst = 0.0
    for x, f in zip(clips, feats):
        got = ex.process(x.reshape(1, -1))[0]
        assert got.shape == (90, 101)
        worst = max(worst, (got - f).abs().max().item())
    assert worst <= 2e-5, worst


def test_fp64_matches_reference_features():
    clips, feats, _ = gio.load_features()
    worst, ill = 0.0, []
    for i, (x, f) in enumerate(zip(clips, feats)):
        d = np.abs(co.features_fp64(x.numpy()) - f.numpy()).max()
        if co.ill_conditioned_fraction(co.conditioned_window(x.numpy())) > 0:
            ill.append(i)           # DC-like clips: the reference's own fp32 noise exceeds 1e-4
            assert d <= 5e-3, (i, d)
        else:
            worst = max(worst, d)
    assert worst <= 2e-5, worst
    assert 22 in ill and len(ill) <= 3, ill


def test_port_is_bit_identical_to_reference():
    clips, feats, _ = gio.load_features()
    pp = rp.PortPreprocessor()
    for x, f in zip(clips, feats):
        assert torch.equal(pp.process(x.reshape(1, -1))[0], f)


def test_built_reference_reproduces_goldens_and_port():
    """oracle/_ref (the unmodified reference's bytecode, oracle/ref_build.py) must give the committed goldens bit for bit,
    the port must equal it on fresh seeded clips, and both models must agree from one state dict.  Skipped only where
    oracle/_ref has not been built (no /root/reference and no prebuilt files)."""
    import pytest
    from cough_detector_b200 import synth
    ref = ref_build.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    pp = ref["preprocessing"].AudioPreprocessor(use_pcen=False, use_pre_emphasis=False, use_delta_delta=False,
                                                use_spectral_contrast=False)
    clips, feats, _ = gio.load_features()
    for x, f in zip(clips[:8], feats[:8]):
        assert torch.equal(pp.process(x.reshape(1, -1), 16000)[0], f)
    port = rp.PortPreprocessor()
    fresh = synth.synth_batch(6, seed=4242, device="cpu")
    got = []
    for w in fresh:
        w = pp.pad_or_trim(pp.normalize(w.reshape(1, -1)))          # BASELINE.md §5's per-clip sequence
        got.append(pp.extract_features(w))
        assert torch.equal(got[-1], port.process(w))
    torch.manual_seed(0)
    m_ref = ref["model"].create_model("residual", n_mels=90, num_classes=2, in_channels=1).eval()
    m_port = rp.PortResidual().eval()
    m_port.load_state_dict(m_ref.state_dict())
    with torch.no_grad():
        x = torch.stack(got)
        assert torch.equal(m_ref(x), m_port(x))


def test_cnn_forward_matches_reference_logits():
    _, feats, _ = gio.load_features()
    sd0, sd1, z = gio.load_model()
    x = feats.unsqueeze(1)
    for name, sd in (("default", sd0), ("random", sd1)):
        lg = co.cnn_forward(sd, x)
        ref = torch.from_numpy(z[f"logits_{name}"])
        assert (lg - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
        preds, probs = co.predict(lg)
        assert np.array_equal(preds.numpy(), z[f"preds_{name}"])
        assert np.abs(probs.numpy() - z[f"probs_{name}"]).max() <= 1e-5
        pm = rp.PortResidual()
        pm.load_state_dict(sd)
        with torch.no_grad():
            assert torch.allclose(pm.eval()(x), ref, atol=1e-5, rtol=1e-5)


def test_random_weights_straddle_threshold():
    _, _, z = gio.load_model()
    p = z["probs_random"][:, 1]
    assert p.min() < 0.5 < p.max() and (p >= 0.7).any() and (p < 0.7).any()


def test_bf16_emulation_close_to_fp32():
    _, feats, _ = gio.load_features()
    _, sd1, z = gio.load_model()
    lg = co.cnn_forward_bf16(sd1, feats.unsqueeze(1))
    ref = torch.from_numpy(z["logits_random"])
    assert (lg - ref).abs().max().item() <= 0.05 * max(1.0, ref.abs().max().item())


def test_stream_matches_reference_realtime_preprocessor():
    stream, z = gio.load_stream()
    for hop_ms in (250, 100):
        ref = torch.from_numpy(z[f"feat_hop{hop_ms}"])
        got = co.stream_features(stream, hop_ms * 16)
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() <= 2e-5
    assert co.stream_windows(15999, 1600) == 0 and co.stream_windows(16000, 1600) == 1
    assert co.stream_windows(57_600_000, 1600) == 35991


def test_engine_decisions_match_reference():
    stream, zs = gio.load_stream()
    e = gio.load_engine()
    _, sd1, _ = gio.load_model()
    feats = torch.from_numpy(zs["feat_hop250"])
    probs = co.predict(co.cnn_forward(sd1, feats))[1][:, 1].numpy()
    assert np.abs(probs - e["probs"]).max() <= 1e-5
    # reference was driven with one 250 ms chunk per call; the first 3 calls yield no window
    n_calls = len(e["fired"])
    chunk_windows = [0] * 3 + [1] * (n_calls - 3)
    fired = co.decisions_stream(e["probs"], 0.25, threshold=0.5, smoothing_window=3, debounce_s=0.0,
                                chunk_windows=chunk_windows)
    ref_idx = [i - 3 for i in np.nonzero(e["fired"])[0]]
    assert [j for j, _ in fired] == ref_idx
    assert np.allclose([c for _, c in fired], e["confs"][e["fired"]], atol=1e-6)


def test_resample_is_bit_identical_to_reference():
    for sr, (x, ref) in gio.load_resample().items():
        got = co.resample(x, sr)
        assert got.shape == ref.shape and torch.equal(got, ref), sr
    assert co.resample(torch.ones(1, 5), 16000) .shape == (1, 5)


def test_pre_emphasis_and_pcen_oracles_match_reference_goldens():
    """features_flags.npz (reference process() with use_pre_emphasis / use_pcen): the port is bit-identical, the explicit
    fp32 restatement within 2e-6, float64 within 1e-4 on well-conditioned clips."""
    from oracle.reference_port import PortPreprocessor
    clips, _, _ = gio.load_features()
    want, stream_ref = gio.load_feature_flags()
    combos = {"pre": (False, True), "pcen": (True, False), "pre_pcen": (True, True)}
    for name, (pcen, pre) in combos.items():
        port, ex = PortPreprocessor(pcen, pre), co.Explicit32(pcen, pre)
        for i in range(0, len(clips), 3):
            x = clips[i].reshape(1, -1)
            assert torch.equal(port.process(x)[0], want[name][i]), (name, i)
            assert (ex.process(x)[0] - want[name][i]).abs().max().item() <= 2e-6, (name, i)
            if co.ill_conditioned_fraction(co.conditioned_window(clips[i].numpy())) == 0:
                d = np.abs(co.features_fp64(clips[i].numpy(), pcen, pre) - want[name][i].numpy()).max()
                assert d <= 1e-4, (name, i, d)
    stream, _ = gio.load_stream()
    got = co.stream_features(stream, 1600, co.Explicit32(True, True))
    assert got.shape == stream_ref.shape and (got - stream_ref).abs().max().item() <= 2e-6


def test_delta_delta_oracles_match_reference_goldens():
    from oracle.reference_port import PortPreprocessor
    clips, base, _ = gio.load_features()
    ddrows = gio.load_delta_delta_rows()
    for name, (pcen, pre) in {"dd": (False, False), "pre_pcen_dd": (True, True)}.items():
        port, ex = PortPreprocessor(pcen, pre, use_delta_delta=True), co.Explicit32(pcen, pre, use_delta_delta=True)
        for i in range(0, len(clips), 4):
            x = clips[i].reshape(1, -1)
            p = port.process(x)[0]
            assert p.shape == (103, 101) and torch.equal(p[90:], ddrows[name][i]), (name, i)
            if name == "dd":
                assert torch.equal(p[:90], base[i])
            assert (ex.process(x)[0][90:] - ddrows[name][i]).abs().max().item() <= 2e-6
            if co.ill_conditioned_fraction(co.conditioned_window(clips[i].numpy())) == 0:
                d = np.abs(co.features_fp64(clips[i].numpy(), pcen, pre, use_delta_delta=True)[90:] - ddrows[name][i].numpy()).max()
                assert d <= 1e-4, (name, i, d)


def test_other_model_types_oracle_matches_reference_goldens():
    """'small' (the reference's default model_type) and 'standard': the functional restatements against logits recorded
    from the reference modules, 90- and 103-row inputs."""
    from tests import golden_io as gio
    _, feats, _ = gio.load_features()
    x90 = feats.unsqueeze(1)
    x103 = torch.cat([feats, gio.load_delta_delta_rows()["dd"]], dim=1).unsqueeze(1)
    for kind, fn in (("small", co.cnn_forward_small), ("standard", co.cnn_forward_standard)):
        sd, l90, l103 = gio.load_other_model(kind)
        scale = max(1.0, l90.abs().max().item())
        assert (fn(sd, x90) - l90).abs().max().item() <= 2e-5 * scale, kind
        assert (fn(sd, x103) - l103).abs().max().item() <= 2e-5 * scale, kind
        p = torch.softmax(l90, dim=1)[:, 1]
        assert p.min() < 0.65 and p.max() > 0.75, (kind, p.min(), p.max())     # the goldens straddle the 0.7 threshold


def test_reference_spectral_contrast_rows_are_nan():
    """Why AudioPreprocessor(use_spectral_contrast=True) -- and with it the 110-feature constructor default -- raises here
    instead of being built: run unmodified, the reference returns NaN in ALL seven spectral-contrast rows for every input
    (golden recorded from it), because its first band holds a single FFT bin (restated in oracle.spectral_contrast)."""
    import os
    import numpy as np
    from tests import golden_io as gio
    z = np.load(os.path.join(gio.GOLD, "contrast_nan.npz"))
    assert z["nan_rows"].shape == (4, 110)
    assert z["nan_rows"][:, 103:].all() and not z["nan_rows"][:, :103].any()      # rows 103..109 NaN, the rest fine
    assert z["finite_rows"][:, :103].all()
    assert co.spectral_contrast_band_edges() == [1, 2, 4, 10, 23, 52, 116, 256]
    clips, _, _ = gio.load_features()
    sc = co.spectral_contrast(co.pad_or_trim(co.normalize(clips[0].reshape(1, -1))))
    assert sc.shape == (1, 7, 101) and torch.isnan(sc).all()
