"""Parity of the CUDA CoughDetectorResidual forward / predict / engine with the reference goldens."""
import os
import tempfile

import numpy as np
import pytest
import torch

from cough_detector_b200 import _lib, synth, create_model, CoughDetectorInference
from cough_detector_b200.model import _PREC
from oracle import cough_oracle as co
from tests import golden_io as gio

pytestmark = pytest.mark.gpu
# Stated tolerances (north_star: "logits within a stated bf16 tolerance, decisions at 0.7 identical outside the band").
#   fp32 path : |logit - ref| <= TOL_FP32 * scale,  scale = max(1, max|ref logit|)
#   bf16 path : |logit - ref| <= TOL_BF16_REL * scale  (bf16 weights + activations, fp32 accumulate; the measured worst
#               case over the goldens and the 4 096-clip seeded batch is printed by the tests, the bound is ~3x that)
#   decisions : P(cough) >= 0.7 must equal the fp32 reference's wherever |p_ref - 0.7| > TOL_P_BF16 (fixed band, not
#               derived from the measured error); the number of clips inside the band is printed and bounded.
TOL_FP32 = 2e-4
TOL_BF16_REL = 1e-2
TOL_P_FP32 = 1e-4
TOL_P_BF16 = 2e-2
PRECISIONS = ["fp32", "bf16"]


def _model(sd, precision):
    m = create_model("residual", n_mels=90, num_classes=2, in_channels=1).eval()
    m.load_state_dict(sd)
    m.precision = precision
    return m


@pytest.mark.parametrize("precision", PRECISIONS)
def test_golden_logits_and_decisions(precision):
    _, feats, _ = gio.load_features()
    sd0, sd1, z = gio.load_model()
    x = feats.unsqueeze(1)
    for name, sd in (("default", sd0), ("random", sd1)):
        m = _model(sd, precision)
        ref = torch.from_numpy(z[f"logits_{name}"])
        scale = max(1.0, ref.abs().max().item())
        logits = m(x)
        assert logits.shape == (x.shape[0], 2) and logits.device.type == "cpu"
        err = (logits - ref).abs().max().item()
        tol = TOL_FP32 * scale if precision == "fp32" else TOL_BF16_REL * scale
        assert err <= tol, (name, precision, err, tol)
        preds, probs = m.predict(x)
        assert preds.dtype == torch.int64 and probs.shape == (x.shape[0], 2)
        pref = z[f"probs_{name}"]
        # decisions at 0.7 identical outside the FIXED tolerance band |p_ref - 0.7| <= band
        band = TOL_P_FP32 if precision == "fp32" else TOL_P_BF16
        assert float(np.abs(probs.numpy() - pref).max()) <= band, (name, precision)
        dec, dref = probs[:, 1].numpy() >= 0.7, pref[:, 1] >= 0.7
        outside = np.abs(pref[:, 1] - 0.7) > band
        assert np.array_equal(dec[outside], dref[outside])
        amb = np.abs(pref[:, 1] - 0.5) > band
        assert np.array_equal(preds.numpy()[amb], z[f"preds_{name}"][amb])
        print(f"{name}/{precision}: logit err {err:.3e} (scale {scale:.2f}), prob err {np.abs(probs.numpy() - pref).max():.3e}")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_seeded_batch_vs_oracle(precision):
    _, sd1, _ = gio.load_model()
    h = _lib.get_handle(0)
    x = torch.stack(synth.synth_clips(300, seed=21, fixed_length=True))
    feat = h.feat_forward(x.cuda())
    m = _model(sd1, precision)
    logits = m(feat)
    ref = co.cnn_forward(sd1, feat.cpu())
    scale = max(1.0, ref.abs().max().item())
    err = (logits.cpu() - ref).abs().max().item()
    assert err <= (TOL_FP32 if precision == "fp32" else TOL_BF16_REL) * scale, err
    # fused audio -> logits equals the two-step path bit for bit, and batch position does not matter
    fused = m.forward_audio(x.cuda())
    assert torch.equal(fused, logits)
    perm = torch.randperm(300, device="cuda")
    assert torch.equal(m.forward_audio(x.cuda()[perm]), fused[perm])
    assert m(feat[:0]).shape == (0, 2)


def test_bf16_within_emulated_bf16_band():
    """The tcgen05 path should agree with a CPU emulation of bf16 operands / fp32 accumulate much more
    tightly than with the fp32 reference."""
    _, feats, _ = gio.load_features()
    _, sd1, _ = gio.load_model()
    x = feats.unsqueeze(1)
    emu = co.cnn_forward_bf16(sd1, x)
    got = _model(sd1, "bf16")(x)
    scale = max(1.0, emu.abs().max().item())
    assert (got - emu).abs().max().item() <= 1e-3 * scale


def test_weight_swap_and_reload():
    _, feats, _ = gio.load_features()
    sd0, sd1, z = gio.load_model()
    x = feats[:4].unsqueeze(1)
    m = _model(sd0, "fp32")
    a = m(x)
    m.load_state_dict(sd1)                      # in-place parameter update must trigger a re-upload
    b = m(x)
    assert (a - torch.from_numpy(z["logits_default"][:4])).abs().max() <= TOL_FP32
    assert (b - torch.from_numpy(z["logits_random"][:4])).abs().max() <= TOL_FP32 * max(1.0, float(np.abs(z["logits_random"]).max()))
    m2 = _model(sd0, "fp32")
    assert torch.equal(m2(x), a)


def test_host_buffer_e2e_equals_device_path():
    _, sd1, _ = gio.load_model()
    h = _lib.get_handle(0)
    m = _model(sd1, "fp32")
    x = torch.stack(synth.synth_clips(5000, seed=2, fixed_length=True))      # > 2 pipeline chunks
    dev = m.forward_audio(x.cuda()).cpu()
    for prec in PRECISIONS:
        m.precision = prec
        dev = m.forward_audio(x.cuda()).cpu()
        host = h.e2e_forward_host(x.pin_memory(), _PREC[prec])
        assert torch.equal(host, dev)
        assert torch.equal(h.e2e_forward_host(x, _PREC[prec]), dev)           # pageable buffer


def test_engine_checkpoint_roundtrip_and_chunk_loop():
    stream, zs = gio.load_stream()
    e = gio.load_engine()
    _, sd1, _ = gio.load_model()
    cfg = dict(model_type="residual", use_pcen=False, use_pre_emphasis=False, use_delta_delta=False,
               use_spectral_contrast=False)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "best_model.pt")
        torch.save({"epoch": 0, "model_state_dict": sd1, "optimizer_state_dict": {}, "metrics": {}, "config": cfg}, path)
        eng = CoughDetectorInference(path, device="cuda", confidence_threshold=0.5, smoothing_window=3,
                                     debounce_seconds=0.0, verbose=False)
        eng.model.precision = "fp32"
        fired, confs = [], []
        for c in range(0, stream.numel(), 4000):
            r = eng.process_audio_chunk(stream[c:c + 4000].numpy())
            fired.append(r is not None)
            confs.append(-1.0 if r is None else float(r[1]))
        assert fired == list(e["fired"])
        assert np.allclose(confs, e["confs"], atol=2e-4)
        probs = eng.stream_probabilities(stream, 0.25).cpu().numpy()
        assert np.abs(probs - e["probs"]).max() <= 2e-4
        ev = eng.detect_stream(stream, 0.25)
        assert [round((t - 1.0) / 0.25) for t, _ in ev] == [i - 3 for i in np.nonzero(e["fired"])[0]]
        with pytest.raises(NotImplementedError):          # checkpoint without feature flags = 110-feature config
            torch.save({"model_state_dict": sd1, "config": {"model_type": "residual"}}, path)
            CoughDetectorInference(path, device="cuda", verbose=False)


@pytest.mark.parametrize("rows", [90, 103])
def test_ragged_batch_sizes_are_prefixes_of_a_big_batch(rows):
    """Last-tile handling of every persistent kernel (tiles of 128 GEMM rows over clips of 648 / 168 / 56 ... rows, the
    CTA pair's odd tile, zero padding shared between neighbouring clips): a batch of n clips must give exactly the first n
    logits of a larger batch, whatever n is."""
    _, sd1, _ = gio.load_model()
    m = _model(sd1, "bf16")
    g = torch.Generator().manual_seed(rows)
    x = torch.rand(1031, 1, rows, 101, generator=g).cuda()
    x[:, :, 64:] = (x[:, :, 64:] - 0.5) * 6.0                 # MFCC-like rows are signed
    big = m(x)
    for n in (1, 2, 3, 5, 7, 16, 31, 33, 127, 129, 255, 257, 1000):
        assert torch.equal(m(x[:n]), big[:n]), n
    ref = co.cnn_forward(sd1, x[:64].cpu())
    assert (big[:64].cpu() - ref).abs().max().item() <= TOL_BF16_REL * max(1.0, ref.abs().max().item())


def test_engine_with_all_supported_feature_flags():
    """A checkpoint whose config turns on pre-emphasis, PCEN and delta-delta (103 rows): the engine's window
    probabilities (chunk loop, fused stream call) against the oracle pipeline (explicit fp32 features -> functional CNN)."""
    stream, _ = gio.load_stream()
    _, sd1, _ = gio.load_model()
    cfg = dict(model_type="residual", use_pcen=True, use_pre_emphasis=True, pre_emphasis_coef=0.95, use_delta_delta=True,
               use_spectral_contrast=False)
    ex = co.Explicit32(True, True, 0.95, use_delta_delta=True)
    feats = co.stream_features(stream, 4000, ex)
    assert feats.shape[1:] == (1, 103, 101)
    want = torch.softmax(co.cnn_forward(sd1, feats), dim=1)[:, 1].numpy()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "best_model.pt")
        torch.save({"epoch": 0, "model_state_dict": sd1, "optimizer_state_dict": {}, "metrics": {}, "config": cfg}, path)
        eng = CoughDetectorInference(path, device="cuda", confidence_threshold=0.5, smoothing_window=1,
                                     debounce_seconds=0.0, verbose=False)
        assert eng.preprocessor.get_num_features() == 103
        eng.model.precision = "fp32"
        got = eng.stream_probabilities(stream, 0.25).cpu().numpy()
        assert got.shape == want.shape and np.abs(got - want).max() <= 2e-4
        loop = []
        for c in range(0, stream.numel(), 4000):
            for f in eng.preprocessor.add_audio(stream[c:c + 4000]):
                loop.append(eng.predict(f)[1])
        assert np.abs(np.array(loop) - want).max() <= 2e-4
        eng.model.precision = "bf16"
        assert np.abs(eng.stream_probabilities(stream, 0.25).cpu().numpy() - want).max() <= 3e-2


@pytest.mark.parametrize("n,k", [(64, 64), (64, 320), (128, 256), (256, 128), (32, 64)])
def test_tcgen05_tile_selftest_bf16(n, k):
    """One UMMA tile through the conv kernels' descriptor / swizzle helpers vs a float matmul of bf16-rounded operands."""
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=g)
    b = torch.randn(n, k, generator=g)
    d = _lib.get_handle(0).selftest_umma(a, b)
    ref = a.bfloat16().double() @ b.bfloat16().double().t()
    err, scale = (d.double() - ref).abs().max().item(), ref.abs().max().item()
    assert err <= 2e-3 * scale


@pytest.mark.parametrize("mode,shift", [(0, 1), (0, 7), (0, 17), (0, 32), (2, 1), (2, 9), (2, 27), (2, 31)])
def test_tcgen05_shifted_descriptor_views(mode, shift):
    """The convolutions never build an A tile: tap (ky, kx) is the staged rows viewed `shift` rows further down, i.e. an
    operand descriptor whose start address is not a multiple of the 8-row swizzle atom.  mode 0: SWIZZLE_128B rows
    (3x3 convs), mode 2: SWIZZLE_32B rows (the stem's 4x4-pixel blocks; K = 16 per row, so K is fixed by the tile)."""
    k = 128
    g = torch.Generator().manual_seed(100 * mode + shift)
    a = torch.randn(128 + shift, k, generator=g)
    b = torch.randn(64, k, generator=g)
    d = _lib.get_handle(0).selftest_umma(a, b, shift, mode)
    ref = a[shift:shift + 128].bfloat16().double() @ b.bfloat16().double().t()
    err, scale = (d.double() - ref).abs().max().item(), ref.abs().max().item()
    assert err <= 2e-3 * scale


@pytest.mark.parametrize("n,k", [(128, 64), (128, 256), (64, 128), (256, 128)])
def test_tcgen05_cta_pair_tile(n, k):
    """cta_group::2: one MMA stream over a CTA pair (M = 256; each CTA holds its own 128 A rows and N/2 of the B rows),
    the mode block 1's conv2 runs in."""
    g = torch.Generator().manual_seed(7 * n + k)
    a = torch.randn(256, k, generator=g)
    b = torch.randn(n, k, generator=g)
    d = _lib.get_handle(0).selftest_umma(a, b, 0, 3)
    ref = a.bfloat16().double() @ b.bfloat16().double().t()
    err, scale = (d.double() - ref).abs().max().item(), ref.abs().max().item()
    assert err <= 2e-3 * scale


def _dd_features():
    _, base, _ = gio.load_features()
    return torch.cat([base, gio.load_delta_delta_rows()["dd"]], dim=1).unsqueeze(1)        # (32, 1, 103, 101)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_103_row_network_vs_reference_goldens(precision):
    """use_delta_delta feature sets: the same network on 103 rows (heights 26 / 13 / 7), reference logits as goldens."""
    sd0, sd1, _ = gio.load_model()
    want = gio.load_delta_delta_logits()
    x = _dd_features()
    for name, sd in (("default", sd0), ("random", sd1)):
        got = _model(sd, precision)(x)
        ref = want[name]
        scale = max(1.0, ref.abs().max().item())
        tol = TOL_FP32 * scale if precision == "fp32" else TOL_BF16_REL * scale
        assert got.shape == ref.shape and (got - ref).abs().max().item() <= tol, (name, (got - ref).abs().max().item(), scale)
        if precision == "bf16":
            emu = co.cnn_forward_bf16(sd, x)
            assert (got - emu).abs().max().item() <= 5e-3 * scale, (name, (got - emu).abs().max().item())
        print(f"103 rows {name}/{precision}: logit err {(got - ref).abs().max().item():.3e} (scale {scale:.2f})")


def test_switching_between_90_and_103_row_networks():
    """The tensor-core workspace is laid out per shape family and its zero padding sits in different places: going
    90 -> 103 -> 90 rows (and a big batch in between) must reproduce the first answers bit for bit."""
    from cough_detector_b200 import AudioPreprocessor
    _, sd1, _ = gio.load_model()
    m = _model(sd1, "bf16")
    _, base, _ = gio.load_features()
    x90, x103 = base.unsqueeze(1).cuda(), _dd_features().cuda()
    a90, a103 = m(x90), m(x103)
    big = torch.stack(synth.synth_clips(700, seed=8, fixed_length=True)).cuda()
    kw = dict(use_pcen=False, use_pre_emphasis=False, use_spectral_contrast=False)
    f103 = m.forward_audio(big, preprocessor=AudioPreprocessor(use_delta_delta=True, **kw))
    f90 = m.forward_audio(big)
    assert torch.equal(m(x90), a90) and torch.equal(m(x103), a103)
    # fused audio -> logits equals features -> network for both families
    pp = AudioPreprocessor(use_delta_delta=True, **kw)
    feats = pp.extract_features_batch(big)
    assert feats.shape[2] == 103 and torch.equal(m(feats), f103)
    assert torch.equal(m.forward_audio(big), f90) and (f90 - f103).abs().max().item() > 1e-4
    # fp32 path takes the family from the call's options too
    m32 = _model(sd1, "fp32")
    assert (m32.forward_audio(big[:64], preprocessor=pp) - f103[:64]).abs().max().item() <= 3e-2 * f103.abs().max().item()


def test_every_entry_point_on_small_shapes():
    """tools/sanitize_case.py: all kernels / C-ABI entry points on tiny, ragged inputs (the case written for
    compute-sanitizer; that tool is closed on this pool, so it runs as a plain smoke test)."""
    import runpy
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, "tools", "sanitize_case.py"), run_name="__main__")


def test_cpp_example_runs():
    """The torch-free C++ caller: cd_create -> cd_cnn_load -> cd_e2e_forward -> cd_e2e_forward_host, both paths bit-identical."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run(["make", "-C", root, "examples"], capture_output=True, text=True).returncode == 0
    r = subprocess.run([os.path.join(root, "build", "e2e_minimal")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host path agrees: bit for bit" in r.stdout, r.stdout


def test_bf16_decisions_at_0p7_on_4096_seeded_clips():
    """End to end on the SURVEY section-8d mixture: audio -> fp32 oracle features -> fp32 oracle network is the reference;
    the GPU runs audio -> fused bf16 path.  The classifier is re-centred (synth.spread_fc) so that P(cough) straddles
    0.7 -- with raw random weights every clip lands on one side and 'identical decisions' says nothing."""
    _, sd1, _ = gio.load_model()
    n = 4096
    x = synth.synth_batch(n, seed=77, device="cpu")
    ex = co.Explicit32()
    feats = torch.stack([ex.process(c.reshape(1, -1)) for c in x])           # (n, 1, 90, 101), fp32 oracle
    sd = synth.spread_fc(sd1, co.cnn_forward(sd1, feats))
    ref = co.cnn_forward(sd, feats)
    p_ref = torch.softmax(ref, dim=1)[:, 1].numpy()
    scale = max(1.0, ref.abs().max().item())
    h = _lib.get_handle(0)
    for precision, tol_l, tol_p in (("fp32", TOL_FP32, TOL_P_FP32), ("bf16", TOL_BF16_REL, TOL_P_BF16)):
        m = _model(sd, precision)
        logits = m.forward_audio(x.cuda())
        _, probs, dec = h.predict(logits, threshold=0.7)
        p = probs[:, 1].cpu().numpy()
        lerr = (logits.cpu() - ref).abs().max().item()
        perr = float(np.abs(p - p_ref).max())
        outside = np.abs(p_ref - 0.7) > tol_p
        inside = int((~outside).sum())
        flips_in = int((dec.cpu().numpy()[~outside] != (p_ref[~outside] >= 0.7)).sum())
        print(f"{precision}: {n} clips, logit err {lerr:.3e} (scale {scale:.2f}, rel {lerr / scale:.2e}), prob err {perr:.3e}, "
              f"{inside} clips inside |p-0.7| <= {tol_p} ({flips_in} flipped there), "
              f"{int((p_ref >= 0.7).sum())} positive / {int((p_ref < 0.7).sum())} negative")
        assert lerr <= tol_l * scale, (precision, lerr, scale)
        assert perr <= tol_p, (precision, perr)
        assert np.array_equal(dec.cpu().numpy()[outside], p_ref[outside] >= 0.7)
        # the check is not vacuous: both classes are well populated and the band holds few clips
        assert (p_ref[outside] >= 0.7).sum() >= n // 5 and (p_ref[outside] < 0.7).sum() >= n // 5
        assert inside <= n // 20


def test_two_handles_on_one_gpu_are_independent():
    """include/cough_b200.h: calls on different handles are independent -- also two handles on the same GPU.  Round 1
    kept the bf16 weights and the resampling tables per device index; here two private handles load different weights,
    are used alternately, and one outlives the other."""
    sd0, sd1, z = gio.load_model()
    _, feats, _ = gio.load_features()
    x = feats.unsqueeze(1).cuda()
    shared = _lib.get_handle(0)
    h0, h1 = _lib.Handle(0), _lib.Handle(0)
    try:
        h0.load_state_dict(sd0)
        h1.load_state_dict(sd1)                                  # must not touch h0's weights (either precision)
        want = {}
        for prec in (0, 1):
            want[prec] = (h0.cnn_forward(x, prec).clone(), h1.cnn_forward(x, prec).clone())
            assert (want[prec][0].cpu() - torch.from_numpy(z["logits_default"])).abs().max().item() <= (TOL_FP32 if prec == 0 else TOL_BF16_REL)
            sc = max(1.0, float(np.abs(z["logits_random"]).max()))
            assert (want[prec][1].cpu() - torch.from_numpy(z["logits_random"])).abs().max().item() <= (TOL_FP32 if prec == 0 else TOL_BF16_REL) * sc
        for _ in range(3):
            for prec in (1, 0):
                assert torch.equal(h1.cnn_forward(x, prec), want[prec][1])
                assert torch.equal(h0.cnn_forward(x, prec), want[prec][0])
        w = torch.randn(2, 30000, device="cuda")
        r0 = h0.resample(w, 44100)
        assert torch.equal(h1.resample(w, 44100), r0) and torch.equal(shared.resample(w, 44100), r0)
        h1.close()                                               # frees ITS weights, tables and workspaces only
        torch.cuda.synchronize()
        junk = torch.full((64 << 20,), 7.0, device="cuda")       # recycle the freed memory
        for prec in (0, 1):
            assert torch.equal(h0.cnn_forward(x, prec), want[prec][0])
        assert torch.equal(h0.resample(w, 44100), r0)
        del junk
    finally:
        h0.close()
        h1.close()


def test_calls_on_different_streams_share_the_workspace_safely():
    """One handle, several caller streams and the host pipeline's private streams: the library orders successive users
    of its workspace with an event, so un-synchronised back-to-back calls give the serial answers."""
    _, sd1, _ = gio.load_model()
    h = _lib.get_handle(0)
    m = _model(sd1, "bf16")
    a = synth.synth_batch(6000, seed=5, device="cuda")
    b = synth.synth_batch(6000, seed=6, device="cuda")
    want_a, want_b = m.forward_audio(a).clone(), m.forward_audio(b).clone()
    feat_b = h.feat_forward(b)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    host_a = a.cpu().pin_memory()
    for _ in range(5):
        with torch.cuda.stream(s1):
            la = h.e2e_forward(a, 1)
        with torch.cuda.stream(s2):
            lb = h.e2e_forward(b, 1)                             # no sync in between: same workspace, other stream
            lc = h.cnn_forward(feat_b, 1)
        lh = h.e2e_forward_host(host_a, 1)                       # private pipeline streams, right behind both
        torch.cuda.synchronize()
        assert torch.equal(la, want_a) and torch.equal(lb, want_b) and torch.equal(lc, want_b)
        assert torch.equal(lh.cuda(), want_a)


def test_int16_and_host_audio_through_forward_audio():
    """forward_audio: CPU audio takes the library's pipelined host path, int16 PCM (CPU or CUDA) stays int16 on the wire;
    all of them give the float device path's logits bit for bit (the 1 / 32768 scale cancels in the peak normalisation).
    Write-combined host memory from cd_host_alloc works as a source."""
    _, sd1, _ = gio.load_model()
    m = _model(sd1, "bf16")
    x = torch.stack(synth.synth_clips(3000, seed=4, fixed_length=True))
    pcm = (x / x.abs().amax(dim=1, keepdim=True).clamp(min=1e-9) * 30000.0).round().to(torch.int16)
    xf = pcm.float() / 32768.0
    want = m.forward_audio(xf.cuda())
    assert torch.equal(m.forward_audio(pcm.cuda()), want)
    assert torch.equal(m.forward_audio(pcm).cuda(), want) and m.forward_audio(pcm).device.type == "cpu"
    assert torch.equal(m.forward_audio(xf).cuda(), want)
    wc = _lib.host_alloc(xf.numel(), torch.float32, write_combined=True).view_as(xf)
    wc.copy_(xf)
    assert torch.equal(_lib.get_handle(0).e2e_forward_host(wc, 1).cuda(), want)
    lg, ft = m.forward_audio(xf[:64], return_features=True)
    assert lg.device.type == "cpu" and ft.shape == (64, 1, 90, 101) and torch.equal(lg.cuda(), want[:64])


@pytest.mark.parametrize("kind", ["small", "standard"])
def test_other_model_types_vs_reference_goldens(kind):
    """create_model('small' | 'standard') (src/model.py:43-207; 'small' is the reference's default model_type): fp32
    CUDA-core kernels against logits recorded from the reference, 90- and 103-row inputs, ragged batch sizes."""
    sd, l90, l103 = gio.load_other_model(kind)
    _, feats, _ = gio.load_features()
    x90 = feats.unsqueeze(1)
    x103 = _dd_features()
    m = create_model(kind, n_mels=90, num_classes=2, in_channels=1).eval()
    m.load_state_dict(sd)
    scale = max(1.0, l90.abs().max().item())
    for x, want in ((x90, l90), (x103, l103)):
        got = m(x)
        assert got.shape == want.shape and got.device.type == "cpu"
        assert (got - want).abs().max().item() <= TOL_FP32 * scale, (kind, (got - want).abs().max().item(), scale)
    preds, probs = m.predict(x90)
    assert preds.dtype == torch.int64 and torch.equal(preds, torch.softmax(l90, dim=1).argmax(dim=1))
    big = torch.cat([x90] * 40).cuda()                                   # 1280 clips: more than one workspace chunk
    out = m(big)
    assert torch.equal(out[:32], out[1248:]) and (out[:32].cpu() - l90).abs().max().item() <= TOL_FP32 * scale
    for n in (1, 3, 33):
        assert torch.equal(m(big[:n]), out[:n])
    assert m(big[:0]).shape == (0, 2)


def test_engine_with_the_reference_default_model_type():
    """A checkpoint written by the reference's stock command line has no model_type in its config, i.e. 'small'
    (src/inference.py:146, src/train.py:218): it must load and run, chunk loop and whole-recording call."""
    stream, _ = gio.load_stream()
    sd, _, _ = gio.load_other_model("small")
    cfg = dict(use_pcen=False, use_pre_emphasis=False, use_delta_delta=False, use_spectral_contrast=False)
    feats = co.stream_features(stream, 4000, co.Explicit32())
    want = torch.softmax(co.cnn_forward_small(sd, feats), dim=1)[:, 1].numpy()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "best_model.pt")
        torch.save({"epoch": 0, "model_state_dict": sd, "optimizer_state_dict": {}, "metrics": {}, "config": cfg}, path)
        eng = CoughDetectorInference(path, device="cuda", confidence_threshold=0.5, smoothing_window=1, debounce_seconds=0.0, verbose=False)
        assert type(eng.model).__name__ == "CoughDetectorSmall"
        got = eng.stream_probabilities(stream, 0.25).cpu().numpy()
        assert got.shape == want.shape and np.abs(got - want).max() <= 2e-4
        loop = []
        for c in range(0, stream.numel(), 4000):
            for f in eng.preprocessor.add_audio(stream[c:c + 4000]):
                loop.append(eng.predict(f)[1])
        assert np.abs(np.array(loop) - want).max() <= 2e-4
