"""CPU-side checks: the C-ABI library loads and exports every declared symbol, its tables equal
torchaudio's, and the host-side mirror keeps the reference's semantics and error behaviour."""
import ctypes
import re
import os

import numpy as np
import pytest
import torch

from cough_detector_b200 import _lib, synth
from cough_detector_b200 import (AudioPreprocessor, RealtimePreprocessor, create_preprocessor, create_model,
                                 count_parameters, smooth_and_debounce)
from oracle import cough_oracle as co
from tests import golden_io as gio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = dict(use_pcen=False, use_pre_emphasis=False, use_delta_delta=False, use_spectral_contrast=False)


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "cough_b200.h")).read()
    declared = set(re.findall(r"\b(cd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.cd_version()
    assert lib.cd_stream_windows(57_600_000, 1600) == 35991
    assert lib.cd_stream_windows(15999, 1600) == 0 and lib.cd_stream_windows(16000, 0) == 0
    assert lib.cd_resample_length(16330, 44100, 16000) == 5925 and lib.cd_resample_length(2973, 8000, 16000) == 5946


def test_library_tables_are_the_reference_tables():
    hann, fb, dct = _lib.get_tables()
    _, _, z = gio.load_features()
    assert np.array_equal(hann.numpy(), z["hann"])
    assert np.array_equal(fb.numpy(), z["fb"])
    assert np.array_equal(dct.numpy(), z["dct"])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_no_fallback():
    lib = _lib.load_library()
    h = ctypes.c_void_p()
    assert lib.cd_create(0, ctypes.byref(h)) != 0
    assert lib.cd_last_error(None)
    pp = AudioPreprocessor(**FLAGS)
    with pytest.raises(_lib.CoughB200Error):
        pp.extract_features(torch.zeros(1, 16000))
    m = create_model("residual").eval()
    with pytest.raises(_lib.CoughB200Error):
        m(torch.zeros(1, 1, 90, 101))


def test_preprocessor_surface_and_errors():
    pp = AudioPreprocessor(**FLAGS)
    assert pp.get_num_features() == 90 and pp.get_expected_time_frames() == 101 and pp.segment_samples == 16000
    with pytest.raises(NotImplementedError):
        AudioPreprocessor()                      # reference defaults = 110-feature config: not built, no fallback
    with pytest.raises(NotImplementedError):
        AudioPreprocessor(n_mels=80, **FLAGS)
    with pytest.raises(NotImplementedError):
        AudioPreprocessor(**{**FLAGS, "use_spectral_contrast": True})               # NaN rows in the reference; not built
    assert AudioPreprocessor(**{**FLAGS, "use_delta_delta": True}).get_num_features() == 103    # features-only option
    both = AudioPreprocessor(**{**FLAGS, "use_pcen": True, "use_pre_emphasis": True, "pre_emphasis_coef": 0.95})
    assert both.get_num_features() == 90 and both.use_pcen and both.use_pre_emphasis and both.pre_emphasis_coef == 0.95
    with pytest.raises(ValueError):
        AudioPreprocessor(**{**FLAGS, "use_pre_emphasis": True, "pre_emphasis_coef": 1.5})
    with pytest.raises(ValueError):
        pp.resample(torch.zeros(1, 10), 44100.5)
    assert pp.resample(torch.ones(1, 4), 16000).shape == (1, 4)
    assert isinstance(create_preprocessor(True, hop_duration=0.1, **FLAGS), RealtimePreprocessor)
    with pytest.raises(ValueError):
        pp.extract_features(torch.zeros(16000))
    with pytest.raises(ValueError):
        pp.extract_features(torch.zeros(1, 8000))
    for x in synth.edge_clips() + [torch.randn(2, 500)]:
        w = x.reshape(1, -1) if x.dim() == 1 else x
        assert torch.equal(pp.to_mono(w), co.to_mono(w))
        m = pp.to_mono(w)
        assert torch.equal(pp.normalize(m), co.normalize(m))
        assert torch.equal(pp.pad_or_trim(m), co.pad_or_trim(m))
    f = torch.randn(1, 13, 101)
    ref = torch.nn.functional.pad(f, (1, 1), mode="replicate")
    assert torch.equal(pp.compute_deltas(f), (ref[:, :, 2:] - ref[:, :, :-2]) / 2)


def test_model_surface_matches_reference_state_dict():
    sd0, sd1, _ = gio.load_model()
    m = create_model("residual", n_mels=90, num_classes=2, in_channels=1)
    assert count_parameters(m) == 290370
    msd = m.state_dict()
    assert list(msd.keys()) == list(sd0.keys())
    assert all(msd[k].shape == sd0[k].shape for k in msd)
    m.load_state_dict(sd1)                       # strict
    with pytest.raises(ValueError):
        create_model("nope")
    for kind, n_params in (("small", None), ("standard", None)):        # the reference's other model types: same state_dict
        ref_sd, _, _ = gio.load_other_model(kind)
        mk = create_model(kind, n_mels=90, num_classes=2, in_channels=1)
        assert list(mk.state_dict().keys()) == [k for k in mk.state_dict().keys() if k in ref_sd]
        assert set(mk.state_dict().keys()) == set(ref_sd.keys())
        assert all(mk.state_dict()[k].shape == ref_sd[k].shape for k in ref_sd)
        mk.load_state_dict(ref_sd)                                       # strict
    m.train()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 90, 101))


def test_same_seed_gives_reference_initialisation():
    """Module construction order equals the reference's, so torch.manual_seed(0) reproduces its init."""
    sd0, _, _ = gio.load_model()
    torch.manual_seed(0)
    m = create_model("residual", n_mels=90, num_classes=2, in_channels=1)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd0[k]), k


def test_decision_postpass_matches_oracle_and_reference():
    e = gio.load_engine()
    probs = e["probs"]
    got = smooth_and_debounce(probs, 0.25, 0.5, 3, 0.0)
    want = co.decisions_stream(probs, 0.25, 0.5, 3, 0.0)
    assert got == want
    ref_idx = [i - 3 for i in np.nonzero(e["fired"])[0]]        # one window per reference call: no early-return loss
    assert [j for j, _ in got] == ref_idx
    rng = np.random.default_rng(0)
    p = rng.random(500)
    for thr, win, deb in ((0.7, 3, 0.5), (0.5, 1, 0.0), (0.6, 5, 2.0)):
        assert smooth_and_debounce(p, 0.1, thr, win, deb) == co.decisions_stream(p, 0.1, thr, win, deb)


def test_bind_to_gpu_numa_is_harmless_without_nvml():
    # no GPU / no NVML here: must report False and leave the process affinity alone
    import os

    from cough_detector_b200 import sharding

    before = os.sched_getaffinity(0)
    assert sharding.bind_to_gpu_numa(0) in (False, True)
    if not __import__("torch").cuda.is_available():
        assert os.sched_getaffinity(0) == before


def test_cpp_example_builds_against_the_c_abi():
    """examples/e2e_minimal.cu is a torch-free C++ caller of include/cough_b200.h; it must compile and link here."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["make", "-C", root, "examples"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(os.path.join(root, "build", "e2e_minimal"))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: one JSON line with the metric keys of the b200 arm, `impl`, a `cpu_baseline`
    describing this run and an `e2e` that repeats the value with zero copies (the driver computes the ratio from it)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["metric"] == "1-s clips/sec (mel+MFCC+CNN)" and line["value"] > 0 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a rank other than 0 of a torchrun launch exits 0 without work or output
    env["RANK"], env["WORLD_SIZE"] = "1", "2"
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=root)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout, out.stderr[-500:])
