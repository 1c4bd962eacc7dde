"""Feature accuracy of a library build against the fp32 oracle and the float64 truth (CD_LIB=<path> picks the build):
max |error| per row block over the seeded mixture and over the random-length / random-gain clips of the test-suite."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, AudioPreprocessor
if os.environ.get("CD_LIB"):
    _lib.LIB_PATH = os.environ["CD_LIB"]
from oracle import cough_oracle as co

FLAGS = dict(use_pcen=False, use_pre_emphasis=False, use_delta_delta=False, use_spectral_contrast=False)
ALL = dict(use_pcen=True, use_pre_emphasis=True, use_delta_delta=True, use_spectral_contrast=False)


def random_clips():
    g = torch.Generator().manual_seed(1234)
    clips = []
    for i in range(96):
        n = int(torch.randint(1, 50001, (1,), generator=g)) if i % 5 else [1, 2, 399, 16000, 15999][i // 5 % 5]
        x = torch.randn(n, generator=g) * (0.2 + 0.8 * torch.rand(1, generator=g))
        if i % 3 == 0:
            t = torch.arange(n) / 16000.0
            x = x * 0.05 + torch.sin(2 * 3.14159265 * float(torch.randint(80, 3900, (1,), generator=g)) * t)
        x = x * 10.0 ** float(torch.empty(1).uniform_(-6, 3, generator=g)) + (0.1 * float(torch.randn(1, generator=g)) if i % 4 == 0 else 0.0)
        clips.append(x.float())
    return clips


def blocks(d):
    return "mel %.2e  mfcc %.2e  delta %.2e" % (d[:64].max(), d[64:77].max(), d[77:90].max()) + ("  dd %.2e" % d[90:].max() if d.shape[0] > 90 else "")


for name, clips in (("mixture", synth.synth_clips(120, seed=11)), ("random-length", random_clips())):
    for cfg, kw, ex in (("plain", FLAGS, co.Explicit32()), ("pcen+pre+dd", ALL, co.Explicit32(True, True, use_delta_delta=True))):
        got = AudioPreprocessor(**kw).process_batch(clips)[:, 0]
        w32 = np.zeros(got.shape[1]); w64 = np.zeros(got.shape[1]); worst_i = -1; worst = 0
        n_ill = 0
        for i, x in enumerate(clips):
            if co.ill_conditioned_fraction(co.conditioned_window(x.numpy())) > 0:
                n_ill += 1
                continue
            ref = ex.process(x.reshape(1, -1))[0]
            d32 = (got[i] - ref).abs().amax(dim=1).numpy()
            if d32.max() > worst:
                worst, worst_i = d32.max(), i
            w32 = np.maximum(w32, d32)
            if i % 2 == 0:
                t64 = co.features_fp64(x.numpy(), kw["use_pcen"], kw["use_pre_emphasis"], 0.97, kw["use_delta_delta"])
                w64 = np.maximum(w64, np.abs(got[i].numpy() - t64).max(axis=1))
                o64 = np.abs(ref.numpy() - t64).max()
        print(f"{name:14s} {cfg:12s} vs fp32 oracle: {blocks(w32)} (worst clip {worst_i}, len {len(clips[worst_i])})")
        print(f"{'':14s} {'':12s} vs float64    : {blocks(w64)}   [{n_ill} ill-conditioned clips skipped]")
