"""Small shapes through every kernel, for compute-sanitizer (memcheck / racecheck), one tool per run."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, create_model, AudioPreprocessor, RealtimePreprocessor
FLAGS = dict(use_pcen=False, use_pre_emphasis=False, use_delta_delta=False, use_spectral_contrast=False)
h = _lib.get_handle(0)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="s")
clips = synth.synth_clips(12, seed=4) + synth.edge_clips()[:6]
pp = AudioPreprocessor(**FLAGS)
f = pp.process_batch(clips)                                   # var-length path
x = torch.stack(synth.synth_clips(9, seed=5, fixed_length=True)).cuda()
f2 = h.feat_forward(x)                                        # bulk-copy path
rp = RealtimePreprocessor(window_duration=1.0, hop_duration=0.1, **FLAGS)
s = torch.cat([c for c in synth.synth_clips(3, seed=6, fixed_length=True)])
f3 = rp.stream_features(s)                                    # shared-STFT path
f4 = h.feat_forward_stream(s.cuda(), 777)                     # per-window path, unaligned
for prec in (0, 1):
    lg = h.e2e_forward(x, prec)
    lg2 = h.cnn_forward(f2, prec)
    ls = h.e2e_forward_stream(s.cuda(), 1600, prec)
for opts in (dict(use_pcen=True, use_pre_emphasis=False), dict(use_pcen=False, use_pre_emphasis=True),
             dict(use_pcen=True, use_pre_emphasis=True), dict(use_pcen=False, use_pre_emphasis=False, use_delta_delta=True),
             dict(use_pcen=True, use_pre_emphasis=True, use_delta_delta=True)):   # the optional stages, every entry point again
    kw = dict(use_spectral_contrast=False, **{"use_delta_delta": False, **opts})
    po = AudioPreprocessor(**kw)
    g1 = po.process_batch(clips)
    g2 = po.extract_features_batch(x)
    g3 = RealtimePreprocessor(window_duration=1.0, hop_duration=0.1, **kw).stream_features(s)
    g4 = RealtimePreprocessor(window_duration=1.0, hop_duration=0.0485625, **kw).stream_features(s[:20000])
    g5 = m.forward_audio(x, preprocessor=po)
    assert all(torch.isfinite(t).all() for t in (g1, g2, g3, g4, g5))
host = h.e2e_forward_host(x.cpu().pin_memory(), 1)
d = h.selftest_umma(torch.randn(128 + 9, 128), torch.randn(64, 128), 9, 0)
torch.cuda.synchronize()
print("ok", f.shape, f3.shape, f4.shape, float(lg.sum()), float(host.sum()))
