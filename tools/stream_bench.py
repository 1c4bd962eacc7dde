"""BASELINE config 4 on one GPU: 1 h of 16 kHz audio, 1-s window, 100 ms hop (35 991 windows), audio -> P(cough)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, create_model

h = _lib.get_handle(0)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="s")
n, hop = 57_600_000, 1600
x = synth.synth_batch(n // 16000, seed=9, device="cuda").reshape(-1)
out = {}
for name, hp in (("shared_stft_hop1600", 1600), ("per_window_hop1599", 1599)):
    for _ in range(2):
        lg = h.e2e_forward_stream(x, hp, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lg = h.e2e_forward_stream(x, hp, 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[name] = {"windows": lg.shape[0], "ms": ms, "windows_per_s": lg.shape[0] / ms * 1e3, "x_realtime": 3600.0 / (ms * 1e-3)}
    for _ in range(3):
        ft = h.feat_forward_stream(x[: 16000 + hp * 8191], hp)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(5):
        ft = h.feat_forward_stream(x[: 16000 + hp * 8191], hp)
    f1.record()
    torch.cuda.synchronize()
    out[name]["features_only_windows_per_s"] = 8192 / (f0.elapsed_time(f1) / 5) * 1e3
print(json.dumps(out))
