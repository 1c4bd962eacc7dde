"""Latency of one 1-s window (the real-time use of the reference: src/inference.py's chunk loop): audio -> P(cough)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, create_model
h = _lib.get_handle(0)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="l")
out = {}
for B in (1, 8, 64, 512):
    x = synth.synth_batch(B, seed=3, device="cuda")
    host = x.cpu().pin_memory()
    res = torch.empty(B, 2).pin_memory()
    for _ in range(20):
        h.e2e_forward(x, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        h.e2e_forward(x, 1)
    e1.record()
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) / 200 * 1e3
    for _ in range(20):
        h.e2e_forward_host(host, 1, res)
    t0 = time.perf_counter()
    for _ in range(200):
        h.e2e_forward_host(host, 1, res)          # synchronous: host audio in, host logits out
    host_us = (time.perf_counter() - t0) / 200 * 1e6
    out[f"B{B}"] = {"device_resident_us": round(dev_us, 1), "host_in_host_out_us": round(host_us, 1)}
print(json.dumps(out))
