"""Times cd_feat_forward alone (CUDA events, 5 launches after 2 warm-ups).  CD_LIB=<path> picks an experimental build."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth
if os.environ.get("CD_LIB"):
    _lib.LIB_PATH = os.environ["CD_LIB"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
h = _lib.get_handle(0)
x = synth.synth_batch(B, seed=100, device="cuda")
for _ in range(2):
    f = h.feat_forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f = h.feat_forward(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{os.environ.get('CD_LIB', 'default')}: {ms:.3f} ms  {B / ms * 1e3 / 1e6:.2f} M clips/s  {ms * 1e6 / B:.2f} ns/clip  checksum {float(f.double().sum()):.6e}")
