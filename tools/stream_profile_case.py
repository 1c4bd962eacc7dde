import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth
h = _lib.get_handle(0)
x = synth.synth_batch(900, seed=9, device="cuda").reshape(-1)
for _ in range(2):
    ft = h.feat_forward_stream(x[: 16000 + 1600 * 8191], 1600)
torch.cuda.synchronize()
print("ok", ft.shape)
