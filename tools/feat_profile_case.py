"""The bench's roofline launch under ncu: cd_feat_forward over 16 384 resident clips (fp32 features out)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth
if os.environ.get("CD_LIB"):
    _lib.LIB_PATH = os.environ["CD_LIB"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
h = _lib.get_handle(0)
x = synth.synth_batch(B, seed=100, device="cuda")
for _ in range(2):
    f = h.feat_forward(x)
torch.cuda.synchronize()
print("ok", float(f[0].sum()))
