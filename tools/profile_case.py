"""Small fixed case for ncu: two e2e passes over B clips (pass 1 = warm-up).  Launch order per 16 384-clip slab:
feat_kernel, stem_tc_kernel, conv3x3_s2_tc_kernel, conv3x3_s1_tc_kernel, conv3x3_s2_tc_kernel, conv3x3_s1_pair_tc_kernel,
head_kernel (7 launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, create_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = {"fp32": 0, "bf16": 1}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
h = _lib.get_handle(0)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="p")
x = synth.synth_batch(B, seed=5, device="cuda")
for _ in range(2):
    out = h.e2e_forward(x, prec)
torch.cuda.synchronize()
print("ok", float(out.sum()), h.launches)
