"""Do the feature kernel (FP32 / shared-memory bound) and the CNN kernels (tensor / operand-fetch bound) overlap
when issued on two streams?  Sequential vs concurrent time for the same work."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, create_model
h = _lib.get_handle(0)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="s")
B = 8192
x = synth.synth_batch(B, seed=3, device="cuda")
feat = h.feat_forward(x)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(concurrent, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if concurrent:
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s1):
                f = h.feat_forward(x)
            with torch.cuda.stream(s2):
                lg = h.cnn_forward(feat, 1)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        else:
            f = h.feat_forward(x)
            lg = h.cnn_forward(feat, 1)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for _ in range(2):
    run(False, 2); run(True, 2)
print("sequential ms", run(False), "concurrent ms", run(True))
