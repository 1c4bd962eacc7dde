"""A/B of the stem's pruned taps (CD_STEM_PRUNE, read per call by the library): cd_cnn_forward alone and the end-to-end
call over 16 384 resident clips, alternating settings; logits must be bit-identical (the pruned terms are + 0)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, create_model
B = 16384
h = _lib.get_handle(0)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="p")
x = synth.synth_batch(B, seed=5, device="cuda")
feat = h.feat_forward(x)
ref = {}


def timed(fn, reps=30):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for prune in ("1", "0", "1", "0", "1", "0"):
    os.environ["CD_STEM_PRUNE"] = prune
    for name, fn in (("cnn", lambda: h.cnn_forward(feat, 1)), ("e2e", lambda: h.e2e_forward(x, 1))):
        ms, out = timed(fn)
        ref.setdefault(name, out.clone())
        print(f"prune {prune} {name}: {ms:.4f} ms  {ms * 1e6 / B:.2f} ns/clip  identical={bool(torch.equal(out, ref[name]))}", flush=True)
