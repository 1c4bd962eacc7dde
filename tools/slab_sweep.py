"""e2e (audio resident -> logits) time over 16 384 clips as a function of the slab the call is cut into (CD_E2E_SLAB, read per
call by the library).  Smaller slabs keep a slab's activations in the 126 MB L2; larger ones amortise the per-launch ramp."""
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
ref = None
for slab in (16384, 8192, 4096, 2048, 1024, 592):
    os.environ["CD_E2E_SLAB"] = str(slab)
    for _ in range(3):
        out = h.e2e_forward(x, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = h.e2e_forward(x, 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ref = out.clone() if ref is None else ref
    print(f"slab {slab:6d}: {ms:.3f} ms  {ms * 1e6 / B:.1f} ns/clip  {B / ms / 1e3:.2f} M clips/s  identical={bool(torch.equal(out, ref))}", flush=True)
