"""e2e (audio resident -> logits) time over 16 384 clips as a function of the slab the call is cut into (CD_E2E_SLAB, read per
call by the library).  Smaller slabs keep a slab's activations in the 126 MB L2; larger ones amortise the per-launch ramp."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, create_model
if os.environ.get("CD_LIB"):
    _lib.LIB_PATH = os.environ["CD_LIB"]
B = 16384
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"      # three repeats of the default configuration only
h = _lib.get_handle(0)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="p")
x = synth.synth_batch(B, seed=5, device="cuda")
ref = None
CASES = ((16384, "1"),) * 3 if QUICK else ((16384, "1"), (16384, "0"), (16384, "1"), (16384, "0"), (8192, "1"), (4096, "1"), (2048, "1"), (1024, "1"))
for slab, order in CASES:
    os.environ["CD_E2E_SLAB"] = str(slab)
    os.environ["CD_TC_ORDER"] = order          # 1: tile order alternates down the CNN chain, 0: every kernel forward
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
    print(f"slab {slab:6d} order {order}: {ms:.3f} ms  {ms * 1e6 / B:.1f} ns/clip  {B / ms / 1e3:.2f} M clips/s  identical={bool(torch.equal(out, ref))}", flush=True)
