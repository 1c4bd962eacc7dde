"""Does capturing one e2e slab (7 launches) in a CUDA graph buy anything over stream launches?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, synth, create_model
h = _lib.get_handle(0)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="g")
B = 16384
x = synth.synth_batch(B, seed=5, device="cuda")


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ref = h.e2e_forward(x, 1)
t_stream = timed(lambda: h.e2e_forward(x, 1))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    h.e2e_forward(x, 1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out = h.e2e_forward(x, 1)
torch.cuda.synchronize()
t_graph = timed(g.replay)
print(f"stream {t_stream:.4f} ms  graph {t_graph:.4f} ms  same {torch.equal(out, ref)}")
