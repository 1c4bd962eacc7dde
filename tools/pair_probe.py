import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib
h = _lib.get_handle(0)
g = torch.Generator().manual_seed(0)
for n, k in ((128, 64), (128, 256), (64, 128), (256, 128)):
    a = torch.randn(256, k, generator=g); b = torch.randn(n, k, generator=g)
    try:
        d = h.selftest_umma(a, b, 0, 3)
    except Exception as e:
        print(n, k, "ERR", e); continue
    ref = a.bfloat16().double() @ b.bfloat16().double().t()
    err = (d.double() - ref).abs() / ref.abs().max()
    print(n, k, "rel err %.2e" % err.max().item(), "rows0-127 %.1e rows128-255 %.1e" % (err[:128].max().item(), err[128:].max().item()),
          "cols lo %.1e hi %.1e" % (err[:, : n // 2].max().item(), err[:, n // 2:].max().item()))
