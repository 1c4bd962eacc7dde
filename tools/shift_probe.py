"""Does a SWIZZLE_128B K-major A descriptor tolerate a start address that is not a multiple of 8 rows?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib
h = _lib.get_handle(0)
g = torch.Generator().manual_seed(0)
for mode in (0, 2):
    for shift in (0, 1, 2, 7, 8, 9, 15, 16, 17, 30, 31, 32):
        a = torch.randn(128 + shift, 128, generator=g)
        b = torch.randn(64, 128, generator=g)
        try:
            d = h.selftest_umma(a, b, shift, mode)
        except Exception as e:
            print("mode", mode, "shift", shift, "ERR", e); continue
        ref = (a[shift:shift + 128].bfloat16().double() @ b.bfloat16().double().t())
        err = (d.double() - ref).abs().max().item() / ref.abs().max().item()
        print("mode", mode, "shift", shift, "rel err %.2e" % err, "OK" if err < 2e-3 else "BAD")
