"""A/B of the {I,dx,dy} packing kernel: direct stencil (k_pixelinfo) vs TMA-staged tile (k_pixelinfo_tma), per pyramid level.
python tools/pixelinfo_ab.py   (on the GPU box)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsopp_b200 import capi

lib = capi.load_library()
print("| plane | direct stencil (us / launch) | TMA-staged tile (us / launch) | outputs |")
print("|---|---:|---:|---|")
for W, H in ((640, 480), (320, 240), (160, 120), (80, 60), (1280, 960), (1920, 1080)):
    ms = (C.c_double * 2)()
    bad = C.c_int64(-1)
    rc = lib.dpba_debug_pixelinfo_ab(W, H, 200, ms, C.addressof(bad))
    print(f"| {W}x{H} | {ms[0] * 1e3:.2f} | {ms[1] * 1e3:.2f} | {'bit-identical' if bad.value == 0 else str(bad.value) + ' words differ'} (rc {rc}) |")
