"""Golden vectors from the REFERENCE'S OWN image pyramid (oracle/build_ref_pba.py: photometrically_corrected_image.cpp,
downscale_image.hpp, pixel_data_frame.cpp, pixel_map.cpp, calculate_pixelinfo.cpp compiled from /root/reference)
-> tests/golden/ref_pyramid.npz.  Run in the build container; the GPU box uses the committed file.

    python tools/make_ref_pyramid_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_pyramid_cases as PC  # noqa: E402
from oracle import build_ref_pba, ref_pba  # noqa: E402


def main():
    assert build_ref_pba.have_reference(), "needs /root/reference"
    out = {}
    for name in PC.CASES:
        gray, lut, vign, levels = PC.make(name)
        pyr = ref_pba.pixel_data_frame(gray, lut, vign, levels)
        for l, a in enumerate(pyr):
            out[f"{name}::level{l}"] = a
        out[f"{name}::corrected"] = ref_pba.photometric_correction(gray, lut, vign)
        out[f"{name}::half"] = ref_pba.downscale(out[f"{name}::corrected"])
        print(name, [a.shape for a in pyr])
    path = os.path.join(ROOT, "tests", "golden", "ref_pyramid.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
