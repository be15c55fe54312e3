"""Raw frames for the pyramid pin (tests/test_reference_pyramid.py, tools/make_ref_pyramid_golden.py): an 8-bit image, a
response table, a vignette, the number of levels asked for."""
import numpy as np

# Every level's width is a multiple of 8: the domain of the reference's double build.  Its dispatch (calculate_pixelinfo.cpp:
# 388, `width % 8 == 0 && is_aligned(input, 32), is_aligned(output, 32)` -- a comma, not an &&) sends EVERY aligned image to
# the AVX2 routine, which only walks whole groups of 8 columns; on other widths it leaves the remaining columns unwritten
# (tests/test_reference_pyramid.py::test_reference_avx2_dispatch_quirk).  Real frames (640 x 480, 1280 x 720, 5 levels) stay
# inside the domain.
W, H = 128, 48
CASES = ("lut_vignette", "lut_only", "identity_seven_levels")


def make(name):
    rng = np.random.default_rng({"lut_vignette": 11, "lut_only": 12, "identity_seven_levels": 13}[name])
    yy, xx = np.mgrid[0:H, 0:W]
    scene = 120 + 70 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + 25 * rng.standard_normal((H, W))
    gray = np.clip(np.rint(scene), 0, 255).astype(np.uint8)
    gray[3, 5], gray[10, 20] = 0, 255  # both ends of the table
    lut = (np.arange(256, dtype=np.float32) ** 1.1 * 0.6).astype(np.float32)  # a monotone response curve, float32-exact
    vign = (255 - 140 * ((xx - W / 2) ** 2 + (yy - H / 2) ** 2) / ((W / 2) ** 2 + (H / 2) ** 2)).astype(np.uint8)
    vign[0, 0] = 0  # the reference divides by (v + 1): a black vignette pixel is legal
    if name == "lut_vignette":
        return gray, lut, vign, 4
    if name == "lut_only":
        return gray, lut, None, 3
    return gray, np.arange(256, dtype=np.float32), None, 7  # asks for more than kMaxPyramidDepth = 5
