"""CPU oracle (NumPy) of the reference's image preparation.  TEST INFRASTRUCTURE ONLY.

Restates, paths relative to /root/reference/src/features/:
  photometricallyCorrectedImage   src/photometrically_corrected_image.cpp:9-29
  downscaleImage                  internal/features/camera/downscale_image.hpp:16-33
  PixelDataFrame (pyramid)        src/pixel_data_frame.cpp:12-31
  {I,dx,dy} packing               src/calculate_pixelinfo.cpp:340-374 (scalar definition; the AVX2 path is tested equal to
                                  it by test/test/features/test_dxdy_accelerated.cpp:43-80)
PARITY PINNED, all four: the reference's own sources compile here (oracle/build_ref_pba.py: photometrically_corrected_image
.cpp, pixel_data_frame.cpp with downscale_image.hpp, pixel_map.cpp, calculate_pixelinfo.cpp, against the stand-in Eigen /
OpenCV headers of oracle/ref_stubs_full) and these functions, run in float64 like the reference's build, equal what they
return BIT FOR BIT on whole pyramids (tests/test_reference_pyramid.py, tests/golden/ref_pyramid.npz); pixel_info is held
in addition to the AVX2 and the plain-C routine in double and float (tests/test_reference_parts.py, tests/golden/
ref_parts.npz).  Domain: level widths that are multiples of 8 -- on other widths the reference's double build leaves columns
unwritten (the comma in the dispatch at calculate_pixelinfo.cpp:388; test_reference_avx2_dispatch_quirk) and the scalar
definition restated here is what its own test names as the truth (test/test/features/test_dxdy_accelerated.cpp:43-80).
All arithmetic is exact in the given dtype (float32 = the reference's USE_FLOAT build), so the device result must be
bit-identical: the only operations are a table look-up, one multiply by max / (v + 1), sums of four and halves.
"""
import numpy as np


def photometrically_corrected_image(gray_u8, lut, vignetting_u8=None, dtype=np.float32):
    lut = np.asarray(lut, dtype=dtype)
    out = lut[np.asarray(gray_u8, dtype=np.uint8)]
    if vignetting_u8 is not None:
        v = np.asarray(vignetting_u8)
        max_v = dtype(np.float64(v.max()))  # cv::minMaxLoc returns a double; the product is taken in Precision
        out = out * (max_v / (v.astype(dtype) + dtype(1)))
    return out.astype(dtype)


def downscale_image(I):
    """0.25 * (A + B + C + D) with A = (even, even), B = (odd, odd), C = (even, odd), D = (odd, even), summed in that
    order (an Eigen expression evaluates left to right)."""
    q = I.dtype.type(0.25)
    H2, W2 = I.shape[0] // 2, I.shape[1] // 2
    a = I[0:2 * H2:2, 0:2 * W2:2]
    b = I[1:2 * H2:2, 1:2 * W2:2]
    c = I[0:2 * H2:2, 1:2 * W2:2]
    d = I[1:2 * H2:2, 0:2 * W2:2]
    return q * (((a + b) + c) + d)


def pixel_info(I):
    I = np.asarray(I)
    H, W = I.shape
    out = np.empty((H, W, 3), dtype=I.dtype)
    half = I.dtype.type(0.5)
    out[..., 0] = I
    out[:, 1:-1, 1] = half * (I[:, 2:] - I[:, :-2])
    out[:, 0, 1] = I[:, 1] - I[:, 0]
    out[:, -1, 1] = I[:, -1] - I[:, -2]
    out[1:-1, :, 2] = half * (I[2:, :] - I[:-2, :])
    out[0, :, 2] = I[1, :] - I[0, :]
    out[-1, :, 2] = I[-1, :] - I[-2, :]
    return out


def pixel_data_frame(gray_u8, lut, vignetting_u8, levels, dtype=np.float32):
    """PixelDataFrame: list of {I,dx,dy} maps, level 0 first (levels capped at kMaxPyramidDepth = 5)."""
    levels = min(levels, 5)
    I = photometrically_corrected_image(gray_u8, lut, vignetting_u8, dtype)
    out = [pixel_info(I)]
    for _ in range(1, levels):
        I = downscale_image(I)
        out.append(pixel_info(I))
    return out
