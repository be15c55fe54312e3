"""The image pyramid pinned against the REFERENCE'S OWN code.

oracle/build_ref_pba.py compiles photometrically_corrected_image.cpp, pixel_data_frame.cpp (with downscale_image.hpp),
pixel_map.cpp and calculate_pixelinfo.cpp from /root/reference; tests/golden/ref_pyramid.npz holds what they return on the
frames of tests/ref_pyramid_cases.py (tools/make_ref_pyramid_golden.py).  The reference build computes in double, so:

  * oracle/features_oracle.py run in float64 must equal the reference BIT FOR BIT (a table look-up, one product by
    max / (v + 1), sums of four, halves and differences: no reordering is possible);
  * the same oracle code run in float32 is what the device is held to bit for bit (tests/test_image_preparation.py); the
    device is also compared with the reference's numbers directly, at float32 rounding of a four-step chain (4 ulp of the
    largest value).
"""
import os

import numpy as np
import pytest

import ref_pyramid_cases as PC
from oracle import features_oracle as F
from oracle import ref_pba

GOLDEN_PATH = os.path.join(os.path.dirname(__file__), "golden", "ref_pyramid.npz")
needs_ref = pytest.mark.skipif(not ref_pba.available(), reason="neither /root/reference nor oracle/_ref is present")


@pytest.fixture(scope="module")
def golden():
    g = np.load(GOLDEN_PATH)
    return {k: g[k] for k in g.files}


@pytest.mark.parametrize("name", PC.CASES)
def test_oracle_equals_the_reference_golden_bit_for_bit(golden, name):
    gray, lut, vign, levels = PC.make(name)
    pyr = F.pixel_data_frame(gray, lut.astype(np.float64), vign, levels, dtype=np.float64)
    n_ref = sum(1 for k in golden if k.startswith(name + "::level"))
    assert len(pyr) == n_ref == min(levels, 5)  # kMaxPyramidDepth
    for l, a in enumerate(pyr):
        assert np.array_equal(a, golden[f"{name}::level{l}"]), (name, l)
    corrected = F.photometrically_corrected_image(gray, lut.astype(np.float64), vign, np.float64)
    assert np.array_equal(corrected, golden[f"{name}::corrected"])
    assert np.array_equal(F.downscale_image(corrected), golden[f"{name}::half"])
    if name == "lut_vignette":  # the black vignette pixel: lut * max / (0 + 1)
        assert corrected[0, 0] == np.float64(lut[gray[0, 0]]) * (np.float64(vign.max()) / 1.0)


def _float32_bounds(got, ref, l):
    ulp = np.spacing(np.float32(np.abs(ref[..., 0]).max()))
    assert np.abs(got[..., 0].astype(np.float64) - ref[..., 0]).max() <= (1 + l) * ulp, (l, "I")
    assert np.abs(got[..., 1:].astype(np.float64) - ref[..., 1:]).max() <= 2 * (1 + l) * ulp, (l, "dx dy")


@pytest.mark.parametrize("name", PC.CASES)
def test_float32_oracle_within_rounding_of_the_reference(golden, name):
    """The float32 run of the same oracle code -- the one the device is bit-identical to -- stays within (1 + level) ulp of
    the level's largest intensity (twice that for the gradients) of the reference's double pyramid."""
    gray, lut, vign, levels = PC.make(name)
    for l, a in enumerate(F.pixel_data_frame(gray, lut, vign, levels, dtype=np.float32)):
        _float32_bounds(a, golden[f"{name}::level{l}"], l)


@needs_ref
def test_golden_is_what_the_reference_computes(golden):
    for name in PC.CASES:
        gray, lut, vign, levels = PC.make(name)
        for l, a in enumerate(ref_pba.pixel_data_frame(gray, lut, vign, levels)):
            assert np.array_equal(a, golden[f"{name}::level{l}"])


@needs_ref
def test_oracle_equals_the_reference_live_on_other_frames():
    rng = np.random.default_rng(5)
    for h, w, levels in ((48, 64, 4), (120, 160, 3), (30, 48, 2)):  # widths % 8 == 0 on every level (ref_pyramid_cases.py)
        gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
        lut = np.sort(rng.uniform(0, 255, 256))
        vign = rng.integers(0, 256, (h, w), dtype=np.uint8)
        for v in (None, vign):
            got = F.pixel_data_frame(gray, lut, v, levels, dtype=np.float64)
            ref = ref_pba.pixel_data_frame(gray, lut, v, levels)
            assert all(np.array_equal(a, b) for a, b in zip(got, ref)), (h, w, v is None)
        im = rng.uniform(0, 255, (h, w))
        assert np.array_equal(F.downscale_image(im), ref_pba.downscale(im))


@needs_ref
def test_reference_avx2_dispatch_quirk():
    """Outside its domain the reference's double build does not compute the definition it tests
    (test/test/features/test_dxdy_accelerated.cpp:43-80 holds the AVX2 routine to calculate_pixelinfo_c): on a 12-pixel-wide
    level the comma in calculate_pixelinfo.cpp:388 routes the image to the AVX2 routine, which writes the first group of 8
    columns, closes it with the right-border formula, and never writes columns 8..11.  The oracle and the device follow the scalar definition everywhere; this
    test records that the difference is the reference's, column by column."""
    rng = np.random.default_rng(6)
    gray = rng.integers(0, 256, (8, 12), dtype=np.uint8)
    lut = np.arange(256, dtype=np.float64)
    ref = ref_pba.pixel_data_frame(gray, lut, None, 1)[0]
    want = F.pixel_info(lut[gray])
    assert np.array_equal(ref[:, :7], want[:, :7])          # whole columns of the first group, bar its right edge
    assert np.array_equal(ref[:, 7, 0], want[:, 7, 0]) and np.array_equal(ref[:, 7, 2], want[:, 7, 2])
    img = lut[gray]
    assert np.array_equal(ref[:, 7, 1], img[:, 7] - img[:, 6])  # the group's last column takes the image-border formula for dx
    assert not np.array_equal(ref[:, 7, 1], want[:, 7, 1])
    # columns 8..11 of `ref` are whatever the allocation held (never written): nothing to assert on them


@pytest.mark.gpu
@pytest.mark.parametrize("name", PC.CASES)
def test_device_pyramid_against_the_reference_golden(golden, name):
    """dpba_build_pyramid (float32 on the device) against the reference's double-precision pyramid."""
    from dsopp_b200 import capi
    gray, lut, vign, levels = PC.make(name)
    levels = min(levels, 5)
    h = capi.Handle(2, 16, PC.W, PC.H)
    got = h.build_pyramid(gray, lut, vign, levels=levels)
    assert len(got) == levels
    for l, g in enumerate(got):
        ref = golden[f"{name}::level{l}"]
        assert g.shape == ref.shape
        _float32_bounds(g, ref, l)
    h.close()
