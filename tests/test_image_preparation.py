"""photometric correction, 2x2 box pyramid and {I,dx,dy} packing.

CPU: the oracle restatement (oracle/features_oracle.py) agrees with the generator's helpers and with the exactness
property the reference tests (test/test/features/test_dxdy_accelerated.cpp:43-80: vectorised == scalar definition).
GPU: the device build is BIT-IDENTICAL to the oracle in float32 (the reference's USE_FLOAT build), and a frame pushed raw
gives the same residuals as the same frame pushed as {I,dx,dy}."""
import numpy as np
import pytest

from dsopp_b200 import synth
from oracle import features_oracle as F


def raw_case(seed=0, W=160, H=120):
    rng = np.random.default_rng(seed)
    gray = rng.integers(0, 256, (H, W), dtype=np.uint8)
    lut = (np.arange(256, dtype=np.float32) ** 1.1 * 0.6).astype(np.float32)  # a monotone response curve
    yy, xx = np.mgrid[0:H, 0:W]
    vign = (255 - 90 * ((xx - W / 2) ** 2 + (yy - H / 2) ** 2) / ((W / 2) ** 2 + (H / 2) ** 2)).astype(np.uint8)
    return gray, lut, vign


def test_oracle_matches_definitions():
    gray, lut, vign = raw_case()
    I = F.photometrically_corrected_image(gray, lut, vign)
    y, x = 17, 33
    assert I[y, x] == np.float32(lut[gray[y, x]]) * (np.float32(vign.max()) / (np.float32(vign[y, x]) + np.float32(1)))
    assert (F.photometrically_corrected_image(gray, lut, None) == lut[gray]).all()
    d = F.downscale_image(I)
    assert d.shape == (60, 80)
    assert d[5, 7] == np.float32(0.25) * (((I[10, 14] + I[11, 15]) + I[10, 15]) + I[11, 14])
    # same definitions as the synthetic generator uses (up to the summation order of the box filter)
    assert (F.pixel_info(I) == synth.pixelinfo(I)).all()
    assert np.allclose(d, synth.downscale(I), rtol=1e-6)
    pyr = F.pixel_data_frame(gray, lut, vign, 7)
    assert len(pyr) == 5 and pyr[4].shape == (7, 10, 3)  # kMaxPyramidDepth


@pytest.mark.gpu
@pytest.mark.parametrize("with_lut,with_vign", [(True, True), (True, False), (False, False)])
def test_device_pyramid_is_bit_identical(with_lut, with_vign):
    from dsopp_b200 import capi
    gray, lut, vign = raw_case(seed=1)
    h = capi.Handle(2, 16, 160, 120)
    got = h.build_pyramid(gray, lut if with_lut else None, vign if with_vign else None, levels=4)
    want = F.pixel_data_frame(gray, lut if with_lut else np.arange(256, dtype=np.float32), vign if with_vign else None, 4)
    for l, (g, wv) in enumerate(zip(got, want)):
        assert g.shape == wv.shape
        assert (g == wv).all(), (l, np.abs(g - wv).max())
    h.close()


@pytest.mark.gpu
def test_raw_push_equals_float_push():
    from dsopp_b200 import capi
    win = synth.make_window(n_frames=3, points_per_frame=120, seed=4, ab_scale=0.0)
    # 8-bit versions of the rendered frames; the float path gets the oracle's preparation of the same bytes
    lut = np.arange(256, dtype=np.float32) * np.float32(1.0)
    grays = [np.clip(np.rint(f.image[..., 0]), 0, 255).astype(np.uint8) for f in win.frames]
    a = capi.Handle(3, 120, win.width, win.height)
    b = capi.Handle(3, 120, win.width, win.height)
    for f, g in zip(win.frames, grays):
        img = F.pixel_info(F.photometrically_corrected_image(g, lut, None))
        a.push_frame(f.frame_id, img, f.mask, f.T_w_lin, f.exposure, f.ab0, f.intr, f.fixed)
        b.push_frame_raw(f.frame_id, g, lut, None, f.mask, f.T_w_lin, f.exposure, f.ab0, f.intr, f.fixed)
    for h in (a, b):
        for i, f in enumerate(win.frames):
            h.set_landmarks(i, f.uv, f.idepth, f.patch, f.flags)
        h.set_state(np.concatenate([f.state_eps for f in win.frames]), np.zeros(24))
        h.first_estimate()
    ea, na = a.evaluate(20.0, True, True)
    eb, nb = b.evaluate(20.0, True, True)
    assert (ea, na) == (eb, nb)  # identical images on the device -> identical sweep
    a.close(), b.close()
