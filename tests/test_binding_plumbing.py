"""The ctypes binding (dsopp_b200/capi.py) against a stand-in library that records its inputs and fills its outputs with
patterns (tests/emu/fake_capi.c): every array must reach the parameter it belongs to, 2-D outputs must get the right row
pointers, skipped entries must be NULL.  Runs in a subprocess: the stand-in is never loaded next to the real library.
No compute is involved -- the real library has no CPU path and keeps none."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import ctypes as C, os, subprocess, sys
    import numpy as np
    sys.path.insert(0, ROOT)
    from dsopp_b200 import capi, synth

    out = os.path.join(ROOT, "tests", "emu", "_build")
    os.makedirs(out, exist_ok=True)
    src = open(os.path.join(ROOT, "tests", "emu", "fake_capi.c")).read()
    stubs = ["/* generated: every other entry point of the header returns 0 */"]
    for name in capi.SIGNATURES:
        if (name + "(") not in src:
            stubs.append("int %s(void* h, ...) { return 0; }" % name)
    gen = os.path.join(out, "fake_capi_stubs.c")
    open(gen, "w").write("\\n".join(stubs) + "\\n")
    lib_path = os.path.join(out, "libfake_capi.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", lib_path, os.path.join(ROOT, "tests", "emu", "fake_capi.c"), gen])
    lib = capi.load_library(lib_path)
    lib.fake_sum.restype = C.c_double
    lib.fake_sum.argtypes = [C.c_int, C.c_int, C.c_int]

    win = synth.make_window(n_frames=4, points_per_frame=37, width=64, height=48, seed=2)
    rng = np.random.default_rng(0)
    for f in win.frames:
        f.flags[:] = rng.integers(0, 4, len(f.flags))
    for k in win.statuses:
        win.statuses[k][:] = rng.integers(0, 5, len(win.statuses[k]))
    h = capi.upload_window(win)            # push_frame, set_landmarks, set_statuses / set_frame_statuses, set_state
    assert h.n_frames == 4
    for s, f in enumerate(win.frames):
        img = np.asarray(f.image, np.float32)
        expect = (float(img.astype(np.float64).sum()) + f.mask.flat[0] + f.exposure + f.ab0[0] + 2 * f.ab0[1] + f.intr[0]
                  + f.intr[3] + 100 * int(f.fixed) + 1000 * f.frame_id)
        assert abs(lib.fake_sum(0, s, 0) - expect) < 1e-6 * abs(expect), ("image / scalars", s)
        T = np.asarray(f.T_w_lin)[:3, :4].reshape(12)
        assert abs(lib.fake_sum(1, s, 0) - float((np.arange(1, 13) * T).sum())) < 1e-9, ("pose", s)
        assert h.num_landmarks(s) == len(f.idepth)
        assert abs(lib.fake_sum(2, s, 0) - float(np.asarray(f.uv, np.float32).astype(np.float64).sum())) < 1e-6
        assert abs(lib.fake_sum(3, s, 0) - float(np.asarray(f.idepth, np.float32).astype(np.float64).sum())) < 1e-9
        assert abs(lib.fake_sum(4, s, 0) - float(np.asarray(f.patch, np.float32).astype(np.float64).sum())) < 1e-3
        assert lib.fake_sum(5, s, 0) == float(f.flags.sum())

    # set_frame_statuses: dict and list forms, the reference frame's own row and missing targets are NULL
    for r in range(4):
        h.set_frame_statuses(r, {t: win.statuses[(r, t)] for t in range(4) if t != r})
        for t in range(4):
            want = -1.0 if t == r else float(win.statuses[(r, t)].sum())
            assert lib.fake_sum(6, r, t) == want, ("statuses dict", r, t)
    h.set_frame_statuses(1, [win.statuses[(1, 0)], None, None, win.statuses[(1, 3)]])
    assert [lib.fake_sum(6, 1, t) for t in range(4)] == [float(win.statuses[(1, 0)].sum()), -1.0, -1.0,
                                                         float(win.statuses[(1, 3)].sum())]

    # get_frame_statuses: row t of the result is what the library wrote through pointer t
    for r in range(4):
        st, cd = h.get_frame_statuses(r)
        n = h.num_landmarks(r)
        assert st.shape == (4, n) and cd.shape == (4, n)
        for t in range(4):
            if t == r:
                assert not st[t].any() and not cd[t].any()
            else:
                assert np.array_equal(st[t], (10 * t + np.arange(n) % 7).astype(np.uint8)), (r, t)
                assert np.all(cd[t] == 100 + t)

    # get_landmarks: seven outputs, each in its own parameter
    lm = h.get_landmarks(2)
    n = h.num_landmarks(2)
    i = np.arange(n)
    for key, base in (("idepth", 1), ("idepth_step", 2), ("inv_hdd", 3), ("b_d", 4), ("rel_baseline", 7)):
        assert np.array_equal(lm[key], (base + i).astype(np.float32)), key
    assert np.array_equal(lm["flags"], (i % 5).astype(np.uint8)) and np.array_equal(lm["n_inliers"], (6 + i).astype(np.uint32))

    # state round trip
    eps = rng.normal(size=32)
    step = rng.normal(size=32)
    h.set_state(eps, step)
    e2, s2 = h.get_state()
    assert np.array_equal(e2, eps) and np.array_equal(s2, step)
    h.set_state(eps * 2, None)             # NULL keeps the other half
    e3, s3 = h.get_state()
    assert np.array_equal(e3, eps * 2) and np.array_equal(s3, step)
    # every other wrapper: the argument conversion must go through (a pointer handed to the wrong kind of parameter raises
    # ctypes.ArgumentError here instead of on the GPU box)
    f0 = win.frames[0]
    n0 = h.num_landmarks(0)
    h.set_frame_linearization(0, f0.T_w_lin, f0.ab0)
    h.set_frame_flags(0, True, False)
    h.set_landmarks(0, f0.uv, f0.idepth, f0.patch, None)
    h.set_landmarks(0, f0.uv[:3], f0.idepth[:3], f0.patch[:3], f0.flags[:3], append=True)
    h.set_landmarks(0, f0.uv, f0.idepth, f0.patch, f0.flags)
    h.set_landmark_flags(0, f0.flags)
    assert h.get_pose_idepth_blocks(0).shape == (n0, 8 * 4)
    h.set_statuses(0, 1, win.statuses[(0, 1)])
    st, cd = h.get_statuses(0, 1)
    assert st.shape == (n0,) and cd.shape == (n0,)
    h.first_estimate()
    assert h.evaluate(20.0) == (0.0, 0)
    h.evaluate_jacobians(20.0)
    blk = h.download_residual_block(0, 1)
    assert blk["J_ref"].shape == (n0, 8, 8)
    for mat in (False, True):
        Hp, bp, Hs, bs = h.linearize(20.0, True, True, False, materialized=mat)
        assert Hp.shape == (32, 32) and bs.shape == (32,)
    h.back_substitute(np.zeros(32), 1e-5)
    h.accept(), h.reject(), h.change_residual_statuses(True), h.landmarks_energy(), h.update_point_statuses(1, 20.0)
    out, act, nv = h.refine_immature_landmarks(0, f0.uv[:5], f0.idepth[:5], f0.patch[:5, 0], 2)
    assert out.shape == (5,) and act.dtype == bool and nv.shape == (5,)
    maps = h.create_reference_depth_maps(3, 1e-5)
    assert [m[0].shape for m in maps] == [(48, 64), (24, 32), (12, 16)]
    h.solve_lm(20.0)
    h.solve_lm(20.0, H_marg=np.eye(32), b_marg=np.zeros(32), energy_marg=1.0)
    h.set_option("cuda_graph", 1)
    h.profile_enable(True)
    assert set(h.profile_read()) == set(h.PROFILE_KINDS)
    h.comm_init(bytes(128), 0, 1)
    assert len(h.peer_export()) == 64
    h.peer_attach(bytes(128), 0, 2)
    gray = np.zeros((48, 64), np.uint8)
    h.push_frame_raw(9, gray, np.arange(256, dtype=np.float32), None, f0.mask, f0.T_w_lin, 1.0, f0.ab0, f0.intr, False)
    assert [o.shape for o in h.build_pyramid(gray, np.arange(256, dtype=np.float32), gray, 3)] == [(48, 64, 3), (24, 32, 3), (12, 16, 3)]
    h.push_frame(10, np.zeros((48, 64), np.float32), f0.mask, f0.T_w_lin, 1.0, f0.ab0, f0.intr, False)   # intensity-only form
    while h.n_frames:
        h.remove_frame(0)
    assert len(capi.comm_unique_id()) == 128 and capi.launch_count() == 0
    print("PLUMBING OK")
''')


def test_binding_passes_every_array_to_its_parameter():
    run = subprocess.run([sys.executable, "-c", "ROOT = %r\n" % ROOT + SCRIPT], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "PLUMBING OK" in run.stdout, run.stdout[-2000:] + run.stderr[-3000:]
