"""The per-thread bodies of the kernels written in round 1 before they could run on a GPU, executed on the CPU (tests/emu/kernel_emu.cpp
calls the same __host__ __device__ functions once per CUDA thread, over the launcher's grid) and compared with the
oracles.  This checks indexing, predicates, rounding and buffer layout of

  dsopp_b200/csrc/depth_maps.cu       (createReferenceDepthMaps, create_depth_maps.cpp:19-146)
  dsopp_b200/csrc/energy_quantile.cu  (updatePointStatuses' nth_element, photometric_bundle_adjustment.cpp:325-361)

without a device; concurrency (atomic order, warp intrinsics) remains for the -m gpu tests.  The emulation library is
test infrastructure: nothing under dsopp_b200/ loads it.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from dsopp_b200 import synth
from oracle import depth_map_oracle as D
from oracle import pba_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "kernel_emu.cpp")
OUT = os.path.join(ROOT, "tests", "emu", "_build")
LIB = os.path.join(OUT, "libkernel_emu.so")
MAXF = 16  # PBA_MAXF


@pytest.fixture(scope="module")
def emu():
    os.makedirs(OUT, exist_ok=True)
    csrc = os.path.join(ROOT, "dsopp_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("depth_maps_body.h", "energy_quantile_body.h", "optical_flow_body.h",
                                                    "pba_internal.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                               "-I/usr/local/cuda/include", "-I", csrc, "-I", os.path.join(ROOT, "include"),
                               "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.emu_depth_maps_floats.restype = C.c_size_t
    lib.emu_depth_maps_floats.argtypes = [C.c_int] * 3
    lib.emu_reference_depth_maps.restype = None
    lib.emu_reference_depth_maps.argtypes = [C.c_int] * 4 + [C.c_void_p] * 8 + [C.c_int, C.c_float, C.c_void_p]
    lib.emu_energy_quantile.restype = C.c_uint
    lib.emu_energy_quantile.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_double, C.c_void_p]
    lib.emu_optical_flow.restype = C.c_int
    lib.emu_optical_flow.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def device_arrays(win, frames, phys, max_pts):
    """The handle's HBM layout on host arrays: landmark slot (phys[f], l), residual slot ((phys[r] * 16 + phys[t]), l)."""
    n = win.n_frames
    n_slots = max(phys) + 1
    lmk = np.zeros((n_slots * max_pts, 4), np.float32)
    flags = np.zeros(n_slots * max_pts, np.uint8)
    inv_hdd = np.zeros(n_slots * max_pts, np.float32)
    status = np.full(n_slots * MAXF * max_pts, 4, np.uint8)  # kUnknown everywhere a pair does not exist
    energy = np.zeros(n_slots * MAXF * max_pts, np.float32)
    for f, (sf, fr) in enumerate(zip(win.frames, frames)):
        m = len(fr.idepth)
        b = phys[f] * max_pts
        lmk[b:b + m, 0:2] = fr.uv
        lmk[b:b + m, 2] = fr.idepth
        lmk[b:b + m, 3] = fr.idepth
        flags[b:b + m] = sf.flags
        inv_hdd[b:b + m] = fr.inv_hdd
        for t, tg in enumerate(frames):
            if t == f:
                continue
            rb = (phys[f] * MAXF + phys[t]) * max_pts
            res = fr.residuals[tg.id]
            status[rb:rb + m] = res.status
            if res.e is not None and len(np.atleast_1d(res.e)) == m:
                energy[rb:rb + m] = res.e
    return lmk, flags, inv_hdd, status, energy


def pair_constants(frames):
    """A = reproject_, M = transform_unproject_ per ordered pair, computed in double and rounded to fp32 as k_pair_setup
    does (camera_reproject.hpp:256-258)."""
    n = len(frames)
    A = np.zeros((n, n, 12), np.float32)
    M = np.zeros((n, n, 12), np.float32)
    for r, ref in enumerate(frames):
        for t, tgt in enumerate(frames):
            if r == t:
                continue
            T = O.se3_inv(tgt.t_world_agent()) @ ref.t_world_agent()
            rp = O.Reprojector(ref, tgt, T)
            A[r, t] = rp.reproject_.reshape(12)
            M[r, t] = rp.transform_unproject_.reshape(12)
    return A, M


def split_levels(buf, W, H, n_levels):
    out, off = [], 0
    for l in range(n_levels):
        nl = (W >> l) * (H >> l)
        shape = (H >> l, W >> l)
        out.append((buf[off + 2 * nl:off + 3 * nl].reshape(shape), buf[off + 3 * nl:off + 4 * nl].reshape(shape)))
        off += 4 * nl
    return out


def compare_maps(got, ref, rtol):
    for lvl, ((gi, gw), (ri, rw)) in enumerate(zip(got, ref)):
        assert gi.shape == ri.shape
        g_on, r_on = gw > 0, rw > 0
        # fp32 reprojection vs float64: a landmark within ~1e-4 px of a rounding boundary may land on the next pixel
        assert (g_on != r_on).sum() <= 12 * (1 + (lvl == 0)), (lvl, int((g_on != r_on).sum()))
        both = g_on & r_on
        bad_w = np.abs(gw[both] - rw[both]) > rtol * np.abs(rw[both])
        bad_i = np.abs(gi[both] - ri[both]) > rtol * np.abs(ri[both])
        assert bad_w.sum() <= 12 and bad_i.sum() <= 12, (lvl, int(bad_w.sum()), int(bad_i.sum()))
    return sum(int((gw > 0).sum()) for _, gw in got)


@pytest.mark.parametrize("phys", [[0, 1, 2, 3, 4], [3, 0, 4, 1, 2]])
def test_depth_map_kernels_on_the_cpu(emu, phys):
    win = synth.make_window(n_frames=5, points_per_frame=400, seed=8, ab_scale=0.0)
    win.frames[1].flags[::7] |= synth.FLAG_OUTLIER
    win.frames[0].flags[::9] |= synth.FLAG_MARGINALIZED
    win.statuses[(2, 4)][::5] = 1  # kOutlier towards the newest keyframe
    win.frames[3].idepth[::11] = -0.01     # negative inverse depth: an outlier in the track (updateFrame), not splatted
    win.frames[3].idepth[1::11] = 3e-9     # below kIdepthEps: splatted as idepth 0
    frames = O.frames_from_window(win)
    O.first_estimate_jacobians(frames)
    O.Problem(frames, 20.0).linearize()  # inv_hdd per landmark
    max_pts = 512
    lmk, flags, inv_hdd, status, _ = device_arrays(win, frames, phys, max_pts)
    A, M = pair_constants(frames)
    n_lm = np.array([len(f.idepth) for f in frames], np.int32)
    ph = np.array(phys, np.int32)
    W, H, L = win.width, win.height, 4
    for const_var, var in ((1e-5, None), (-1.0, [f.inv_hdd for f in frames])):
        buf = np.full(emu.emu_depth_maps_floats(W, H, L), np.nan, np.float32)
        emu.emu_reference_depth_maps(5, W, H, max_pts, _p(n_lm), _p(ph), _p(lmk), _p(flags), _p(status), _p(inv_hdd),
                                     _p(A), _p(M), L, const_var, _p(buf))
        assert np.isfinite(buf).all()
        got = split_levels(buf, W, H, L)
        ref = D.create_reference_depth_maps(frames, L, var)
        assert compare_maps(got, ref, 2e-4 if var is None else 2e-3) > 10000


def test_depth_map_kernels_with_ragged_and_empty_frames(emu):
    win = synth.make_window(n_frames=3, points_per_frame=40, seed=1, ab_scale=0.0)
    f1 = win.frames[1]
    for name in ("uv", "idepth", "idepth_true", "patch", "flags"):
        setattr(f1, name, getattr(f1, name)[:0])
    win.statuses[(1, 0)] = win.statuses[(1, 0)][:0]
    win.statuses[(1, 2)] = win.statuses[(1, 2)][:0]
    frames = O.frames_from_window(win)
    lmk, flags, inv_hdd, status, _ = device_arrays(win, frames, [0, 1, 2], 64)
    A, M = pair_constants(frames)
    n_lm = np.array([len(f.idepth) for f in frames], np.int32)
    W, H = win.width, win.height
    buf = np.zeros(emu.emu_depth_maps_floats(W, H, 2), np.float32)
    emu.emu_reference_depth_maps(3, W, H, 64, _p(n_lm), _p(np.arange(3, dtype=np.int32)), _p(lmk), _p(flags), _p(status),
                                 _p(inv_hdd), _p(A), _p(M), 2, 1e-5, _p(buf))
    compare_maps(split_levels(buf, W, H, 2), D.create_reference_depth_maps(frames, 2), 2e-4)


@pytest.mark.parametrize("phys,marg_frame", [([0, 1, 2, 3], None), ([2, 0, 3, 1], 0)])
def test_energy_quantile_on_the_cpu(emu, phys, marg_frame):
    rng = np.random.default_rng(7)
    win = synth.make_window(n_frames=4, points_per_frame=257, seed=6, ab_scale=0.0)
    win.frames[2].flags[::3] |= synth.FLAG_MARGINALIZED
    frames = O.frames_from_window(win)
    for fr in frames:
        for tid, res in fr.residuals.items():
            m = len(fr.idepth)
            res.e = rng.exponential(40.0, m).astype(np.float32).astype(np.float64)
            res.status = rng.choice([0, 0, 0, 1, 3], m).astype(np.uint8)
    if marg_frame is not None:
        frames[marg_frame].is_marginalized = True
    max_pts = 300
    _, flags, _, status, energy = device_arrays(win, frames, phys, max_pts)
    n_lm = np.array([len(f.idepth) for f in frames], np.int32)
    fm = np.array([int(f.is_marginalized) for f in frames], np.int32)
    value = np.zeros(1, np.float32)
    count = emu.emu_energy_quantile(4, max_pts, _p(n_lm), _p(np.array(phys, np.int32)), _p(fm), _p(flags), _p(status),
                                    _p(energy), 0.75, _p(value))
    # the oracle's selection (update_point_statuses, first half)
    es = []
    for ref in frames:
        act = ~ref.lm_marginalized
        for tgt in frames:
            if tgt.is_marginalized or tgt.id == ref.id:
                continue
            res = ref.residuals[tgt.id]
            es.append(res.e[act & (res.status == O.K_OK)])
    es = np.concatenate(es).astype(np.float32)
    assert count == len(es) and count > 1000
    k = int(float(len(es)) * 0.75)
    assert value[0] == np.partition(es, k)[k]


def test_energy_quantile_without_any_eligible_residual(emu):
    win = synth.make_window(n_frames=3, points_per_frame=20, seed=6, ab_scale=0.0)
    for f in win.frames:
        f.flags[:] = synth.FLAG_MARGINALIZED
    frames = O.frames_from_window(win)
    _, flags, _, status, energy = device_arrays(win, frames, [0, 1, 2], 32)
    n_lm = np.array([len(f.idepth) for f in frames], np.int32)
    value = np.ones(1, np.float32)
    count = emu.emu_energy_quantile(3, 32, _p(n_lm), _p(np.arange(3, dtype=np.int32)), _p(np.zeros(3, np.int32)),
                                    _p(flags), _p(status), _p(energy), 0.75, _p(value))
    assert count == 0 and value[0] == 0.0


@pytest.mark.parametrize("motion", ["small", "no_rotation", "leaves_the_image", "behind_the_camera"])
def test_mean_square_optical_flow_on_the_cpu(emu, motion):
    """calculateMeanSquareOpticalFlow (monocular_tracker.cpp:104-133): the per-landmark body of k_optical_flow and the
    fp32 reprojection constants over the landmark list the aligner compacts from a depth map, against the oracle."""
    from oracle import pose_alignment_oracle as PA
    win = synth.make_window(n_frames=5, points_per_frame=500, seed=11, pose_noise=0.0, idepth_noise=0.0, eps_scale=0.0,
                            ab_scale=0.0)
    frames = O.frames_from_window(win)
    idw, wgt = D.create_reference_depth_maps(frames, 1)[0]
    uv, idepth, _ = PA.landmarks_from_depth_map(idw, wgt, np.zeros(wgt.shape + (1,)))
    lm = np.zeros((len(idepth), 4), np.float32)
    lm[:, :2], lm[:, 2] = uv, idepth
    xi = {"small": [0.03, -0.02, 0.05, 0.01, -0.015, 0.008], "no_rotation": [0.03, -0.02, 0.05, 0, 0, 0],
          "leaves_the_image": [1.5, 0.4, 0.0, 0.0, 0.3, 0.0], "behind_the_camera": [0, 0, -30.0, 0, 0, 0]}[motion]
    T = O.se3_exp(np.array(xi))
    intr = np.asarray(win.frames[-1].intr, np.float64)
    ref_flow, ref_n = PA.mean_square_optical_flow(idw, wgt, T, intr)
    T34 = np.ascontiguousarray(T[:3, :4]).reshape(12)
    flow = np.zeros(1)
    n = emu.emu_optical_flow(len(lm), _p(lm), _p(T34), _p(intr), win.width, win.height, _p(flow))
    assert len(lm) > 5000
    if motion == "behind_the_camera":
        assert ref_n == 0 and n == 0 and np.isnan(ref_flow) and np.isnan(flow[0])
        return
    assert abs(n - ref_n) <= 3               # landmarks within fp32 rounding of the ROI border
    assert ref_n > (100 if motion == "leaves_the_image" else 5000)
    assert abs(flow[0] - ref_flow) <= 2e-4 * ref_flow


# ---- NVLink mailbox all-reduce with real threads ---------------------------------------------------------------------
PEER_SRC = os.path.join(ROOT, "tests", "emu", "peer_emu.cpp")
PEER_LIB = os.path.join(OUT, "libpeer_emu.so")
EMU_FLAGS = ["-std=c++17", "-O1", "-g", "-pthread", "-Wall", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include",
             "-I", os.path.join(ROOT, "dsopp_b200", "csrc"), "-I", os.path.join(ROOT, "include")]


def lm_exchange_calls(n_frames=8):
    """(offset, count) of the exchanges dpba_solve_lm / dpba_linearize / dpba_evaluate issue: the whole block (system +
    scalars), the system alone, the 8 scalars alone -- different grid sizes over overlapping parts of the mailbox."""
    D = 8 * n_frames
    sysn = 2 * (D * D + D)
    calls = [(0, sysn + 8)] * 3 + [(sysn, 8), (0, sysn), (sysn, 8), (sysn, 8)] + [(0, sysn + 8)] * 4 + [(sysn, 8)]
    return np.array([c[0] for c in calls], np.int64), np.array([c[1] for c in calls], np.int64), sysn + 8


@pytest.fixture(scope="module")
def peer_emu():
    os.makedirs(OUT, exist_ok=True)
    deps = [PEER_SRC, os.path.join(ROOT, "dsopp_b200", "csrc", "peer_exchange_body.h"),
            os.path.join(ROOT, "dsopp_b200", "csrc", "pba_internal.h")]
    if not os.path.exists(PEER_LIB) or any(os.path.getmtime(d) > os.path.getmtime(PEER_LIB) for d in deps):
        subprocess.check_call(["g++"] + EMU_FLAGS + ["-fPIC", "-shared", "-o", PEER_LIB, PEER_SRC])
    lib = C.CDLL(PEER_LIB)
    lib.emu_peer_run.restype = C.c_longlong
    lib.emu_peer_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_uint, C.c_int]
    return lib


@pytest.mark.parametrize("world,n_frames", [(2, 8), (3, 5), (8, 8), (4, 16)])
def test_peer_exchange_code_with_real_threads(peer_emu, world, n_frames):
    """peer_exchange_body.h (the kernel's own steps) with one host thread per CTA and rank: every rank gets the
    rank-ordered sum of every exchange, bit for bit, whatever the interleaving and however far one rank runs ahead."""
    offs, ns, slot = lm_exchange_calls(n_frames)
    for seed in range(4):
        bad = peer_emu.emu_peer_run(world, len(offs), _p(offs), _p(ns), slot, seed, 150 if seed else 0)
        assert bad == 0, (world, seed, bad)


def test_peer_exchange_code_is_race_free_under_thread_sanitizer():
    exe = os.path.join(OUT, "peer_tsan")
    main = os.path.join(ROOT, "tests", "emu", "peer_tsan_main.cpp")
    os.makedirs(OUT, exist_ok=True)
    build = subprocess.run(["g++"] + EMU_FLAGS + ["-fsanitize=thread", "-Wno-tsan", "-o", exe, main, PEER_SRC],
                           capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime here: " + build.stderr[-200:])
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66"))
    if "FATAL: ThreadSanitizer" in run.stderr:  # e.g. unsupported address-space layout inside a sandbox
        pytest.skip(run.stderr.strip().splitlines()[0])
    assert "WARNING: ThreadSanitizer" not in run.stderr, run.stderr[-2000:]
    assert run.returncode == 0 and "wrong elements: 0" in run.stdout, (run.returncode, run.stdout, run.stderr[-500:])
