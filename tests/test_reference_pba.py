"""The oracle pinned against the REFERENCE'S OWN photometric bundle adjustment.

oracle/build_ref_pba.py compiles the reference's sources where they lie under /root/reference -- LocalFrame, PixelMap, the
pinhole ArrayReprojector, evaluateJacobians, firstEstimateJacobians_, the Hessian block evaluation, the Problem class with
its priors, updateMarginalizedLinearSystem, the LM driver, NormalLinearSystem::solve / reduce_system (SURVEY 8a rows a1-a18,
a20) -- against minimal stand-ins of the absent third-party libraries, into oracle/_ref/libdsopp_ref_pba.so.  Here every
array those functions leave behind on the windows of tests/ref_pba_cases.py (out-of-bounds, masked, invalid-depth,
non-OK-status, marginalised landmarks, pending steps, Huber weights, FEJ on / off, forced and regular LM, marginalisation of
the oldest keyframe followed by a solve on the prior) is compared with what oracle/pba_oracle.py computes:

  * statuses, status candidates, jacobian-valid flags, ill-conditioned flags, landmark flags: EXACT;
  * everything else: 1e-9 of the array's largest magnitude (both sides are double; the stand-in linear algebra and NumPy
    differ in summation order only).

Each check runs against the committed golden vectors made from the reference (tests/golden/ref_pba.npz,
tools/make_ref_pba_golden.py) -- always -- and against the library itself when it can be built or was shipped.  The same
golden vectors bound the C++ restatement (oracle/cpu_ref, double build) and, in tests/test_gpu_reference_golden.py, the CUDA
path itself.
"""
import os

import numpy as np
import pytest

import ref_pba_cases as RC
from oracle import ref_pba

GOLDEN_PATH = os.path.join(os.path.dirname(__file__), "golden", "ref_pba.npz")
needs_ref = pytest.mark.skipif(not ref_pba.available(), reason="neither /root/reference nor oracle/_ref is present")
RTOL = 1e-9


@pytest.fixture(scope="module")
def golden():
    g = np.load(GOLDEN_PATH)
    out = {}
    for key in g.files:
        run, k = key.split("::", 1)
        out.setdefault(run, {})[k] = g[key]
    return out


def test_golden_covers_every_run(golden):
    assert set(golden) == set(RC.RUNS) | set(RC.PA_CASES)
    golden = {k: v for k, v in golden.items() if k in RC.RUNS}
    for name, arrays in golden.items():
        leaves = {k.rsplit("/", 1)[-1] for k in arrays}
        assert {"status", "cand", "r", "w", "e", "idepth", "state_eps"} & leaves, name


@pytest.mark.parametrize("name", list(RC.RUNS))
def test_oracle_reproduces_the_reference_golden(golden, name):
    case, seq, kw = RC.RUNS[name]
    got = RC.run(RC.OracleBackend, case, seq, **kw)
    ref = golden[name]
    bad = RC.compare(ref, {k: got[k] for k in ref if k in got}, RTOL)
    assert not bad, bad[:10]
    # the cases must actually reach the branches they were built for
    if name == "edge_lin_fej":
        cand = np.concatenate([v for k, v in ref.items() if k.startswith("lin/res") and k.endswith("/cand")])
        hist = np.bincount(cand, minlength=5)
        assert hist[RC.O.K_OOB] > 100 and hist[RC.O.K_OUTLIER] > 10 and hist[RC.O.K_UNKNOWN] > 10, hist
        w = np.concatenate([v for k, v in ref.items() if k.startswith("lin/res") and k.endswith("/w")])
        assert (w < 1).sum() > 100
        ill = np.concatenate([v for k, v in ref.items() if k.startswith("schur/lm") and k.endswith("/ill")])
        assert ill.sum() > 10
    if name == "marginalise":
        assert int(ref["marg/n_frames"][0]) == 3 and ref["marg/H"].shape == (24, 24) and ref["marg/energy"][0] > 0


@needs_ref
@pytest.mark.parametrize("name", list(RC.RUNS))
def test_oracle_reproduces_the_reference_live(name):
    """Every array, every pair -- not only the cross-section the golden file keeps."""
    case, seq, kw = RC.RUNS[name]
    ref = RC.run(RC.ReferenceBackend, case, seq, **kw)
    got = RC.run(RC.OracleBackend, case, seq, **kw)
    bad = RC.compare(ref, got, RTOL)
    assert not bad, bad[:10]
    assert len(ref) >= 100


@needs_ref
def test_golden_is_what_the_reference_computes(golden):
    for name in ("edge_lin_fej", "plain_solve_fej", "marginalise"):
        case, seq, kw = RC.RUNS[name]
        ref = RC.run(RC.ReferenceBackend, case, seq, **kw)
        bad = RC.compare(golden[name], {k: ref[k] for k in golden[name]}, 1e-12)
        assert not bad, (name, bad[:10])


@needs_ref
def test_other_seeds_live():
    """The pin does not hang on the seeds the golden file was made with."""
    import functools
    for seed in (21, 22):
        for case_fn, seq, kw in ((RC.case_edge, RC.seq_linearize, dict(fej=True)), (RC.case_plain, RC.seq_solve, dict(fej=True)),
                                 (RC.case_edge, RC.seq_solve, dict(fej=False))):
            outs = []
            for backend in (RC.ReferenceBackend, RC.OracleBackend):
                win, raws, extra = functools.partial(case_fn, seed=seed)()
                outs.append(seq(backend(win, raws, extra), **kw))
            bad = RC.compare(outs[0], outs[1], RTOL)
            assert not bad, (seed, case_fn.__name__, bad[:10])


@needs_ref
def test_normal_linear_system_solve_live():
    """NormalLinearSystem<>::solve (normal_linear_system.cpp:52-60) on systems of the size and conditioning of a window."""
    rng = np.random.default_rng(0)
    for n in (8, 24, 64):
        A = rng.normal(size=(n + 5, n)) * np.logspace(0, 3, n)[None, :]
        H = A.T @ A + np.diag(rng.uniform(1, 10, n))
        b = rng.normal(size=n) * 1e3
        x_ref = ref_pba.normal_solve(H, b)
        x = RC.O.normal_solve(H, b)
        assert np.abs(x - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
        assert np.abs(H @ x_ref - b).max() <= 1e-8 * np.abs(b).max()


@needs_ref
def test_landmark_patch_sampling_live():
    """PatternPatch::getIntensities over the reference's PixelMap (pattern_patch.hpp:52-64, pixel_map.hpp:25-45): the 8-pixel
    patch a tracker stores with a landmark, at integer and at fractional positions."""
    win, raws, _ = RC.case_plain()
    rw = ref_pba.window_from_synth(win, raws)
    rng = np.random.default_rng(1)
    uv = np.concatenate([win.frames[0].uv[:20], rng.uniform(8, 100, (20, 2))])
    got = rw.get_intensities(0, uv)
    img = RC.FO.pixel_info(raws[0])
    pts = uv[:, None, :] + RC.O.PATTERN[None, :, :]
    want = RC.O.interpolate_linear(img, pts[..., 0], pts[..., 1])[..., 0]
    assert np.abs(got - want).max() <= 1e-12 * 255
    # the synthetic generator's integer-position patches are exactly these samples
    assert np.abs(got[:20] - win.frames[0].patch[:20]).max() <= 1e-4  # float32 image in the generator


def _pa_compare(ref, got, subsample):
    assert np.array_equal(np.asarray(ref["pa/uv"], np.int64), np.asarray(got["pa/uv"], np.int64))  # same landmarks, same order
    gi, gp = (got["pa/idepth"][::7], got["pa/patch"][::7]) if subsample else (got["pa/idepth"], got["pa/patch"])
    assert np.array_equal(ref["pa/idepth"], gi) and np.array_equal(ref["pa/patch"], gp)  # idepth / weight and I(x, y): exact
    assert ref["pa/result"][1] == got["pa/result"][1] and ref["pa/result"][2] == got["pa/result"][2]
    assert abs(ref["pa/result"][0] - got["pa/result"][0]) <= RTOL * ref["pa/result"][0]
    assert np.abs(ref["pa/T_t_r"] - got["pa/T_t_r"]).max() <= RTOL
    assert np.abs(ref["pa/ab_eps"] - got["pa/ab_eps"]).max() <= RTOL * max(1.0, np.abs(ref["pa/ab_eps"]).max())
    assert np.abs(ref["pa/H"] - got["pa/H"]).max() <= RTOL * np.abs(ref["pa/H"]).max()


@pytest.mark.parametrize("name", list(RC.PA_CASES))
def test_pose_alignment_oracle_reproduces_the_reference_golden(golden, name):
    """The coarse tracker's image alignment: the reference's depth-map LocalFrame constructor
    (local_frame.hpp:350-393) and its class PoseAlignerProblem (eigen_pose_alignment.cpp:24-242) under its LM driver, set up
    as EigenPoseAlignment::solve does, against oracle/pose_alignment_oracle.py: the landmark list exactly, energy / pose /
    affine increment / Hessian at 1e-9."""
    _pa_compare(golden[name], RC.pa_run_oracle(name), subsample=True)


@needs_ref
@pytest.mark.parametrize("name", list(RC.PA_CASES))
def test_pose_alignment_oracle_reproduces_the_reference_live(name):
    _pa_compare(RC.pa_run_reference(name), RC.pa_run_oracle(name), subsample=False)
