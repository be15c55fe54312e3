"""Recipe for oracle/_ref/libdsopp_ref_pba.so: the REFERENCE'S OWN photometric bundle adjustment, compiled from its sources.

Test infrastructure only.  The reference's hot path (LocalFrame, PixelMap, the pinhole
ArrayReprojector, evaluateJacobians, firstEstimateJacobians_, the Hessian block evaluation, the Problem class, the LM
driver, NormalLinearSystem) is header-heavy C++ over Eigen, Sophus, oneTBB, glog, OpenCV, Ceres and protobuf-generated
headers.  None of those libraries is in this image and there is no network, so the reference's own build cannot run.
What CAN be done -- and is done here -- is to compile the reference's sources UNCHANGED, where they lie under
/root/reference, against minimal stand-ins of the third-party interfaces (oracle/ref_stubs_full/: an eager mini-Eigen, SE3
/ SO3 with Sophus' formulas, serial tbb::parallel_for, aborting CHECKs, a plain cv::Mat, name-only Ceres / protobuf
types).  Every stand-in says "this is NOT <library>" in its first line.  The arithmetic that the restatements in oracle/
must reproduce -- which residual is evaluated when, the operation order of the Jacobians, the accumulation and
symmetrisation of the blocks, the Schur complement, the priors, the LM loop, the status bookkeeping -- is the reference's
code; the stand-ins only supply matrix products, quaternion algebra and containers.

Reference sources compiled (never copied):
  src/energy/camera_model/src/camera_model_base.cpp
  src/features/src/pixel_map.cpp, src/features/src/calculate_pixelinfo.cpp
  src/features/src/photometrically_corrected_image.cpp, src/features/src/pixel_data_frame.cpp (+ downscale_image.hpp)
  src/sensors/camera_calibration/src/camera_mask.cpp
  src/energy/problems/src/normal_linear_system.cpp
  src/energy/problems/src/eigen_pose_alignment.cpp, lines 1-242 (class PoseAlignerProblem; see POSE_ALIGNMENT_SRC below)
  src/energy/problems/src/photometric_bundle_adjustment.cpp, lines 1-413 (updatePointStatuses, relinearizeSystem; PBA_BASE_SRC)
plus every header they and oracle/ref_shims/ref_pba.cpp include (40 reference headers; listed by `--deps`).

The library is git-ignored and travels to the GPU box with the snapshot; /root/reference does not exist there, so the
tests fall back to the golden vectors made from it (tests/golden/ref_pba_*.npz, tools/make_ref_pba_golden.py).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SRC = os.path.join(REF, "src")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libdsopp_ref_pba.so")
STUBS = os.path.join(HERE, "ref_stubs_full")
SHIMS = [os.path.join(HERE, "ref_shims", "ref_pba.cpp"), os.path.join(HERE, "ref_shims", "ref_link_stubs.cpp"),
         os.path.join(HERE, "ref_shims", "ref_pose_alignment.cpp"), os.path.join(HERE, "ref_shims", "ref_stub_checks.cpp"),
         os.path.join(HERE, "ref_shims", "ref_pyramid.cpp")]
# The coarse-tracker aligner's algorithm is a class in an anonymous namespace of this file (lines 24-242); the members of
# EigenPoseAlignment that follow need the track subsystem.  The compiler is given the file's own lines up to the end of that
# namespace through a temporary copy OUTSIDE the repository, removed after the build (oracle/ref_shims/ref_pose_alignment.cpp).
POSE_ALIGNMENT_SRC = os.path.join(SRC, "energy/problems/src/eigen_pose_alignment.cpp")
# Likewise the member-function templates of PhotometricBundleAdjustment (updatePointStatuses, relinearizeSystem) up to, not
# including, the explicit instantiations at the end of the file (oracle/ref_shims/ref_pba.cpp).
PBA_BASE_SRC = os.path.join(SRC, "energy/problems/src/photometric_bundle_adjustment.cpp")
REF_SOURCES = [os.path.join(SRC, p) for p in (
    "energy/camera_model/src/camera_model_base.cpp",
    "features/src/pixel_map.cpp",
    "features/src/calculate_pixelinfo.cpp",
    "features/src/photometrically_corrected_image.cpp",
    "features/src/pixel_data_frame.cpp",
    "sensors/camera_calibration/src/camera_mask.cpp",
    "energy/problems/src/normal_linear_system.cpp",
)]


def include_dirs():
    """Every `include/` and `internal/` directory of the reference's source tree (what its CMake targets export)."""
    dirs = []
    for d, sub, _ in os.walk(SRC):
        for s in sub:
            if s in ("include", "internal"):
                dirs.append(os.path.join(d, s))
    return sorted(dirs)


def have_reference():
    return all(os.path.exists(p) for p in REF_SOURCES)


def available():
    return os.path.exists(LIB) or have_reference()


def _stub_files():
    out = []
    for d, _, fs in os.walk(STUBS):
        out += [os.path.join(d, f) for f in fs]
    return out


def pose_alignment_prefix(directory):
    """The reference file's lines up to the `}  // namespace` that closes its anonymous namespace, plus the three closing
    braces of dsopp::energy::problem -> path of the temporary file."""
    lines = open(POSE_ALIGNMENT_SRC).read().split("\n")
    end = next(i for i, l in enumerate(lines) if l.strip() == "}  // namespace" and i > 200)
    assert "class PoseAlignerProblem" in "\n".join(lines[:end])
    path = os.path.join(directory, "eigen_pose_alignment_prefix.inc")
    with open(path, "w") as f:
        f.write("\n".join(lines[:end + 1]) + "\n}  // namespace problem\n}  // namespace energy\n}  // namespace dsopp\n")
    return path


def pba_base_prefix(directory):
    """photometric_bundle_adjustment.cpp up to the `#define PBAInstantiation` line, plus the three closing braces."""
    lines = open(PBA_BASE_SRC).read().split("\n")
    end = next(i for i, l in enumerate(lines) if l.startswith("#define PBAInstantiation"))
    assert "::updatePointStatuses(" in "\n".join(lines[:end])
    path = os.path.join(directory, "photometric_bundle_adjustment_prefix.inc")
    with open(path, "w") as f:
        f.write("\n".join(lines[:end]) + "\n}  // namespace problem\n}  // namespace energy\n}  // namespace dsopp\n")
    return path


def command(prefix="eigen_pose_alignment_prefix.inc", pba_prefix="photometric_bundle_adjustment_prefix.inc"):
    cmd = ["g++", "-std=c++20", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-I", STUBS,
           "-DREF_POSE_ALIGNMENT_PREFIX=\"%s\"" % prefix, "-DREF_PBA_PREFIX=\"%s\"" % pba_prefix]
    for d in include_dirs():
        cmd += ["-I", d]
    return cmd + ["-o", LIB] + SHIMS + REF_SOURCES


def build():
    """Returns the library path, or None when neither the reference checkout nor a prebuilt library is here."""
    if not have_reference():
        return LIB if os.path.exists(LIB) else None
    deps = REF_SOURCES + SHIMS + _stub_files() + [POSE_ALIGNMENT_SRC, PBA_BASE_SRC, os.path.abspath(__file__)]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    import shutil
    import tempfile
    tmp = tempfile.mkdtemp(prefix="dsopp_ref_pa_")
    try:
        cmd = command(pose_alignment_prefix(tmp), pba_base_prefix(tmp))
        print("+", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    if "--deps" in sys.argv:
        out = subprocess.run(command()[:-len(SHIMS + REF_SOURCES) - 2] + ["-MM"] + SHIMS[:2] + REF_SOURCES,
                             capture_output=True, text=True).stdout
        for tok in sorted(set(t for t in out.replace("\\\n", " ").split() if t.startswith(REF))):
            print(tok)
    else:
        print(build())
