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
  src/sensors/camera_calibration/src/camera_mask.cpp
  src/energy/problems/src/normal_linear_system.cpp
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
SHIMS = [os.path.join(HERE, "ref_shims", "ref_pba.cpp"), os.path.join(HERE, "ref_shims", "ref_link_stubs.cpp")]
REF_SOURCES = [os.path.join(SRC, p) for p in (
    "energy/camera_model/src/camera_model_base.cpp",
    "features/src/pixel_map.cpp",
    "features/src/calculate_pixelinfo.cpp",
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


def command():
    cmd = ["g++", "-std=c++20", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-I", STUBS]
    for d in include_dirs():
        cmd += ["-I", d]
    return cmd + ["-o", LIB] + SHIMS + REF_SOURCES


def build():
    """Returns the library path, or None when neither the reference checkout nor a prebuilt library is here."""
    if not have_reference():
        return LIB if os.path.exists(LIB) else None
    deps = REF_SOURCES + SHIMS + _stub_files()
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = command()
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    if "--deps" in sys.argv:
        out = subprocess.run(command()[:-len(SHIMS + REF_SOURCES) - 2] + ["-MM"] + SHIMS + REF_SOURCES,
                             capture_output=True, text=True).stdout
        for tok in sorted(set(t for t in out.replace("\\\n", " ").split() if t.startswith(REF))):
            print(tok)
    else:
        print(build())
