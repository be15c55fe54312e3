"""Recipe for oracle/_ref/libdsopp_ref_tracker.so: the REFERENCE'S OWN tracker arithmetic either side of the bundle
adjustment, compiled from its sources.  Test infrastructure only.

  src/tracker/tracker/src/create_depth_maps.cpp                       the whole file, unchanged
  src/tracker/landmarks_activator/src/landmarks_activator.cpp:122-316  class LandmarkActivationProblem and
                                                                      optimizeImmatureLandmark (ACTIVATOR_SRC below)
  + camera_model_base.cpp, pixel_map.cpp, calculate_pixelinfo.cpp, camera_mask.cpp and every header they include

Like oracle/build_ref_pba.py the third-party libraries are the stand-ins of oracle/ref_stubs_full.  In addition the track
subsystem's CONTAINERS (ActiveKeyframe, landmark records, FrameConnection, CameraCalibration: protobuf-backed storage with
no arithmetic on this path) are the plain records of oracle/ref_stubs_track, put in front of the reference's include
directories for this library only -- which is why it is a separate .so.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import build_ref_pba as B  # noqa: E402

LIB = os.path.join(B.OUT, "libdsopp_ref_tracker.so")
TRACK_STUBS = os.path.join(HERE, "ref_stubs_track")
SHIM = os.path.join(HERE, "ref_shims", "ref_tracker.cpp")
LINK_STUBS = os.path.join(HERE, "ref_shims", "ref_link_stubs.cpp")  # SemanticFilter::filtered (named by camera_mask, never called)
ACTIVATOR_SRC = os.path.join(B.SRC, "tracker/landmarks_activator/src/landmarks_activator.cpp")
REF_SOURCES = [os.path.join(B.SRC, p) for p in (
    "tracker/tracker/src/create_depth_maps.cpp",
    "energy/camera_model/src/camera_model_base.cpp",
    "features/src/pixel_map.cpp",
    "features/src/calculate_pixelinfo.cpp",
    "sensors/camera_calibration/src/camera_mask.cpp",
)]


def have_reference():
    return all(os.path.exists(p) for p in REF_SOURCES + [ACTIVATOR_SRC])


def available():
    return os.path.exists(LIB) or have_reference()


def activator_prefix(directory):
    """landmarks_activator.cpp from `class LandmarkActivationProblem`'s template header to the closing brace of
    optimizeImmatureLandmark -> path of the temporary file (outside the repository, removed after the build)."""
    lines = open(ACTIVATOR_SRC).read().split("\n")
    cls = next(i for i, l in enumerate(lines) if l.startswith("class LandmarkActivationProblem"))
    start = cls - 1
    assert lines[start].startswith("template <")
    nxt = next(i for i, l in enumerate(lines) if l.startswith("void optimizeImmatureLandmarks("))
    end = nxt - 1  # the `template <...>` line of optimizeImmatureLandmarks
    assert lines[end].startswith("template <")
    body = "\n".join(lines[start:end])
    assert "optimizeImmatureLandmark(" in body and "calculateEnergy" in body
    path = os.path.join(directory, "landmarks_activator_prefix.inc")
    with open(path, "w") as f:
        f.write(body + "\n")
    return path


def _stub_files():
    out = []
    for root in (B.STUBS, TRACK_STUBS):
        for d, _, fs in os.walk(root):
            out += [os.path.join(d, f) for f in fs]
    return out


def command(prefix):
    cmd = ["g++", "-std=c++20", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-I", TRACK_STUBS, "-I", B.STUBS,
           "-DREF_ACTIVATOR_PREFIX=\"%s\"" % prefix]
    for d in B.include_dirs():
        cmd += ["-I", d]
    return cmd + ["-o", LIB, SHIM, LINK_STUBS] + REF_SOURCES


def build():
    """Returns the library path, or None when neither the reference checkout nor a prebuilt library is here."""
    if not have_reference():
        return LIB if os.path.exists(LIB) else None
    deps = REF_SOURCES + [SHIM, LINK_STUBS, ACTIVATOR_SRC, os.path.abspath(__file__)] + _stub_files()
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(B.OUT, exist_ok=True)
    import shutil
    import tempfile
    tmp = tempfile.mkdtemp(prefix="dsopp_ref_tracker_")
    try:
        r = subprocess.run(command(activator_prefix(tmp)), capture_output=True, text=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building %s failed" % LIB)
    return LIB


if __name__ == "__main__":
    print(build())
