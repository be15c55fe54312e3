"""Recipe for oracle/_ref: the parts of the REFERENCE ITSELF that compile here from their own sources.

Almost all of the reference's hot path needs Eigen, Sophus, TBB and glog (absent, no network), so the path as a whole
cannot be built by its own build system; oracle/build_ref_pba.py and oracle/build_ref_tracker.py compile it against
stand-ins of those libraries instead.  Two files need no stand-in at all -- they use nothing but the standard library and AVX2 intrinsics, apart from including common/settings.hpp,
whose only third-party line is a type alias on Eigen::aligned_allocator that neither file uses (oracle/ref_stubs/Eigen/Core
supplies that one name):

  src/features/src/calculate_pixelinfo.cpp                    -> ref_pixelinfo_f64 / ref_pixelinfo_f32
  .../levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp (header-only template) -> ref_lm_solve
  src/common/pattern/include/common/pattern/pattern.hpp (constants; needs only the NAME Eigen::Matrix)  -> ref_pattern

They are compiled WHERE THEY LIE under /root/reference (never copied), together with oracle/ref_shims/ref_parts.cpp, into
oracle/_ref/libdsopp_ref_parts.so (git-ignored, travels to the GPU box with the snapshot).  /root/reference does not
exist on the GPU box: there the prebuilt library is used if present, and the tests fall back to the golden vectors made
from it (tests/golden/ref_parts.npz, tools/make_ref_golden.py).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libdsopp_ref_parts.so")
SHIM = os.path.join(HERE, "ref_shims", "ref_parts.cpp")
REF_SOURCES = [os.path.join(REF, "src/features/src/calculate_pixelinfo.cpp")]
REF_HEADERS = [os.path.join(REF, "src/energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp"),
               os.path.join(REF, "src/common/include/common/settings.hpp"),
               os.path.join(REF, "src/features/internal/features/camera/calculate_pixelinfo.hpp"),
               os.path.join(REF, "src/common/pattern/include/common/pattern/pattern.hpp")]


def available():
    return os.path.exists(LIB) or all(os.path.exists(p) for p in REF_SOURCES + REF_HEADERS)


def build():
    """Returns the library path, or None when neither the reference checkout nor a prebuilt library is here."""
    have_ref = all(os.path.exists(p) for p in REF_SOURCES + REF_HEADERS)
    if not have_ref:
        return LIB if os.path.exists(LIB) else None
    deps = REF_SOURCES + REF_HEADERS + [SHIM, os.path.join(HERE, "ref_stubs", "Eigen", "Core"),
                                        os.path.join(HERE, "ref_stubs", "Eigen", "Dense")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    # the reference's own flags are -O3 -march=native (CMakeLists.txt:28); x86-64-v3 keeps its AVX2 path and lets the
    # library run on another host
    cmd = ["g++", "-std=c++20", "-O3", "-march=x86-64-v3", "-fPIC", "-shared",
           "-I", os.path.join(HERE, "ref_stubs"),
           "-I", os.path.join(REF, "src/common/include"),
           "-I", os.path.join(REF, "src/common/pattern/include"),
           "-I", os.path.join(REF, "src/features/internal"),
           "-I", os.path.join(REF, "src/energy/problems/include"),
           "-o", LIB, SHIM] + REF_SOURCES
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build())
