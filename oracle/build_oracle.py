"""Builds oracle/cpu_ref (the C++ restatement used as checker and as the timed CPU baseline).

  oracle/_build/libpba_cpu_ref.so         -O3 -march=x86-64-v3 -ffp-contract=off  (portable, bit-stable predicates)
  oracle/_build/libpba_cpu_ref_native_<cpu>.so  -O3 -march=native                 (timing; the reference's own flags,
                                          CMakeLists.txt:28) -- built on the machine that runs the benchmark, the name
                                          carries a digest of that machine's CPU flags

The reference itself (oracle/_ref) cannot be built: its path needs Eigen, Sophus, TBB and glog, none of which
is installed, so there is no recipe for it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_ref", "pba_cpu_ref.cpp")
OUT = os.path.join(HERE, "_build")
PORTABLE = os.path.join(OUT, "libpba_cpu_ref.so")


def _build(target, flags):
    os.makedirs(OUT, exist_ok=True)
    if os.path.exists(target) and os.path.getmtime(target) >= os.path.getmtime(SRC):
        return target
    cmd = ["g++", "-std=c++17", "-O3", "-fopenmp", "-fPIC", "-shared", "-Wall"] + flags + ["-o", target, SRC]
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return target


PA_SRC = os.path.join(HERE, "cpu_ref", "pose_alignment_cpu_ref.cpp")
PA_LIB = os.path.join(OUT, "libpose_alignment_cpu_ref.so")


def build_pose_alignment():
    """oracle/_build/libpose_alignment_cpu_ref.so: serial C++ restatement of the coarse-tracker aligner.  Portable flags
    (x86-64-v3): the library may be built in one container and loaded on another host."""
    os.makedirs(OUT, exist_ok=True)
    if os.path.exists(PA_LIB) and os.path.getmtime(PA_LIB) >= os.path.getmtime(PA_SRC):
        return PA_LIB
    cmd = ["g++", "-std=c++17", "-O3", "-fPIC", "-shared", "-Wall", "-march=x86-64-v3", "-o", PA_LIB, PA_SRC]
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return PA_LIB


def build_portable():
    return _build(PORTABLE, ["-march=x86-64-v3", "-ffp-contract=off"])


def _host_tag():
    """-march=native code only runs on the CPU it was built for: the file name carries a digest of this host's CPU
    flags, so a library that travelled here from another machine is never loaded."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def build_native():
    return _build(os.path.join(OUT, f"libpba_cpu_ref_native_{_host_tag()}.so"), ["-march=native"])


if __name__ == "__main__":
    build_portable()
    if "--native" in sys.argv:
        build_native()
