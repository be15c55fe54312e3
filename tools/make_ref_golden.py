"""Writes tests/golden/ref_parts.npz from the REFERENCE ITSELF (oracle/_ref/libdsopp_ref_parts.so, oracle/build_ref.py):
outputs of dsopp::features::calculate_pixelinfo<1> and call traces of levenberg_marquardt_algorithm::solve on scripted
problems.  Run here, where /root/reference exists; the fixture travels, the reference does not.

    python tools/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_parts as R  # noqa: E402


def lm_cases(seed=0, n=160):
    """Scripted LM runs: random energy sequences (descents, plateaus within the tolerances, increases), valid counts
    (now and then 0), step norms (now and then below the parameter tolerance) and the option sets the solver uses."""
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n):
        max_it = int(rng.choice([1, 3, 7, 12, 50]))
        m = max_it + 3
        e = np.empty(m)
        e[0] = rng.uniform(1e2, 1e6)
        for k in range(1, m):
            u = rng.random()
            if u < 0.55:
                e[k] = e[k - 1] * rng.uniform(0.3, 0.999)
            elif u < 0.70:
                e[k] = e[k - 1] * (1.0 - rng.uniform(0, 2e-8))   # inside the function tolerance
            elif u < 0.75:
                e[k] = e[k - 1]
            else:
                e[k] = e[k - 1] * rng.uniform(1.0001, 3.0)
        valid = rng.integers(1, 100000, m).astype(np.int32)
        if rng.random() < 0.15:
            valid[rng.integers(0, m)] = 0
        norms = np.stack([rng.uniform(1e-2, 1e4, m), 10.0 ** rng.uniform(-14, 0, m)], axis=1)
        force = bool(rng.random() < 0.5)
        opts = dict(max_it=max_it, lambda0=float(rng.choice([1e-5, 1e-2, 0.1])), ftol=float(rng.choice([1e-8, 1e-5, 0.0])),
                    ptol=float(rng.choice([1e-8, 1e-5, 0.0])), force_accept=force,
                    min_it=int(rng.choice([0, 3])) if force else 0,
                    dec=float(rng.choice([1.0, 2.0])), inc=float(rng.choice([1.0, 5.0, 10.0])))
        cases.append((e, valid, norms, opts))
    return cases


def main():
    out = {}
    rng = np.random.default_rng(1)
    for name, (H, W) in {"a": (48, 64), "b": (37, 53), "c": (2, 8), "d": (16, 4), "e": (5, 16)}.items():
        I = rng.uniform(0.0, 255.0, (H, W))
        out[f"px_{name}_in"] = I
        out[f"px_{name}_f64"] = R.pixelinfo(I, aligned=False)                    # plain-C path, any width
        out[f"px_{name}_f32"] = R.pixelinfo(I.astype(np.float32), aligned=False)
        if W % 8 == 0:                                                          # AVX2 path (double, aligned, width % 8 == 0)
            assert np.array_equal(out[f"px_{name}_f64"], R.pixelinfo(I, aligned=True))
    pat, center = R.pattern()
    out["pattern_xy"], out["pattern_center"] = pat, np.array(center)
    cases = lm_cases()
    out["lm_n"] = np.array(len(cases))
    for i, (e, v, nr, o) in enumerate(cases):
        calls, lams, E, nv, conv = R.lm_solve(e, v, nr, **o)
        out[f"lm_{i}_e"], out[f"lm_{i}_v"], out[f"lm_{i}_norms"] = e, v, nr
        out[f"lm_{i}_opts"] = np.array([o["max_it"], o["lambda0"], o["ftol"], o["ptol"], float(o["force_accept"]),
                                        o["min_it"], o["dec"], o["inc"]])
        out[f"lm_{i}_calls"] = np.array([R.CALL_NAMES.index(c) for c in calls], np.int8)
        out[f"lm_{i}_lams"] = lams
        out[f"lm_{i}_result"] = np.array([E, nv, float(conv)])
    path = os.path.join(ROOT, "tests", "golden", "ref_parts.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
