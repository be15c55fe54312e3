"""Timing of the coarse-tracker alignment (BASELINE.json configs[2]) on the GPU box: the one-cluster device LM solve
vs the NumPy oracle on the host, sparse (reference-realistic) and dense (synthetic full-frame bound) depth maps."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsopp_b200 import pose_alignment as G, synth
from oracle import pose_alignment_oracle as PA

out = []
for name, W, H, density in (("sparse 640x480 (2% of pixels)", 640, 480, 0.02), ("dense 640x480 (every pixel)", 640, 480, 1.0)):
    case = synth.make_alignment_case(seed=3, width=W, height=H, density=density, pose_noise=4e-3)
    r, t = case.reference, case.target
    ids32, w32 = case.idepth_sum.astype(np.float32), case.weight.astype(np.float32)
    al = G.Aligner(W * H, W, H)
    n = al.set_reference_depth_map(r.image, ids32, w32, r.T_w_true, r.exposure, r.ab0, r.intr)
    al.set_target(t.image, t.mask, case.T_w_target_guess, t.exposure, t.ab0, t.intr)
    stream = torch.cuda.ExternalStream(al.stream)
    res = None
    ms = []
    for i in range(13):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = al.solve()
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    sweeps = res["iterations"] + 1
    uv, idepth, patch = PA.landmarks_from_depth_map(ids32.astype(np.float64), w32.astype(np.float64), r.image)
    ref = PA.PAFrame(r.T_w_true, r.exposure, r.ab0, r.intr, r.image, r.mask)
    tgt = PA.PAFrame(case.T_w_target_guess, t.exposure, t.ab0, t.intr, t.image, t.mask)
    t0 = time.perf_counter()
    tr = []
    PA.solve(ref, tgt, uv, idepth, patch, trace=tr)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    cpp = min((PA.solve_cpp(ref, tgt, uv, idepth, patch) for _ in range(3)), key=lambda o: o["seconds"])
    med = float(np.median(ms))
    out.append(dict(case=name, landmarks=n, lm_iterations=res["iterations"], gpu_ms_per_solve=med,
                    gpu_point_residuals_per_s=n * sweeps / (med * 1e-3), numpy_oracle_ms_per_solve=cpu_ms,
                    cpp_serial_ms_per_solve=cpp["seconds"] * 1e3, cpp_lm_iterations=cpp["iterations"], rmse=res["rmse"]))
    al.close()
print(json.dumps(out))
