"""Prints the headline numbers of a bench.py JSON line (file argument)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f M/s  us/iter %.1f  e2e %.1f M/s (%.2f ms/step)" % (d["value"] / 1e6, d["us_per_gn_iter"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]))
print(" ".join("%s=%.1f" % (k, 1e3 * v["ms_total"] / v["launches"]) for k, v in d["kernel_ms"].items()))
print("roofline frac fused %.3f sweep %.3f big %s" % (d["roofline"]["frac"], d["roofline_sweep"]["frac"], d["roofline_sweep_big"] and round(d["roofline_sweep_big"]["frac"], 3)))
if d.get("parity_check"):
    print("parity_check", d["parity_check"])
if d.get("e2e_raw_frames"):
    print("e2e raw frames %.1f M/s (%.2f ms/step)" % (d["e2e_raw_frames"]["value"] / 1e6, d["e2e_raw_frames"]["ms_per_step"]))
if d.get("config3_strong"):
    c = d["config3_strong"]
    print("config3 strong: %.1f M/s, %.1f us/iter at N=%d, check %s" % (c["value"] / 1e6, c["us_per_gn_iter"], c["n_gpus"], c["multi_gpu_check"]))
if d.get("config2_pose_alignment"):
    for c in d["config2_pose_alignment"].get("cases", []):
        print("config2 %s: gpu %.3f ms, cpu serial %.1f ms (x%.0f)" % (c["case"][:6], c["gpu_ms_coarse_to_fine"], c["cpu_serial_cpp_ms_coarse_to_fine"], c["speedup"]))
    if "error" in d["config2_pose_alignment"]:
        print("config2 error", d["config2_pose_alignment"]["error"])
if d.get("config4_marginalisation"):
    print("config4", {k: v for k, v in d["config4_marginalisation"].items() if k in ("gpu_ms", "cpu_port_ms", "speedup", "prior_rel_diff_gpu_vs_cpu_port", "error")})
if d.get("e2e_one_new_keyframe"):
    print("e2e one new keyframe per solve: %.1f M/s (%.2f ms/step), energy %.4f" % (d["e2e_one_new_keyframe"]["value"] / 1e6, d["e2e_one_new_keyframe"]["ms_per_step"], d["e2e_one_new_keyframe"]["energy"]))
if d.get("e2e_pixelinfo_records"):
    print("e2e with {I,dx,dy} records uploaded: %.1f M/s (%.2f ms/step)" % (d["e2e_pixelinfo_records"]["value"] / 1e6, d["e2e_pixelinfo_records"]["ms_per_step"]))
