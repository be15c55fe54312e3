#!/usr/bin/env python
"""Benchmark of the photometric bundle-adjustment hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one Levenberg-Marquardt solve of the sliding window = GN_ITERS Gauss-Newton iterations, each =
linearize (residual+Jacobian sweep, H_pp, per-point Schur) + reduced solve + back-substitution + residual-only
energy sweep + accept (the loop body of levenberg_marquardt_algorithm.hpp:85-114), executed by dpba_solve_lm
(the whole loop on the device, one host synchronisation per solve; --host-lm drives the same kernels from the C++
LevenbergMarquardtProblem adapter instead).  metric = patch-residuals per second per GN iteration = units * GN_ITERS * steps / time.

  value : window resident in HBM when the timed region starts.
  e2e   : every step additionally re-uploads the whole window (images, masks, landmarks, statuses, state) from
          pinned host memory through dpba_push_frame / dpba_set_* and reads the result (state, idepths, statuses)
          back -- conservative: the tracker uploads ONE new keyframe per solve.

Multi-GPU (weak scaling): every rank holds PTS_PER_GPU landmarks per keyframe of a window with
PTS_PER_GPU * N landmarks per keyframe; frames are replicated; one NCCL allreduce of the packed reduced system
per linearisation and one of (energy, n) per energy evaluation.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAMES = 8
PTS_PER_GPU = 2000
GN_ITERS = 7
SIGMA = 20.0
AB_REG = (1e12, 1e8)
FIXED_REG = 1e16
METRIC = "patch-residuals/sec per GN iter"
UNIT = "patch-residuals/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_window(world):
    from dsopp_b200 import synth
    return synth.make_window(n_frames=N_FRAMES, points_per_frame=PTS_PER_GPU * world, seed=0, ab_scale=0.0)


def config_dict(world, extra=None):
    c = {
        "workload": f"configs[1]: {N_FRAMES}-keyframe sliding window, {PTS_PER_GPU} active points/KF per GPU, 8-px patch, "
                    f"640x480 level 0, Huber sigma=20, FEJ, affine a/b per frame, frame 0 fixed",
        "units_per_gn_iter": N_FRAMES * (N_FRAMES - 1) * PTS_PER_GPU * world,
        "gn_iters_per_step": GN_ITERS,
        "lm": "levenberg_marquardt_algorithm::solve, force_accept, min_it=max_it=7, tolerances 0 (fixed work per step)",
        "parallelism": f"landmarks sharded over {world} GPU(s), frames replicated",
        "l2": "256 MiB buffer written between timed steps (L2 flush); within a step the 79 MB image set (32-byte texel-pair records) stays in the 126 MB L2",
    }
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's algorithm and dataflow (oracle/cpu_ref: the reference binary cannot be built here --
    Eigen/Sophus/TBB absent) on this box's host cores, same window, same 7-iteration step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the restatement asks for its thread count per parallel region
    # (num_threads clause), which overrides that, but the OpenMP runtime must not have been capped at load time either
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    from oracle import cpu_ref
    win = build_window(1)
    cw = cpu_ref.CpuWindow(win, use_float=False, threads=threads, native=True)
    cw.first_estimate()
    units = win.units
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for _ in range(GN_ITERS):
            cw.gn_iteration(SIGMA, True, 1e-5, AB_REG, FIXED_REG)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = units * GN_ITERS * len(times) / total
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(1, {"parallelism": f"{threads} OpenMP threads on the host"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{len(times)} full steps of {GN_ITERS} GN iterations on the whole window "
                                   f"({units} patch-residuals), double precision, reference three-pass dataflow"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def _time_cpu_window(cwin, use_float, threads, budget_s, max_iters):
    """Median per-phase seconds of GN iterations of oracle/cpu_ref on `cwin` (bounded by time and count)."""
    from oracle import cpu_ref
    cw = cpu_ref.CpuWindow(cwin, use_float=use_float, threads=threads, native=True)
    cw.first_estimate()
    ts, t_begin = [], time.perf_counter()
    first_step = None
    for i in range(max(2, GN_ITERS) + max_iters):
        e, tm, _ = cw.gn_iteration(SIGMA, True, 1e-5, AB_REG, FIXED_REG)
        if i == GN_ITERS - 1:  # the first GN_ITERS iterations from the fresh window ARE one bench step: keep its result
            first_step = {"energy": float(e), "eps": cw.get_state()[0].copy()}
        if i >= 2:
            ts.append(tm.copy())
        if time.perf_counter() - t_begin > budget_s and len(ts) >= 3 and first_step is not None:
            break
    cw.close()
    return np.median(np.array(ts), axis=0), len(ts), first_step


def cpu_baseline_leg():
    """SURVEY 8(d) 'CPU reference timing': the reference's three-pass dataflow (oracle/cpu_ref) on this box's host
    cores -- double on min(nproc, 8) - 1 threads (the reference's TBB cap, dsopp_main.cpp:114-117) is the headline;
    the float build and the 1-thread figure ride along on shorter samples.  Reported, never required."""
    try:
        from dsopp_b200 import synth
        cwin = synth.make_window(n_frames=N_FRAMES, points_per_frame=PTS_PER_GPU, seed=0, ab_scale=0.0)
        threads = max(1, min(os.cpu_count() or 1, 8) - 1)
        med, n, first_step = _time_cpu_window(cwin, False, threads, 20.0, 40)
        cpu = {"value": cwin.units / med[5], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n} GN iterations of the same {cwin.units}-unit window, double precision, "
                         f"reference three-pass dataflow (oracle/cpu_ref, -O3 -march=native), median",
               "phase_ms": {"sweep_K1": 1e3 * med[0], "posepose_K3": 1e3 * med[1], "schur_K4": 1e3 * med[2],
                            "solve_K5": 1e3 * med[3], "energy_K2": 1e3 * med[4], "iteration": 1e3 * med[5]},
               "host_cores_total": os.cpu_count(), "_first_step": first_step}
        variants = {}
        for name, use_float, th in (("float_%dthreads" % threads, True, threads), ("double_1thread", False, 1)):
            try:
                m, k, _ = _time_cpu_window(cwin, use_float, th, 6.0, 12)
                variants[name] = {"value": cwin.units / m[5], "cores": th, "iterations": k, "iteration_ms": 1e3 * m[5]}
            except Exception as ex:
                variants[name] = {"value": None, "error": str(ex)}
        cpu["variants"] = variants
        return cpu
    except Exception as ex:  # the baseline is reported, never required
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}


# --------------------------------------------------------------------------------------------------
def pose_alignment_leg(torch):
    from dsopp_b200 import pose_alignment as G, synth
    from oracle import pose_alignment_oracle as PA
    out = {"workload": "configs[2]: direct image alignment of a new frame to the last keyframe, 640x480, SE3 + affine "
                       "(6 + 2 DoF), 1-pixel residuals, LM lambda0 = 1e-2 x/÷ 2, <= 50 iterations, 4 pyramid levels coarse to fine",
           "cases": []}
    for name, density in (("dense (every pixel carries depth: BASELINE's full-frame bound)", 1.0),
                          ("sparse (2 % of the pixels: the reference's splatted depth map)", 0.02)):
        case = synth.make_alignment_case(seed=3, width=640, height=480, density=density, pose_noise=4e-3)
        r, t = case.reference, case.target
        ref_I, tgt_I = [r.image[..., 0]], [t.image[..., 0]]
        ids, w = [case.idepth_sum.astype(np.float32)], [case.weight.astype(np.float32)]
        for _ in range(1, 4):
            ref_I.append(synth.downscale(ref_I[-1]))
            tgt_I.append(synth.downscale(tgt_I[-1]))
            a, b = ids[-1], w[-1]
            ids.append(a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2])
            w.append(b[0::2, 0::2] + b[1::2, 0::2] + b[0::2, 1::2] + b[1::2, 1::2])
        al = G.Aligner(640 * 480, 640, 480)
        stream = torch.cuda.ExternalStream(al.stream)
        T = case.T_w_target_guess
        levels, gpu_total, cpu_total, units = [], 0.0, 0.0, 0
        for lvl in range(3, -1, -1):
            intr = r.intr / (2 ** lvl)
            Hh, Ww = ref_I[lvl].shape
            ref_img, tgt_img = synth.pixelinfo(ref_I[lvl]), synth.pixelinfo(tgt_I[lvl])
            mask = np.full((Hh, Ww), 255, np.uint8)
            n = al.set_reference_depth_map(ref_img, ids[lvl], w[lvl], r.T_w_true, r.exposure, r.ab0, intr)
            al.set_target(tgt_img, mask, T, t.exposure, t.ab0, intr)
            ms, res = [], None
            for i in range(3 + 10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                res = al.solve()
                e1.record(stream)
                torch.cuda.synchronize()
                if i >= 3:
                    ms.append(e0.elapsed_time(e1))
            uv, idepth, patch = PA.landmarks_from_depth_map(ids[lvl].astype(np.float64), w[lvl].astype(np.float64), ref_img)
            ref = PA.PAFrame(r.T_w_true, r.exposure, r.ab0, intr, ref_img, np.full((Hh, Ww), 255, np.uint8))
            tgt = PA.PAFrame(T, t.exposure, t.ab0, intr, tgt_img, mask)
            cpp = min((PA.solve_cpp(ref, tgt, uv, idepth, patch) for _ in range(2)), key=lambda o: o["seconds"])
            med = float(np.median(ms))
            sweeps = res["iterations"] + 1
            levels.append({"level": lvl, "size": [Ww, Hh], "landmarks": int(n), "lm_iterations": int(res["iterations"]),
                           "gpu_ms": med, "cpu_serial_cpp_ms": cpp["seconds"] * 1e3, "cpu_lm_iterations": int(cpp["iterations"]),
                           "rmse": float(res["rmse"]), "rmse_cpu": float(cpp.get("rmse", float("nan")))})
            gpu_total += med
            cpu_total += cpp["seconds"] * 1e3
            units += int(n) * sweeps
            T = res["T_w_target"]
        al.close()
        out["cases"].append({"case": name, "levels": levels, "gpu_ms_coarse_to_fine": gpu_total,
                             "cpu_serial_cpp_ms_coarse_to_fine": cpu_total, "speedup": cpu_total / gpu_total,
                             "gpu_point_residuals_per_s": units / (gpu_total * 1e-3),
                             "timing": "CUDA events around dpa_solve (one kernel launch + result readback per level), median of 10"})
    return out


def marginalisation_leg():
    """configs[4]: Schur-eliminate the oldest keyframe (2000 landmarks) of the 8-keyframe window into the dense prior.
    GPU: CudaPhotometricBundleAdjustment::marginalizeNow (firstEstimateJacobians + fused linearise over the flagged landmarks
    + landmark energy on the device, the 64x64 algebra and reduce_system in fp64 on the host), wall clock incl. every
    synchronisation.  CPU: the same update with oracle/cpu_ref (double, reference dataflow) + NumPy for the 64x64 part."""
    from dsopp_b200 import host, synth
    from oracle import cpu_ref, pba_oracle as O
    # eps_scale = 0: keyframes enter the solver class at their linearisation point
    win = synth.make_window(n_frames=N_FRAMES, points_per_frame=PTS_PER_GPU, seed=0, ab_scale=0.0, eps_scale=0.0)
    gpu_ms = []
    for rep in range(4):
        pba = host.CudaPhotometricBundleAdjustment(win.width, win.height, max_frames=9, max_points=2048, estimate_uncertainty=False)
        ids = []
        for f in win.frames:
            pba.push_frame(f.frame_id, f.timestamp, f.T_w_lin, f.exposure, f.ab0, f.intr, f.image, f.mask, f.uv, f.idepth,
                           f.patch, f.flags, fixed=f.fixed, other_ids=ids)
            ids.append(f.frame_id)
        f = win.frames[0]
        pba.update_local_frame(f.frame_id, f.timestamp, f.T_w_lin, f.exposure, f.ab0, f.intr, f.uv, f.idepth, f.patch,
                               np.full(len(f.idepth), synth.FLAG_MARGINALIZED, np.uint8), is_marginalized=True)
        t0 = time.perf_counter()
        pba.marginalize_now()
        dt = time.perf_counter() - t0
        if rep:
            gpu_ms.append(dt * 1e3)
        Hm, bm, em = pba.marginalized_system()
        pba.close()
    # CPU: same flags, same update
    win.frames[0].flags[:] = synth.FLAG_MARGINALIZED | synth.FLAG_TO_MARGINALIZE
    threads = max(1, min(os.cpu_count() or 1, 8) - 1)
    cpu_ms = []
    for rep in range(3):
        cw = cpu_ref.CpuWindow(win, use_float=False, threads=threads, native=True)
        t0 = time.perf_counter()
        cw.first_estimate()
        cw.evaluate(SIGMA, True, True)
        cw.change_statuses(True)
        Hp, bp = cw.pose_pose(True)
        Hs, bs = cw.schur(True)
        e_l, _ = cw.landmarks_energy(True)
        eps, _ = cw.get_state()
        H, b = Hp - Hs, bp - bs
        e_marg = e_l + eps @ (H @ eps) - eps @ b
        b = b - H @ eps
        prior_H, prior_b = np.zeros_like(H), np.zeros_like(b)
        prior_H[:8, :8] += np.eye(8) * FIXED_REG  # the dropped frame is the fixed one
        prior_b[:8] += FIXED_REG * eps[:8]
        prior_b -= prior_H @ eps
        Hr, br = O.reduce_system(H + prior_H, b + prior_b, list(range(8)))
        cpu_ms.append((time.perf_counter() - t0) * 1e3)
        cw.close()
    n = 8 * (N_FRAMES - 1)
    err_h = float(np.abs(Hm[:n, :n] - Hr).max() / np.abs(Hr).max())
    return {"workload": f"configs[4]: marginalisation of the oldest keyframe ({PTS_PER_GPU} landmarks) of the {N_FRAMES}-keyframe "
                        f"window into the dense {n}x{n} prior (updateMarginalizedLinearSystem + reduce_system)",
            "gpu_ms": float(np.median(gpu_ms)), "cpu_port_ms": float(np.median(cpu_ms)), "cpu_threads": threads,
            "speedup": float(np.median(cpu_ms) / np.median(gpu_ms)),
            "landmarks_marginalised": int(len(win.frames[0].idepth)),
            "patch_residuals_per_s": float(len(win.frames[0].idepth) * (N_FRAMES - 1) / (np.median(gpu_ms) * 1e-3)),
            "prior_rel_diff_gpu_vs_cpu_port": err_h,
            "timing": "wall clock around CudaPhotometricBundleAdjustment::marginalizeNow (synchronous), median of 3"}


def pin(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


def run_ours(args):
    import torch
    import torch.distributed as dist
    from dsopp_b200 import capi, host

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peers_attached = False
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    win = build_window(world)
    n = win.n_frames
    shard = [np.arange(rank, len(f.idepth), world) for f in win.frames]
    units_local = sum(len(s) for s in shard) * (n - 1)
    units_global = win.units
    h = capi.upload_window(win, device=local, rank=rank, world_size=world)
    if args.fused_min_blocks:
        h.set_option("fused_min_blocks", args.fused_min_blocks)
    if args.fused_version:
        h.set_option("fused_version", args.fused_version)
    if args.merged_tail >= 0:
        h.set_option("merged_tail", args.merged_tail)
    if args.pdl >= 0:
        h.set_option("pdl", args.pdl)
    if args.three_branch >= 0:
        h.set_option("three_branch", args.three_branch)
    if args.speculative_multi_gpu >= 0:
        h.set_option("speculative_multi_gpu", args.speculative_multi_gpu)
    if args.fused_prefetch >= 0:
        h.set_option("fused_prefetch", args.fused_prefetch)
    if args.fused_epilogue >= 0:
        h.set_option("fused_epilogue", args.fused_epilogue)
    for kv in args.opt:  # any dpba_set_option switch, for A/B runs
        k_, v_ = kv.split("=")
        h.set_option(k_, int(v_))
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        h.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        h.first_estimate()
        h.evaluate(SIGMA, True, True)  # first collective outside any graph capture: NCCL sets its channels up here
        # exchange of the reduced system: the library's own NVLink mailbox kernel, scalars and system in two concurrent
        # exchanges (measured on B200 at N = 2: 102.4 us per iteration against 114.4 with ncclAllReduce in the graph; at
        # N = 8 NCCL's latency grows with the rank count, the one-shot mailbox exchange does not; profiles/r02_ab.md).
        # --peer-exchange 0 selects NCCL.
        use_peer = args.peer_exchange != 0
        peers_attached = use_peer
        if use_peer:
            capi.attach_peers(h, rank, world, dev)
            h.set_option("peer_exchange", 1)
            h.set_option("peer_fused", args.peer_fused if args.peer_fused >= 0 else 0)
            h.evaluate(SIGMA, True, True)

    ab0 = np.stack([f.ab0 for f in win.frames])
    fixed = [int(f.fixed) for f in win.frames]
    eps0 = np.concatenate([f.state_eps for f in win.frames])
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def solve():
        # device-resident LM (dpba_solve_lm): same loop / problem methods as the C++ host LM, one host sync per solve
        if args.host_lm:
            return host.lm_solve(h, ab0, fixed, SIGMA, AB_REG, FIXED_REG, max_it=GN_ITERS, min_it=GN_ITERS, ftol=0.0,
                                 ptol=0.0, force_accept=True, lambda0=1e-5)
        h.first_estimate()
        r = h.solve_lm(SIGMA, AB_REG, FIXED_REG, max_it=GN_ITERS, min_it=GN_ITERS, ftol=0.0, ptol=0.0,
                       force_accept=True, lambda0=1e-5)
        return r[0], r[1]

    def reset_resident():
        # same starting point for every step (untimed): landmarks + state back to the initial estimate
        for i, f in enumerate(win.frames):
            h.set_landmarks(i, f.uv[shard[i]], f.idepth[shard[i]], f.patch[shard[i]], f.flags[shard[i]])
        for (r, t), st in win.statuses.items():
            h.set_statuses(r, t, st[shard[r]])
        h.set_state(eps0, np.zeros_like(eps0))

    def timed(fn, count_launches=False):
        flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if world > 1 and peers_attached and args.device_rendezvous:
            # the processes leave torch.distributed's barrier tens of microseconds apart; a device-side rendezvous over the
            # NVLink mailboxes starts the timed region on all ranks at once (dpba_peer_barrier), so that a 0.6 ms step is
            # not charged with host-side process jitter.  --device-rendezvous 0 switches it off (A/B).
            h.peer_barrier()
        l0 = capi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), capi.launch_count() - l0

    # ---- value: resident window (per-kernel event profiling OFF: the graph holds kernels only) ------------------
    sampler = ClockSampler(local)
    step_ms, launches, solved = [], 0, []
    for i in range(args.warmup + args.steps):
        reset_resident()
        if i == args.warmup:
            sampler.start()
        ms, nl = timed(lambda: solved.append(solve()))
        if i >= args.warmup:
            step_ms.append(ms)
            launches += nl
    iters_seen = [r_[1] for r_ in solved]
    assert all(it == GN_ITERS for it in iters_seen), iters_seen
    energy_dev, eps_dev = float(solved[-1][0]), h.get_state()[0].copy()
    t_local = sum(step_ms)
    t = torch.tensor([t_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = units_global * GN_ITERS * args.steps / (total_ms * 1e-3)

    # ---- the same steps again with CUDA events around every kernel inside the graph (roofline / kernel_ms) --------
    h.profile_enable(True)
    for i in range(2 + min(args.steps, 10)):
        reset_resident()
        if i == 2:
            h.profile_enable(True)  # resets the sums after the re-capture + warm-up
        timed(solve)
    prof = h.profile_read()
    h.profile_enable(False)

    # ---- e2e: host buffers in, results out, every step, through ONE C++ call (dpbah_solve_window) ------------------
    keep = []

    def pinned(a):
        x, t_ = pin(a)
        keep.append(t_)
        return x

    def pinned_alloc(shape, dtype):
        return pinned(np.zeros(shape, dtype))

    host_frames = []
    for i, f in enumerate(win.frames):
        s_ = shard[i]
        host_frames.append(dict(frame_id=f.frame_id, image=pinned(f.image.astype(np.float32)), mask=pinned(f.mask),
                                T_w_lin=f.T_w_lin, exposure=f.exposure, ab0=f.ab0, intr=f.intr, fixed=f.fixed,
                                uv=pinned(f.uv[s_].astype(np.float32)), idepth=pinned(f.idepth[s_].astype(np.float32)),
                                patch=pinned(f.patch[s_].astype(np.float32)), flags=pinned(f.flags[s_])))
    host_status = {(r, tt): pinned(win.statuses[(r, tt)][shard[r]]) for r in range(n) for tt in range(n) if tt != r}
    lm_kw = dict(sigma=SIGMA, ab_reg=AB_REG, fixed_reg=FIXED_REG, max_it=GN_ITERS, min_it=GN_ITERS, ftol=0.0, ptol=0.0,
                 force_accept=True, lambda0=1e-5)
    step_rec = host.WindowStep(h, host_frames, host_status, eps0, alloc=pinned_alloc, **lm_kw)
    # the step a binding makes: the intensity plane of every keyframe (what PixelMap::data() holds, pixel_map.hpp:117) crosses
    # PCIe, {I,dx,dy} is built on the device with the reference's gradient definition (bit-identical records,
    # tests/test_gpu_parity.py::test_intensity_upload_builds_the_same_pixel_map) -- a third of the bytes of the record upload
    host_frames_int = [dict(f_, image=pinned(np.ascontiguousarray(f_["image"][..., 0]))) for f_ in host_frames]
    step_io = host.WindowStep(h, host_frames_int, host_status, eps0, alloc=pinned_alloc, **lm_kw)
    raw_frames = [pinned(np.clip(np.rint(f.image[..., 0]), 0, 255).astype(np.uint8)) for f in win.frames]
    step_raw = host.WindowStep(h, host_frames, host_status, eps0, alloc=pinned_alloc, raw_gray=raw_frames,
                               photometric_lut=np.arange(256, dtype=np.float32), **lm_kw)

    def time_e2e(io, steps, warm):
        out_ms = []
        for i in range(warm + steps):
            ms, _ = timed(io.run)
            if i >= warm:
                out_ms.append(ms)
        t_ = torch.tensor([sum(out_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item()) / steps

    e2e_ms_per_step = time_e2e(step_io, args.steps, args.warmup)
    e2e_value = units_global * GN_ITERS / (e2e_ms_per_step * 1e-3)
    e2e_energy, e2e_iters = step_io.io.energy, step_io.io.iterations
    h2d = int(step_io.io.h2d_bytes)
    # what actually crosses PCIe on the way back: the first getter after the solve mirrors ALL landmark arrays and status
    # rows of the handle in one bulk readback (5 floats + float4 + flag per landmark slot, 2 x 16 status rows per frame
    # slot); the arrays handed to the caller are a subset of it
    mpp = max(len(s_) for s_ in shard)
    d2h = max(int(step_io.io.d2h_bytes), n * mpp * (5 * 4 + 16 + 1) + 2 * n * 16 * mpp + 2 * 8 * n * 8)
    # the same step with the {I,dx,dy} RECORDS uploaded (the PixelMap's pixel-info storage, 12 bytes per pixel: round 1's and
    # early round 2's `e2e`) -- informational
    e2e_rec_ms = time_e2e(step_rec, min(args.steps, 10), 3)
    e2e_rec = {"value": units_global * GN_ITERS / (e2e_rec_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_rec_ms,
               "h2d_bytes_per_step": int(step_rec.io.h2d_bytes), "energy": float(step_rec.io.energy)}
    # the same step fed with RAW 8-bit frames (dpba_push_frame_raw: photometric table + gradients on the device,
    # SURVEY 8f-4) -- informational
    e2e_raw_ms = time_e2e(step_raw, min(args.steps, 10), 3)
    e2e_raw_value = units_global * GN_ITERS / (e2e_raw_ms * 1e-3)
    h2d_raw = int(step_raw.io.h2d_bytes)
    # the tracker's steady state (monocular_tracker.cpp:491-507): ONE new keyframe per solve -- the oldest leaves, one arrives
    # from pinned host memory with its landmarks and connection statuses, solve, every frame's results back
    step_io.run()
    e2e_sliding_ms = []
    n_warm = n + 1  # one full rotation first: every (logical -> physical slot) map of the cycle has its captured graph
    for i in range(n_warm + min(args.steps, 16)):
        ms, _ = timed(step_io.run_sliding)
        if i >= n_warm:
            e2e_sliding_ms.append(ms)
    t_sl = torch.tensor([sum(e2e_sliding_ms) / len(e2e_sliding_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_sl, op=dist.ReduceOp.MAX)
    e2e_sliding = {"value": units_global * GN_ITERS / (float(t_sl.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(t_sl.item()),
                   "h2d_bytes_per_step": int(step_io.io.h2d_bytes), "d2h_bytes_per_step": int(d2h),
                   "energy": float(step_io.io.energy),
                   "path": "dpbah_solve_sliding (one call per step): dpba_remove_frame(oldest) + dpba_push_frame(ONE keyframe) + its "
                           "landmarks / statuses from pinned host buffers, state reset, dpba_first_estimate + dpba_solve_lm, "
                           "readback of every frame"}
    step_io.run()  # leave the handle with the float window in its original order for the sweeps timed below
    # clocks / throttle reasons were sampled from the first timed `value` step to the last timed e2e step (the timed
    # `value` region alone lasts ~20 ms, less than one nvidia-smi sampling period)
    clocks = sampler.stop()

    # ---- materialising sweep (K1, reference-surface mode) timed alone with an L2 flush before each launch --------
    def time_sweep(hh):
        hh.first_estimate()
        hh.profile_enable(True)
        for i in range(3 + 10):
            flush.fill_(1)
            torch.cuda.synchronize()
            if i == 3:
                hh.profile_enable(True)
            hh.evaluate_jacobians(SIGMA, True, True)
        ms, cnt = hh.profile_read()["materialise_sweep"]
        hh.profile_enable(False)
        return ms / max(cnt, 1)

    # pure-write HBM bandwidth of this device for context: the materialising sweep is ~80 % stores, and a store stream
    # does not reach the copy figure MEASURED_PEAKS.json holds (read + write bytes of a copy)
    write_gbs = None
    if rank == 0:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        for _ in range(2):
            big.fill_(1)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(5):
            big.fill_(2)
        ev1.record()
        torch.cuda.synchronize()
        write_gbs = 5 * big.numel() / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
        del big
    sweep = sweep4 = None
    units4 = 0
    if rank == 0:
        reset_resident()
        sweep = time_sweep(h)
        if world == 1 and not args.no_big_sweep:
            # configs[3]'s per-GPU shape at G=1 (20000 points/KF, 1.12 M patch-residuals): the size at which the
            # HBM roofline of the sweep is meaningful; same kernel, same code path
            from dsopp_b200 import synth
            win4 = synth.make_window(n_frames=N_FRAMES, points_per_frame=20000, seed=1, ab_scale=0.0)
            h4 = capi.upload_window(win4, device=local)
            units4 = win4.units
            sweep4 = time_sweep(h4)
            h4.close()
            del win4

    # ---- configs[3]: the large window north_star shards -- 8 keyframes x 20000 points (1.12 M patch-residuals), STRONG
    # scaling: the same window on every N, landmarks dealt over the ranks, one exchange of the reduced system per GN
    # iteration.  Every --gpus N line carries it, so the N = 1, 2, 4, 8 lines give the strong curve.  With N > 1 rank 0 also
    # linearises the UNSHARDED window on its own GPU and compares: the exchanged system must equal it.
    config3 = None
    if not args.no_config3:
        from dsopp_b200 import synth
        win3 = synth.make_window(n_frames=N_FRAMES, points_per_frame=20000, seed=1, ab_scale=0.0)
        h3 = capi.upload_window(win3, device=local, rank=rank, world_size=world)
        shard3 = [np.arange(rank, len(f.idepth), world) for f in win3.frames]
        eps3 = np.concatenate([f.state_eps for f in win3.frames])
        if world > 1:
            uid3 = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                uid3.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid3, 0)
            h3.comm_init(bytes(uid3.cpu().numpy().tobytes()), rank, world)
            if use_peer:
                capi.attach_peers(h3, rank, world, dev)
                h3.set_option("peer_exchange", 1)
                h3.set_option("peer_fused", 0)
        h3.first_estimate()
        sys3 = h3.linearize(SIGMA, True, True, False)  # also the first collective of this communicator (outside capture)
        check3 = None
        if world > 1 and rank == 0:
            href = capi.upload_window(win3, device=local)
            href.first_estimate()
            ref3 = href.linearize(SIGMA, True, True, False)
            href.close()
            errs = {nm: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
                    for nm, a, b in zip(("H_pose", "b_pose", "H_schur", "b_schur"), sys3, ref3)}
            # the per-chunk partial sums are fp32 (other landmarks share a chunk when the window is sharded), their sum
            # over chunks and ranks is fp64: agreement to fp32 partial-sum noise, 1e-6 of the largest entry
            check3 = {"sharded_vs_unsharded_rel_max": errs, "ok": bool(max(errs.values()) <= 1e-6), "tolerance": 1e-6,
                      "what": "rank 0: dpba_linearize of the window sharded over N ranks (exchanged sum) against the same "
                              "window unsharded on one GPU"}
        stream3 = torch.cuda.ExternalStream(h3.stream, device=dev)

        def reset3():
            for i, f in enumerate(win3.frames):
                h3.set_landmarks(i, f.uv[shard3[i]], f.idepth[shard3[i]], f.patch[shard3[i]], f.flags[shard3[i]])
            for (r_, t_), st_ in win3.statuses.items():
                h3.set_statuses(r_, t_, st_[shard3[r_]])
            h3.set_state(eps3, np.zeros_like(eps3))

        def solve3():
            h3.first_estimate()
            return h3.solve_lm(SIGMA, AB_REG, FIXED_REG, max_it=GN_ITERS, min_it=GN_ITERS, ftol=0.0, ptol=0.0, force_accept=True,
                               lambda0=1e-5)

        ms3, res3 = [], None
        for i in range(3 + 10):
            reset3()
            flush.fill_(1)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream3)
            res3 = solve3()
            a1.record(stream3)
            torch.cuda.synchronize()
            if i >= 3:
                ms3.append(a0.elapsed_time(a1))
        t3 = torch.tensor([sum(ms3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        ms_step3 = float(t3.item()) / len(ms3)
        config3 = {"workload": f"configs[3]: {N_FRAMES} keyframes x 20000 points/KF ({win3.units} patch-residuals), landmarks "
                               f"sharded over {world} GPU(s) (strong scaling), one exchange of the reduced 8N x 8N system per GN iteration",
                   "value": win3.units * GN_ITERS / (ms_step3 * 1e-3), "unit": UNIT, "ms_per_step": ms_step3,
                   "us_per_gn_iter": 1e3 * ms_step3 / GN_ITERS, "steps": len(ms3), "energy": float(res3[0]),
                   "iterations": int(res3[1]), "n_gpus": world, "scaling": "strong", "multi_gpu_check": check3}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        h3.close()
        del win3

    # ---- configs[2]: coarse-tracker direct image alignment, 640x480, SE3 + affine (6 + 2), full-frame residual, 4 pyramid
    # levels coarse to fine (kNumberOfPyramidLevels = 4; monocular_tracker.cpp:199-214).  GPU: one kernel launch per level
    # runs the whole LM solve (whole-chip cooperative kernel for the dense levels); CPU: the serial C++ restatement of
    # EigenPoseAlignment (the reference runs it on one core), timed on this box.  Also the reference-realistic SPARSE depth
    # map (2 % of the pixels carry depth).  Rank 0 only.
    config2 = None
    if rank == 0 and not args.no_config2:
        try:
            config2 = pose_alignment_leg(torch)
        except Exception as ex:  # reported, never required for the headline metric
            config2 = {"error": repr(ex)}

    config4 = None
    if rank == 0 and world == 1 and not args.no_config4:
        try:
            config4 = marginalisation_leg()
        except Exception as ex:
            config4 = {"error": repr(ex)}

    def teardown():
        # ordered shutdown, then a normal interpreter exit: the handle first (dpba_destroy destroys the captured LM graph
        # BEFORE its NCCL communicator -- a graph that still references the communicator's kernels is what used to block
        # the exit), then torch's process group.  All ranks have finished their work when they get here.
        sys.stdout.flush()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        h.close()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return

    peaks, peak_kind = measured_peaks()
    W, H = win.width, win.height
    D = 8 * n
    img_bytes = 12 * n * W * H
    fused_ms, fused_n = prof["linearize_fused"]
    fused_avg = fused_ms / max(fused_n, 1)
    fused_bytes = 51 * units_local + img_bytes + 2 * (D * D + D) * 8
    achieved = fused_bytes / (fused_avg * 1e-3) / 1e9 if fused_avg > 0 else 0.0
    # FLOPs of one launch from the committed SASS mix of the same configuration (profiles/r02b_k_linearize_fused2.md: FFMA
    # 3.46 M, FMUL 1.50 M, FADD 1.15 M warp instructions per launch at 112 000 patch-residuals): 2736 FLOP per patch-residual
    FLOP_PER_UNIT = (3462640 * 64 + 1500152 * 32 + 1148864 * 32) / 112000.0
    FP32_PEAK_TFLOPS = 72.6  # measured on this pool's B200 with tools/fp32_probe.cu (profiles/r02a_fp32_probe.txt)
    fused_tflops = FLOP_PER_UNIT * units_local / (fused_avg * 1e-3) / 1e12 if fused_avg > 0 else 0.0
    roofline = {"kernel": "k_linearize_fused2 (K1+K3+K4a, one thread per patch-residual, nothing materialised)", "bound": "hbm",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                # dram__bytes_read + dram__bytes_write of one launch from the committed ncu --set full capture of this
                # configuration (profiles/r02b_k_linearize_fused2.md; cold L2: the 79 MB image set is read in part)
                "traffic": (60.92e6 + 1.85e6) if (world == 1 and N_FRAMES == 8 and PTS_PER_GPU == 2000) else None,
                "traffic_source": "profiles/r02b_k_linearize_fused2.md",
                "peak_source": f"{peak_kind} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
                "algorithmic_bytes_per_launch": fused_bytes, "avg_launch_ms": fused_avg, "launches_timed": fused_n,
                "note": "the fused linearise is bound by fp32 issue + L2 gather latency, not HBM (SURVEY 8d: ~55 FLOP/B): see "
                        "roofline_fp32 for the other ceiling and roofline_sweep for the HBM-bound materialising sweep the 60% "
                        "target is stated on (a kernel that is NOT on the production solve path)"}
    roofline_fp32 = {"kernel": roofline["kernel"], "bound": "fp32 CUDA-core issue", "achieved": fused_tflops,
                     "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": fused_tflops / FP32_PEAK_TFLOPS,
                     "peak_source": "measured: tools/fp32_probe.cu, 148 SMs x 125 FFMA/clk (profiles/r02a_fp32_probe.txt)",
                     "flop_per_patch_residual": FLOP_PER_UNIT,
                     "flop_source": "SASS opcode mix of the committed capture (FFMA x 2 + FMUL + FADD), profiles/r02b_k_linearize_fused2.md",
                     "issue_slots_note": "10.9 M warp instructions per launch = 9.4 us at one instruction per scheduler per "
                                         "clock; the kernel runs at 0.36 of that rate while active (long-scoreboard: L2 gathers)"}
    sweep_bytes = 595 * units_local + img_bytes
    sweep_ach = sweep_bytes / (sweep * 1e-3) / 1e9 if sweep else 0.0
    roofline_sweep = {"kernel": "k_materialise_sweep (K1, reference-surface mode, 595 B/unit)", "bound": "hbm",
                      "achieved": sweep_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                      "frac": sweep_ach / peaks["hbm_gbs"],
                      "traffic": (60.51e6 + 22.97e6) if (N_FRAMES == 8 and PTS_PER_GPU == 2000) else None,  # profiles/r01h_k_materialise_sweep.md (65 MB of the stores are still in L2 when the kernel ends)
                      "algorithmic_bytes_per_launch": sweep_bytes,
                      "avg_launch_ms": sweep, "timing": "alone, L2 flushed before each launch, 10 launches",
                      "workload": f"configs[1], {units_local} patch-residuals"}
    roofline_sweep_big = None
    if sweep4:
        b4 = 595 * units4 + img_bytes
        a4 = b4 / (sweep4 * 1e-3) / 1e9
        roofline_sweep_big = {"kernel": "k_materialise_sweep (K1, reference-surface mode, 595 B/unit)", "bound": "hbm",
                              "achieved": a4, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a4 / peaks["hbm_gbs"],
                              "traffic": 155.45e6 + 600.48e6,  # profiles/r01h_k_materialise_sweep_big.md
                              "algorithmic_bytes_per_launch": b4, "avg_launch_ms": sweep4,
                              "timing": "alone, L2 flushed before each launch, 10 launches",
                              "workload": f"configs[3] per-GPU shape at G=1: 8 KF x 20000 points, {units4} patch-residuals",
                              "patch_residuals_per_s": units4 / (sweep4 * 1e-3),
                              "hbm_write_only_gbs_measured": write_gbs,
                              "note": "traffic = 750 MB per launch (ncu, profiles/): 600 MB of stores + 150 MB of loads; a pure "
                                      "store stream (torch fill of 1 GiB) reaches hbm_write_only_gbs_measured on this device, "
                                      "so launch time ~ stores / write rate + loads / read rate"}
    kernel_ms = {k: {"ms_total": v[0], "launches": v[1]} for k, v in prof.items() if v[1]}

    # ---- CPU baseline on this box's host cores (bounded sample) ----------------------------------------
    cpu = cpu_baseline_leg() if world == 1 and not args.no_cpu else None
    # the timed step against the CPU restatement (double) of the same step: 7 forced GN iterations from the same window.
    # Tolerances as in tests/test_gpu_baseline_sizes.py (fp32 device arithmetic vs float64): energy 2e-5 relative, pose
    # increments / affine gain 2e-5 absolute, affine offset 1e-4 (intensities 0..255 in fp32)
    parity_check = None
    if cpu and cpu.get("_first_step"):
        fs = cpu.pop("_first_step")
        d = np.abs(eps_dev - fs["eps"]).reshape(n, 8)
        rel_e = abs(energy_dev - fs["energy"]) / abs(fs["energy"])
        rel_e2e = abs(e2e_energy - fs["energy"]) / abs(fs["energy"])
        parity_check = {"against": "oracle/cpu_ref double, same window, same 7 forced GN iterations",
                        "energy_device": energy_dev, "energy_e2e_call": e2e_energy, "energy_cpu_ref": fs["energy"],
                        "energy_rel_diff": rel_e, "energy_rel_diff_e2e": rel_e2e,
                        "state_max_abs_diff_pose_and_a": float(d[:, :7].max()), "state_max_abs_diff_b": float(d[:, 7].max()),
                        "ok": bool(rel_e <= 2e-5 and rel_e2e <= 2e-5 and d[:, :7].max() <= 2e-5 and d[:, 7].max() <= 1e-4)}
    elif cpu:
        cpu.pop("_first_step", None)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(world, {"exchange": "NVLink mailbox all-reduce (peer_exchange.cu)" if world > 1 and use_peer
                                      else ("ncclAllReduce of the packed system, one per GN iteration" if world > 1 else "none")}),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms_per_step,
                "path": "dpbah_solve_window (C++ host library, one call per step): dpba_remove_frame / push_frame_intensity "
                        "(every keyframe's intensity plane -- PixelMap::data() -- from pinned host memory, {I,dx,dy} built on the "
                        "device with the reference's gradient definition) / set_landmarks / set_frame_statuses / set_state, "
                        "dpba_first_estimate + dpba_solve_lm, dpba_get_state / get_landmarks / get_frame_statuses into host arrays"},
        "e2e_pixelinfo_records": dict(e2e_rec, path="as e2e, but the 12-byte {I,dx,dy} records of every keyframe are uploaded "
                                                     "(dpba_push_frame): the e2e of round 1 and of the round-2 lines before this one"),
        "e2e_raw_frames": {"value": e2e_raw_value, "unit": UNIT, "ms_per_step": e2e_raw_ms, "h2d_bytes_per_step": h2d_raw,
                           "path": "as e2e, but dpba_push_frame_raw: 8-bit frames in, photometric table + {I,dx,dy} on the device"},
        "e2e_one_new_keyframe": e2e_sliding,
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_fp32": roofline_fp32, "roofline_sweep": roofline_sweep,
        "roofline_sweep_big": roofline_sweep_big,
        "kernel_ms": kernel_ms,
        "cpu_baseline": cpu,
        "parity_check": parity_check,
        "config3_strong": config3,
        "config2_pose_alignment": config2,
        "config4_marginalisation": config4,
        "us_per_gn_iter": 1e3 * total_ms / args.steps / GN_ITERS,
    }
    print(json.dumps(out), flush=True)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-config4", action="store_true", help="skip the configs[4] marginalisation leg")
    ap.add_argument("--no-config2", action="store_true", help="skip the configs[2] pose-alignment leg")
    ap.add_argument("--no-config3", action="store_true", help="skip the configs[3] strong-scaling leg (8 KF x 20000 points)")
    ap.add_argument("--no-big-sweep", action="store_true", help="skip the 20000-points/KF materialising-sweep roofline")
    ap.add_argument("--fused-min-blocks", type=int, default=0, help="tuning A/B: 3 or 4 resident CTAs/SM for the fused linearise")
    ap.add_argument("--fused-version", type=int, default=0, help="A/B: 1 = first-generation fused linearise (8 lanes per patch), 2 = one thread per patch-residual")
    ap.add_argument("--three-branch", type=int, default=-1, help="A/B: 1 = three graph branches, energy decision from the sweep's records (default), 0 = the round-1 two-branch sequence")
    ap.add_argument("--pdl", type=int, default=-1, help="A/B: programmatic dependent launch between the device-LM kernels (default 1)")
    ap.add_argument("--merged-tail", type=int, default=-1, help="A/B: 1 = three launches per LM iteration (default), 0 = the eight-kernel sequence")
    ap.add_argument("--speculative-multi-gpu", type=int, default=-1, help="A/B: one-allreduce speculative device LM for N > 1")
    ap.add_argument("--peer-exchange", type=int, default=-1, nargs="?", const=1,
                    help="N > 1: 1 = sum the exchange block with the library's NVLink mailbox kernel, 0 = ncclAllReduce "
                         "(default: mailbox kernel from 4 GPUs on)")
    ap.add_argument("--peer-fused", type=int, default=-1, help="N > 1 with --peer-exchange: 1 = exchange fused into the producers / consumers (default), 0 = stand-alone mailbox kernel")
    ap.add_argument("--fused-prefetch", type=int, default=-1, help="tuning A/B: 0/1 L1 prefetch of the next taps in the fused linearise")
    ap.add_argument("--fused-epilogue", type=int, default=-1, help="A/B: 1 = second-generation epilogue of the fused sweep (default), 0 = the first generation's")
    ap.add_argument("--device-rendezvous", type=int, default=1, help="N > 1 with the mailbox exchange: start each timed step after a device-side rendezvous of the ranks (default 1)")
    ap.add_argument("--opt", action="append", default=[], help="A/B: name=value passed to dpba_set_option (repeatable)")
    ap.add_argument("--host-lm", action="store_true", help="drive the LM loop from the C++ host adapter instead of dpba_solve_lm")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()
    sys.stderr.flush()


if __name__ == "__main__":
    main()
