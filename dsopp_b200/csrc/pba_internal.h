// Internal device-side data model and kernel launchers of the photometric BA path (sm_100a).
// Reference data model being replaced: LocalFrame / Landmark / ResidualPoint,
// src/energy/problems/internal/energy/problems/photometric_bundle_adjustment/local_frame.hpp:173-584.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PBA_MAXF 16
#define PBA_P 8
#define PBA_B 8
#define PBA_CORE 48  // 36 (upper triangle of the 8x8 core) + 8 (core^T r) padded to 48 doubles per ordered pair

// Per ordered pair (reference r, target t) constants, computed in double on the device from the frame
// state (k_pair_setup) and rounded to fp32.  Replaces the ArrayReprojector members
// (src/energy/projector/include/energy/projector/camera_reproject.hpp:235-260,369-377) and the per-pair
// prologue of evaluateJacobians (evaluate_jacobians.hpp:36-66).
struct __align__(16) PairConst {   // 512 bytes, every member 16-byte aligned for LDS.128 / LDG.128
  float A[12];     // reproject_            = K_t [R|t] Kr^-1  at the CURRENT state (eps + step)
  float M[12];     // transform_unproject_  = [R|t] Kr^-1      at the current state
  float M0[12];    // transform_unproject_ at the linearisation point (first_estimate_jacobians.hpp:28-32)
  float tr[4];     // translation_ at the current state (+ pad)
  float t0[4];     // translation_ at the linearisation point (+ pad)
  float adj[36];   // Adj(T_t_r)   row-major (rightLogTransformer, se3_motion.hpp:245)
  float adj0[36];  // Adj(T_t_r0)
  float s;         // brightness_change_scale at the current state (evaluate_jacobians.hpp:56-57)
  float s0;        // ... at the linearisation point (first_estimate_jacobians.hpp:36-37)
  float s0_last;   // s0 towards the LAST target of this reference frame (quirk Q1: corrected_intensities)
  float b_t;       // target affine shift, current
  float b_r;       // reference affine shift, current
  float b_r0;      // reference affine shift at the linearisation point
  float fx_t, fy_t;
  float cx_t, cy_t, pad0, pad1;
};
static_assert(sizeof(PairConst) == 512, "PairConst layout");

// blockdiag(Adj^T, 1, s') per pair in double for the final assembly (J_ref = U B^T, J_tgt = -U)
struct PairAssemble {
  double adj_fej[36];
  double adj_cur[36];
  double s0, s;
};

struct FrameParams {   // double-precision frame state on the device
  double T_lin[12];    // world<-agent 3x4 row-major
  double eps[8];
  double step[8];
  double exposure;
  double ab0[2];
  double intr[4];
};

struct WindowDev {
  int n_frames, W, H, max_pts, hpd_stride;
  int n_lm[PBA_MAXF];
  int fixed[PBA_MAXF];
  int frame_marg[PBA_MAXF];  // LocalFrame::is_marginalized
  int phys[PBA_MAXF];        // logical slot -> physical storage slot
  int mask_all[PBA_MAXF];    // 1: the frame's mask has no zero, the lookup can be skipped
  const float4* img[PBA_MAXF];     // per pixel x a 32-byte record {texel(x), texel(x+1)}, texel = {I, dx, dy, 0}
  const uint8_t* mask[PBA_MAXF];
  // landmark arrays, frame f at [phys[f] * max_pts, phys[f] * max_pts + n_lm[f])
  float4* lmk;             // {u, v, idepth, idepth at the FEJ linearisation point}: one LDG.128 per landmark
  float* idepth_step;
  const float* patch;      // [lm][8]
  uint8_t* flags;
  float* inv_hdd;
  float* b_d;
  float* hpd;              // [lm][hpd_stride], hpd_stride = 8 * n_frames
  float* rel_baseline;
  uint32_t* n_inliers;
  // per residual (r, t, l) -> ((phys[r] * PBA_MAXF + phys[t]) * max_pts + l)
  uint8_t* status;
  uint8_t* cand;
  uint8_t* jac_valid;      // ResidualPoint::reprojection_jacobians_valid of the FEJ pass (K6)
  float* energy;
  const PairConst* pairs;          // [PBA_MAXF * PBA_MAXF]
  const PairAssemble* pairs_asm;
  // materialised ResidualPoint arrays (reference-surface mode), may be null
  float* m_r;      // [res][8]
  float* m_jref;   // [res][8][8]
  float* m_jtgt;   // [res][8][8]
  float* m_did;    // [res][8]
  float* m_w;      // [res]
};

// reduction buffer (doubles), also the multi-GPU exchange buffer: [Hp | bp | Hs | bs | scal], D = 8 n_frames
struct ReduceBuf {
  double* Hp;   // assembled pose-pose H [D*D]
  double* bp;   // [D]
  double* Hs;   // Schur complement [D*D]
  double* bs;   // [D]
  double* scal; // [8]: 0 energy, 1 n_valid, 2 state_sq, 3 step_sq
  // first-stage partials (one slot per CTA / warp, never exchanged): hot-spot atomics are avoided on purpose --
  // 10^5 fp64 atomics into a few KB serialise in a handful of L2 slices and cost more than the sweep itself
  float* core_part;    // [host frame][chunk][target warp][PBA_CORE]
  double* core;        // [PBA_MAXF * PBA_MAXF][PBA_CORE] per ordered pair, summed over the chunks
  float* fschur_part;  // fused path: [host frame][chunk][ntri*16 + D] per-chunk Schur partial
  double* schur_part;  // stand-alone SYRK kernel: [CTA][D*D] upper triangle
  double* bs_part;     // [SYRK CTA][D]
  double* e_part;      // [residual-sweep CTA] (energy, n_valid)
  double* n_part;      // [back-substitution CTA] (state_sq, step_sq)
};
struct FusedShape {
  int lpb, chunks;  // landmarks per CTA and CTAs per host frame of the last k_linearize_fused launch
};

// ---- device-resident Levenberg-Marquardt (levenberg_marquardt_algorithm.hpp:77-128 on the device) --------------
struct LmOptionsDev {
  int max_it, min_it, force_accept, fej, huber;
  double lambda0, ftol, ptol, dec, inc, sigma;
  double ab_reg[2], fixed_reg, energy_marg;
};
struct LmCtl {
  double lambda, energy, next_energy;
  double state_sq, step_sq;
  int n_valid, next_n;
  int converged, done, system_valid, accept, iteration, iterations_executed;
  int apply;  // the last trial-energy kernel ran: k_accept_landmarks must apply ctl->accept
  int relin;  // speculative mode: a trial step was rejected after its linearisation overwrote the landmark fields
};

namespace pba {
// kinds for launch_lm_energy
enum { LM_ENERGY_INITIAL = 0, LM_ENERGY_TRIAL = 1, LM_ENERGY_FINAL = 2 };
void launch_lm_init(LmCtl* ctl, const LmOptionsDev* opt, cudaStream_t s);
// peer_collect (fused exchange): 1 = wait for every rank's scalars and sum them in rank order into `scal` first;
// 2 = the same and close the exchange (no calculateStep follows)
void launch_lm_energy(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, int N, double* scal,
                      const double* Hmarg, const double* bmarg, int kind, cudaStream_t s, const double* e_part = nullptr,
                      int n_e = 0, const double* n_part = nullptr, int n_n = 0, int from_core = 0, int peer_collect = 0);
// peer_expected > 0 (fused exchange): wait for that many system arrivals per rank, sum the ranks' blocks in rank order
// into rb (contiguous [H_pp | b_p | H_s | b_s]) and close the exchange
void launch_lm_step(const LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, const int* fixed, int N, ReduceBuf rb,
                    const double* Hmarg, const double* bmarg, double* step_dev, cudaStream_t s, int peer_expected = 0);
void launch_pair_setup(const FrameParams* frames, int n_frames, PairConst* pairs, PairAssemble* pasm, cudaStream_t s);
// round 2: energy decision + calculateStep + per-pair constants of the trial state in one single-CTA launch
void launch_lm_solve(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, const int* fixed, int N, double* scal,
                     ReduceBuf rb, const double* Hmarg, const double* bmarg, double* step_dev, int kind,
                     const double* e_part, int n_e, const double* n_part, int n_n, int from_core, int with_step,
                     PairConst* pairs, PairAssemble* pasm, cudaStream_t s);
// round 2: core reduction + block assembly + Schur partial reduction in one launch
void launch_reduce_system(const WindowDev& w, int fej, ReduceBuf rb, FusedShape shape, int with_system, cudaStream_t s,
                          const LmCtl* ctl = nullptr, int with_schur = 1);
// energy decision reading the pair energies straight from the sweep's per-chunk records (no k_core_reduce in front)
void launch_lm_energy_from_records(const WindowDev& w, LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, double* scal,
                                   const double* Hmarg, const double* bmarg, int kind, ReduceBuf rb, FusedShape shape,
                                   const double* n_part, int n_n, cudaStream_t s);
void launch_clear_frame_rows(uint8_t* status, uint8_t* cand, uint8_t* jac_valid, float* energy, int phys, int mp,
                             int max_frames, cudaStream_t s);
void launch_pack_image(const float* src3, float4* dst, int n_px, int W, cudaStream_t s);
void launch_pixelinfo(const float* I, float4* dst, int W, int H, cudaStream_t s);
// image_tma.cu: the same records with the intensity tile staged by TMA; false = no tensor map could be made, nothing launched
bool launch_pixelinfo_tma(const float* I, float4* dst, int W, int H, cudaStream_t s);
void set_pixelinfo_tma(bool on);
bool get_pixelinfo_tma();  // option "pixelinfo_tma": launch_pixelinfo routes through the TMA variant
void launch_photometric(const uint8_t* gray, const float* lut, const uint8_t* vignetting, float max_v, float* out, int n,
                        cudaStream_t s);
void launch_downscale(const float* src, float* dst, int W, int H, cudaStream_t s);
void launch_pixelinfo3(const float* I, float* dst, int W, int H, cudaStream_t s);
int launch_residual_sweep(const WindowDev& w, float sigma, int huber, int fej, double* e_part, cudaStream_t s,
                          const LmCtl* ctl = nullptr, int ctl_mode = 0);
// peer_push != 0 (fused exchange): the kernel also stores its results into every rank's mailbox and counts its arrival
void launch_reduce_scal(const LmCtl* ctl, int ctl_mode, const double* e_part, int n_e, const double* n_part, int n_n,
                        double* scal, cudaStream_t s, int core_frames = 0, int peer_push = 0);
void launch_materialise_sweep(const WindowDev& w, float sigma, int huber, int fej, cudaStream_t s);
// fold_step (second-generation kernel, device LM only): the pose step whose back-substitution -- together with the
// acceptStep / rejectStep of the previous trial -- this sweep performs for its own landmarks before evaluating; the
// per-CTA norm partials go to rb.n_part [n_frames * shape.chunks] (double2)
FusedShape launch_linearize_fused(const WindowDev& w, float sigma, int huber, int fej, int for_marg, ReduceBuf rb,
                                  cudaStream_t s, const LmCtl* ctl = nullptr, int ctl_mode = 2,
                                  const double* fold_step = nullptr);
int fused_version();
void debug_stamps(int enable, long long out[64]);
void debug_cta_times(long long* out, int n);
void debug_kernel_times(long long out[32]);
void stamps_off_async(cudaStream_t s);
void peer_stamps_off_async(cudaStream_t s);
void set_peer_fence_all(int v);
void debug_peer_times(int enable, long long out[4]);  // peer_exchange.cu: entry, latest exit, wait start, wait end of the last mailbox exchange
void launch_linearize_from_materialized(const WindowDev& w, int for_marg, ReduceBuf rb, cudaStream_t s);
int launch_schur(const WindowDev& w, int for_marg, ReduceBuf rb, cudaStream_t s, const LmCtl* ctl = nullptr);
void launch_assemble(const WindowDev& w, int fej, ReduceBuf rb, FusedShape shape, cudaStream_t s,
                     const LmCtl* ctl = nullptr);
void launch_core_reduce(const WindowDev& w, ReduceBuf rb, FusedShape shape, cudaStream_t s, const LmCtl* ctl = nullptr);
void launch_assemble_blocks(const WindowDev& w, int fej, ReduceBuf rb, FusedShape shape, cudaStream_t s,
                            const LmCtl* ctl = nullptr, int peer_push = 0);
// number of CTAs of launch_assemble_blocks + launch_finish_fused for a window of N frames (arrivals per rank, kind 1)
int system_producer_ctas(int n_frames);
void launch_finish_system(int D, ReduceBuf rb, int nsb, cudaStream_t s, const LmCtl* ctl = nullptr);
void launch_finish_fused(const WindowDev& w, ReduceBuf rb, FusedShape shape, cudaStream_t s, const LmCtl* ctl = nullptr,
                         int peer_push = 0);
void launch_back_substitute(const WindowDev& w, const double* step_pose_dev, double lambda, cudaStream_t s,
                            const LmCtl* ctl = nullptr, double* norms = nullptr);
void launch_accept(const WindowDev& w, int accept, double* scal, cudaStream_t s, const LmCtl* ctl = nullptr,
                   int with_statuses = 0);
void launch_change_statuses(const WindowDev& w, int accept, cudaStream_t s, const LmCtl* ctl = nullptr);
void launch_landmarks_energy(const WindowDev& w, int for_marg, double* scal, cudaStream_t s);
void launch_first_estimate(const WindowDev& w, cudaStream_t s);
void launch_refine_immature(const WindowDev& w, int r, int n, const float* proj, const float* idepth_in, const float* patch8,
                            int min_inliers, float sigma, float* idepth_out, uint8_t* activate, int* n_valid_out,
                            cudaStream_t s);  // K6: idepth snapshot + reprojection_jacobians_valid
void launch_apply_point_statuses(const WindowDev& w, float threshold, int min_valid, const float* pair_dist,
                                 cudaStream_t s);
int sm_count();
long long launch_count();
void add_launches(long long n);
void set_schur_mma(bool on);
void set_fused_min_blocks(int b);
void set_fused_version(int v);
void set_pdl(bool on);
void set_fused_prefetch(bool on);
void set_fused_epilogue(int v);
void set_fused_lpb_max(int v);
void set_fused_reduce_once(int v);
void set_fused2_min_blocks(int b);

// ---- peer-memory exchange (peer_exchange.cu): one-shot all-reduce over NVLink mailboxes -------------------------------
constexpr int PEER_MAXW = 8;   // ranks of one NVSwitch domain
constexpr int PEER_MAXC = 32;  // CTAs of one exchange kernel (each with its own arrival flag per source rank)
struct PeerDev {
  double* data[PEER_MAXW];    // mailbox data of rank r: [2 parities][PEER_MAXW sources][slot] doubles
  unsigned* flag[PEER_MAXW];  // mailbox flags of rank r: [PEER_MAXW sources][PEER_MAXC] epochs
  unsigned* seq;              // local: number of completed exchanges (the epoch of the last one)
  unsigned* done;             // local: CTAs of the running exchange that have finished
  int* error;                 // local device word: 1 after a time-out (later exchanges do not wait again)
  int* error_host;            // the same flag in mapped pinned host memory, for the host to read after a synchronisation
  int rank, world;
  size_t slot;                // doubles per [parity][source] slot
  // fused exchange (round 2): arrival COUNTERS of rank r's mailbox, [2 parities][2 kinds][PEER_MAXW sources]; every
  // producer CTA of source s adds one to its counter in every mailbox after its pushes (kind 0: the 8 scalars, kind 1:
  // the system [H_pp | b_p | H_s | b_s])
  unsigned* cnt[PEER_MAXW];
  const double* red_base;     // local exchange block the producers write: element index = pointer - red_base
};
// context of the fused exchange, copied to a __device__ symbol when the peers are attached (process-wide)
void set_peer_context(const PeerDev& pd);
void launch_peer_allreduce(const PeerDev& pd, const double* in, double* out, size_t off, size_t n, cudaStream_t s);

// ---- device-side quantile of updatePointStatuses (energy_quantile.cu) -------------------------------------------------
struct SelectState {
  unsigned hist[256];
  unsigned prefix, mask;     // bytes of the k-th key fixed so far
  unsigned long long k;      // rank of the wanted element among the keys that match the prefix
  unsigned count;            // residuals that took part
  float value;               // result: the k-th smallest energy
};
// `after_hist` (may be null) runs after every histogram pass, before the byte is picked: with several ranks it sums
// st->hist over the ranks (256 unsigned), which makes the select exact over the union of the shards
void launch_energy_quantile(const WindowDev& w, int nmax, double frac, SelectState* st, cudaStream_t s,
                            int (*after_hist)(void*) = nullptr, void* after_hist_arg = nullptr);

// ---- reference depth maps of the coarse tracker (depth_maps.cu) --------------------------------------------------------
size_t dm_level_offset(int W, int H, int level);
void launch_reference_depth_maps(const WindowDev& w, int n_levels, float const_var, float* buf, cudaStream_t s);
}  // namespace pba
