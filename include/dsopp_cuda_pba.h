/*
 * dsopp_cuda_pba.h -- C ABI of the B200-native photometric bundle-adjustment hot path.
 *
 * Drop-in boundary for RoadlyInc/DSOPP @ a4af2aa (paths relative to the reference's src/):
 * the device owns the LocalFrame-equivalent data and runs the data-parallel loops of
 *   energy/problems/internal/energy/problems/photometric_bundle_adjustment/  ("PBA/" below)
 * while the host keeps levenberg_marquardt_algorithm::solve, the priors and the 8Nx8N solve.
 * Every entry point replaces one reference interface, cited beside it.  The C++ adapter that
 * satisfies the reference's LevenbergMarquardtProblem concept on top of this header is
 * dsopp_b200/csrc/host/cuda_pba_problem.hpp; INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every call returns 0 on success or a negative DPBA_E_* code; dpba_last_error() gives text.
 *     Nothing throws, aborts or logs (reference: LOG(ERROR)+return / CHECK abort,
 *     energy/problems/src/photometric_bundle_adjustment.cpp:66-69,101).
 *   - the library COPIES landmarks, statuses and state when it is given them (LocalFrame copies them,
 *     PBA/local_frame.hpp:309-335): the caller may reuse those buffers as soon as the call returns (they are staged
 *     in pinned memory and DMA'd asynchronously; no set_* call synchronises the stream).  Images and masks in
 *     PAGE-LOCKED memory are BORROWED until the next synchronising call (dpba_solve_lm, dpba_linearize,
 *     dpba_evaluate, any dpba_get_*), as LocalFrame borrows its PixelMap pointers (PBA/local_frame.hpp:44,325);
 *     pageable images are copied before the call returns.
 *   - frames live in dense slots 0..N-1 in ascending timestamp order
 *     (photometric_bundle_adjustment.cpp:101); only slot 0 may be fixed
 *     (PBA/hessian_block_evaluation.hpp:143); single sensor, C = 1, pinhole + SE3
 *     (energy/problems/src/eigen_photometric_bundle_adjustment.cpp:64,155).
 *   - per-frame state block = [tx ty tz rx ry rz a b] (Sophus tangent order + affine brightness),
 *     matrices are row-major, poses are 3x4 row-major [R|t] world<-agent, doubles.
 *   - a handle is NOT thread-safe (the tracker is single-threaded around the solver).
 *   - landmark index == residual index == host landmark order (PBA/evaluate_jacobians.hpp:77).
 */
#ifndef DSOPP_CUDA_PBA_H_
#define DSOPP_CUDA_PBA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPBA_MAX_FRAMES 16 /* steady state is maximum_size+1 = 9 (standart.yaml), 16 with dense.yaml */
#define DPBA_PATTERN 8     /* common/pattern/include/common/pattern/pattern.hpp:17 */
#define DPBA_BLOCK 8       /* Motion::DoF + 2, PBA/eigen_photometric_bundle_adjustment_problem.hpp:259 */

/* track/connections/include/track/connections/frame_connection.hpp:19-25 */
enum { DPBA_OK_STATUS = 0, DPBA_OUTLIER = 1, DPBA_OCCLUDED = 2, DPBA_OOB = 3, DPBA_UNKNOWN = 4 };

/* LocalFrame::Landmark flags, PBA/local_frame.hpp:276-293 */
enum {
  DPBA_LM_MARGINALIZED = 1,
  DPBA_LM_TO_MARGINALIZE = 2,
  DPBA_LM_OUTLIER = 4,
  DPBA_LM_ILL_CONDITIONED = 8
};

enum {
  DPBA_SUCCESS = 0,
  DPBA_E_INVALID = -1, /* bad argument / precondition of the reference violated */
  DPBA_E_CUDA = -2,    /* CUDA runtime error, no CPU fallback exists */
  DPBA_E_CAPACITY = -3,
  DPBA_E_STATE = -4, /* call order (e.g. back_substitute before linearize) */
  DPBA_E_COMM = -5   /* multi-GPU exchange failed */
};

typedef struct dpba_handle dpba_handle;

typedef struct dpba_config {
  int32_t max_frames;           /* <= DPBA_MAX_FRAMES */
  int32_t max_points_per_frame; /* landmarks hosted by one frame ON THIS HANDLE (shard) */
  int32_t width, height;        /* level-0 image size */
  int32_t device;               /* CUDA ordinal */
  int32_t rank, world_size;     /* landmark shard: this handle holds landmarks l with l % world_size == rank */
} dpba_config;

/* ResidualPoint-shaped view of one (reference, target) pair, the four arrays that
 * BundleAdjustmentPhotometricCostFunctorAnalytic::Evaluate memcpy's
 * (energy/problems/internal/energy/problems/cost_functors/bundle_adjustment_photometric_cost_functor_analytic.hpp:46-68)
 * plus huber_weight / energy / statuses (PBA/local_frame.hpp:173-220).  Any pointer may be NULL. */
typedef struct dpba_residual_view {
  int32_t n;                     /* in: capacity (landmarks); out: landmarks written */
  float* residuals;              /* [n][8] */
  float* d_reference_state_eps;  /* [n][8][8] row-major, as ResidualPoint stores it */
  float* d_target_state_eps;     /* [n][8][8] row-major */
  float* d_idepth;               /* [n][8] */
  float* huber_weight;           /* [n] */
  float* energy;                 /* [n] */
  uint8_t* connection_status;            /* [n] */
  uint8_t* connection_status_candidate;  /* [n] */
} dpba_residual_view;

/* ---- lifetime -------------------------------------------------------------------------- */
int dpba_create(const dpba_config* cfg, dpba_handle** out);
int dpba_destroy(dpba_handle* h);
const char* dpba_last_error(const dpba_handle* h); /* never NULL */
const char* dpba_version(void);
/* cudaStream_t all kernels of this handle are launched on (for CUDA-event timing by the caller) */
void* dpba_stream(dpba_handle* h);

/* ---- window: PhotometricBundleAdjustment::pushFrame / LocalFrame ctor ------------------- */
/* Appends a frame at slot N (photometric_bundle_adjustment.cpp:98-106, PBA/local_frame.hpp:309-335).
 * image_I_dx_dy: HxWx3 float, the PixelMap<1> storage {I,dx,dy} interleaved
 * (features/include/features/camera/pixel_map.hpp:126-131); mask: HxW uchar or NULL (= all valid),
 * sensors/camera_calibration/include/sensors/camera_calibration/mask/camera_mask.hpp:48-89.
 * intr = fx, fy, cx, cy of the level-0 pinhole model.  Returns the slot (>= 0) or an error. */
int dpba_push_frame(dpba_handle* h, int32_t frame_id, const float* image_I_dx_dy, const uint8_t* mask,
                    const double T_w_agent_lin[12], double exposure_time, const double affine_brightness0[2],
                    const double intr[4], int32_t fixed);
/* As dpba_push_frame but from the raw intensity plane (HxW float): {I,dx,dy} is built on the device
 * with the reference's gradient definition (features/src/calculate_pixelinfo.cpp:340-374). */
int dpba_push_frame_intensity(dpba_handle* h, int32_t frame_id, const float* image_I, const uint8_t* mask,
                              const double T_w_agent_lin[12], double exposure_time,
                              const double affine_brightness0[2], const double intr[4], int32_t fixed);
/* As dpba_push_frame but from the RAW 8-bit gray frame: photometric correction (table of 256 floats or NULL = identity,
 * vignetting image or NULL; features/src/photometrically_corrected_image.cpp:9-29) and the {I,dx,dy} packing run on
 * the device -- 0.3 MB cross PCIe per VGA keyframe instead of 3.7 MB. */
int dpba_push_frame_raw(dpba_handle* h, int32_t frame_id, const uint8_t* gray, const float* photometric_calibration,
                        const uint8_t* vignetting, const uint8_t* mask, const double T_w_agent_lin[12],
                        double exposure_time, const double affine_brightness0[2], const double intr[4],
                        int32_t fixed);
/* PixelDataFrame (features/src/pixel_data_frame.cpp:12-31) on the device: corrected level 0, 2x2 box pyramid
 * (features/internal/features/camera/downscale_image.hpp:16-33, same summation order), gradients per level; level l is
 * written to out_I_dx_dy[l] as (H >> l) x (W >> l) x 3 floats (NULL entries are skipped).  levels is capped at 5
 * (kMaxPyramidDepth).  Bit-identical to the reference's float build.  Returns the number of levels. */
int dpba_build_pyramid(dpba_handle* h, const uint8_t* gray, const float* photometric_calibration,
                       const uint8_t* vignetting, int32_t levels, float* const* out_I_dx_dy);
/* frames.erase(...) of a marginalised frame (PBA/eigen_photometric_bundle_adjustment_problem.hpp:201-202);
 * later slots shift down by one. */
int dpba_remove_frame(dpba_handle* h, int32_t slot);
int dpba_num_frames(const dpba_handle* h);
/* Drains the handle's stream: every upload queued by the set_* / push_* calls has landed and borrowed page-locked
 * buffers may be reused (the reference has no counterpart: its pushFrame copies synchronously) */
int dpba_synchronize(dpba_handle* h);
/* relinearizeSystem (photometric_bundle_adjustment.cpp:311-316): new linearisation point / affine0, eps := 0 */
int dpba_set_frame_linearization(dpba_handle* h, int32_t slot, const double T_w_agent_lin[12],
                                 const double affine_brightness0[2]);
int dpba_set_frame_flags(dpba_handle* h, int32_t slot, int32_t fixed, int32_t to_marginalize);
/* LocalFrame::is_marginalized (PBA/local_frame.hpp:571; set from ActiveKeyframe::isMarginalized, :320): a marginalised
 * frame is skipped as a target by updatePointStatuses (photometric_bundle_adjustment.cpp:339,377) */
int dpba_set_frame_marginalized(dpba_handle* h, int32_t slot, int32_t is_marginalized);

/* Landmarks hosted by `slot` (LocalFrame::active_landmarks, PBA/local_frame.hpp:327-333): replaces all n
 * landmarks.  proj_xy [n][2], idepth [n], patch [n][8], flags [n] (DPBA_LM_*).  idepth_step := 0. */
int dpba_set_landmarks(dpba_handle* h, int32_t slot, int32_t n, const float* proj_xy, const float* idepth,
                       const float* patch, const uint8_t* flags);
/* LocalFrame::update appends freshly matured landmarks (PBA/local_frame.hpp:498-505) */
int dpba_append_landmarks(dpba_handle* h, int32_t slot, int32_t n, const float* proj_xy, const float* idepth,
                          const float* patch, const uint8_t* flags);
int dpba_set_landmark_flags(dpba_handle* h, int32_t slot, int32_t n, const uint8_t* flags);
int dpba_num_landmarks(const dpba_handle* h, int32_t slot);
/* Read back what updateFrame needs (photometric_bundle_adjustment.cpp:223-262).  Any pointer may be NULL. */
int dpba_get_landmarks(dpba_handle* h, int32_t slot, int32_t n, float* idepth, float* idepth_step,
                       float* inv_hessian_idepth_idepth, float* b_idepth, uint8_t* flags,
                       uint32_t* number_of_inlier_residuals, float* relative_baseline);
/* hessian_poses_idepth_block of the landmarks of `slot`: [n][8N] (PBA/local_frame.hpp:297) */
int dpba_get_pose_idepth_blocks(dpba_handle* h, int32_t slot, int32_t n, float* out);

/* connection statuses of the residual vector (ref_slot -> tgt_slot), one per host landmark
 * (ResidualPoint ctor, PBA/local_frame.hpp:212-213: status and candidate both := given) */
int dpba_set_statuses(dpba_handle* h, int32_t ref_slot, int32_t tgt_slot, int32_t n, const uint8_t* statuses);
int dpba_get_statuses(dpba_handle* h, int32_t ref_slot, int32_t tgt_slot, int32_t n, uint8_t* statuses,
                      uint8_t* candidates);
/* LocalFrame::update (PBA/local_frame.hpp:505-518): statuses of freshly matured landmarks are APPENDED to the residual
 * vector (ref_slot -> tgt_slot) from index `first` on; residuals [0, first) -- which the solver owns by then (kOutlier /
 * kOOB set by updatePointStatuses / changeResidualStatuses) -- are not touched */
int dpba_append_statuses(dpba_handle* h, int32_t ref_slot, int32_t tgt_slot, int32_t first, int32_t n,
                         const uint8_t* statuses);
/* Per-residual scalars the sweeps keep resident, for the residual vector (ref_slot -> tgt_slot): ResidualPoint::energy of
 * the last evaluation (PBA/local_frame.hpp:203, PBA/evaluate_jacobians.hpp:136-146,184-194) and
 * ResidualPoint::reprojection_jacobians_valid of the last firstEstimateJacobians pass (PBA/local_frame.hpp:197,
 * PBA/first_estimate_jacobians.hpp:52-54).  Either output may be NULL.  Plain device reads, no mode required. */
int dpba_get_residual_scalars(dpba_handle* h, int32_t ref_slot, int32_t tgt_slot, int32_t n, float* energy,
                              uint8_t* reprojection_jacobians_valid);

/* All residual vectors of one reference frame in one call -- what the LocalFrame ctor / LocalFrame::update do with
 * frame.connections() (PBA/local_frame.hpp:336-347,506-520).  per_target[t] -> [n] statuses towards slot t;
 * entry ref_slot is ignored and NULL entries are skipped.  The getter fills statuses[t] / candidates[t] (either
 * array, or single entries, may be NULL) with ONE stream synchronisation for the whole frame. */
int dpba_set_frame_statuses(dpba_handle* h, int32_t ref_slot, int32_t n, const uint8_t* const* per_target);
int dpba_get_frame_statuses(dpba_handle* h, int32_t ref_slot, int32_t n, uint8_t* const* statuses,
                            uint8_t* const* candidates);

/* The whole window in one call: dpba_set_landmarks for every slot and (when `statuses` is not NULL) dpba_set_frame_statuses
 * for every reference slot -- what PhotometricBundleAdjustment::pushFrame's LocalFrame constructors do frame by frame
 * (PBA/local_frame.hpp:314-347) when a tracker hands a window over.  n[f], uv[f], idepth[f], patch[f], flags[f] as in
 * dpba_set_landmarks (flags, or single entries of it, may be NULL); statuses[r * num_frames + t] -> [n[r]] statuses of the
 * residual vector (r -> t), NULL on the diagonal, NULL entries leave kOk.  The arrays are packed in device layout in the
 * handle's pinned arena and every device array is written by ONE asynchronous copy (5 DMAs instead of ~7 per frame); no
 * stream synchronisation, the caller's buffers are free on return. */
int dpba_set_window_landmarks(dpba_handle* h, const int32_t* n, const float* const* uv, const float* const* idepth,
                              const float* const* patch, const uint8_t* const* flags, const uint8_t* const* statuses);

/* state_eps / state_eps_step of all frames, [8N] each (PBA/local_frame.hpp:561-563) */
int dpba_set_state(dpba_handle* h, const double* state_eps, const double* state_eps_step);
int dpba_get_state(dpba_handle* h, double* state_eps, double* state_eps_step);

/* ---- the sweeps ------------------------------------------------------------------------ */
/* firstEstimateJacobians_ (PBA/first_estimate_jacobians.hpp:14-71): freezes the linearisation point used by
 * the FEJ reprojection Jacobians (idepth snapshot, T_t_r0, brightness scale, corrected intensities). */
int dpba_first_estimate(dpba_handle* h);

/* evaluateJacobians<..., EVALUATE_JACOBIANS=false, NEW_EVALUATION_POINT=true, APPLY_HUBER_LOSS=huber>
 * followed by calculateLandmarksEnergy (PBA/evaluate_jacobians.hpp:20-202;
 * PBA/eigen_photometric_bundle_adjustment_problem.hpp:93-144): residual-only sweep at state_eps+step,
 * idepth+idepth_step.  energy = sum over non-marginalised landmarks, n_valid = #(energy > 0). */
int dpba_evaluate(dpba_handle* h, double sigma_huber, int32_t huber, int32_t fej, double* energy,
                  int32_t* n_valid);

/* evaluateJacobians<..., true, true, huber> materialising every ResidualPoint (reference-surface mode):
 * residuals, d_reference_state_eps, d_target_state_eps, d_idepth, huber_weight, energy, candidates. */
int dpba_evaluate_jacobians(dpba_handle* h, double sigma_huber, int32_t huber, int32_t fej);
int dpba_download_residual_block(dpba_handle* h, int32_t ref_slot, int32_t tgt_slot, dpba_residual_view* view);

/* linearize(): evaluateJacobians<true,true,huber> + evaluateLinearSystemPosePose<for_marginalized> +
 * evaluateLinearSystemPoseDepthSchurComplement<for_marginalized>
 * (PBA/eigen_photometric_bundle_adjustment_problem.hpp:322-336; PBA/hessian_block_evaluation.hpp:96-236),
 * fused: no Jacobian is written to memory.  Outputs are [8N][8N] row-major / [8N] doubles, WITHOUT the priors
 * (evaluateLinearSystemPrior stays on the host).  With world_size > 1 the outputs are already summed over
 * ranks.  Also stores per landmark hessian_poses_idepth_block, b_idepth_block, inv_hessian_idepth_idepth,
 * ill_conditioned. */
int dpba_linearize(dpba_handle* h, double sigma_huber, int32_t huber, int32_t fej, int32_t for_marginalized,
                   double* H_pose, double* b_pose, double* H_schur, double* b_schur);
/* Same result through the reference's three-pass dataflow on the device: K1 materialise, then PosePose and
 * Schur passes that re-read it.  Exists to cross-check the fused path and to measure the materialising sweep. */
int dpba_linearize_materialized(dpba_handle* h, double sigma_huber, int32_t huber, int32_t fej,
                                int32_t for_marginalized, double* H_pose, double* b_pose, double* H_schur,
                                double* b_schur);

/* calculateIdepths (PBA/hessian_block_evaluation.hpp:238-263): step_pose is the solution of the reduced
 * system (the caller sets state_eps_step = -step_pose itself, as calculateStep does). */
int dpba_back_substitute(dpba_handle* h, const double* step_pose, double levenberg_marquardt_lambda);

/* acceptStep / rejectStep (PBA/eigen_photometric_bundle_adjustment_problem.hpp:366-402) for landmarks and
 * statuses AND the frame state held by the handle.  Norms include frames and landmarks; with
 * world_size > 1 the landmark part is summed over ranks. */
int dpba_accept(dpba_handle* h, double* state_squared_norm, double* step_squared_norm);
int dpba_reject(dpba_handle* h);
/* changeResidualStatuses(frames, accept) alone (PBA/eigen_photometric_bundle_adjustment_problem.hpp:20-35) */
int dpba_change_residual_statuses(dpba_handle* h, int32_t accept);
/* calculateLandmarksEnergy<for_marginalized> over the stored per-residual energies */
int dpba_landmarks_energy(dpba_handle* h, int32_t for_marginalized, double* energy, int32_t* n_valid);

/* updatePointStatuses (photometric_bundle_adjustment.cpp:322-406): 75-percentile energy + sigma^2/2
 * threshold, outlier reset, inlier counts, relative baseline, is_outlier. */
int dpba_update_point_statuses(dpba_handle* h, int32_t minimum_valid_reprojections, double sigma_huber,
                               double* energy_threshold);

/* ---- immature-landmark activation refine ------------------------------------------------- */
/* optimizeImmatureLandmark for n candidates hosted by frame `ref_slot`
 * (tracker/landmarks_activator/src/landmarks_activator.cpp:122-316): per candidate a 1-D Levenberg-Marquardt on the
 * inverse depth over all other frames of the window at their CURRENT state (tWorldAgent, affine brightness), lambda0 =
 * 0.1, <= 3 iterations.  proj_xy [n][2], idepth [n], patch [n][8].  Outputs (any may be NULL): refined idepth (-1 when
 * the problem stopped), activate = 1 (kActivate) / 0 (kDelete: fewer valid residuals than min(minimum_inliers, N - 1)
 * or negative idepth), valid residuals of the accepted state.  The frames must have been pushed (dpba_push_frame*). */
int dpba_refine_immature_landmarks(dpba_handle* h, int32_t ref_slot, int32_t n, const float* proj_xy,
                                   const float* idepth, const float* patch, int32_t minimum_inliers,
                                   double sigma_huber, float* idepth_out, uint8_t* activate,
                                   int32_t* number_of_valid_residuals);

/* ---- device-resident solve --------------------------------------------------------------- */
/* energy::levenberg_marquardt_algorithm::Options
 * (energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp:38-57) plus the
 * problem constants of PhotometricBundleAdjustmentProblem (PBA/eigen_photometric_bundle_adjustment_problem.hpp:271-284) */
typedef struct dpba_lm_options {
  int32_t max_num_iterations;
  int32_t min_num_iterations;
  int32_t force_accept;
  int32_t first_estimate_jacobians;
  double initial_levenberg_marquardt_regularizer;
  double function_tolerance;
  double parameter_tolerance;
  double levenberg_marquardt_regularizer_decrease_on_accept;
  double levenberg_marquardt_regularizer_increase_on_reject;
  double sigma_huber_loss;
  double affine_brightness_regularizer[2];
  double fixed_state_regularizer;
} dpba_lm_options;

/* levenberg_marquardt_algorithm::Result (:62-69) + the number of loop bodies executed */
typedef struct dpba_lm_result {
  double energy;
  int32_t number_of_valid_residuals;
  int32_t converged;
  int32_t iterations;
} dpba_lm_result;

/* levenberg_marquardt_algorithm::solve(problem, options) (:77-128) with PhotometricBundleAdjustmentProblem's
 * calculateEnergy / linearize / calculateStep / acceptStep / rejectStep (problem.hpp:290-402) executed ENTIRELY on
 * the device: sweeps, priors, marginalised-prior terms, the Jacobi-preconditioned LDL^T of the 8N x 8N system and
 * the accept/reject decisions, as one stream of kernel launches with a single host synchronisation at the end.
 * H_marg [8N][8N] / b_marg [8N] may be NULL (no marginalised prior).  The caller runs dpba_first_estimate first,
 * as EigenPhotometricBundleAdjustment::solve does.  On return the handle's frame state (dpba_get_state) is the
 * accepted state with zero step. */
int dpba_solve_lm(dpba_handle* h, const dpba_lm_options* options, const double* H_marg, const double* b_marg,
                  double energy_marg, dpba_lm_result* result);

/* createReferenceDepthMaps (src/tracker/tracker/src/create_depth_maps.cpp:122-146; the tracker calls it right after
 * every BA solve, monocular_tracker.cpp:465,509) from the window resident in the handle: every landmark of the older
 * keyframes with status kOk towards the NEWEST keyframe, neither outlier nor marginalised, is reprojected into the
 * newest keyframe and {idepth / depth_scale * w, w}, w = sqrt(1e-3 / (variance + 1e-12)), is accumulated at the rounded
 * pixel (:19-58); coarser levels are 2x2 sums (:70-88); empty interior pixels are dilated from their non-empty
 * neighbours (:90-120).  idepth_variance < 0 takes each landmark's inv_hessian_idepth_idepth (the reference with
 * estimate_uncertainty, photometric_bundle_adjustment.cpp:252), otherwise the given constant (1e-5 in the reference).
 * idepth_sum[l] / weight[l] (either array or any entry may be NULL): host buffers of (height >> l) * (width >> l)
 * floats, row-major [y][x] (= the reference's map(x, y)) -- the two planes dpa_set_reference_depth_map takes. */
int dpba_create_reference_depth_maps(dpba_handle* h, int32_t n_levels, double idepth_variance, float* const* idepth_sum,
                                     float* const* weight);

/* Tuning knobs without a reference counterpart.  "cuda_graph" (default 1): dpba_solve_lm replays its launch
 * sequence as one CUDA graph.  "speculative_linearize" (default 1): under force_accept (single GPU) dpba_solve_lm
 * evaluates the trial energy with the fused linearise, which then serves as the next iteration's linearize(); state,
 * idepths, statuses and energy are unchanged, the per-landmark hpd / b_d / inv_hdd left behind are those of the final
 * state (the reference's are one accepted step older; its uncertainty pass recomputes them).  "speculative_multi_gpu"
 * (default 1): the same with world_size > 1, the pair energies travelling in the scalar slots behind the system so that
 * ONE allreduce per iteration carries both.  "fused_epilogue" (default 1, process-wide): second-generation epilogue of
 * the fused sweep (reference block of H_pd formed per target inside the sweep, two block barriers instead of four); 0 selects
 * the first generation's.  "pixelinfo_tma" (default 0, process-wide): the {I,dx,dy} packing of uploaded frames stages its
 * tile with TMA (dpba_debug_pixelinfo_ab has the A/B).  "fused_prefetch" (default 0): L1 prefetch A/B switch.  "fused_min_blocks"
 * (3 or 4, process-wide): resident CTAs per SM the fused linearise is built for.  "schur_tensor_cores" (default 1, process-wide): the Schur-complement SYRK runs as
 * 3xTF32 mma.sync; 0 selects the fp32 FFMA kernel.  "peer_exchange" (default 0, needs dpba_peer_attach): the sum over
 * ranks runs as the library's own NVLink mailbox kernel instead of ncclAllReduce; inside dpba_solve_lm the 8 scalars of an
 * iteration (energies, norms) and its linear system then travel in two concurrent exchanges over two mailbox channels
 * ("split_exchange", default 1), so that the energy decision does not wait for the assembly of the system.  "peer_fused"
 * (default 0): the exchange rides in the producers' epilogues and the consumers' prologues instead (measured slower).
 * "peer_fence_all" (default 0): A/B switch, a system-scope fence in every pushing thread instead of one per signalling thread.  "device_quantile" (default 0):
 * dpba_update_point_statuses finds the 75 % energy quantile with an exact radix select on the device instead of
 * reading the rows back for std::nth_element. */
int dpba_set_option(dpba_handle* h, const char* name, int64_t value);

/* ---- measurement hooks (no reference counterpart) --------------------------------------- */
/* kernels launched by this process so far (every launch wrapper counts itself) */
int64_t dpba_launch_count(void);
/* Diagnostics: clock64() stamps taken by thread 0 at the phase boundaries of the single-CTA LM kernel (energy decision,
 * system fill, the block steps of the LDL^T, back substitution, pair constants) during the LAST launch; enable = 1 arms
 * the stamping (process-wide), out (may be NULL) receives the 64 stamps of the last launch. */
int dpba_debug_stamps(int32_t enable, int64_t out[64]);
/* Diagnostics: per-CTA timeline of the last fused sweep taken while the stamps were on -- out[4 c + {0,1,2,3}] = entry, end of
 * the sweep proper, end of the CTA (%globaltimer, ns) and the SM id of CTA c = blockIdx.y * gridDim.x + blockIdx.x; n <= 4096. */
int dpba_debug_cta_times(int64_t* out, int32_t n);
/* Diagnostics: out[2 k], out[2 k + 1] = entry of block 0 and exit of the latest block (%globaltimer, ns) of the last launch of
 * LM-loop kernel k (0 fused sweep, 1 core reduce, 2 energy decision, 3 Schur reduce, 4 block assembly, 5 LM step,
 * 6 back-substitution, 7 pair constants, 8 landmark accept) taken while the stamps were on. */
int dpba_debug_kernel_times(int64_t out[32]);
/* A/B of the {I, dx, dy} gradient packing (calculate_pixelinfo, features/src/calculate_pixelinfo.cpp:340-374) on a W x H
 * pseudo-random plane: variant 0 reads the stencil straight from global memory (k_pixelinfo, the default), variant 1 stages
 * the tile in shared memory with TMA (cp.async.bulk.tensor.2d + mbarrier, k_pixelinfo_tma; option "pixelinfo_tma" routes the
 * product path through it).  ms_per_launch[v] = device time per launch over `reps` launches (CUDA events, after warm-up);
 * mismatching_words = 32-bit words in which the two outputs differ (0: bit-identical). */
int dpba_debug_pixelinfo_ab(int32_t W, int32_t H, int32_t reps, double ms_per_launch[2], int64_t* mismatching_words);
/* Per-kernel device timing with CUDA events recorded on the handle's stream around each launch.
 * kinds: 0 fused linearise sweep, 1 Schur SYRK (three-pass path), 2 residual-only sweep, 3 materialising sweep,
 *        4 assemble+symmetrise (three-pass path), 5 back-substitution, 6 per-pair constants,
 *        7 device LM step (priors + LDL^T), 8 per-pair core reduction, 9 H_pp block assembly,
 *        10 Schur partial reduction, 11 LM control kernels (energy tail, accept, loop bookkeeping).
 *        dpba_profile_read synchronises, returns per kind the
 *        summed milliseconds and launch counts since the last dpba_profile_enable(h, 1), and keeps profiling on. */
#define DPBA_PROFILE_KINDS 12
int dpba_profile_enable(dpba_handle* h, int32_t on);
int dpba_profile_read(dpba_handle* h, double ms[DPBA_PROFILE_KINDS], int32_t launches[DPBA_PROFILE_KINDS]);

/* ---- multi-GPU (one process per GPU; landmarks sharded, frames replicated) -------------- */
/* 128-byte NCCL unique id created on rank 0, broadcast by the caller (torch.distributed / MPI / files). */
int dpba_comm_unique_id(uint8_t id[128]);
int dpba_comm_init(dpba_handle* h, const uint8_t id[128], int32_t rank, int32_t world_size);
/* Optional: sum the exchange block with the library's own one-shot kernel over NVLink peer memory instead of
 * ncclAllReduce (66.6 KB per GN iteration: latency, not bandwidth).  Every rank exports its mailbox (a 64-byte
 * cudaIpcMemHandle_t), the caller all-gathers the handles (rank-major, world_size * 64 bytes), every rank attaches,
 * the caller BARRIERS, then dpba_set_option(h, "peer_exchange", 1).  All ranks must make the same sequence of solver
 * calls (they do: frames are replicated); a rank that never arrives makes the others fail with DPBA_E_COMM after ~15 s
 * instead of hanging.  Barrier again before dpba_destroy.  One node (one NVSwitch domain), 2..8 ranks. */
int dpba_peer_export(dpba_handle* h, uint8_t ipc_handle[64]);
int dpba_peer_attach(dpba_handle* h, const uint8_t* ipc_handles, int32_t rank, int32_t world_size);
/* Device-side rendezvous of the attached ranks on the handle's stream (one mailbox exchange of the scalar slots): work
 * enqueued after it starts on every rank within a flag's flight time of the last rank's arrival.  bench.py uses it to start
 * the timed region of a sharded step on all ranks at once instead of after torch.distributed's host-side barrier, whose
 * exit jitter between processes (tens of microseconds) would otherwise be charged to a 0.6 ms step. */
int dpba_peer_barrier(dpba_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* DSOPP_CUDA_PBA_H_ */
