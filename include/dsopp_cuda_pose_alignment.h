/*
 * dsopp_cuda_pose_alignment.h -- C ABI of the B200-native coarse-tracker direct image alignment
 * (BASELINE.json configs[2]).
 *
 * Drop-in boundary for RoadlyInc/DSOPP @ a4af2aa (paths relative to the reference's src/): replaces
 *   EigenPoseAlignment<SE3, Pinhole, PatternSize = 1, PixelMap, C = 1, OPTIMIZE_AFFINE_BRIGHTNESS = true>
 *   (energy/problems/src/eigen_pose_alignment.cpp:28-329; interface
 *    energy/problems/include/energy/problems/pose_alignment/pose_alignment.hpp:22-63)
 * as MonocularTracker::estimatePose drives it per pyramid level
 * (tracker/tracker/src/monocular_tracker.cpp:199-214): reset(), pushFrame(reference with depth map, kFixed),
 * pushFrame(new frame, kFree), solve().  The host keeps the coarse-to-fine loop and the re-try schedule
 * (monocular_tracker.cpp:137-172,193-243); ONE call of dpa_solve runs the whole Levenberg-Marquardt solve of a
 * level on the device (a single thread-block cluster, no host round trip per iteration).
 *
 * Same conventions as dsopp_cuda_pba.h: 0 / negative DPBA_E_* codes, dpa_last_error() text, poses are 3x4 row-major
 * doubles, the library copies what it is given, a handle is not thread-safe.
 */
#ifndef DSOPP_CUDA_POSE_ALIGNMENT_H_
#define DSOPP_CUDA_POSE_ALIGNMENT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpa_handle dpa_handle;

typedef struct dpa_config {
  int32_t max_points;            /* landmarks of the reference depth map at the finest level used */
  int32_t max_width, max_height; /* level-0 image size */
  int32_t device;
} dpa_config;

/* TrustRegionPhotometricBundleAdjustmentOptions as createPoseAlignment fills it
 * (tracker/tracker/src/fabric.cpp:127-147) + the LM factors of EigenPoseAlignment::solve (:298-305) */
typedef struct dpa_options {
  int32_t max_num_iterations;          /* 50 */
  double initial_trust_region_radius;  /* 1e2 -> lambda0 = 1e-2 */
  double function_tolerance;           /* 1e-5 */
  double parameter_tolerance;          /* 1e-5 */
  double sigma_huber_loss;             /* monocular_tracker.hpp:55 */
  double affine_brightness_regularizer[2];
  double regularizer_decrease_on_accept; /* 2 */
  double regularizer_increase_on_reject; /* 2 */
} dpa_options;

typedef struct dpa_result {
  double rmse;   /* sqrt(energy / n_valid / PatternSize), eigen_pose_alignment.cpp:328 */
  double energy; /* levenberg_marquardt_algorithm::Result */
  int32_t number_of_valid_residuals;
  int32_t converged;
  int32_t iterations;                /* loop bodies executed */
  double T_target_reference[12];     /* t_t_r after the solve */
  double T_world_target[12];         /* reference.T_w * t_t_r^-1, what solve() stores in the target frame (:325) */
  double affine_brightness_eps[2];   /* added to the target's affine_brightness0 (:326) */
  double hessian[64];                /* problem.hessian(): the last linearised 8x8 system incl. priors (:320-323) */
} dpa_result;

int dpa_create(const dpa_config* cfg, dpa_handle** out);
int dpa_destroy(dpa_handle* h);
const char* dpa_last_error(const dpa_handle* h);
void* dpa_stream(dpa_handle* h);

/* pushFrame(reference keyframe, ..., reference_frame_depth_map, level, model, kFixed): the 1-pixel landmarks the
 * depth-map LocalFrame constructor creates (PBA/local_frame.hpp:367-392), in its order (y outer, x inner), given
 * directly.  xy [n][2], idepth [n], patch [n] (= I(x, y) of the reference level image). */
int dpa_set_reference_landmarks(dpa_handle* h, int32_t n, const float* xy, const float* idepth, const float* patch,
                                const double T_world_agent[12], double exposure_time,
                                const double affine_brightness[2], const double intr[4], int32_t width,
                                int32_t height);
/* Same, but the landmarks are built ON THE DEVICE from the depth-map accumulators create_depth_maps.cpp leaves
 * (idepth_sum and weight, [height][width] floats) and the reference level image ({I,dx,dy} interleaved): every pixel
 * inside the 4-px border with weight > 0 and idepth_sum / weight >= 1e-6, ordered y outer / x inner.  Returns the
 * number of landmarks (>= 0) or an error. */
int dpa_set_reference_depth_map(dpa_handle* h, const float* image_I_dx_dy, const float* idepth_sum,
                                const float* weight, const double T_world_agent[12], double exposure_time,
                                const double affine_brightness[2], const double intr[4], int32_t width,
                                int32_t height);
int dpa_num_landmarks(const dpa_handle* h);
/* read back the landmarks (any pointer may be NULL) */
int dpa_get_reference_landmarks(dpa_handle* h, int32_t n, float* xy, float* idepth, float* patch);

/* pushFrame(new frame, t_w_t guess, pyramids, masks, exposure, affine brightness, level, model, kFree) */
int dpa_set_target(dpa_handle* h, const float* image_I_dx_dy, const uint8_t* mask, const double T_world_agent[12],
                   double exposure_time, const double affine_brightness[2], const double intr[4], int32_t width,
                   int32_t height);

/* EigenPoseAlignment::solve (eigen_pose_alignment.cpp:275-329).  prior_rotation_t_r: 3x3 row-major or NULL
 * (setRotationPrior, :254-258). */
int dpa_solve(dpa_handle* h, const dpa_options* options, const double* prior_rotation_t_r, dpa_result* result);

/* calculateMeanSquareOpticalFlow (src/tracker/tracker/src/monocular_tracker.cpp:104-133; the tracker's keyframe
 * decision evaluates it on level 0 with the aligned pose and once more with the rotation set to identity, :474-480)
 * over the reference landmarks of the last dpa_set_reference_* call: sqrt(mean |unproject(x) - unproject(reproject(x))|^2)
 * over the landmarks that reproject successfully under T_target_reference (3x4 row-major).  *flow is NaN when none does
 * (0 / 0 in the reference); *n_used (may be NULL) is their number. */
int dpa_mean_square_optical_flow(dpa_handle* h, const double T_target_reference[12], double* flow, int32_t* n_used);
/* Trace of the last dpa_solve: trial energy, regulariser and accept decision of every executed loop body (at most 64).
 * Returns the number of entries written.  Any pointer may be NULL. */
int dpa_get_trace(dpa_handle* h, int32_t capacity, double* energies, double* lambdas, int32_t* accepted);
/* Tuning knob without a reference counterpart: from `min_points` reference landmarks on, dpa_solve runs the sweeps on a
 * cooperative grid with one CTA per SM (partial sums through global memory, grid barrier per sweep) instead of one
 * 8-CTA cluster (distributed shared memory).  Default 32768: the sparse depth maps of the reference (~6 k points) stay
 * on the cluster kernel, the dense raster of BASELINE configs[2] (298 k points) uses the whole chip.  min_points < 0:
 * never.  Same arithmetic, same summation order inside a CTA; the order ACROSS CTAs differs, so results agree to fp64
 * rounding of the partial sums, not bit for bit. */
int dpa_set_grid_threshold(dpa_handle* h, int32_t min_points);

#ifdef __cplusplus
}
#endif
#endif /* DSOPP_CUDA_POSE_ALIGNMENT_H_ */
