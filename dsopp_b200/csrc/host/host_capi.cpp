// Flat C shim over the C++ host side (CudaPhotometricBundleAdjustment, the LM driver and NormalLinearSystem) so
// that pytest / bench.py can drive exactly the code a C++ caller would link.  Not part of the drop-in boundary:
// that is include/dsopp_cuda_pba.h.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "cuda_photometric_bundle_adjustment.hpp"
#include "cuda_pose_alignment.hpp"

using namespace dsopp_b200;

namespace {
thread_local std::string g_err;
struct Solver {
  std::unique_ptr<CudaPhotometricBundleAdjustment> pba;
};

KeyframeView make_view(int id, long long ts, const double* T, double exposure, const double* ab, const double* intr,
                       const float* image, const uint8_t* mask, int is_marg, int n, const float* proj,
                       const float* idepth, const float* patch, const uint8_t* flags, int n_other,
                       const int* other_ids, const uint8_t* ref_statuses, const uint8_t* tgt_statuses,
                       const int* tgt_counts) {
  KeyframeView v;
  v.keyframe_id = id;
  v.timestamp = ts;
  std::memcpy(v.t_world_agent, T, sizeof(v.t_world_agent));
  v.exposure_time = exposure;
  v.affine_brightness[0] = ab[0];
  v.affine_brightness[1] = ab[1];
  std::memcpy(v.intrinsics, intr, sizeof(v.intrinsics));
  v.image_I_dx_dy = image;
  v.mask = mask;
  v.is_marginalized = is_marg != 0;
  v.n_landmarks = n;
  v.projections = proj;
  v.idepths = idepth;
  v.patches = patch;
  v.landmark_flags = flags;
  size_t off = 0;
  for (int i = 0; i < n_other; ++i) {
    if (ref_statuses) v.statuses_as_reference[other_ids[i]].assign(ref_statuses + (size_t)i * n, ref_statuses + (size_t)(i + 1) * n);
    if (tgt_statuses && tgt_counts) {
      v.statuses_as_target[other_ids[i]].assign(tgt_statuses + off, tgt_statuses + off + tgt_counts[i]);
      off += tgt_counts[i];
    }
  }
  return v;
}
}  // namespace

#define GUARD(...)                  \
  try {                             \
    __VA_ARGS__;                    \
  } catch (const DpbaFailure& e) {  \
    g_err = e.what();               \
    return e.code;                  \
  } catch (const std::exception& e) { \
    g_err = e.what();               \
    return -100;                    \
  }

extern "C" {

__attribute__((visibility("default"))) const char* dpbah_last_error() { return g_err.c_str(); }

__attribute__((visibility("default"))) void* dpbah_create(int width, int height, int max_frames, int max_points,
                                                         int device, int estimate_uncertainty, int force_accept,
                                                         int max_iterations, int min_iterations, double radius,
                                                         double ftol, double ptol, double ab_reg0, double ab_reg1,
                                                         double fixed_reg, double sigma, int device_lm) {
  try {
    TrustRegionPhotometricBundleAdjustmentOptions o;
    o.max_iterations = (size_t)max_iterations;
    o.min_iterations = (size_t)min_iterations;
    o.initial_trust_region_radius = radius;
    o.function_tolerance = ftol;
    o.parameter_tolerance = ptol;
    o.affine_brightness_regularizer[0] = ab_reg0;
    o.affine_brightness_regularizer[1] = ab_reg1;
    o.fixed_state_regularizer = fixed_reg;
    o.sigma_huber_loss = sigma;
    auto* s = new Solver();
    s->pba = std::make_unique<CudaPhotometricBundleAdjustment>(o, estimate_uncertainty != 0, force_accept != 0, width,
                                                               height, max_frames, max_points, device, device_lm != 0);
    return s;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

__attribute__((visibility("default"))) void dpbah_destroy(void* s) { delete (Solver*)s; }

__attribute__((visibility("default"))) void* dpbah_handle(void* s) { return ((Solver*)s)->pba->handle(); }

__attribute__((visibility("default"))) int dpbah_push_frame(
    void* s, int id, long long ts, const double* T, double exposure, const double* ab, const double* intr,
    const float* image, const uint8_t* mask, int is_marg, int n, const float* proj, const float* idepth,
    const float* patch, const uint8_t* flags, int fixed, int n_other, const int* other_ids,
    const uint8_t* ref_statuses, const uint8_t* tgt_statuses, const int* tgt_counts) {
  GUARD(((Solver*)s)->pba->pushFrame(
      make_view(id, ts, T, exposure, ab, intr, image, mask, is_marg, n, proj, idepth, patch, flags, n_other, other_ids,
                ref_statuses, tgt_statuses, tgt_counts),
      0, fixed ? FrameParameterization::kFixed : FrameParameterization::kFree));
  return 0;
}

__attribute__((visibility("default"))) int dpbah_update_local_frame(
    void* s, int id, long long ts, const double* T, double exposure, const double* ab, const double* intr,
    int is_marg, int n, const float* proj, const float* idepth, const float* patch, const uint8_t* flags, int n_other,
    const int* other_ids, const uint8_t* ref_statuses, const uint8_t* tgt_statuses, const int* tgt_counts) {
  GUARD(((Solver*)s)->pba->updateLocalFrame(make_view(id, ts, T, exposure, ab, intr, nullptr, nullptr, is_marg, n, proj,
                                                      idepth, patch, flags, n_other, other_ids, ref_statuses,
                                                      tgt_statuses, tgt_counts)));
  return 0;
}

__attribute__((visibility("default"))) int dpbah_solve(void* s, double* energy, int* iterations) {
  GUARD({
    auto& p = *((Solver*)s)->pba;
    const double e = p.solve(1);
    if (energy) *energy = e;
    if (iterations) *iterations = (int)p.lastIterations();
  });
  return 0;
}

__attribute__((visibility("default"))) int dpbah_marginalize_now(void* s) {
  GUARD(((Solver*)s)->pba->marginalizeNow());
  return 0;
}

__attribute__((visibility("default"))) int dpbah_num_frames(void* s) { return (int)((Solver*)s)->pba->frames().size(); }

__attribute__((visibility("default"))) int dpbah_frame_ids(void* s, int* ids) {
  const auto& f = ((Solver*)s)->pba->frames();
  for (size_t i = 0; i < f.size(); ++i) ids[i] = f[i].id;
  return (int)f.size();
}

__attribute__((visibility("default"))) int dpbah_update_frame(void* s, long long ts, double* T, double* ab, int n,
                                                             float* idepth, float* variance, float* baseline,
                                                             uint8_t* outlier, uint32_t* inliers) {
  GUARD({
    LandmarkResult r;
    std::map<int, std::vector<uint8_t>> st;
    ((Solver*)s)->pba->updateFrame(ts, T, ab, r, st);
    const int m = std::min<int>(n, (int)r.idepth.size());
    for (int l = 0; l < m; ++l) {
      if (idepth) idepth[l] = r.idepth[l];
      if (variance) variance[l] = r.idepth_variance[l];
      if (baseline) baseline[l] = r.relative_baseline[l];
      if (outlier) outlier[l] = r.is_outlier[l];
      if (inliers) inliers[l] = r.number_of_inlier_residuals[l];
    }
  });
  return 0;
}

__attribute__((visibility("default"))) int dpbah_marginalized_system(void* s, double* H, double* b, double* energy) {
  const auto& p = *((Solver*)s)->pba;
  const auto& m = p.systemMarginalized();
  if (H) std::memcpy(H, m.H.a.data(), m.H.a.size() * sizeof(double));
  if (b) std::memcpy(b, m.b.data(), m.b.size() * sizeof(double));
  if (energy) *energy = p.energyMarginalized();
  return m.size();
}

__attribute__((visibility("default"))) int dpbah_covariance(void* s, int ref_id, int tgt_id, double* out36) {
  const auto& c = ((Solver*)s)->pba->covariances();
  auto it = c.find({ref_id, tgt_id});
  if (it == c.end()) return -1;
  std::memcpy(out36, it->second.data(), 36 * sizeof(double));
  return 0;
}

// One whole solver step from HOST buffers in ONE call: drop the resident window, push every keyframe (images are DMA'd
// straight from the caller's buffers when those are page-locked), landmarks, connection statuses and state, run
// firstEstimateJacobians + the device-resident levenberg_marquardt_algorithm::solve, and write the results updateFrame
// reads (photometric_bundle_adjustment.cpp:182-264) back into the caller's arrays.  Exactly the C-ABI call sequence a
// host program makes for pushFrame x N / solve / updateFrame x N, without an interpreter between the calls; bench.py
// times it as `e2e`.  Every pointer is a host pointer; rows of `statuses` / `statuses_out` are indexed [r * n + t].
struct dpbah_window_io {
  int32_t n_frames;
  const int32_t* frame_ids;
  const float* const* images;       // [n] H*W*3 {I,dx,dy}
  const uint8_t* const* masks;      // [n] H*W (entries may be null)
  const double* T_w_lin;            // [n][12]
  const double* exposure;           // [n]
  const double* ab0;                // [n][2]
  const double* intr;               // [n][4]
  const int32_t* fixed;             // [n]
  const int32_t* n_landmarks;       // [n]
  const float* const* uv;           // [n] -> [m][2]
  const float* const* idepth;       // [n] -> [m]
  const float* const* patch;        // [n] -> [m][8]
  const uint8_t* const* flags;      // [n] -> [m]
  const uint8_t* const* statuses;   // [n*n] -> [m_r] (null on the diagonal)
  const double* eps0;               // [8n]
  dpba_lm_options lm;
  // results
  double* eps_out;                  // [8n]
  float* const* idepth_out;         // [n] -> [m]
  float* const* inv_hdd_out;
  float* const* rel_baseline_out;
  uint8_t* const* flags_out;
  uint32_t* const* n_inliers_out;
  uint8_t* const* statuses_out;     // [n*n] -> [m_r]
  double energy;
  int32_t iterations, n_valid, converged;
  int64_t h2d_bytes, d2h_bytes;     // counted from the arrays handed over / written back
  // optional: RAW 8-bit frames instead of `images` (dpba_push_frame_raw: photometric table, gradients on the device)
  const uint8_t* const* raw_gray;   // [n] H*W, or null
  const float* photometric_lut;     // [256]
  // host wall-clock per phase [ms]: 0 remove, 1 push frames, 2 landmarks + statuses + state, 3 firstEstimate + solve
  // (includes waiting for the queued uploads), 4 readback.  sync_phases != 0 drains the stream after every phase so
  // that the device time of the asynchronous uploads is attributed to its own phase (diagnostics only).
  double phase_ms[5];
  int32_t sync_phases;
  // 0 / 3: `images` are H*W*3 {I,dx,dy} records (the PixelMap's pixel-info storage); 1: `images` are H*W intensity planes
  // (what PixelMap::data() returns, pixel_map.hpp:117) and {I,dx,dy} is built on the device with the reference's gradient
  // definition (dpba_push_frame_intensity) -- a third of the bytes cross PCIe
  int32_t image_channels;
};

__attribute__((visibility("default"))) int dpbah_solve_window(dpba_handle* h, dpbah_window_io* io, int width, int height) {
  GUARD({
    const int n = io->n_frames;
    int64_t h2d = 0, d2h = 0;
    using clk = std::chrono::steady_clock;
    auto t_prev = clk::now();
    int phase = 0;
    auto mark = [&]() {
      if (io->sync_phases) dpba_check(h, dpba_synchronize(h));
      const auto t = clk::now();
      io->phase_ms[phase++] = std::chrono::duration<double, std::milli>(t - t_prev).count();
      t_prev = t;
    };
    while (dpba_num_frames(h) > 0) dpba_check(h, dpba_remove_frame(h, 0));
    mark();
    const size_t npx = (size_t)width * height;
    for (int f = 0; f < n; ++f) {
      if (io->raw_gray) {
        dpba_check(h, dpba_push_frame_raw(h, io->frame_ids[f], io->raw_gray[f], io->photometric_lut, nullptr, io->masks[f],
                                          io->T_w_lin + 12 * f, io->exposure[f], io->ab0 + 2 * f, io->intr + 4 * f,
                                          io->fixed[f]));
        h2d += (int64_t)npx + 1024 + (io->masks[f] ? (int64_t)npx : 0);
      } else if (io->image_channels == 1) {
        dpba_check(h, dpba_push_frame_intensity(h, io->frame_ids[f], io->images[f], io->masks[f], io->T_w_lin + 12 * f,
                                                io->exposure[f], io->ab0 + 2 * f, io->intr + 4 * f, io->fixed[f]));
        h2d += (int64_t)npx * 4 + (io->masks[f] ? (int64_t)npx : 0);
      } else {
        dpba_check(h, dpba_push_frame(h, io->frame_ids[f], io->images[f], io->masks[f], io->T_w_lin + 12 * f,
                                      io->exposure[f], io->ab0 + 2 * f, io->intr + 4 * f, io->fixed[f]));
        h2d += (int64_t)npx * 12 + (io->masks[f] ? (int64_t)npx : 0);
      }
    }
    mark();
    std::vector<double> zero(8 * (size_t)n, 0.0);
    // every landmark array and every residual vector's statuses in one call: one DMA per device array
    dpba_check(h, dpba_set_window_landmarks(h, io->n_landmarks, io->uv, io->idepth, io->patch, io->flags, io->statuses));
    for (int f = 0; f < n; ++f) h2d += (int64_t)io->n_landmarks[f] * (8 + 4 + 32 + 1) + (int64_t)io->n_landmarks[f] * (n - 1);
    dpba_check(h, dpba_set_state(h, io->eps0, zero.data()));
    h2d += 2 * 8 * (int64_t)n * 8;
    mark();
    dpba_check(h, dpba_first_estimate(h));
    dpba_lm_result r;
    dpba_check(h, dpba_solve_lm(h, &io->lm, nullptr, nullptr, 0.0, &r));
    io->energy = r.energy;
    io->iterations = r.iterations;
    io->n_valid = r.number_of_valid_residuals;
    io->converged = r.converged;
    mark();
    dpba_check(h, dpba_get_state(h, io->eps_out, zero.data()));
    d2h += 2 * 8 * (int64_t)n * 8;
    for (int f = 0; f < n; ++f) {
      const int m = io->n_landmarks[f];
      dpba_check(h, dpba_get_landmarks(h, f, m, io->idepth_out[f], nullptr, io->inv_hdd_out[f], nullptr, io->flags_out[f],
                                       io->n_inliers_out[f], io->rel_baseline_out[f]));
      dpba_check(h, dpba_get_frame_statuses(h, f, m, io->statuses_out + (size_t)f * n, nullptr));
      d2h += (int64_t)m * (4 * 4 + 1) + (int64_t)m * (n - 1);
    }
    mark();
    io->h2d_bytes = h2d;
    io->d2h_bytes = d2h;
  });
  return 0;
}

// The tracker's steady state in ONE call: the oldest keyframe leaves the window, ONE new keyframe is pushed from host
// buffers (image, landmarks, its connection statuses in both directions), then firstEstimateJacobians + the device LM and the
// results of EVERY frame come back -- marginalisation policy aside, this is monocular_tracker.cpp:491-507 per keyframe.
// The benchmark recycles the frame that left as the one that arrives, so the window always holds the same n keyframes (in
// rotating order) and every step does the same work.  `order` (n ints, owned by the caller, initialised 0..n-1) maps slot
// -> index into the io arrays and is rotated by the call.
__attribute__((visibility("default"))) int dpbah_solve_sliding(dpba_handle* h, dpbah_window_io* io, int32_t* order, int width,
                                                              int height) {
  GUARD({
    const int n = io->n_frames;
    if (dpba_num_frames(h) != n) throw DpbaFailure(DPBA_E_STATE, "dpbah_solve_sliding: load the window with dpbah_solve_window first");
    int64_t h2d = 0, d2h = 0;
    const size_t npx = (size_t)width * height;
    const int f = order[0];  // leaves as the oldest, arrives as the newest
    dpba_check(h, dpba_remove_frame(h, 0));
    for (int k = 0; k + 1 < n; ++k) order[k] = order[k + 1];
    order[n - 1] = f;
    dpba_check(h, dpba_set_frame_flags(h, 0, 1, 0));  // the new oldest keyframe holds the gauge (frame 0 is the fixed one)
    if (io->image_channels == 1)
      dpba_check(h, dpba_push_frame_intensity(h, io->frame_ids[f], io->images[f], io->masks[f], io->T_w_lin + 12 * f,
                                              io->exposure[f], io->ab0 + 2 * f, io->intr + 4 * f, 0));
    else
      dpba_check(h, dpba_push_frame(h, io->frame_ids[f], io->images[f], io->masks[f], io->T_w_lin + 12 * f, io->exposure[f],
                                    io->ab0 + 2 * f, io->intr + 4 * f, 0));
    h2d += (int64_t)npx * (io->image_channels == 1 ? 4 : 12) + (io->masks[f] ? (int64_t)npx : 0);
    // The new frame's landmarks and its residual vectors in both directions; and -- same starting state every step, the
    // benchmark's steps must do the same work -- the landmarks and statuses of the frames that stayed and the pose
    // increments go back to the initial estimate: the whole window in one call, pointers permuted into slot order.
    std::vector<double> eps(8 * (size_t)n), zero(8 * (size_t)n, 0.0);
    std::vector<int32_t> cnt((size_t)n);
    std::vector<const float*> uv((size_t)n), idp((size_t)n), pat((size_t)n);
    std::vector<const uint8_t*> flg((size_t)n), sts((size_t)n * n, nullptr);
    for (int t = 0; t < n; ++t) {
      const int g = order[t];
      for (int k = 0; k < 8; ++k) eps[8 * t + k] = io->eps0[8 * g + k];
      cnt[t] = io->n_landmarks[g], uv[t] = io->uv[g], idp[t] = io->idepth[g], pat[t] = io->patch[g], flg[t] = io->flags[g];
      for (int u = 0; u < n; ++u)
        if (u != t) sts[(size_t)t * n + u] = io->statuses[(size_t)g * n + order[u]];
    }
    dpba_check(h, dpba_set_window_landmarks(h, cnt.data(), uv.data(), idp.data(), pat.data(), flg.data(), sts.data()));
    {
      const int m = io->n_landmarks[f];
      h2d += (int64_t)m * (8 + 4 + 32 + 1);
      for (int t = 0; t + 1 < n; ++t) h2d += m + io->n_landmarks[order[t]];
    }
    dpba_check(h, dpba_set_state(h, eps.data(), zero.data()));
    h2d += 2 * 8 * (int64_t)n * 8;
    dpba_check(h, dpba_first_estimate(h));
    dpba_lm_result r;
    dpba_check(h, dpba_solve_lm(h, &io->lm, nullptr, nullptr, 0.0, &r));
    io->energy = r.energy;
    io->iterations = r.iterations;
    io->n_valid = r.number_of_valid_residuals;
    io->converged = r.converged;
    dpba_check(h, dpba_get_state(h, io->eps_out, zero.data()));
    d2h += 2 * 8 * (int64_t)n * 8;
    for (int t = 0; t < n; ++t) {
      const int g = order[t], mg = io->n_landmarks[g];
      dpba_check(h, dpba_get_landmarks(h, t, mg, io->idepth_out[g], nullptr, io->inv_hdd_out[g], nullptr, io->flags_out[g],
                                       io->n_inliers_out[g], io->rel_baseline_out[g]));
      uint8_t* rows[DPBA_MAX_FRAMES] = {};
      for (int u = 0; u < n; ++u) rows[u] = u == t ? nullptr : io->statuses_out[(size_t)g * n + order[u]];
      dpba_check(h, dpba_get_frame_statuses(h, t, mg, rows, nullptr));
      d2h += (int64_t)mg * (4 * 4 + 1) + (int64_t)mg * (n - 1);
    }
    io->h2d_bytes = h2d;
    io->d2h_bytes = d2h;
  });
  return 0;
}

// levenberg_marquardt_algorithm::solve over a window that was uploaded through the C ABI directly.
// trace (optional): per loop body [accepted, energy, n_valid, step(8N)...], row stride 3 + 8N.
__attribute__((visibility("default"))) int dpbah_lm_solve(dpba_handle* h, int n_frames, const double* ab0,
                                                         const int* fixed, double sigma, double ab_reg0,
                                                         double ab_reg1, double fixed_reg, int max_it, int min_it,
                                                         double ftol, double ptol, int force_accept, double lambda0,
                                                         double decrease, double increase, double* energy,
                                                         int* iterations) {
  GUARD({
    namespace lm = levenberg_marquardt_algorithm;
    std::vector<FrameMeta> frames(n_frames);
    for (int i = 0; i < n_frames; ++i) {
      frames[i].id = i;
      frames[i].fixed = fixed[i] != 0;
      frames[i].affine_brightness0[0] = ab0[2 * i];
      frames[i].affine_brightness0[1] = ab0[2 * i + 1];
    }
    const double reg[2] = {ab_reg0, ab_reg1};
    NormalLinearSystem marg(kBlockSize * n_frames);
    CudaPhotometricBundleAdjustmentProblem problem(h, frames, sigma, marg, 0.0, reg, fixed_reg, true);
    lm::Options o;
    o.max_num_iterations = (size_t)max_it;
    o.min_num_iterations = (size_t)min_it;
    o.function_tolerance = ftol;
    o.parameter_tolerance = ptol;
    o.force_accept = force_accept != 0;
    o.initial_levenberg_marquardt_regularizer = lambda0;
    o.levenberg_marquardt_regularizer_decrease_on_accept = decrease;
    o.levenberg_marquardt_regularizer_increase_on_reject = increase;
    dpba_check(h, dpba_first_estimate(h));
    const lm::Result r = lm::solve(problem, o);
    if (energy) *energy = r.energy;
    if (iterations) *iterations = (int)r.iterations;
  });
  return 0;
}

// The LM driver on a SCRIPTED problem (energies, valid counts and step norms come from arrays; every call is recorded):
// lets the tests compare its control flow call by call with the reference's own driver (tests/test_reference_parts.py).
// calls_out codes: 0 calculateEnergy, 1 linearize, 2 calculateStep, 3 acceptStep, 4 rejectStep.  Returns the call count.
__attribute__((visibility("default"))) int dpbah_lm_scripted(int max_it, double lambda0, double ftol, double ptol,
                                                            int force_accept, int min_it, double decrease,
                                                            double increase, const double* energies, const int* valid,
                                                            int n_energy, const double* norms, int n_norms,
                                                            int* calls_out, int cap_calls, double* lambdas_out,
                                                            int cap_lambdas, double* energy_out, int* valid_out,
                                                            int* converged_out) {
  namespace lm = levenberg_marquardt_algorithm;
  struct Scripted {
    const double* e;
    const int* v;
    int ne;
    const double* nr;
    int nn;
    int ie = 0, ia = 0;
    std::vector<int> calls;
    std::vector<double> lambdas;
    std::pair<lm::Precision, int> calculateEnergy() {
      calls.push_back(0);
      const int i = ie < ne ? ie : ne - 1;
      ++ie;
      return {e[i], v[i]};
    }
    void linearize() { calls.push_back(1); }
    void calculateStep(const lm::Precision lambda) {
      calls.push_back(2);
      lambdas.push_back(lambda);
    }
    std::pair<lm::Precision, lm::Precision> acceptStep() {
      calls.push_back(3);
      const int i = ia < nn ? ia : nn - 1;
      ++ia;
      return {nr[2 * i], nr[2 * i + 1]};
    }
    void rejectStep() { calls.push_back(4); }
    bool stop() { return false; }
  } p{energies, valid, n_energy, norms, n_norms, 0, 0, {}, {}};
  lm::Options o;
  o.max_num_iterations = (size_t)max_it;
  o.min_num_iterations = (size_t)min_it;
  o.function_tolerance = ftol;
  o.parameter_tolerance = ptol;
  o.force_accept = force_accept != 0;
  o.initial_levenberg_marquardt_regularizer = lambda0;
  o.levenberg_marquardt_regularizer_decrease_on_accept = decrease;
  o.levenberg_marquardt_regularizer_increase_on_reject = increase;
  const lm::Result r = lm::solve(p, o);
  for (int i = 0; i < (int)p.calls.size() && i < cap_calls; ++i) calls_out[i] = p.calls[i];
  for (int i = 0; i < (int)p.lambdas.size() && i < cap_lambdas; ++i) lambdas_out[i] = p.lambdas[i];
  *energy_out = r.energy;
  *valid_out = r.number_of_valid_residuals;
  *converged_out = r.converged ? 1 : 0;
  return (int)p.calls.size();
}

__attribute__((visibility("default"))) void dpbah_normal_solve(int n, const double* H, const double* b, double* x) {
  NormalLinearSystem s(n);
  std::memcpy(s.H.a.data(), H, (size_t)n * n * sizeof(double));
  std::memcpy(s.b.data(), b, n * sizeof(double));
  const dense::Vec r = s.solve();
  std::memcpy(x, r.data(), n * sizeof(double));
}

__attribute__((visibility("default"))) int dpbah_reduce_system(int n, double* H, double* b, int n_elim, const int* elim) {
  NormalLinearSystem s(n);
  std::memcpy(s.H.a.data(), H, (size_t)n * n * sizeof(double));
  std::memcpy(s.b.data(), b, n * sizeof(double));
  s.reduce_system(std::vector<int>(elim, elim + n_elim));
  std::memcpy(H, s.H.a.data(), s.H.a.size() * sizeof(double));
  std::memcpy(b, s.b.data(), s.b.size() * sizeof(double));
  return s.size();
}

__attribute__((visibility("default"))) void dpbah_sym_pinv(int n, const double* A, int n_null, double* out) {
  dense::Mat M(n, n);
  std::memcpy(M.a.data(), A, (size_t)n * n * sizeof(double));
  const dense::Mat P = dense::sym_pinv(M, n_null);
  std::memcpy(out, P.a.data(), (size_t)n * n * sizeof(double));
}

}  // extern "C"

// ---- CudaPoseAlignment (coarse tracker), the call sequence of monocular_tracker.cpp:199-214 ------------------------
extern "C" {
__attribute__((visibility("default"))) void* dpah_create(int max_width, int max_height, int max_iterations, double sigma,
                                                         double reg_a, double reg_b) {
  try {
    dsopp_b200::PoseAlignmentOptions o;
    o.max_iterations = (size_t)max_iterations;
    o.sigma_huber_loss = sigma;
    o.affine_brightness_regularizer[0] = reg_a;
    o.affine_brightness_regularizer[1] = reg_b;
    return new dsopp_b200::CudaPoseAlignment(o, max_width, max_height);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
__attribute__((visibility("default"))) void dpah_destroy(void* s) { delete (dsopp_b200::CudaPoseAlignment*)s; }
// one level: reset(); pushFrame(reference, depth map); pushFrame(target); solve().  Returns rmse (< 0: error / kZeroCost)
__attribute__((visibility("default"))) double dpah_align_level(
    void* s, int width, int height, const double* intr, const float* ref_image, const float* idepth_sum, const float* weight,
    const double* ref_T, double ref_exposure, const double* ref_ab, const float* tgt_image, const uint8_t* tgt_mask,
    const double* tgt_T_guess, double tgt_exposure, const double* tgt_ab, const double* prior_rotation, double* T_out,
    double* ab_out, double* cov_out, int* n_landmarks) {
  try {
    auto& a = *(dsopp_b200::CudaPoseAlignment*)s;
    a.reset();
    if (prior_rotation) a.setRotationPrior(prior_rotation);
    dsopp_b200::AlignmentFrameView r, t;
    r.timestamp = 1;
    t.timestamp = 2;
    std::memcpy(r.t_world_agent, ref_T, sizeof(r.t_world_agent));
    std::memcpy(t.t_world_agent, tgt_T_guess, sizeof(t.t_world_agent));
    r.exposure_time = ref_exposure;
    t.exposure_time = tgt_exposure;
    for (int k = 0; k < 2; ++k) r.affine_brightness[k] = ref_ab[k], t.affine_brightness[k] = tgt_ab[k];
    for (int k = 0; k < 4; ++k) r.intrinsics[k] = t.intrinsics[k] = intr[k];
    r.width = t.width = width;
    r.height = t.height = height;
    r.image_I_dx_dy = ref_image;
    r.depth_idepth_sum = idepth_sum;
    r.depth_weight = weight;
    t.image_I_dx_dy = tgt_image;
    t.mask = tgt_mask;
    const int n = a.pushReferenceFrame(r);
    if (n_landmarks) *n_landmarks = n;
    a.pushTargetFrame(t);
    const double rmse = a.solve(1);
    std::memcpy(T_out, a.targetPose(), 12 * sizeof(double));
    ab_out[0] = a.targetAffineBrightness()[0];
    ab_out[1] = a.targetAffineBrightness()[1];
    if (cov_out) std::memcpy(cov_out, a.tTargetReferenceCovariance(), 36 * sizeof(double));
    return rmse;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -2.0;
  }
}
}
