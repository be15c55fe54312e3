// Per-landmark body of the mean-square optical flow (pose_alignment.cu, dpa_mean_square_optical_flow) as a
// __host__ __device__ function: tests/emu runs it on the CPU against oracle/pose_alignment_oracle.py.
// Reference: calculateMeanSquareOpticalFlow, src/tracker/tracker/src/monocular_tracker.cpp:104-133.
#pragma once
#include <cuda_runtime.h>

namespace pba {

struct FlowConst {
  float A[12];        // reproject_ = K [R|t] K^-1 acting on [u, v, 1, rho]   (camera_reproject.hpp:256)
  float inv_fx, inv_fy;
  float xmax, ymax;   // W - 5, H - 5  (insideCameraROI with the 4-px border)
};

#ifdef __CUDA_ARCH__
#define OF_MUL(a, b) __fmul_rn((a), (b))
#define OF_ADD(a, b) __fadd_rn((a), (b))
#define OF_RCP(a) __frcp_rn(a)
#else
#define OF_MUL(a, b) ((a) * (b))
#define OF_ADD(a, b) ((a) + (b))
#define OF_RCP(a) (1.0f / (a))
#endif

// reproject_ = K [R|t] K^-1 (3x4 on [u, v, 1, rho]) for a 3x4 row-major T_target_reference and one pinhole camera for
// both frames, formed in double and rounded once to fp32
inline FlowConst make_flow_const(const double* T, const double* intr, int W, int H) {
  const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  double KT[12];
  for (int j = 0; j < 4; ++j) {
    KT[0 * 4 + j] = fx * T[0 * 4 + j] + cx * T[2 * 4 + j];
    KT[1 * 4 + j] = fy * T[1 * 4 + j] + cy * T[2 * 4 + j];
    KT[2 * 4 + j] = T[2 * 4 + j];
  }
  FlowConst c;
  for (int i = 0; i < 3; ++i) {
    c.A[i * 4 + 0] = (float)(KT[i * 4 + 0] / fx);
    c.A[i * 4 + 1] = (float)(KT[i * 4 + 1] / fy);
    c.A[i * 4 + 2] = (float)(KT[i * 4 + 2] - KT[i * 4 + 0] * cx / fx - KT[i * 4 + 1] * cy / fy);
    c.A[i * 4 + 3] = (float)KT[i * 4 + 3];
  }
  c.inv_fx = (float)(1.0 / fx);
  c.inv_fy = (float)(1.0 / fy);
  c.xmax = (float)(W - 5);
  c.ymax = (float)(H - 5);
  return c;
}

// lm = {x, y, idepth, -}: returns false when the scalar reproject fails (:121), else the squared ray difference (:122-125)
__host__ __device__ inline bool flow_term(const FlowConst& c, float4 lm, float& sq) {
  const float u = lm.x, v = lm.y, rho = lm.z;
  if (!(rho > -1e-4f && rho < 1010.f)) return false;
  if (!(u >= 4.f && v >= 4.f && u <= c.xmax && v <= c.ymax)) return false;
  const float X = OF_ADD(OF_ADD(OF_MUL(c.A[0], u), OF_MUL(c.A[1], v)), OF_ADD(c.A[2], OF_MUL(c.A[3], rho)));
  const float Y = OF_ADD(OF_ADD(OF_MUL(c.A[4], u), OF_MUL(c.A[5], v)), OF_ADD(c.A[6], OF_MUL(c.A[7], rho)));
  const float Z = OF_ADD(OF_ADD(OF_MUL(c.A[8], u), OF_MUL(c.A[9], v)), OF_ADD(c.A[10], OF_MUL(c.A[11], rho)));
  if (!(Z > 0.f)) return false;
  const float rz = OF_RCP(Z);
  const float tu = OF_MUL(X, rz), tv = OF_MUL(Y, rz);
  if (!(tu >= 4.f && tv >= 4.f && tu <= c.xmax && tv <= c.ymax)) return false;
  const float dx = (u - tu) * c.inv_fx, dy = (v - tv) * c.inv_fy;   // unproject: ((x - cx) / fx, (y - cy) / fy, 1)
  sq = dx * dx + dy * dy;
  return true;
}

}  // namespace pba
