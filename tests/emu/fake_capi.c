/* A stand-in for libdsopp_pba_cuda.so WITHOUT any compute (test infrastructure): it records what the Python binding
 * (dsopp_b200/capi.py) hands over and fills the output buffers with recognisable patterns, so that the binding's pointer
 * plumbing -- which array goes to which parameter, row pointers of 2-D outputs, NULL for skipped entries -- is checked on
 * a machine without a GPU (tests/test_binding_plumbing.py, run in a subprocess so the real library is never mixed up
 * with this one).  Entry points that are not written out here are generated as `return 0` stubs by the test. */
#include <stdint.h>
#include <string.h>

#define MAXF 16
static int n_frames = 0;
static int n_lm[MAXF];
static double sum_image[MAXF], sum_pose[MAXF], sum_uv[MAXF], sum_idepth[MAXF], sum_patch[MAXF], sum_flags[MAXF];
static double sum_status[MAXF][MAXF];
static double state_eps[8 * MAXF], state_step[8 * MAXF];
static int width = 0, height = 0;

typedef struct { int32_t max_frames, max_points_per_frame, width, height, device, rank, world_size; } cfg_t;

int dpba_create(const cfg_t* cfg, void** out) {
  static int handle;
  width = cfg->width;
  height = cfg->height;
  n_frames = 0;
  memset(n_lm, 0, sizeof(n_lm));
  *out = &handle;
  return 0;
}
int dpba_destroy(void* h) { return 0; }
const char* dpba_last_error(void* h) { return "fake"; }
const char* dpba_version(void) { return "fake sm_100a"; }
void* dpba_stream(void* h) { return 0; }
int64_t dpba_launch_count(void) { return 0; }
int dpba_num_frames(void* h) { return n_frames; }

int dpba_push_frame(void* h, int32_t id, const float* image, const uint8_t* mask, const double* T, double exposure,
                    const double* ab, const double* intr, int32_t fixed) {
  const int s = n_frames++;
  double a = 0;
  for (long i = 0; i < (long)width * height * 3; ++i) a += image[i];
  sum_image[s] = a + (mask ? mask[0] : -1) + exposure + ab[0] + 2 * ab[1] + intr[0] + intr[3] + 100 * fixed + 1000 * id;
  sum_pose[s] = 0;
  for (int i = 0; i < 12; ++i) sum_pose[s] += (i + 1) * T[i];
  return s;
}
int dpba_remove_frame(void* h, int32_t slot) {
  --n_frames;
  return 0;
}
int dpba_set_landmarks(void* h, int32_t slot, int32_t n, const float* uv, const float* idepth, const float* patch,
                       const uint8_t* flags) {
  n_lm[slot] = n;
  sum_uv[slot] = sum_idepth[slot] = sum_patch[slot] = sum_flags[slot] = 0;
  for (int i = 0; i < 2 * n; ++i) sum_uv[slot] += uv[i];
  for (int i = 0; i < n; ++i) sum_idepth[slot] += idepth[i];
  for (int i = 0; i < 8 * n; ++i) sum_patch[slot] += patch[i];
  for (int i = 0; i < n; ++i) sum_flags[slot] += flags ? flags[i] : 0;
  return 0;
}
int dpba_num_landmarks(void* h, int32_t slot) { return n_lm[slot]; }
int dpba_set_frame_statuses(void* h, int32_t r, int32_t n, const uint8_t* const* rows) {
  for (int t = 0; t < n_frames; ++t) {
    sum_status[r][t] = -1;
    if (!rows[t]) continue;
    sum_status[r][t] = 0;
    for (int i = 0; i < n; ++i) sum_status[r][t] += rows[t][i];
  }
  return 0;
}
int dpba_get_frame_statuses(void* h, int32_t r, int32_t n, uint8_t* const* st, uint8_t* const* cd) {
  for (int t = 0; t < n_frames; ++t) {
    if ((st[t] == 0) != (t == r) && n) return -1;  /* exactly row r is NULL */
    if (!st[t]) continue;
    for (int i = 0; i < n; ++i) {
      st[t][i] = (uint8_t)(10 * t + i % 7);
      cd[t][i] = (uint8_t)(100 + t);
    }
  }
  return 0;
}
int dpba_get_landmarks(void* h, int32_t slot, int32_t n, float* idepth, float* idepth_step, float* inv_hdd, float* b_d,
                       uint8_t* flags, uint32_t* n_inl, float* rel_baseline) {
  for (int i = 0; i < n; ++i) {
    idepth[i] = 1.f + i;
    idepth_step[i] = 2.f + i;
    inv_hdd[i] = 3.f + i;
    b_d[i] = 4.f + i;
    flags[i] = (uint8_t)(i % 5);
    n_inl[i] = 6u + (unsigned)i;
    rel_baseline[i] = 7.f + i;
  }
  return 0;
}
int dpba_set_state(void* h, const double* eps, const double* step) {
  if (eps) memcpy(state_eps, eps, sizeof(double) * 8 * n_frames);
  if (step) memcpy(state_step, step, sizeof(double) * 8 * n_frames);
  return 0;
}
int dpba_get_state(void* h, double* eps, double* step) {
  memcpy(eps, state_eps, sizeof(double) * 8 * n_frames);
  memcpy(step, state_step, sizeof(double) * 8 * n_frames);
  return 0;
}

/* what the test reads back */
double fake_sum(int what, int a, int b) {
  switch (what) {
    case 0: return sum_image[a];
    case 1: return sum_pose[a];
    case 2: return sum_uv[a];
    case 3: return sum_idepth[a];
    case 4: return sum_patch[a];
    case 5: return sum_flags[a];
    case 6: return sum_status[a][b];
  }
  return 0;
}
