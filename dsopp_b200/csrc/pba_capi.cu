// C ABI of the B200 photometric bundle-adjustment path (include/dsopp_cuda_pba.h).
// Host side of the handle: owns device memory, the stream and the frame state; no compute happens on the CPU
// here except O(N) frame-state bookkeeping and the nth_element of updatePointStatuses.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>  // types only: the library is resolved at run time (see nccl_api below)
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/dsopp_cuda_pba.h"
#include "pba_internal.h"

namespace {

// NCCL is bound lazily with dlopen so that (a) single-GPU users never load it and (b) under torchrun the
// process-wide libnccl.so.2 that torch already loaded (2.28.x here) is the one we call, instead of pulling a
// second, older copy in through DT_NEEDED.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
  api.AllReduce = (decltype(api.AllReduce))dlsym(lib, "ncclAllReduce");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
  return api;
}

struct FrameHost {
  int id = 0;
  int phys = -1;
  double T_lin[12];
  double exposure = 1;
  double ab0[2] = {0, 0};
  double intr[4];
  int fixed = 0;
  int to_marg = 0;
  int is_marg = 0;
  int n_lm = 0;
  int mask_all = 1;
  double eps[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double step[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

}  // namespace

struct dpba_handle {
  dpba_config cfg;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;            // side branch of the LM launch sequence (fork/join inside the graph)
  cudaStream_t stream3 = nullptr;            // second side branch (block assembly beside the Schur reduction)
  bool final_sweep = false;                  // option "final_sweep": run the (redundant) closing residual sweep of a speculative solve
  int dbg_freeze = -1;                       // option "debug_freeze_stamps": the diagnostic stamps stop after the energy decision of this iteration
  bool three_branch = false;                 // option "three_branch": measured slower (94.7 vs 90.5 us per iteration)
  std::vector<cudaEvent_t> fork_ev;          // dependency-only events of the fork/join edges
  size_t fork_used = 0;
  std::string err = "";
  int n_frames = 0;
  FrameHost fr[PBA_MAXF];
  bool phys_used[PBA_MAXF] = {};
  float4* img[PBA_MAXF] = {};
  uint8_t* mask[PBA_MAXF] = {};
  // landmark arrays
  float4* lmk = nullptr;  // {u, v, idepth, idepth at the FEJ linearisation point}
  float *idepth_step = nullptr, *patch = nullptr;
  float* lm_slab = nullptr;  // owns idepth_step, inv_hdd, b_d, rel_baseline, n_inliers ([5][max_frames * max_pts])
  // Host mirror of everything updateFrame reads (photometric_bundle_adjustment.cpp:182-264), filled by ONE bulk readback
  // on the first dpba_get_* after the device state changed; later getters are plain memcpy's.  The reference always
  // follows solve() by updateFrame for every active frame (monocular_tracker.cpp:251-256).
  bool rb_valid = false;
  float* rb_slab = nullptr;     // pinned [5][max_frames * max_pts]
  float4* rb_lmk = nullptr;     // pinned
  uint8_t* rb_flags = nullptr;  // pinned
  uint8_t *rb_status = nullptr, *rb_cand = nullptr;  // pinned, first max_frames * 16 rows
  // pinned staging arena for the small host arrays (landmarks, statuses): the caller's buffer is copied here and
  // DMA'd asynchronously, so set_* calls return without a stream synchronisation and the caller may reuse its buffer
  char* arena_h = nullptr;
  size_t arena_cap = 0, arena_used = 0;
  uint8_t* flags = nullptr;
  float *inv_hdd = nullptr, *b_d = nullptr, *hpd = nullptr, *rel_baseline = nullptr;
  uint32_t* n_inliers = nullptr;
  uint8_t *status = nullptr, *cand = nullptr, *jac_valid = nullptr;
  float* energy = nullptr;
  PairConst* pairs = nullptr;
  PairAssemble* pasm = nullptr;
  FrameParams* fparams = nullptr;    // device
  FrameParams* fparams_h = nullptr;  // pinned
  double* red = nullptr;             // device reduction buffer (what the kernels accumulate into)
  double* red2 = nullptr;            // world_size > 1: out-of-place allreduce result of the exchanged block
  float* core_part = nullptr;        // first-stage partials (see ReduceBuf)
  float* fschur_part = nullptr;
  double* core = nullptr;
  double *schur_part = nullptr, *bs_part = nullptr, *e_part = nullptr, *n_part = nullptr;
  int sm_count = 148;
  double* red_h = nullptr;           // pinned mirror
  // device-resident LM
  LmCtl* ctl = nullptr;
  LmCtl* ctl_h = nullptr;            // pinned
  LmOptionsDev* lmopt = nullptr;
  LmOptionsDev* lmopt_h = nullptr;   // pinned
  int* fixed_dev = nullptr;
  int* fixed_h = nullptr;            // pinned
  bool use_graph = true;
  bool speculative = true;           // dpba_solve_lm: trial evaluation == next linearisation under force_accept
  // round 2 experiment, option "merged_tail": three launches per iteration (sweep with the accept / back-substitution fold,
  // k_reduce_system, k_lm_solve) instead of eight kernels on two graph branches.  Measured SLOWER on B200 (94 against 89 us
  // per iteration, profiles/r02_ab.md): the single-CTA k_lm_solve serialises what the two-branch graph overlaps, so the
  // eight-kernel sequence stays the default.
  bool merged_tail = false;
  bool speculative_multi = true;     // ... also with world_size > 1 (one allreduce per iteration); 2-GPU check: same state and energy as the two-sweep sequence
  // captured LM launch sequences, keyed by everything the captured kernels hold by value (window shape, physical slots,
  // options).  A sliding window cycles through a handful of (logical -> physical slot) maps: a small cache keeps one
  // graph per map instead of re-capturing on every keyframe.
  struct LmGraph {
    std::vector<long long> key;
    cudaGraphExec_t exec = nullptr;
    size_t events = 0;
    long long kernels = 0;
  };
  std::vector<LmGraph> lm_graphs;
  std::vector<long long> lm_graph_key;  // cleared by dpba_set_option: drops every cached graph at the next solve
  bool lm_graph_fresh = false;
  double* marg_dev = nullptr;        // [MAXD*MAXD + MAXD]
  double* marg_h = nullptr;          // pinned staging
  size_t red_n = 0;
  double* step_dev = nullptr;
  float* pair_dist = nullptr;
  float* pyr_planes = nullptr;   // lazily: intensity planes of the device-side pyramid build (2 * W * H floats)
  uint8_t* raw_u8 = nullptr;     // lazily: raw gray image + vignetting (2 * W * H bytes)
  float* lut_dev = nullptr;      // lazily: photometric calibration table (256 floats)
  float* stage = nullptr;  // image upload staging (device)
  // {I,dx,dy} uploads are pipelined: the DMA of keyframe k + 1 (copy stream, staging buffer k + 1 mod 2) runs while the
  // pack kernel of keyframe k reads its staging buffer on the main stream
  float* stage_alt = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t pack_stream = nullptr;   // the pack kernels run beside the main stream, which only joins before the first
                                        // kernel that reads an image (wait_images): landmark / status uploads and the
                                        // per-pair constants no longer queue behind the image DMA
  cudaEvent_t stage_ready[2] = {nullptr, nullptr}, stage_free[2] = {nullptr, nullptr};
  cudaEvent_t img_done = nullptr, main_mark = nullptr;
  bool img_pending = false;
  bool pack_after_main = false;  // the pack stream already waits for everything the main stream held at the last join
  int stage_idx = 0;
  float* stage_h = nullptr;  // pinned staging for images
  float *m_r = nullptr, *m_jref = nullptr, *m_jtgt = nullptr, *m_did = nullptr, *m_w = nullptr;
  bool linearized = false;
  int lin_frames = 0;
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  // peer-memory exchange (peer_exchange.cu): this rank's mailbox, the peers' mailboxes opened over CUDA IPC
  void* peer_box = nullptr;
  void* peer_open[pba::PEER_MAXW] = {};
  unsigned* peer_ctr = nullptr;  // device: {seq, done, error, pad} of channel A, {seq, done} of channel B, pad
  pba::PeerDev peer_b{};         // second, independent mailbox channel (the 8 scalars of a sharded LM iteration; option "split_exchange")
  bool split_exchange = true;    // sharded speculative LM: scalars and system travel in two concurrent exchanges
  int* peer_err_h = nullptr;     // mapped pinned: set by the kernel after a time-out
  pba::PeerDev peer{};
  bool peer_attached = false;
  bool peer_on = false;          // option "peer_exchange"
  bool peer_fused = false;       // option "peer_fused": inside dpba_solve_lm the exchange rides in the producers' epilogues (measured slower than the stand-alone kernel + split exchange: profiles/r02_ab.md)
                                 // and the consumers' prologues (no exchange kernel); 0 = the stand-alone mailbox kernel
  // device-side quantile of updatePointStatuses (energy_quantile.cu)
  bool device_quantile = true;   // option "device_quantile" (0: host nth_element over rows read back)
  pba::SelectState* sel_dev = nullptr;
  pba::SelectState* sel_h = nullptr;  // pinned
  float* dm_buf = nullptr;       // reference depth maps (depth_maps.cu), lazily: 4 * sum_l (W>>l)(H>>l) floats
  int dm_levels = 0;
  // per-kernel CUDA-event profiling (dpba_profile_*)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;           // pairs: [2i] start, [2i+1] stop
  std::vector<int> ev_kind;                   // kind of pair i
  size_t ev_used = 0;
  double prof_ms[DPBA_PROFILE_KINDS] = {};  // see dpba_profile_read for the kinds
  int prof_n[DPBA_PROFILE_KINDS] = {};
};

namespace {

constexpr size_t MAXD = PBA_MAXF * 8;
// exchange block [Hp | bp | Hs | bs | scal], packed for the CURRENT window (D = 8 n_frames): this is what crosses
// NVLink, 2 (D^2 + D) + 8 doubles = 66.6 KB at 8 keyframes
struct RedLayout {
  size_t hp, bp, hs, bs, scal, n;
};
inline RedLayout red_layout(int n_frames) {
  const size_t D = 8 * (size_t)n_frames;
  RedLayout L;
  L.hp = 0;
  L.bp = D * D;
  L.hs = L.bp + D;
  L.bs = L.hs + D * D;
  L.scal = L.bs + D;
  L.n = L.scal + 8;
  return L;
}
constexpr size_t N_RED = 2 * (MAXD * MAXD + MAXD) + 8;
constexpr size_t N_EXCHANGE = N_RED;
static_assert(N_EXCHANGE % 2 == 0, "the peer exchange moves double2");
constexpr size_t PEER_BOX_DATA = 2 * (size_t)pba::PEER_MAXW * N_EXCHANGE;  // doubles
constexpr size_t PEER_BOX_FLAGS = (size_t)pba::PEER_MAXW * pba::PEER_MAXC;  // words
constexpr size_t PEER_BOX_COUNTERS = 2 * 2 * (size_t)pba::PEER_MAXW;         // words: [parity][kind][source]
constexpr size_t PEER_BOX_A_BYTES = PEER_BOX_DATA * sizeof(double) + (PEER_BOX_FLAGS + PEER_BOX_COUNTERS) * sizeof(unsigned);
// channel B (the 8 scalar slots only): its own data [2 parities][PEER_MAXW sources][8] and flags, behind channel A's region
constexpr size_t PEER_B_SLOT = 8;
constexpr size_t PEER_BOX_B_DATA = 2 * (size_t)pba::PEER_MAXW * PEER_B_SLOT;  // doubles
constexpr size_t PEER_BOX_B_OFFSET = (PEER_BOX_A_BYTES + 255) / 256 * 256;
constexpr size_t PEER_BOX_BYTES = PEER_BOX_B_OFFSET + PEER_BOX_B_DATA * sizeof(double) + PEER_BOX_FLAGS * sizeof(unsigned);

// several ranks AND a way to sum over them (NCCL communicator or attached peer mailboxes)
inline bool multi_gpu(const dpba_handle* h) { return h->world > 1 && (h->comm || h->peer_on); }

int fail(dpba_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

// bump allocation out of the pinned arena; when it is full the stream is drained first (every DMA that reads the
// arena has then completed) and allocation restarts at the bottom
void* arena_alloc(dpba_handle* h, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes > h->arena_cap) return nullptr;
  if (h->arena_used + bytes > h->arena_cap) {
    cudaStreamSynchronize(h->stream);
    h->arena_used = 0;
  }
  void* p = h->arena_h + h->arena_used;
  h->arena_used += bytes;
  return p;
}

#define CK(expr)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (expr);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(h, DPBA_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));               \
  } while (0)

#define REQUIRE(cond, msg)                                  \
  do {                                                      \
    if (!(cond)) return fail(h, DPBA_E_INVALID, (msg));     \
  } while (0)

// RAII scope that brackets one kernel launch with events on the handle's stream when profiling is on
struct ProfScope {
  dpba_handle* h;
  cudaStream_t st;
  size_t idx = (size_t)-1;
  ProfScope(dpba_handle* h_, int kind, cudaStream_t st_ = nullptr) : h(h_), st(st_ ? st_ : h_->stream) {
    if (!h->profiling) return;
    if (h->ev_used * 2 + 2 > h->ev_pool.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      h->ev_pool.push_back(a);
      h->ev_pool.push_back(b);
      h->ev_kind.push_back(kind);
    }
    idx = h->ev_used++;
    h->ev_kind[idx] = kind;
    record(h->ev_pool[2 * idx]);
  }
  ~ProfScope() {
    if (idx != (size_t)-1) record(h->ev_pool[2 * idx + 1]);
  }
  // inside stream capture a timing event must become an event-record NODE (cudaEventRecordExternal); a plain
  // cudaEventRecord would only express a dependency and never be stamped when the graph runs
  void record(cudaEvent_t e) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(e, st, cudaEventRecordExternal);
    else cudaEventRecord(e, st);
  }
};

// edge `from` -> `to` between the handle's two streams (a graph dependency under capture, a real wait otherwise)
int stream_edge(dpba_handle* h, cudaStream_t from, cudaStream_t to) {
  if (h->fork_used == h->fork_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(h, DPBA_E_CUDA, "cudaEventCreate");
    h->fork_ev.push_back(e);
  }
  cudaEvent_t e = h->fork_ev[h->fork_used++];
  if (cudaEventRecord(e, from) != cudaSuccess || cudaStreamWaitEvent(to, e, 0) != cudaSuccess)
    return fail(h, DPBA_E_CUDA, "stream fork/join failed");
  return 0;
}

void profile_collect(dpba_handle* h) {
  if (!h->ev_used) return;
  cudaStreamSynchronize(h->stream);
  for (size_t i = 0; i < h->ev_used; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]) == cudaSuccess) {
      h->prof_ms[h->ev_kind[i]] += ms;
      h->prof_n[h->ev_kind[i]] += 1;
    }
  }
  h->ev_used = 0;
  cudaGetLastError();  // a failed elapsed-time query must not poison later cudaGetLastError() checks
}

// producers write `red` (and the first-stage partial buffers)
ReduceBuf redbuf(dpba_handle* h) {
  ReduceBuf rb;
  const RedLayout L = red_layout(h->n_frames);
  rb.Hp = h->red + L.hp;
  rb.bp = h->red + L.bp;
  rb.Hs = h->red + L.hs;
  rb.bs = h->red + L.bs;
  rb.scal = h->red + L.scal;
  rb.core_part = h->core_part;
  rb.fschur_part = h->fschur_part;
  rb.core = h->core;
  rb.schur_part = h->schur_part;
  rb.bs_part = h->bs_part;
  rb.e_part = h->e_part;
  rb.n_part = h->n_part;
  return rb;
}
// consumers (device solve, energy decision, readback) read the cross-rank sums: `red2` when world_size > 1
// (out-of-place allreduce, so a skipped linearisation re-reduces the same partials instead of summing sums)
ReduceBuf redbuf_out(dpba_handle* h) {
  ReduceBuf rb = redbuf(h);
  if (multi_gpu(h)) {
    const RedLayout L = red_layout(h->n_frames);
    rb.Hp = h->red2 + L.hp;
    rb.bp = h->red2 + L.bp;
    rb.Hs = h->red2 + L.hs;
    rb.bs = h->red2 + L.bs;
    rb.scal = h->red2 + L.scal;
  }
  return rb;
}

// the main stream joins the image uploads (copy + pack streams); called outside stream capture only
void wait_images(dpba_handle* h) {
  // after this join the main stream may be given kernels that read the image slots: the next push marks it again
  h->pack_after_main = false;
  if (!h->img_pending) return;
  cudaEventRecord(h->img_done, h->pack_stream);  // once for all the packs queued since the last join
  cudaStreamWaitEvent(h->stream, h->img_done, 0);
  h->img_pending = false;
}

WindowDev make_window(dpba_handle* h) {
  h->rb_valid = false;  // every kernel-launching path builds its WindowDev here: the device state is about to change
  wait_images(h);       // (dpba_solve_lm joins before it starts capturing; by then nothing is pending here)
  WindowDev w;
  memset(&w, 0, sizeof(w));
  w.n_frames = h->n_frames;
  w.W = h->cfg.width;
  w.H = h->cfg.height;
  w.max_pts = h->cfg.max_points_per_frame;
  w.hpd_stride = 8 * (h->linearized ? h->lin_frames : h->n_frames);
  for (int f = 0; f < h->n_frames; ++f) {
    const FrameHost& F = h->fr[f];
    w.n_lm[f] = F.n_lm;
    w.fixed[f] = F.fixed;
    w.frame_marg[f] = F.is_marg;
    w.phys[f] = F.phys;
    w.mask_all[f] = F.mask_all;
    w.img[f] = h->img[F.phys];
    w.mask[f] = h->mask[F.phys];
  }
  w.lmk = h->lmk;
  w.idepth_step = h->idepth_step;
  w.patch = h->patch;
  w.flags = h->flags;
  w.inv_hdd = h->inv_hdd;
  w.b_d = h->b_d;
  w.hpd = h->hpd;
  w.rel_baseline = h->rel_baseline;
  w.n_inliers = h->n_inliers;
  w.status = h->status;
  w.cand = h->cand;
  w.jac_valid = h->jac_valid;
  w.energy = h->energy;
  w.pairs = h->pairs;
  w.pairs_asm = h->pasm;
  w.m_r = h->m_r;
  w.m_jref = h->m_jref;
  w.m_jtgt = h->m_jtgt;
  w.m_did = h->m_did;
  w.m_w = h->m_w;
  return w;
}

void fill_frame_params(dpba_handle* h) {
  for (int f = 0; f < h->n_frames; ++f) {
    FrameParams& p = h->fparams_h[f];
    const FrameHost& F = h->fr[f];
    memcpy(p.T_lin, F.T_lin, sizeof(p.T_lin));
    memcpy(p.eps, F.eps, sizeof(p.eps));
    memcpy(p.step, F.step, sizeof(p.step));
    p.exposure = F.exposure;
    p.ab0[0] = F.ab0[0];
    p.ab0[1] = F.ab0[1];
    memcpy(p.intr, F.intr, sizeof(p.intr));
  }
}

// upload the frame state and recompute the per-pair constants on the device
int sync_pairs(dpba_handle* h) {
  fill_frame_params(h);
  CK(cudaMemcpyAsync(h->fparams, h->fparams_h, sizeof(FrameParams) * h->n_frames, cudaMemcpyHostToDevice, h->stream));
  {
    ProfScope ps(h, 6);
    pba::launch_pair_setup(h->fparams, h->n_frames, h->pairs, h->pasm, h->stream);
  }
  CK(cudaGetLastError());
  return 0;
}

int ensure_materialized(dpba_handle* h) {
  if (h->m_r) return 0;
  const size_t nres = (size_t)PBA_MAXF * PBA_MAXF * h->cfg.max_points_per_frame;
  // only max_frames^2 blocks are ever addressed, but the index uses the PBA_MAXF stride
  CK(cudaMalloc(&h->m_r, nres * 8 * sizeof(float)));
  CK(cudaMalloc(&h->m_did, nres * 8 * sizeof(float)));
  CK(cudaMalloc(&h->m_w, nres * sizeof(float)));
  CK(cudaMalloc(&h->m_jref, nres * 64 * sizeof(float)));
  CK(cudaMalloc(&h->m_jtgt, nres * 64 * sizeof(float)));
  CK(cudaMemsetAsync(h->m_r, 0, nres * 8 * sizeof(float), h->stream));
  CK(cudaMemsetAsync(h->m_did, 0, nres * 8 * sizeof(float), h->stream));
  CK(cudaMemsetAsync(h->m_jref, 0, nres * 64 * sizeof(float), h->stream));
  CK(cudaMemsetAsync(h->m_jtgt, 0, nres * 64 * sizeof(float), h->stream));
  // huber_weight = 1 in the ResidualPoint ctor (local_frame.hpp:218)
  std::vector<float> ones(nres, 1.f);
  CK(cudaMemcpyAsync(h->m_w, ones.data(), nres * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// sum the [core | Hs | bs | scal] block over ranks on the compute stream: our own one-shot kernel over NVLink peer
// memory when the mailboxes are attached and the option is on, else one NCCL allreduce (out of place: red -> red2)
enum { EX_SYSTEM = 0, EX_SCAL = 1, EX_ALL = 2 };
int exchange_raw(dpba_handle* h, size_t off, size_t n) {
  if (!multi_gpu(h)) return 0;
  if (h->peer_on) {
    pba::launch_peer_allreduce(h->peer, h->red, h->red2, off, n, h->stream);
    return 0;
  }
  NcclApi& nc = nccl_api();
  ncclResult_t r = nc.AllReduce(h->red + off, h->red2 + off, n, ncclDouble, ncclSum, h->comm, h->stream);
  if (r != ncclSuccess) return fail(h, DPBA_E_COMM, std::string("ncclAllReduce: ") + nc.GetErrorString(r));
  return 0;
}

// the 8 scalar slots over the second mailbox channel, on stream `s` (runs beside a system exchange on channel A)
int exchange_scal_b(dpba_handle* h, cudaStream_t s) {
  const RedLayout L = red_layout(h->n_frames);
  pba::launch_peer_allreduce(h->peer_b, h->red + L.scal, h->red2 + L.scal, 0, 8, s);
  return 0;
}
// the linear system [Hp | bp | Hs | bs] over channel A on stream `s`
int exchange_system_on(dpba_handle* h, cudaStream_t s) {
  const RedLayout L = red_layout(h->n_frames);
  pba::launch_peer_allreduce(h->peer, h->red, h->red2, L.hp, L.scal - L.hp, s);
  return 0;
}

// after a stream synchronisation: did a peer exchange kernel give up waiting?
int peer_check(dpba_handle* h) {
  if (h->peer_err_h && *h->peer_err_h)
    return fail(h, DPBA_E_COMM, "peer exchange timed out waiting for another rank (results are invalid)");
  return 0;
}

// one fused in-place-shaped allreduce of either the linear system [Hp | bp | Hs | bs] or the 8 scalars
int exchange(dpba_handle* h, int what) {
  const RedLayout L = red_layout(h->n_frames);
  if (what == EX_ALL) return exchange_raw(h, L.hp, L.n - L.hp);  // system and scalars in ONE allreduce
  return what == EX_SYSTEM ? exchange_raw(h, L.hp, L.scal - L.hp) : exchange_raw(h, L.scal, 8);
}

void host_se3_exp_translation(const double* T_lin, const double* eps, double* t_out) {
  // translation of T_lin * exp(eps[0:6])  (LocalFrame::tWorldAgent, local_frame.hpp:525-527)
  const double* v = eps;
  const double* w = eps + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double b, c;
  if (th < 1e-10) {
    b = 0.5;
    c = 1.0 / 6.0;
  } else {
    b = (1.0 - cos(th)) / th2;
    c = (th - sin(th)) / (th2 * th);
  }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += W[i * 3 + k] * W[k * 3 + j];
      W2[i * 3 + j] = s;
    }
  double tv[3];
  for (int i = 0; i < 3; ++i) {
    tv[i] = 0;
    for (int j = 0; j < 3; ++j) tv[i] += ((i == j ? 1.0 : 0.0) + b * W[i * 3 + j] + c * W2[i * 3 + j]) * v[j];
  }
  for (int i = 0; i < 3; ++i)
    t_out[i] = T_lin[i * 4 + 3] + T_lin[i * 4 + 0] * tv[0] + T_lin[i * 4 + 1] * tv[1] + T_lin[i * 4 + 2] * tv[2];
}

// one bulk device -> pinned host readback of the landmark arrays and residual statuses
int ensure_readback(dpba_handle* h) {
  if (h->rb_valid) return 0;
  wait_images(h);  // a synchronising call: borrowed page-locked images are released afterwards
  const size_t mp = h->cfg.max_points_per_frame, nlm = (size_t)h->cfg.max_frames * mp;
  const size_t nst = (size_t)h->cfg.max_frames * PBA_MAXF * mp;
  CK(cudaMemcpyAsync(h->rb_slab, h->lm_slab, 5 * nlm * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->rb_lmk, h->lmk, nlm * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->rb_flags, h->flags, nlm, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->rb_status, h->status, nst, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->rb_cand, h->cand, nst, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->arena_used = 0;  // the stream is idle
  h->rb_valid = true;
  return 0;
}

int push_frame_common(dpba_handle* h, int32_t frame_id, const float* image, int channels, const uint8_t* mask,
                      const double* T, double exposure, const double* ab0, const double* intr, int32_t fixed) {
  REQUIRE(h, "null handle");
  REQUIRE(image && T && ab0 && intr, "null argument");
  if (h->n_frames >= h->cfg.max_frames) return fail(h, DPBA_E_CAPACITY, "window is full");
  h->rb_valid = false;
  REQUIRE(!fixed || h->n_frames == 0, "only the first frame can be fixed (hessian_block_evaluation.hpp:143)");
  REQUIRE(exposure > 0, "exposure_time must be positive");
  int phys = -1;
  for (int p = 0; p < h->cfg.max_frames; ++p)
    if (!h->phys_used[p]) {
      phys = p;
      break;
    }
  if (phys < 0) return fail(h, DPBA_E_CAPACITY, "no free frame slot");
  const int W = h->cfg.width, H = h->cfg.height;
  const size_t npx = (size_t)W * H;
  // page-locked caller buffers are DMA'd directly; pageable ones go through the handle's pinned staging buffer
  cudaPointerAttributes attr;
  const bool pinned = channels < 0 || (cudaPointerGetAttributes(&attr, image) == cudaSuccess && attr.type == cudaMemoryTypeHost);
  cudaGetLastError();
  const float* src = image;
  const bool mask_all = !mask || memchr(mask, 0, npx) == nullptr;
  if (channels == 3 || channels == 1) {  // records or intensity plane from the host: pipelined over the copy stream
    if (!pinned) {
      CK(cudaStreamSynchronize(h->copy_stream));  // the previous pageable push may still be reading stage_h
      memcpy(h->stage_h, image, npx * channels * sizeof(float));
      src = h->stage_h;
    }
    const int sb = h->stage_idx;
    h->stage_idx ^= 1;
    float* dst = sb ? h->stage_alt : h->stage;
    CK(cudaStreamWaitEvent(h->copy_stream, h->stage_free[sb], 0));  // its last reader (a pack kernel) has finished
    CK(cudaMemcpyAsync(dst, src, npx * channels * sizeof(float), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaEventRecord(h->stage_ready[sb], h->copy_stream));
    // the pack kernel overwrites a physical image slot: everything the main stream has queued so far (a solve that still
    // reads the slot's previous occupant) comes first
    // (readers of image slots are only queued after wait_images(); between two joins the main stream receives nothing but
    // uploads, so one mark per run of pushes is enough)
    if (!h->pack_after_main) {
      CK(cudaEventRecord(h->main_mark, h->stream));
      CK(cudaStreamWaitEvent(h->pack_stream, h->main_mark, 0));
      h->pack_after_main = true;
    }
    CK(cudaStreamWaitEvent(h->pack_stream, h->stage_ready[sb], 0));
    if (channels == 3) pba::launch_pack_image(dst, h->img[phys], (int)npx, W, h->pack_stream);
    else pba::launch_pixelinfo(dst, h->img[phys], W, H, h->pack_stream);  // {I,dx,dy} from the intensity plane on the device
    CK(cudaEventRecord(h->stage_free[sb], h->pack_stream));
    h->img_pending = true;  // wait_images() records img_done behind the last pack
  } else {
    if (channels > 0) {
      if (!pinned) {
        memcpy(h->stage_h, image, npx * channels * sizeof(float));
        src = h->stage_h;
      }
      CK(cudaMemcpyAsync(h->stage, src, npx * channels * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    }
    pba::launch_pixelinfo(channels > 0 ? h->stage : image, h->img[phys], W, H, h->stream);  // -1: device plane
    h->pack_after_main = false;  // the main stream now holds a writer of an image slot
    if (channels > 0) CK(cudaEventRecord(h->stage_free[0], h->stream));  // h->stage is staging buffer 0 of the pipeline
  }
  CK(cudaGetLastError());
  // a mask without a zero is never looked at by the sweeps (WindowDev::mask_all): no need to send it over PCIe
  if (!mask_all) CK(cudaMemcpyAsync(h->mask[phys], mask, npx, cudaMemcpyHostToDevice, h->stream));
  else CK(cudaMemsetAsync(h->mask[phys], 255, npx, h->stream));
  // new residual vectors of this frame start as kOk until dpba_set_statuses says otherwise: the rows (phys -> p) are
  // contiguous, the rows (p -> phys) are one strided 2-D memset per array
  const size_t mp = h->cfg.max_points_per_frame;
  pba::launch_clear_frame_rows(h->status, h->cand, h->jac_valid, h->energy, phys, (int)mp, h->cfg.max_frames, h->stream);
  CK(cudaGetLastError());
  // pageable images went through stage_h, which the next push reuses; page-locked images are BORROWED until the next
  // synchronising call (solve / get_*), exactly as LocalFrame borrows its PixelMap pointers (local_frame.hpp:44,325)
  if (!pinned && channels != 3) CK(cudaStreamSynchronize(h->stream));
  FrameHost& F = h->fr[h->n_frames];
  F = FrameHost();
  F.id = frame_id;
  F.phys = phys;
  memcpy(F.T_lin, T, sizeof(F.T_lin));
  F.exposure = exposure;
  F.ab0[0] = ab0[0];
  F.ab0[1] = ab0[1];
  memcpy(F.intr, intr, sizeof(F.intr));
  F.fixed = fixed ? 1 : 0;
  F.mask_all = mask_all ? 1 : 0;
  h->phys_used[phys] = true;
  h->linearized = false;
  return h->n_frames++;
}

int upload_landmarks(dpba_handle* h, int slot, int first, int n, const float* uv, const float* idepth,
                     const float* patch, const uint8_t* flags) {
  const size_t base = (size_t)h->fr[slot].phys * h->cfg.max_points_per_frame + first;
  h->rb_valid = false;
  if (n == 0) return 0;
  float4* lk = (float4*)arena_alloc(h, sizeof(float4) * n);
  float* pt = (float*)arena_alloc(h, sizeof(float) * 8 * n);
  uint8_t* fl = flags ? (uint8_t*)arena_alloc(h, n) : nullptr;
  if (!lk || !pt || (flags && !fl)) return fail(h, DPBA_E_CAPACITY, "staging arena too small");
  for (int l = 0; l < n; ++l) lk[l] = make_float4(uv[2 * l], uv[2 * l + 1], idepth[l], idepth[l]);
  memcpy(pt, patch, sizeof(float) * 8 * n);
  CK(cudaMemcpyAsync(h->lmk + base, lk, sizeof(float4) * n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->patch + base * 8, pt, sizeof(float) * 8 * n, cudaMemcpyHostToDevice, h->stream));
  if (flags) {
    memcpy(fl, flags, n);
    CK(cudaMemcpyAsync(h->flags + base, fl, n, cudaMemcpyHostToDevice, h->stream));
  } else {
    CK(cudaMemsetAsync(h->flags + base, 0, n, h->stream));
  }
  const size_t nlm = (size_t)h->cfg.max_frames * h->cfg.max_points_per_frame;
  CK(cudaMemset2DAsync(h->lm_slab + base, nlm * sizeof(float), 0, sizeof(float) * n, 5, h->stream));
  return 0;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* dpba_version(void) { return "dsopp_b200 0.1 (sm_100a)"; }

const char* dpba_last_error(const dpba_handle* h) { return h ? h->err.c_str() : "null handle"; }

void* dpba_stream(dpba_handle* h) { return h ? (void*)h->stream : nullptr; }

int dpba_create(const dpba_config* cfg, dpba_handle** out) {
  if (!cfg || !out) return DPBA_E_INVALID;
  *out = nullptr;
  if (cfg->max_frames < 2 || cfg->max_frames > DPBA_MAX_FRAMES || cfg->max_points_per_frame < 1 || cfg->width < 16 ||
      cfg->height < 16)
    return DPBA_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
    // no CPU fallback: the product path needs the GPU
    fprintf(stderr, "dpba_create: CUDA device %d not available (%d devices)\n", cfg->device, ndev);
    return DPBA_E_CUDA;
  }
  dpba_handle* h = new dpba_handle();
  h->cfg = *cfg;
  h->rank = cfg->rank;
  h->world = cfg->world_size > 0 ? cfg->world_size : 1;
#define CKC(expr)                                                                    \
  do {                                                                               \
    cudaError_t e_ = (expr);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      fprintf(stderr, "dpba_create: %s: %s\n", #expr, cudaGetErrorString(e_));       \
      dpba_destroy(h);                                                               \
      return DPBA_E_CUDA;                                                            \
    }                                                                                \
  } while (0)
  CKC(cudaSetDevice(cfg->device));
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking));
  const size_t npx = (size_t)cfg->width * cfg->height;
  const size_t mp = cfg->max_points_per_frame;
  const size_t nlm = (size_t)cfg->max_frames * mp;
  const size_t nres = (size_t)PBA_MAXF * PBA_MAXF * mp;
  for (int p = 0; p < cfg->max_frames; ++p) {
    CKC(cudaMalloc(&h->img[p], npx * 2 * sizeof(float4)));  // {texel(x), texel(x+1)} records
    CKC(cudaMalloc(&h->mask[p], npx));
  }
  CKC(cudaMalloc(&h->lmk, nlm * sizeof(float4)));
  CKC(cudaMemset(h->lmk, 0, nlm * sizeof(float4)));
  // idepth_step | inv_hdd | b_d | rel_baseline | n_inliers carved out of ONE slab, so that a landmark upload clears
  // them with one 2-D memset and a readback fetches them with one 2-D copy
  CKC(cudaMalloc(&h->lm_slab, 5 * nlm * sizeof(float)));
  CKC(cudaMemset(h->lm_slab, 0, 5 * nlm * sizeof(float)));
  h->idepth_step = h->lm_slab;
  h->inv_hdd = h->lm_slab + nlm;
  h->b_d = h->lm_slab + 2 * nlm;
  h->rel_baseline = h->lm_slab + 3 * nlm;
  h->n_inliers = reinterpret_cast<uint32_t*>(h->lm_slab + 4 * nlm);
  CKC(cudaMalloc(&h->patch, nlm * 8 * sizeof(float)));
  CKC(cudaMalloc(&h->flags, nlm));
  CKC(cudaMalloc(&h->hpd, nlm * MAXD * sizeof(float)));
  CKC(cudaMalloc(&h->status, nres));
  CKC(cudaMalloc(&h->cand, nres));
  CKC(cudaMalloc(&h->jac_valid, nres));
  CKC(cudaMemset(h->jac_valid, 0, nres));  // ResidualPoint::reprojection_jacobians_valid = false (local_frame.hpp:191)
  CKC(cudaMalloc(&h->energy, nres * sizeof(float)));
  CKC(cudaMemset(h->status, 0, nres));
  CKC(cudaMemset(h->cand, 0, nres));
  CKC(cudaMemset(h->energy, 0, nres * sizeof(float)));
  CKC(cudaMemset(h->flags, 0, nlm));
  CKC(cudaMemset(h->hpd, 0, nlm * MAXD * sizeof(float)));
  CKC(cudaMalloc(&h->pairs, sizeof(PairConst) * PBA_MAXF * PBA_MAXF));
  CKC(cudaMalloc(&h->pasm, sizeof(PairAssemble) * PBA_MAXF * PBA_MAXF));
  CKC(cudaMalloc(&h->fparams, sizeof(FrameParams) * PBA_MAXF));
  CKC(cudaMallocHost(&h->fparams_h, sizeof(FrameParams) * PBA_MAXF));
  h->red_n = N_RED;
  CKC(cudaMalloc(&h->red, N_RED * sizeof(double)));
  CKC(cudaMallocHost(&h->red_h, N_RED * sizeof(double)));
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  if (h->sm_count <= 0) h->sm_count = 148;
  {
    const size_t maxd = (size_t)8 * cfg->max_frames;
    const size_t chunks = (mp + 7) / 8;  // k_linearize_fused never uses fewer than 8 landmarks per CTA
    CKC(cudaMalloc(&h->core_part, (size_t)cfg->max_frames * chunks * (cfg->max_frames - 1) * PBA_CORE * sizeof(float)));
    const size_t t4 = maxd / 4, nout = t4 * (t4 + 1) / 2 * 16 + maxd;
    CKC(cudaMalloc(&h->fschur_part, (size_t)cfg->max_frames * chunks * nout * sizeof(float)));
    CKC(cudaMalloc(&h->schur_part, (size_t)h->sm_count * maxd * maxd * sizeof(double)));
    CKC(cudaMalloc(&h->core, (size_t)PBA_MAXF * PBA_MAXF * PBA_CORE * sizeof(double)));
    CKC(cudaMalloc(&h->bs_part, (size_t)h->sm_count * maxd * sizeof(double)));
    // k_residual_sweep: one partial per (landmark chunk >= 8, host frame) CTA
    const size_t sweep_ctas = ((mp + 7) / 8) * (size_t)cfg->max_frames;
    CKC(cudaMalloc(&h->e_part, sweep_ctas * 2 * sizeof(double)));
    CKC(cudaMalloc(&h->n_part, ((mp + 31) / 32) * (size_t)cfg->max_frames * 2 * sizeof(double)));
  }
  CKC(cudaMalloc(&h->red2, N_EXCHANGE * sizeof(double)));
  CKC(cudaMemset(h->red2, 0, N_EXCHANGE * sizeof(double)));
  CKC(cudaMalloc(&h->ctl, sizeof(LmCtl)));
  CKC(cudaMallocHost(&h->ctl_h, sizeof(LmCtl)));
  CKC(cudaMalloc(&h->lmopt, sizeof(LmOptionsDev)));
  CKC(cudaMallocHost(&h->lmopt_h, sizeof(LmOptionsDev)));
  CKC(cudaMalloc(&h->fixed_dev, PBA_MAXF * sizeof(int)));
  CKC(cudaMallocHost(&h->fixed_h, PBA_MAXF * sizeof(int)));
  CKC(cudaMalloc(&h->marg_dev, (MAXD * MAXD + MAXD) * sizeof(double)));
  CKC(cudaMallocHost(&h->marg_h, (MAXD * MAXD + MAXD) * sizeof(double)));
  CKC(cudaMalloc(&h->step_dev, MAXD * sizeof(double)));
  CKC(cudaMalloc(&h->pair_dist, PBA_MAXF * PBA_MAXF * sizeof(float)));
  CKC(cudaMalloc(&h->stage, npx * 3 * sizeof(float)));
  CKC(cudaMalloc(&h->stage_alt, npx * 3 * sizeof(float)));
  CKC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&h->pack_stream, cudaStreamNonBlocking));
  CKC(cudaEventCreateWithFlags(&h->img_done, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&h->main_mark, cudaEventDisableTiming));
  for (int k = 0; k < 2; ++k) {
    CKC(cudaEventCreateWithFlags(&h->stage_ready[k], cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->stage_free[k], cudaEventDisableTiming));
  }
  CKC(cudaMallocHost(&h->stage_h, npx * 3 * sizeof(float)));
  // one window's worth of landmark records (52 B each, 256 B aligned per array) and status rows, twice
  h->arena_cap = 2 * ((size_t)cfg->max_frames * (mp * 64 + 1024) + (size_t)cfg->max_frames * cfg->max_frames * (mp + 256)) +
                 4 * npx + 4096;  // + raw 8-bit frames and vignetting of the device-side image preparation
  CKC(cudaMallocHost(&h->arena_h, h->arena_cap));
  CKC(cudaMallocHost(&h->rb_slab, 5 * nlm * sizeof(float)));
  CKC(cudaMallocHost(&h->rb_lmk, nlm * sizeof(float4)));
  CKC(cudaMallocHost(&h->rb_flags, nlm));
  CKC(cudaMallocHost(&h->rb_status, (size_t)cfg->max_frames * PBA_MAXF * mp));
  CKC(cudaMallocHost(&h->rb_cand, (size_t)cfg->max_frames * PBA_MAXF * mp));
#undef CKC
  *out = h;
  return DPBA_SUCCESS;
}

int dpba_destroy(dpba_handle* h) {
  if (!h) return DPBA_E_INVALID;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  // the LM graph holds captured NCCL kernels: it must go before the communicator it references
  for (auto& g : h->lm_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->lm_graphs.clear();
  if (h->comm) nccl_api().CommDestroy(h->comm);
  // peers' mailboxes are closed, ours is freed: the caller puts a barrier between the last solve and dpba_destroy
  for (void* p : h->peer_open)
    if (p) cudaIpcCloseMemHandle(p);
  cudaFree(h->peer_box);
  cudaFree(h->peer_ctr);
  cudaFree(h->sel_dev);
  cudaFree(h->dm_buf);
  if (h->sel_h) cudaFreeHost(h->sel_h);
  if (h->peer_err_h) cudaFreeHost(h->peer_err_h);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : h->fork_ev) cudaEventDestroy(e);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream3) cudaStreamDestroy(h->stream3);
  if (h->copy_stream) {
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamDestroy(h->copy_stream);
  }
  if (h->pack_stream) {
    cudaStreamSynchronize(h->pack_stream);
    cudaStreamDestroy(h->pack_stream);
  }
  if (h->img_done) cudaEventDestroy(h->img_done);
  if (h->main_mark) cudaEventDestroy(h->main_mark);
  for (int k = 0; k < 2; ++k) {
    if (h->stage_ready[k]) cudaEventDestroy(h->stage_ready[k]);
    if (h->stage_free[k]) cudaEventDestroy(h->stage_free[k]);
  }
  cudaFree(h->stage_alt);
  for (int p = 0; p < PBA_MAXF; ++p) {
    cudaFree(h->img[p]);
    cudaFree(h->mask[p]);
  }
  void* dev[] = {h->lmk,     h->jac_valid, h->lm_slab,                  h->patch,  h->flags,
                 h->hpd,     h->status, h->cand,    h->energy,
                 h->pairs,   h->pasm,   h->fparams,      h->red,        h->step_dev, h->pair_dist, h->stage,
                 h->m_r,     h->m_jref, h->m_jtgt,       h->m_did,      h->m_w,    h->red2,    h->ctl,
                 h->lmopt,   h->fixed_dev, h->marg_dev, h->core_part, h->fschur_part, h->core, h->schur_part, h->bs_part, h->e_part,
                 h->n_part, h->pyr_planes, h->raw_u8, h->lut_dev};
  for (void* p : dev) cudaFree(p);
  cudaFreeHost(h->fparams_h);
  cudaFreeHost(h->red_h);
  cudaFreeHost(h->stage_h);
  cudaFreeHost(h->arena_h);
  cudaFreeHost(h->rb_slab);
  cudaFreeHost(h->rb_lmk);
  cudaFreeHost(h->rb_flags);
  cudaFreeHost(h->rb_status);
  cudaFreeHost(h->rb_cand);
  cudaFreeHost(h->ctl_h);
  cudaFreeHost(h->lmopt_h);
  cudaFreeHost(h->marg_h);
  cudaFreeHost(h->fixed_h);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DPBA_SUCCESS;
}

int dpba_push_frame(dpba_handle* h, int32_t frame_id, const float* image_I_dx_dy, const uint8_t* mask,
                    const double T[12], double exposure, const double ab0[2], const double intr[4], int32_t fixed) {
  return push_frame_common(h, frame_id, image_I_dx_dy, 3, mask, T, exposure, ab0, intr, fixed);
}

int dpba_push_frame_intensity(dpba_handle* h, int32_t frame_id, const float* image_I, const uint8_t* mask,
                              const double T[12], double exposure, const double ab0[2], const double intr[4],
                              int32_t fixed) {
  return push_frame_common(h, frame_id, image_I, 1, mask, T, exposure, ab0, intr, fixed);
}

// device-side image preparation: uploads the raw gray frame (+ table, vignetting) and leaves the
// photometrically corrected level-0 intensity plane in h->pyr_planes
static int raw_to_plane(dpba_handle* h, const uint8_t* gray, const float* lut, const uint8_t* vignetting) {
  const size_t npx = (size_t)h->cfg.width * h->cfg.height;
  if (!h->pyr_planes) {
    CK(cudaMalloc(&h->pyr_planes, 2 * npx * sizeof(float)));
    CK(cudaMalloc(&h->raw_u8, 2 * npx));
    CK(cudaMalloc(&h->lut_dev, 256 * sizeof(float)));
  }
  uint8_t* g = (uint8_t*)arena_alloc(h, npx);
  if (!g) return fail(h, DPBA_E_CAPACITY, "staging arena too small");
  memcpy(g, gray, npx);
  CK(cudaMemcpyAsync(h->raw_u8, g, npx, cudaMemcpyHostToDevice, h->stream));
  if (lut) {
    float* l = (float*)arena_alloc(h, 256 * sizeof(float));
    if (!l) return fail(h, DPBA_E_CAPACITY, "staging arena too small");
    memcpy(l, lut, 256 * sizeof(float));
    CK(cudaMemcpyAsync(h->lut_dev, l, 256 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  float max_v = 0.f;
  if (vignetting) {
    uint8_t* v = (uint8_t*)arena_alloc(h, npx);
    if (!v) return fail(h, DPBA_E_CAPACITY, "staging arena too small");
    memcpy(v, vignetting, npx);
    uint8_t m = 0;  // cv::minMaxLoc(vignetting, nullptr, &max), photometrically_corrected_image.cpp:11-14
    for (size_t i = 0; i < npx; ++i) m = vignetting[i] > m ? vignetting[i] : m;
    max_v = (float)m;
    CK(cudaMemcpyAsync(h->raw_u8 + npx, v, npx, cudaMemcpyHostToDevice, h->stream));
  }
  pba::launch_photometric(h->raw_u8, lut ? h->lut_dev : nullptr, vignetting ? h->raw_u8 + npx : nullptr, max_v, h->pyr_planes,
                          (int)npx, h->stream);
  CK(cudaGetLastError());
  return 0;
}

int dpba_push_frame_raw(dpba_handle* h, int32_t frame_id, const uint8_t* gray, const float* photometric_calibration,
                        const uint8_t* vignetting, const uint8_t* mask, const double T[12], double exposure,
                        const double ab0[2], const double intr[4], int32_t fixed) {
  REQUIRE(h, "null handle");
  REQUIRE(gray, "null image");
  int rc = raw_to_plane(h, gray, photometric_calibration, vignetting);
  if (rc) return rc;
  // the plane is already on the device: push_frame_common takes it from there
  return push_frame_common(h, frame_id, h->pyr_planes, -1, mask, T, exposure, ab0, intr, fixed);
}

int dpba_build_pyramid(dpba_handle* h, const uint8_t* gray, const float* photometric_calibration,
                       const uint8_t* vignetting, int32_t levels, float* const* out_I_dx_dy) {
  REQUIRE(h, "null handle");
  REQUIRE(gray && out_I_dx_dy && levels >= 1, "bad argument");
  if (levels > 5) levels = 5;  // kMaxPyramidDepth, features/include/features/camera/pixel_data_frame.hpp:26
  int rc = raw_to_plane(h, gray, photometric_calibration, vignetting);
  if (rc) return rc;
  int W = h->cfg.width, H = h->cfg.height;
  const size_t npx = (size_t)W * H;
  float* cur = h->pyr_planes;
  float* nxt = h->pyr_planes + npx;
  for (int l = 0; l < levels; ++l) {
    if (out_I_dx_dy[l]) {
      pba::launch_pixelinfo3(cur, h->stage, W, H, h->stream);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(out_I_dx_dy[l], h->stage, (size_t)W * H * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));  // h->stage is reused by the next level
      CK(cudaEventRecord(h->stage_free[0], h->stream));
    }
    if (l + 1 < levels) {
      pba::launch_downscale(cur, nxt, W, H, h->stream);
      CK(cudaGetLastError());
      std::swap(cur, nxt);  // the coarser plane is at most a quarter of the finer: both halves of pyr_planes suffice
      W /= 2;
      H /= 2;
    }
  }
  CK(cudaStreamSynchronize(h->stream));
  return levels;
}

int dpba_refine_immature_landmarks(dpba_handle* h, int32_t ref_slot, int32_t n, const float* proj_xy,
                                   const float* idepth, const float* patch, int32_t minimum_inliers,
                                   double sigma_huber, float* idepth_out, uint8_t* activate,
                                   int32_t* number_of_valid_residuals) {
  REQUIRE(h, "null handle");
  REQUIRE(ref_slot >= 0 && ref_slot < h->n_frames && h->n_frames >= 2, "bad reference slot");
  REQUIRE(n >= 0 && (n == 0 || (proj_xy && idepth && patch)), "null argument");
  if (n == 0) return DPBA_SUCCESS;
  // the candidates travel through a scratch allocation sized for this call (activation happens once per keyframe)
  float* dev = nullptr;
  const size_t words = (size_t)n * (2 + 1 + 8 + 1 + 1) + (n + 3) / 4;
  CK(cudaMallocAsync((void**)&dev, words * sizeof(float), h->stream));
  float* d_proj = dev;
  float* d_id = d_proj + 2 * (size_t)n;
  float* d_patch = d_id + n;
  float* d_out = d_patch + 8 * (size_t)n;
  int* d_nv = reinterpret_cast<int*>(d_out + n);
  uint8_t* d_act = reinterpret_cast<uint8_t*>(d_nv + n);
  CK(cudaMemcpyAsync(d_proj, proj_xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_id, idepth, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_patch, patch, sizeof(float) * 8 * n, cudaMemcpyHostToDevice, h->stream));
  int rc = sync_pairs(h);  // t_t_r = T_w_t^-1 T_w_r and the brightness scale at the current frame state (:159-163)
  if (rc) return rc;
  WindowDev w = make_window(h);
  // std::min(minimum_inliers, active frames - 1), :331
  const int min_inl = std::min<int>(minimum_inliers, h->n_frames - 1);
  pba::launch_refine_immature(w, ref_slot, n, d_proj, d_id, d_patch, min_inl, (float)sigma_huber, d_out, d_act, d_nv, h->stream);
  CK(cudaGetLastError());
  if (idepth_out) CK(cudaMemcpyAsync(idepth_out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  if (activate) CK(cudaMemcpyAsync(activate, d_act, n, cudaMemcpyDeviceToHost, h->stream));
  if (number_of_valid_residuals)
    CK(cudaMemcpyAsync(number_of_valid_residuals, d_nv, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaFreeAsync(dev, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

int dpba_remove_frame(dpba_handle* h, int32_t slot) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames, "slot out of range");
  h->rb_valid = false;
  h->phys_used[h->fr[slot].phys] = false;
  for (int f = slot; f + 1 < h->n_frames; ++f) h->fr[f] = h->fr[f + 1];
  --h->n_frames;
  h->linearized = false;
  return DPBA_SUCCESS;
}

int dpba_num_frames(const dpba_handle* h) { return h ? h->n_frames : DPBA_E_INVALID; }

int dpba_synchronize(dpba_handle* h) {
  REQUIRE(h, "null handle");
  wait_images(h);
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

int dpba_set_frame_linearization(dpba_handle* h, int32_t slot, const double T[12], const double ab0[2]) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames && T && ab0, "bad argument");
  FrameHost& F = h->fr[slot];
  memcpy(F.T_lin, T, sizeof(F.T_lin));
  F.ab0[0] = ab0[0];
  F.ab0[1] = ab0[1];
  for (int k = 0; k < 8; ++k) F.eps[k] = 0;
  return DPBA_SUCCESS;
}

int dpba_set_frame_flags(dpba_handle* h, int32_t slot, int32_t fixed, int32_t to_marginalize) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames, "slot out of range");
  REQUIRE(!fixed || slot == 0, "only the first frame can be fixed");
  h->fr[slot].fixed = fixed ? 1 : 0;
  h->fr[slot].to_marg = to_marginalize ? 1 : 0;
  return DPBA_SUCCESS;
}

int dpba_set_frame_marginalized(dpba_handle* h, int32_t slot, int32_t is_marginalized) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames, "slot out of range");
  h->fr[slot].is_marg = is_marginalized ? 1 : 0;
  return DPBA_SUCCESS;
}

int dpba_set_landmarks(dpba_handle* h, int32_t slot, int32_t n, const float* uv, const float* idepth,
                       const float* patch, const uint8_t* flags) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames, "slot out of range");
  REQUIRE(n >= 0 && (n == 0 || (uv && idepth && patch)), "null landmark arrays");
  if (n > h->cfg.max_points_per_frame) return fail(h, DPBA_E_CAPACITY, "too many landmarks for this handle");
  int rc = upload_landmarks(h, slot, 0, n, uv, idepth, patch, flags);
  if (rc) return rc;
  h->fr[slot].n_lm = n;
  return DPBA_SUCCESS;
}

int dpba_append_landmarks(dpba_handle* h, int32_t slot, int32_t n, const float* uv, const float* idepth,
                          const float* patch, const uint8_t* flags) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames, "slot out of range");
  REQUIRE(n >= 0 && (n == 0 || (uv && idepth && patch)), "null landmark arrays");
  const int first = h->fr[slot].n_lm;
  if (first + n > h->cfg.max_points_per_frame) return fail(h, DPBA_E_CAPACITY, "too many landmarks for this handle");
  int rc = upload_landmarks(h, slot, first, n, uv, idepth, patch, flags);
  if (rc) return rc;
  h->fr[slot].n_lm = first + n;
  return DPBA_SUCCESS;
}

int dpba_set_landmark_flags(dpba_handle* h, int32_t slot, int32_t n, const uint8_t* flags) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames && flags, "bad argument");
  REQUIRE(n == h->fr[slot].n_lm, "flag count must equal the landmark count");
  h->rb_valid = false;
  const size_t base = (size_t)h->fr[slot].phys * h->cfg.max_points_per_frame;
  if (n) {
    uint8_t* st = (uint8_t*)arena_alloc(h, n);
    if (!st) return fail(h, DPBA_E_CAPACITY, "staging arena too small");
    memcpy(st, flags, n);
    CK(cudaMemcpyAsync(h->flags + base, st, n, cudaMemcpyHostToDevice, h->stream));
  }
  return DPBA_SUCCESS;
}

int dpba_num_landmarks(const dpba_handle* h, int32_t slot) {
  if (!h || slot < 0 || slot >= h->n_frames) return DPBA_E_INVALID;
  return h->fr[slot].n_lm;
}

int dpba_get_landmarks(dpba_handle* h, int32_t slot, int32_t n, float* idepth, float* idepth_step, float* inv_hdd,
                       float* b_d, uint8_t* flags, uint32_t* n_inl, float* rel_baseline) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames, "slot out of range");
  REQUIRE(n >= 0 && n <= h->fr[slot].n_lm, "n exceeds the landmark count");
  const size_t base = (size_t)h->fr[slot].phys * h->cfg.max_points_per_frame;
  if (n == 0) return DPBA_SUCCESS;
  int rc = ensure_readback(h);
  if (rc) return rc;
  const size_t nlm = (size_t)h->cfg.max_frames * h->cfg.max_points_per_frame;
  const size_t nb = sizeof(float) * n;
  if (idepth)
    for (int l = 0; l < n; ++l) idepth[l] = h->rb_lmk[base + l].z;
  if (idepth_step) memcpy(idepth_step, h->rb_slab + base, nb);
  if (inv_hdd) memcpy(inv_hdd, h->rb_slab + nlm + base, nb);
  if (b_d) memcpy(b_d, h->rb_slab + 2 * nlm + base, nb);
  if (rel_baseline) memcpy(rel_baseline, h->rb_slab + 3 * nlm + base, nb);
  if (n_inl) memcpy(n_inl, h->rb_slab + 4 * nlm + base, nb);
  if (flags) memcpy(flags, h->rb_flags + base, n);
  return DPBA_SUCCESS;
}

int dpba_get_pose_idepth_blocks(dpba_handle* h, int32_t slot, int32_t n, float* out) {
  REQUIRE(h, "null handle");
  REQUIRE(slot >= 0 && slot < h->n_frames && out, "bad argument");
  REQUIRE(n >= 0 && n <= h->fr[slot].n_lm, "n exceeds the landmark count");
  if (!h->linearized) return fail(h, DPBA_E_STATE, "linearize first");
  const size_t base = (size_t)h->fr[slot].phys * h->cfg.max_points_per_frame;
  const size_t stride = 8 * (size_t)h->lin_frames;
  if (n) CK(cudaMemcpyAsync(out, h->hpd + base * stride, sizeof(float) * stride * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

static int set_statuses_row(dpba_handle* h, int r, int t, int n, const uint8_t* st, int first = 0) {
  h->rb_valid = false;
  const size_t base = ((size_t)(h->fr[r].phys * PBA_MAXF + h->fr[t].phys)) * h->cfg.max_points_per_frame + (size_t)first;
  if (!n) return 0;
  uint8_t* stg = (uint8_t*)arena_alloc(h, n);
  if (!stg) return fail(h, DPBA_E_CAPACITY, "staging arena too small");
  memcpy(stg, st, n);
  CK(cudaMemcpyAsync(h->status + base, stg, n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->cand + base, stg, n, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int dpba_set_statuses(dpba_handle* h, int32_t r, int32_t t, int32_t n, const uint8_t* st) {
  REQUIRE(h, "null handle");
  REQUIRE(r >= 0 && r < h->n_frames && t >= 0 && t < h->n_frames && r != t && st, "bad pair");
  REQUIRE(n >= 0 && n <= h->cfg.max_points_per_frame, "n out of range");
  return set_statuses_row(h, r, t, n, st);
}

int dpba_append_statuses(dpba_handle* h, int32_t r, int32_t t, int32_t first, int32_t n, const uint8_t* st) {
  REQUIRE(h, "null handle");
  REQUIRE(r >= 0 && r < h->n_frames && t >= 0 && t < h->n_frames && r != t, "bad pair");
  REQUIRE(first >= 0 && n >= 0 && first + n <= h->cfg.max_points_per_frame, "range out of bounds");
  REQUIRE(n == 0 || st, "null statuses");
  return set_statuses_row(h, r, t, n, st, first);
}

int dpba_set_frame_statuses(dpba_handle* h, int32_t r, int32_t n, const uint8_t* const* per_target) {
  REQUIRE(h, "null handle");
  REQUIRE(r >= 0 && r < h->n_frames && per_target, "bad argument");
  REQUIRE(n >= 0 && n <= h->cfg.max_points_per_frame, "n out of range");
  h->rb_valid = false;
  if (n == 0) return DPBA_SUCCESS;
  // The residual vectors (r -> t) of one reference frame are rows [phys_t] of one contiguous [16][max_pts] block.
  // When every other frame is given, the rows are packed in the arena in device layout and sent with ONE copy per
  // array (the row of r itself is never read by any kernel); otherwise row by row.
  bool all = true;
  int lo = PBA_MAXF, hi = -1;
  for (int t = 0; t < h->n_frames; ++t) {
    if (t != r && !per_target[t]) all = false;
    lo = std::min(lo, h->fr[t].phys);
    hi = std::max(hi, h->fr[t].phys);
  }
  const size_t mp = h->cfg.max_points_per_frame;
  if (all && h->n_frames >= 2) {
    const size_t rows = (size_t)(hi - lo + 1);
    uint8_t* stg = (uint8_t*)arena_alloc(h, rows * mp);
    if (!stg) return fail(h, DPBA_E_CAPACITY, "staging arena too small");
    bool covered[PBA_MAXF] = {};
    memset(stg, 0, rows * mp);  // slots [n, max_pts) of every row and the unused r -> r row: kOk, never stale bytes
    for (int t = 0; t < h->n_frames; ++t) {
      covered[h->fr[t].phys - lo] = true;
      if (t != r) memcpy(stg + (size_t)(h->fr[t].phys - lo) * mp, per_target[t], n);
    }
    bool dense = true;  // a freed physical slot inside [lo, hi] may belong to nobody: leave its rows alone
    for (size_t k = 0; k < rows; ++k) dense = dense && covered[k];
    if (dense) {
      const size_t base = ((size_t)h->fr[r].phys * PBA_MAXF + lo) * mp;
      CK(cudaMemcpyAsync(h->status + base, stg, rows * mp, cudaMemcpyHostToDevice, h->stream));
      CK(cudaMemcpyAsync(h->cand + base, stg, rows * mp, cudaMemcpyHostToDevice, h->stream));
      return DPBA_SUCCESS;
    }
  }
  for (int t = 0; t < h->n_frames; ++t) {
    if (t == r || !per_target[t]) continue;
    const int rc = set_statuses_row(h, r, t, n, per_target[t]);
    if (rc) return rc;
  }
  return DPBA_SUCCESS;
}

int dpba_set_window_landmarks(dpba_handle* h, const int32_t* n, const float* const* uv, const float* const* idepth,
                              const float* const* patch, const uint8_t* const* flags, const uint8_t* const* statuses) {
  REQUIRE(h, "null handle");
  REQUIRE(n && uv && idepth && patch, "null argument");
  const int nf = h->n_frames;
  const size_t mp = h->cfg.max_points_per_frame;
  for (int f = 0; f < nf; ++f) {
    REQUIRE(n[f] >= 0 && (n[f] == 0 || (uv[f] && idepth[f] && patch[f])), "null landmark arrays");
    if ((size_t)n[f] > mp) return fail(h, DPBA_E_CAPACITY, "too many landmarks for this handle");
  }
  if (nf == 0) return DPBA_SUCCESS;
  // The frames' physical slots: when they are one dense run [lo, hi] every device array is written by ONE DMA out of
  // the pinned arena, packed there in device layout; otherwise (holes left by removed frames) frame by frame.
  int lo = PBA_MAXF, hi = -1;
  bool covered[PBA_MAXF] = {};
  for (int f = 0; f < nf; ++f) {
    lo = std::min(lo, h->fr[f].phys);
    hi = std::max(hi, h->fr[f].phys);
    covered[h->fr[f].phys] = true;
  }
  bool dense = true;
  for (int p = lo; p <= hi; ++p) dense = dense && covered[p];
  const size_t rows = (size_t)(hi - lo + 1), cells = rows * mp;
  const size_t need = cells * (sizeof(float4) + 8 * sizeof(float) + 1) + (statuses ? rows * cells : 0) + 4 * 256;
  if (!dense || need > h->arena_cap) {
    for (int f = 0; f < nf; ++f) {
      int rc = dpba_set_landmarks(h, f, n[f], uv[f], idepth[f], patch[f], flags ? flags[f] : nullptr);
      if (rc) return rc;
      if (statuses) {
        rc = dpba_set_frame_statuses(h, f, n[f], statuses + (size_t)f * nf);
        if (rc) return rc;
      }
    }
    return DPBA_SUCCESS;
  }
  h->rb_valid = false;
  if (h->arena_used + need > h->arena_cap) {  // the four blocks below must not wrap around each other
    CK(cudaStreamSynchronize(h->stream));
    h->arena_used = 0;
  }
  float4* lk = (float4*)arena_alloc(h, sizeof(float4) * cells);
  float* pt = (float*)arena_alloc(h, sizeof(float) * 8 * cells);
  uint8_t* fl = (uint8_t*)arena_alloc(h, cells);
  uint8_t* st = statuses ? (uint8_t*)arena_alloc(h, rows * cells) : nullptr;
  if (!lk || !pt || !fl || (statuses && !st)) return fail(h, DPBA_E_CAPACITY, "staging arena too small");
  if (st) memset(st, 0, rows * cells);  // slots [n, max_pts) and the unused r -> r rows: kOk, never stale bytes
  for (int f = 0; f < nf; ++f) {
    const size_t o = (size_t)(h->fr[f].phys - lo) * mp;
    const int m = n[f];
    const float *u = uv[f], *d = idepth[f];
    for (int l = 0; l < m; ++l) lk[o + l] = make_float4(u[2 * l], u[2 * l + 1], d[l], d[l]);
    memcpy(pt + 8 * o, patch[f], sizeof(float) * 8 * m);
    if (flags && flags[f]) memcpy(fl + o, flags[f], m);
    else memset(fl + o, 0, m);
    if ((size_t)m < mp) {  // the tail of the slot: defined bytes (no kernel reads past n_lm)
      memset((void*)(lk + o + m), 0, sizeof(float4) * (mp - m));
      memset(pt + 8 * (o + m), 0, sizeof(float) * 8 * (mp - m));
      memset(fl + o + m, 0, mp - m);
    }
    if (st)
      for (int t = 0; t < nf; ++t) {
        const uint8_t* row = t == f ? nullptr : statuses[(size_t)f * nf + t];
        if (row) memcpy(st + o * rows + (size_t)(h->fr[t].phys - lo) * mp, row, m);
      }
    h->fr[f].n_lm = m;
  }
  const size_t base = (size_t)lo * mp;
  const size_t nlm = (size_t)h->cfg.max_frames * mp;
  CK(cudaMemcpyAsync(h->lmk + base, lk, sizeof(float4) * cells, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->patch + base * 8, pt, sizeof(float) * 8 * cells, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->flags + base, fl, cells, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemset2DAsync(h->lm_slab + base, nlm * sizeof(float), 0, sizeof(float) * cells, 5, h->stream));
  if (st) {
    // residual vectors (r -> t): rows [phys_t] of the [PBA_MAXF][max_pts] block of phys_r -- one 2-D copy per array
    const size_t sbase = ((size_t)lo * PBA_MAXF + lo) * mp;
    CK(cudaMemcpy2DAsync(h->status + sbase, (size_t)PBA_MAXF * mp, st, cells, cells, rows, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpy2DAsync(h->cand + sbase, (size_t)PBA_MAXF * mp, st, cells, cells, rows, cudaMemcpyHostToDevice, h->stream));
  }
  return DPBA_SUCCESS;
}

int dpba_get_frame_statuses(dpba_handle* h, int32_t r, int32_t n, uint8_t* const* statuses, uint8_t* const* candidates) {
  REQUIRE(h, "null handle");
  REQUIRE(r >= 0 && r < h->n_frames, "bad argument");
  REQUIRE(n >= 0 && n <= h->cfg.max_points_per_frame, "n out of range");
  if (n == 0) return DPBA_SUCCESS;
  int rc = ensure_readback(h);
  if (rc) return rc;
  const size_t mp = h->cfg.max_points_per_frame;
  for (int t = 0; t < h->n_frames; ++t) {
    if (t == r) continue;
    const size_t base = ((size_t)(h->fr[r].phys * PBA_MAXF + h->fr[t].phys)) * mp;
    if (statuses && statuses[t]) memcpy(statuses[t], h->rb_status + base, n);
    if (candidates && candidates[t]) memcpy(candidates[t], h->rb_cand + base, n);
  }
  return DPBA_SUCCESS;
}

int dpba_get_statuses(dpba_handle* h, int32_t r, int32_t t, int32_t n, uint8_t* st, uint8_t* cand) {
  REQUIRE(h, "null handle");
  REQUIRE(r >= 0 && r < h->n_frames && t >= 0 && t < h->n_frames && r != t, "bad pair");
  REQUIRE(n >= 0 && n <= h->cfg.max_points_per_frame, "n out of range");
  const size_t base = ((size_t)(h->fr[r].phys * PBA_MAXF + h->fr[t].phys)) * h->cfg.max_points_per_frame;
  int rc = ensure_readback(h);
  if (rc) return rc;
  if (n && st) memcpy(st, h->rb_status + base, n);
  if (n && cand) memcpy(cand, h->rb_cand + base, n);
  return DPBA_SUCCESS;
}

int dpba_get_residual_scalars(dpba_handle* h, int32_t r, int32_t t, int32_t n, float* energy, uint8_t* valid) {
  REQUIRE(h, "null handle");
  REQUIRE(r >= 0 && r < h->n_frames && t >= 0 && t < h->n_frames && r != t, "bad pair");
  REQUIRE(n >= 0 && n <= h->cfg.max_points_per_frame, "n out of range");
  if (!n) return DPBA_SUCCESS;
  const size_t base = ((size_t)(h->fr[r].phys * PBA_MAXF + h->fr[t].phys)) * h->cfg.max_points_per_frame;
  if (energy) CK(cudaMemcpyAsync(energy, h->energy + base, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  if (valid) CK(cudaMemcpyAsync(valid, h->jac_valid + base, n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

int dpba_set_state(dpba_handle* h, const double* eps, const double* step) {
  REQUIRE(h, "null handle");
  for (int f = 0; f < h->n_frames; ++f)
    for (int k = 0; k < 8; ++k) {
      if (eps) h->fr[f].eps[k] = eps[8 * f + k];
      if (step) h->fr[f].step[k] = step[8 * f + k];
    }
  return DPBA_SUCCESS;
}

int dpba_get_state(dpba_handle* h, double* eps, double* step) {
  REQUIRE(h, "null handle");
  for (int f = 0; f < h->n_frames; ++f)
    for (int k = 0; k < 8; ++k) {
      if (eps) eps[8 * f + k] = h->fr[f].eps[k];
      if (step) step[8 * f + k] = h->fr[f].step[k];
    }
  return DPBA_SUCCESS;
}

int dpba_first_estimate(dpba_handle* h) {
  REQUIRE(h, "null handle");
  REQUIRE(h->n_frames >= 2, "need at least two frames");
  int rc = sync_pairs(h);  // the linearisation-point constants (M0, t0) of every pair
  if (rc) return rc;
  WindowDev w = make_window(h);
  pba::launch_first_estimate(w, h->stream);
  CK(cudaGetLastError());
  return DPBA_SUCCESS;
}

int dpba_evaluate(dpba_handle* h, double sigma, int32_t huber, int32_t fej, double* energy, int32_t* n_valid) {
  REQUIRE(h, "null handle");
  REQUIRE(h->n_frames >= 2, "need at least two frames");
  REQUIRE(huber || sigma == 0, "sigma_huber must be 0 without huber (evaluate_jacobians.hpp:25)");
  int rc = sync_pairs(h);
  if (rc) return rc;
  ReduceBuf rb = redbuf(h);
  CK(cudaMemsetAsync(rb.scal, 0, 8 * sizeof(double), h->stream));
  WindowDev w = make_window(h);
  int n_e;
  {
    ProfScope ps(h, 2);
    n_e = pba::launch_residual_sweep(w, (float)sigma, huber, fej, rb.e_part, h->stream);
  }
  pba::launch_reduce_scal(nullptr, 0, rb.e_part, n_e, nullptr, 0, rb.scal, h->stream);
  CK(cudaGetLastError());
  rc = exchange(h, EX_SCAL);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->red_h, redbuf_out(h).scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (peer_check(h)) return DPBA_E_COMM;
  if (energy) *energy = h->red_h[0];
  if (n_valid) *n_valid = (int32_t)llround(h->red_h[1]);
  return DPBA_SUCCESS;
}

int dpba_evaluate_jacobians(dpba_handle* h, double sigma, int32_t huber, int32_t fej) {
  REQUIRE(h, "null handle");
  REQUIRE(h->n_frames >= 2, "need at least two frames");
  REQUIRE(huber || sigma == 0, "sigma_huber must be 0 without huber");
  int rc = ensure_materialized(h);
  if (rc) return rc;
  rc = sync_pairs(h);
  if (rc) return rc;
  WindowDev w = make_window(h);
  {
    ProfScope ps(h, 3);
    pba::launch_materialise_sweep(w, (float)sigma, huber, fej, h->stream);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

int dpba_download_residual_block(dpba_handle* h, int32_t r, int32_t t, dpba_residual_view* v) {
  REQUIRE(h, "null handle");
  REQUIRE(v, "null view");
  REQUIRE(r >= 0 && r < h->n_frames && t >= 0 && t < h->n_frames && r != t, "bad pair");
  if (!h->m_r) return fail(h, DPBA_E_STATE, "dpba_evaluate_jacobians has not run");
  const int n = std::min<int>(v->n, h->fr[r].n_lm);
  const size_t base = ((size_t)(h->fr[r].phys * PBA_MAXF + h->fr[t].phys)) * h->cfg.max_points_per_frame;
  auto dl = [&](void* dst, const void* src, size_t bytes) {
    return dst && bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream) : cudaSuccess;
  };
  CK(dl(v->residuals, h->m_r + base * 8, sizeof(float) * 8 * n));
  CK(dl(v->d_reference_state_eps, h->m_jref + base * 64, sizeof(float) * 64 * n));
  CK(dl(v->d_target_state_eps, h->m_jtgt + base * 64, sizeof(float) * 64 * n));
  CK(dl(v->d_idepth, h->m_did + base * 8, sizeof(float) * 8 * n));
  CK(dl(v->huber_weight, h->m_w + base, sizeof(float) * n));
  CK(dl(v->energy, h->energy + base, sizeof(float) * n));
  CK(dl(v->connection_status, h->status + base, n));
  CK(dl(v->connection_status_candidate, h->cand + base, n));
  CK(cudaStreamSynchronize(h->stream));
  v->n = n;
  return DPBA_SUCCESS;
}

static int linearize_impl(dpba_handle* h, double sigma, int32_t huber, int32_t fej, int32_t for_marg, bool fused,
                          double* H_pose, double* b_pose, double* H_schur, double* b_schur) {
  REQUIRE(h, "null handle");
  REQUIRE(h->n_frames >= 2, "need at least two frames");
  REQUIRE(huber || sigma == 0, "sigma_huber must be 0 without huber");
  int rc;
  if (!fused && (rc = ensure_materialized(h))) return rc;
  if ((rc = sync_pairs(h))) return rc;
  const int N = h->n_frames, D = 8 * N;
  h->linearized = true;
  h->lin_frames = N;
  ReduceBuf rb = redbuf(h);
  CK(cudaMemsetAsync(h->red, 0, N_RED * sizeof(double), h->stream));
  WindowDev w = make_window(h);
  if (fused) {
    FusedShape shape;
    {
      ProfScope ps(h, 0);
      shape = pba::launch_linearize_fused(w, (float)sigma, huber, fej, for_marg, rb, h->stream);
    }
    CK(cudaGetLastError());
    {
      ProfScope ps(h, 4);
      pba::launch_assemble(w, fej, rb, shape, h->stream);
      pba::launch_finish_fused(w, rb, shape, h->stream);
    }
  } else {
    if (multi_gpu(h)) return fail(h, DPBA_E_STATE, "dpba_linearize_materialized is single-GPU only");
    {
      ProfScope ps(h, 3);
      pba::launch_materialise_sweep(w, (float)sigma, huber, fej, h->stream);
    }
    CK(cudaGetLastError());
    pba::launch_linearize_from_materialized(w, for_marg, rb, h->stream);
    CK(cudaGetLastError());
    const int nsb = pba::launch_schur(w, for_marg, rb, h->stream);
    CK(cudaGetLastError());
    pba::launch_finish_system(D, rb, nsb, h->stream);
  }
  CK(cudaGetLastError());
  if ((rc = exchange(h, EX_SYSTEM))) return rc;
  CK(cudaGetLastError());
  const RedLayout L = red_layout(N);
  // [Hp | bp | Hs | bs] is contiguous: one device-to-host copy
  CK(cudaMemcpyAsync(h->red_h, redbuf_out(h).Hp, (L.scal - L.hp) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (peer_check(h)) return DPBA_E_COMM;
  if (H_pose) memcpy(H_pose, h->red_h + L.hp, (size_t)D * D * sizeof(double));
  if (b_pose) memcpy(b_pose, h->red_h + L.bp, D * sizeof(double));
  if (H_schur) memcpy(H_schur, h->red_h + L.hs, (size_t)D * D * sizeof(double));
  if (b_schur) memcpy(b_schur, h->red_h + L.bs, D * sizeof(double));
  return DPBA_SUCCESS;
}

int dpba_linearize(dpba_handle* h, double sigma, int32_t huber, int32_t fej, int32_t for_marg, double* H_pose,
                   double* b_pose, double* H_schur, double* b_schur) {
  return linearize_impl(h, sigma, huber, fej, for_marg, true, H_pose, b_pose, H_schur, b_schur);
}

int dpba_linearize_materialized(dpba_handle* h, double sigma, int32_t huber, int32_t fej, int32_t for_marg,
                                double* H_pose, double* b_pose, double* H_schur, double* b_schur) {
  return linearize_impl(h, sigma, huber, fej, for_marg, false, H_pose, b_pose, H_schur, b_schur);
}

int dpba_back_substitute(dpba_handle* h, const double* step_pose, double lambda) {
  REQUIRE(h, "null handle");
  REQUIRE(step_pose, "null step");
  if (!h->linearized || h->lin_frames != h->n_frames) return fail(h, DPBA_E_STATE, "linearize first");
  const int D = 8 * h->n_frames;
  memcpy(h->red_h, step_pose, D * sizeof(double));  // pinned staging
  CK(cudaMemcpyAsync(h->step_dev, h->red_h, D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  WindowDev w = make_window(h);
  {
    ProfScope ps(h, 5);
    pba::launch_back_substitute(w, h->step_dev, lambda, h->stream);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

int dpba_accept(dpba_handle* h, double* state_sq, double* step_sq) {
  REQUIRE(h, "null handle");
  ReduceBuf rb = redbuf(h);
  CK(cudaMemsetAsync(rb.scal, 0, 8 * sizeof(double), h->stream));
  WindowDev w = make_window(h);
  pba::launch_accept(w, 1, rb.scal, h->stream);
  pba::launch_change_statuses(w, 1, h->stream);
  CK(cudaGetLastError());
  int rc = exchange(h, EX_SCAL);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->red_h, redbuf_out(h).scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (peer_check(h)) return DPBA_E_COMM;
  double st = h->red_h[2], sp = h->red_h[3];
  for (int f = 0; f < h->n_frames; ++f) {  // problem.hpp:369-376
    FrameHost& F = h->fr[f];
    for (int k = 0; k < 8; ++k) st += F.eps[k] * F.eps[k];
    st += F.ab0[0] * F.ab0[0] + F.ab0[1] * F.ab0[1];
    for (int k = 0; k < 8; ++k) {
      F.eps[k] += F.step[k];
      sp += F.step[k] * F.step[k];
      F.step[k] = 0;
    }
  }
  if (state_sq) *state_sq = st;
  if (step_sq) *step_sq = sp;
  return DPBA_SUCCESS;
}

int dpba_reject(dpba_handle* h) {
  REQUIRE(h, "null handle");
  WindowDev w = make_window(h);
  pba::launch_accept(w, 0, nullptr, h->stream);
  pba::launch_change_statuses(w, 0, h->stream);
  CK(cudaGetLastError());
  for (int f = 0; f < h->n_frames; ++f)
    for (int k = 0; k < 8; ++k) h->fr[f].step[k] = 0;
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

int dpba_change_residual_statuses(dpba_handle* h, int32_t accept) {
  REQUIRE(h, "null handle");
  WindowDev w = make_window(h);
  pba::launch_change_statuses(w, accept ? 1 : 0, h->stream);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

int dpba_landmarks_energy(dpba_handle* h, int32_t for_marg, double* energy, int32_t* n_valid) {
  REQUIRE(h, "null handle");
  ReduceBuf rb = redbuf(h);
  CK(cudaMemsetAsync(rb.scal, 0, 8 * sizeof(double), h->stream));
  WindowDev w = make_window(h);
  pba::launch_landmarks_energy(w, for_marg, rb.scal, h->stream);
  CK(cudaGetLastError());
  int rc = exchange(h, EX_SCAL);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->red_h, redbuf_out(h).scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (peer_check(h)) return DPBA_E_COMM;
  if (energy) *energy = h->red_h[0];
  if (n_valid) *n_valid = (int32_t)llround(h->red_h[1]);
  return DPBA_SUCCESS;
}

int dpba_update_point_statuses(dpba_handle* h, int32_t min_valid, double sigma, double* thr_out) {
  REQUIRE(h, "null handle");
  const int N = h->n_frames;
  const size_t mp = h->cfg.max_points_per_frame;
  // first half (photometric_bundle_adjustment.cpp:325-361): energies of kOk residuals of active landmarks
  std::vector<float> energies;
  std::vector<float> e(mp);
  std::vector<uint8_t> st(mp), fl(mp);
  const bool dev_q = h->device_quantile || h->world > 1;  // sharded landmarks: only the device select sees every rank
  if (h->world > 1 && !(h->comm && nccl_api().ok))
    return fail(h, DPBA_E_STATE, "update_point_statuses with world_size > 1 needs the NCCL communicator (dpba_comm_init)");
  for (int r = 0; r < N && !dev_q; ++r) {
    const int n = h->fr[r].n_lm;
    if (!n) continue;
    CK(cudaMemcpyAsync(fl.data(), h->flags + (size_t)h->fr[r].phys * mp, n, cudaMemcpyDeviceToHost, h->stream));
    for (int t = 0; t < N; ++t) {
      if (t == r || h->fr[t].is_marg) continue;
      const size_t base = ((size_t)(h->fr[r].phys * PBA_MAXF + h->fr[t].phys)) * mp;
      CK(cudaMemcpyAsync(e.data(), h->energy + base, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(st.data(), h->status + base, n, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      for (int l = 0; l < n; ++l)
        if (!(fl[l] & DPBA_LM_MARGINALIZED) && st[l] == DPBA_OK_STATUS) energies.push_back(e[l]);
    }
  }
  float thr = 0.f;
  if (dev_q) {
    // exact radix select on the device: no status / energy row crosses PCIe, one small readback
    if (!h->sel_dev) {
      CK(cudaMalloc(&h->sel_dev, sizeof(pba::SelectState)));
      CK(cudaMallocHost(&h->sel_h, sizeof(pba::SelectState)));
    }
    int nmax = 0;
    for (int r = 0; r < N; ++r) nmax = std::max(nmax, h->fr[r].n_lm);
    if (h->world > 1) {  // every rank must run the same number of passes: the largest shard decides
      // (shards differ by at most one landmark per frame; nmax only bounds the enumeration, so the local maximum + 1 is
      // a safe common value without another exchange)
      nmax += 1;
    }
    struct HistExchange {
      dpba_handle* h;
      static int run(void* p) {
        dpba_handle* hh = ((HistExchange*)p)->h;
        const ncclResult_t r = nccl_api().AllReduce(hh->sel_dev->hist, hh->sel_dev->hist, 256, ncclUint32, ncclSum, hh->comm,
                                                    hh->stream);
        if (r != ncclSuccess) {
          hh->err = std::string("ncclAllReduce(histogram): ") + nccl_api().GetErrorString(r);
          return 1;
        }
        return 0;
      }
    } hx{h};
    h->err.clear();
    pba::launch_energy_quantile(make_window(h), nmax, 0.75, h->sel_dev, h->stream, h->world > 1 ? &HistExchange::run : nullptr,
                                &hx);
    if (h->world > 1 && !h->err.empty()) return DPBA_E_COMM;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->sel_h, h->sel_dev, sizeof(pba::SelectState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->sel_h->count) thr = h->sel_h->value + (float)(sigma * sigma / 2);
  } else if (!energies.empty()) {
    const size_t k = (size_t)((double)energies.size() * 0.75);
    std::nth_element(energies.begin(), energies.begin() + (long)k, energies.end());
    thr = energies[k] + (float)(sigma * sigma / 2);
  }
  float dist[PBA_MAXF * PBA_MAXF] = {};
  double tw[PBA_MAXF][3];
  for (int f = 0; f < N; ++f) host_se3_exp_translation(h->fr[f].T_lin, h->fr[f].eps, tw[f]);
  for (int r = 0; r < N; ++r)
    for (int t = 0; t < N; ++t) {
      const double dx = tw[r][0] - tw[t][0], dy = tw[r][1] - tw[t][1], dz = tw[r][2] - tw[t][2];
      dist[r * PBA_MAXF + t] = (float)sqrt(dx * dx + dy * dy + dz * dz);
    }
  CK(cudaMemcpyAsync(h->pair_dist, dist, sizeof(dist), cudaMemcpyHostToDevice, h->stream));
  WindowDev w = make_window(h);
  pba::launch_apply_point_statuses(w, thr, min_valid, h->pair_dist, h->stream);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  if (thr_out) *thr_out = thr;
  return DPBA_SUCCESS;
}

// everything dpba_solve_lm puts on the stream (capturable into a CUDA graph: no host synchronisation inside)
static int lm_enqueue(dpba_handle* h, bool have_marg) {
  const int N = h->n_frames, D = 8 * N;
  const LmOptionsDev& od = *h->lmopt_h;
  cudaStream_t s = h->stream;
  CK(cudaMemcpyAsync(h->lmopt, h->lmopt_h, sizeof(LmOptionsDev), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->fixed_dev, h->fixed_h, PBA_MAXF * sizeof(int), cudaMemcpyHostToDevice, s));
  const double* Hm = nullptr;
  const double* bm = nullptr;
  if (have_marg) {
    CK(cudaMemcpyAsync(h->marg_dev, h->marg_h, ((size_t)D * D + D) * sizeof(double), cudaMemcpyHostToDevice, s));
    Hm = h->marg_dev;
    bm = h->marg_dev + (size_t)D * D;
  }
  // uploads the frame state; from here on the device copy is the master
  CK(cudaMemcpyAsync(h->fparams, h->fparams_h, sizeof(FrameParams) * N, cudaMemcpyHostToDevice, s));
  const WindowDev w = make_window(h);
  ReduceBuf rb = redbuf(h), ro = redbuf_out(h);
  const float sigma = (float)od.sigma;
  const int fej = od.fej;
  int rc;
  auto pairs = [&]() {
    ProfScope ps(h, 6);
    pba::launch_pair_setup(h->fparams, N, h->pairs, h->pasm, s);
  };

  cudaStream_t s2 = h->stream2;
  h->fork_used = 0;
  {
    ProfScope ps(h, 11);
    pba::launch_lm_init(h->ctl, h->lmopt, s);
  }
  pairs();
  // residual-only sweep + calculateEnergy tail.  Single GPU: k_lm_energy sums the per-CTA partials itself; with
  // several ranks the partials are summed first so that the 8 scalars can cross NVLink before the decision.
  const bool multi = multi_gpu(h);
  auto energy_eval = [&](int ctl_mode, int n_norm_parts, int kind) -> int {
    int n_e;
    {
      ProfScope ps(h, 2);
      n_e = pba::launch_residual_sweep(w, sigma, 1, fej, rb.e_part, s, h->ctl, ctl_mode);
    }
    if (multi) {
      pba::launch_reduce_scal(h->ctl, ctl_mode, rb.e_part, n_e, n_norm_parts ? rb.n_part : nullptr, n_norm_parts, rb.scal, s);
      const int rc2 = exchange(h, EX_SCAL);
      if (rc2) return rc2;
      if (kind != pba::LM_ENERGY_FINAL) {
        ProfScope ps(h, 11);
        pba::launch_lm_energy(h->ctl, h->lmopt, h->fparams, N, ro.scal, Hm, bm, kind, s);
      }
    } else if (kind != pba::LM_ENERGY_FINAL) {
      ProfScope ps(h, 11);
      pba::launch_lm_energy(h->ctl, h->lmopt, h->fparams, N, rb.scal, Hm, bm, kind, s, rb.e_part, n_e,
                            n_norm_parts ? rb.n_part : nullptr, n_norm_parts);
    }
    return 0;
  };
  const int m_max = [&]() {
    int m = 0;
    for (int f = 0; f < N; ++f) m = std::max(m, h->fr[f].n_lm);
    return m;
  }();
  const int n_norm_parts = ((m_max + 31) / 32) * N;  // CTAs of k_back_substitute
  // ---- speculative mode (single GPU, force_accept: the production options, fabric.cpp:99) -----------------------------
  // Under force_accept a rejected step ends the loop, so the linear system of the PREVIOUS state is never needed again
  // and the trial-state evaluation can be the next iteration's linearisation: the fused linearise carries the pair
  // energies in the spare slots of its core records, so ONE sweep per iteration yields both E(x + step) and the system
  // at x + step (the residual-only sweep of every iteration disappears).  The solve result (state, idepths, statuses,
  // energy) is what the reference computes; only the auxiliary per-landmark fields (hpd, b_d, inv_hdd) are one accepted
  // step FRESHER than the reference's at the end -- EigenPBA::solve recomputes them in its uncertainty pass anyway
  // (eigen_photometric_bundle_adjustment.cpp:91-98).  After a rejected step they are recomputed at the restored state.
  // ---- round 2: the same speculative sequence in THREE launches per iteration ------------------------------------------
  //   sweep_k        k_linearize_fused2 with the fold: closes trial k - 1 for its landmarks / residuals (accept or reject),
  //                  back-substitutes step k, evaluates + linearises the trial state of step k
  //   reduce_k       k_reduce_system: core reduction, H_pp block assembly, Schur partial reduction
  //   solve_k        k_lm_solve: energy decision of trial k (k = 0: initial energy), calculateStep k + 1, pair constants
  // instead of eight kernels on two graph branches; nothing but the three is on the critical path.
  if (h->speculative && od.force_accept && !multi && h->merged_tail && pba::fused_version() == 2) {
    const int n_pairs = N * (N - 1);
    FusedShape shape{0, 0};
    for (int k = 0; k <= od.max_it; ++k) {
      const bool more = k < od.max_it;
      {
        ProfScope ps(h, 0);
        shape = pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl, 1, k ? h->step_dev : nullptr);
      }
      {
        ProfScope ps(h, 8);
        pba::launch_reduce_system(w, fej, rb, shape, more ? 1 : 0, s, h->ctl);
      }
      {
        ProfScope ps(h, 7);
        pba::launch_lm_solve(h->ctl, h->lmopt, h->fparams, h->fixed_dev, N, rb.scal, ro, Hm, bm, h->step_dev,
                             k == 0 ? pba::LM_ENERGY_INITIAL : pba::LM_ENERGY_TRIAL, rb.core, n_pairs,
                             k ? rb.n_part : nullptr, k ? shape.chunks * N : 0, 1, more ? 1 : 0, h->pairs, h->pasm, s);
      }
    }
    {  // the last trial's acceptStep() / rejectStep() has no following sweep to ride on
      ProfScope ps(h, 11);
      pba::launch_accept(w, 0, nullptr, s, h->ctl, 1);
    }
    pairs();
    {
      ProfScope ps(h, 0);
      pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl, 3);  // only after a rejected step (see below)
    }
    // The closing calculateEnergy() (levenberg_marquardt_algorithm.hpp:119,126) is a no-op here: the last sweep that ran
    // -- the trial evaluation of the last accepted step, or the re-linearisation at the restored state after a rejected
    // one -- already left every residual's energy and candidate at the final state (same per-residual code, same state), and
    // the accept / reject kernel committed the statuses.  Option "final_sweep" brings the redundant sweep back.
    if (h->final_sweep && (rc = energy_eval(0, 0, pba::LM_ENERGY_FINAL))) return rc;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->ctl_h, h->ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->fparams_h, h->fparams, sizeof(FrameParams) * N, cudaMemcpyDeviceToHost, s));
    return 0;
  }
  // ---- round 2 experiment (option "three_branch", off): THREE graph branches, no core reduction on the critical path ----
  // Measured SLOWER than the two-branch sequence below (94.7 against 90.5 us per iteration, profiles/r02_ab.md): on this
  // problem size the small kernels of an iteration do not overlap the way the dependency graph suggests.
  //   main : sweep_k -> energy decision_k (pair energies straight from the sweep's per-chunk records) -> LM step_{k+1}
  //          -> back-substitution_{k+1} -> sweep_{k+1}
  //   s2   : Schur partial reduction_k ; landmark accept_k ; pair constants_{k+1}
  //   s3   : block assembly_k incl. the core reduction it needs (k_reduce_system without its Schur blocks)
  // Round 1 ran core reduce -> decision on the main branch and Schur reduce -> assembly one after the other on ONE side
  // branch (23 us, longer than the main branch's 18.6 us): the LM step started ~10 us later than it does now.
  if (h->speculative && od.force_accept && !multi && h->three_branch && pba::fused_version() == 2) {
    cudaStream_t s3 = h->stream3;
    FusedShape shape{0, 0};
    for (int k = 0; k <= od.max_it; ++k) {
      const bool more = k < od.max_it;
      {
        ProfScope ps(h, 0);
        shape = pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl, 1);
      }
      if (more) {
        if ((rc = stream_edge(h, s, s2))) return rc;
        if ((rc = stream_edge(h, s, s3))) return rc;
        {
          ProfScope ps(h, 10, s2);
          pba::launch_finish_fused(w, rb, shape, s2, h->ctl);
        }
        {
          ProfScope ps(h, 9, s3);
          pba::launch_reduce_system(w, fej, rb, shape, 1, s3, h->ctl, 0);
        }
      }
      {
        ProfScope ps(h, 11);
        pba::launch_lm_energy_from_records(w, h->ctl, h->lmopt, h->fparams, rb.scal, Hm, bm,
                                           k == 0 ? pba::LM_ENERGY_INITIAL : pba::LM_ENERGY_TRIAL, rb, shape,
                                           k ? rb.n_part : nullptr, k ? n_norm_parts : 0, s);
      }
      if (more) {
        if ((rc = stream_edge(h, s2, s))) return rc;  // Schur reduction_k and assembly_k before LM step_{k+1}
        if ((rc = stream_edge(h, s3, s))) return rc;
      }
      if (k > 0) {  // acceptStep() / rejectStep() of the landmarks incl. changeResidualStatuses
        if ((rc = stream_edge(h, s, s2))) return rc;
        ProfScope ps(h, 11, s2);
        pba::launch_accept(w, 0, nullptr, s2, h->ctl, 1);
      }
      if (!more) {
        if (k > 0 && (rc = stream_edge(h, s2, s))) return rc;
        break;
      }
      {
        ProfScope ps(h, 7);
        pba::launch_lm_step(h->ctl, h->lmopt, h->fparams, h->fixed_dev, N, ro, Hm, bm, h->step_dev, s);
      }
      if (k > 0 && (rc = stream_edge(h, s2, s))) return rc;  // accept_k before back_substitute_{k+1}
      if ((rc = stream_edge(h, s, s2))) return rc;
      {
        ProfScope ps(h, 6, s2);
        pba::launch_pair_setup(h->fparams, N, h->pairs, h->pasm, s2);
      }
      {
        ProfScope ps(h, 5);
        pba::launch_back_substitute(w, h->step_dev, 0.0, s, h->ctl, rb.n_part);
      }
      if ((rc = stream_edge(h, s2, s))) return rc;
    }
    pairs();
    {
      ProfScope ps(h, 0);
      pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl, 3);  // only after a rejected step
    }
    // The closing calculateEnergy() (levenberg_marquardt_algorithm.hpp:119,126) is a no-op here: the last sweep that ran
    // -- the trial evaluation of the last accepted step, or the re-linearisation at the restored state after a rejected
    // one -- already left every residual's energy and candidate at the final state (same per-residual code, same state), and
    // the accept / reject kernel committed the statuses.  Option "final_sweep" brings the redundant sweep back.
    if (h->final_sweep && (rc = energy_eval(0, 0, pba::LM_ENERGY_FINAL))) return rc;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->ctl_h, h->ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->fparams_h, h->fparams, sizeof(FrameParams) * N, cudaMemcpyDeviceToHost, s));
    return 0;
  }
  if (h->speculative && od.force_accept && !multi) {
    const int n_pairs = N * (N - 1);
    auto spec_linearize = [&](int ctl_mode) {
      FusedShape shape;
      {
        ProfScope ps(h, 0);
        shape = pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl, ctl_mode);
      }
      return shape;
    };
    // Linearisation k (k = 0: the initial state, k >= 1: the trial state of step k) is followed by
    //   main branch : core_reduce_k -> lm_energy_k (k = 0: result.energy; k >= 1: accept / reject of step k) -> lm_step_{k+1}
    //                 -> back_substitute_{k+1} -> fused_{k+1}
    //   side branch : finish_fused_k, assemble_k (needed by lm_step_{k+1}); accept_landmarks_k (needed by
    //                 back_substitute_{k+1}); pair_setup_{k+1} (needed by fused_{k+1})
    // so the critical path of an iteration is fused -> core_reduce -> energy -> step -> back-substitution.
    FusedShape shape;
    for (int k = 0; k <= od.max_it; ++k) {
      const bool more = k < od.max_it;
      shape = spec_linearize(1);
      if (more) {
        if ((rc = stream_edge(h, s, s2))) return rc;
        ProfScope ps(h, 10, s2);
        pba::launch_finish_fused(w, rb, shape, s2, h->ctl);
      }
      {
        ProfScope ps(h, 8);
        pba::launch_core_reduce(w, rb, shape, s, h->ctl);
      }
      if (more) {
        if ((rc = stream_edge(h, s, s2))) return rc;
        ProfScope ps(h, 9, s2);
        pba::launch_assemble_blocks(w, fej, rb, shape, s2, h->ctl);
      }
      {
        ProfScope ps(h, 11);
        pba::launch_lm_energy(h->ctl, h->lmopt, h->fparams, N, rb.scal, Hm, bm,
                              k == 0 ? pba::LM_ENERGY_INITIAL : pba::LM_ENERGY_TRIAL, s, rb.core, n_pairs,
                              k ? rb.n_part : nullptr, k ? n_norm_parts : 0, 1);
      }
      if (k == h->dbg_freeze) {
        pba::stamps_off_async(s);
        pba::peer_stamps_off_async(s);
      }
      if (more && (rc = stream_edge(h, s2, s))) return rc;  // finish_fused_k, assemble_k before lm_step_{k+1}
      if (k > 0) {  // acceptStep() / rejectStep() of the landmarks incl. changeResidualStatuses
        if ((rc = stream_edge(h, s, s2))) return rc;
        ProfScope ps(h, 11, s2);
        pba::launch_accept(w, 0, nullptr, s2, h->ctl, 1);
      }
      if (!more) {
        if (k > 0 && (rc = stream_edge(h, s2, s))) return rc;
        break;
      }
      {
        ProfScope ps(h, 7);
        pba::launch_lm_step(h->ctl, h->lmopt, h->fparams, h->fixed_dev, N, ro, Hm, bm, h->step_dev, s);
      }
      if (k > 0 && (rc = stream_edge(h, s2, s))) return rc;  // accept_k before back_substitute_{k+1}
      if ((rc = stream_edge(h, s, s2))) return rc;
      {
        ProfScope ps(h, 6, s2);
        pba::launch_pair_setup(h->fparams, N, h->pairs, h->pasm, s2);
      }
      {
        ProfScope ps(h, 5);
        pba::launch_back_substitute(w, h->step_dev, 0.0, s, h->ctl, rb.n_part);
      }
      if ((rc = stream_edge(h, s2, s))) return rc;
    }
    pairs();
    spec_linearize(3);  // only after a rejected step: landmark fields back to the (restored) final state
    // The closing calculateEnergy() (levenberg_marquardt_algorithm.hpp:119,126) is a no-op here: the last sweep that ran
    // -- the trial evaluation of the last accepted step, or the re-linearisation at the restored state after a rejected
    // one -- already left every residual's energy and candidate at the final state (same per-residual code, same state), and
    // the accept / reject kernel committed the statuses.  Option "final_sweep" brings the redundant sweep back.
    if (h->final_sweep && (rc = energy_eval(0, 0, pba::LM_ENERGY_FINAL))) return rc;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->ctl_h, h->ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->fparams_h, h->fparams, sizeof(FrameParams) * N, cudaMemcpyDeviceToHost, s));
    return 0;
  }
  // ---- speculative mode on several GPUs (option "speculative_multi_gpu") ----------------------------------------------
  // Same idea as above, with ONE NCCL allreduce per iteration: the pair energies and landmark norms of this rank are
  // reduced into the scalar slots, which sit right behind [H_pp | b_p | H_s | b_s] in the exchange buffer, so the system
  // of linearisation k and the energy of the state it was taken at cross NVLink together; every rank then takes the
  // same decision from the same sums.
  if (h->speculative && h->speculative_multi && od.force_accept && multi) {
    // fused exchange (peer mailboxes attached, option peer_exchange + peer_fused): no collective call and no exchange
    // kernel -- the producers push into every rank's mailbox, k_lm_energy / k_lm_step wait, sum in rank order and close
    const int fx = (h->peer_on && h->peer_attached && h->peer_fused) ? 1 : 0;
    const int sys_ctas = pba::system_producer_ctas(N);
    FusedShape shape;
    // Split exchange (mailbox kernel, option "split_exchange", default on): the energy decision needs the 8 scalars of every
    // rank, the LM step needs the summed system -- so they travel separately and concurrently.  Main branch: core reduce,
    // scalar reduce, scalar exchange (channel B), energy decision; side branch: Schur reduce, block assembly, system exchange
    // (channel A); the LM step joins both.  The decision no longer waits for the assembly of the system and its 66 KB round.
    const bool split = h->split_exchange && h->peer_on && h->peer_attached && !fx;
    for (int k = 0; split && k <= od.max_it; ++k) {
      const bool more = k < od.max_it;
      {
        ProfScope ps(h, 0);
        shape = pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl, 1);
      }
      if (more) {
        if ((rc = stream_edge(h, s, s2))) return rc;
        ProfScope ps(h, 10, s2);
        pba::launch_finish_fused(w, rb, shape, s2, h->ctl, 0);
      }
      {
        ProfScope ps(h, 8);
        pba::launch_core_reduce(w, rb, shape, s, h->ctl);
      }
      if (more) {
        if ((rc = stream_edge(h, s, s2))) return rc;  // the block assembly reads the reduced cores
        {
          ProfScope ps(h, 9, s2);
          pba::launch_assemble_blocks(w, fej, rb, shape, s2, h->ctl, 0);
        }
        exchange_system_on(h, s2);
      }
      pba::launch_reduce_scal(h->ctl, 1, rb.core, N * (N - 1), k ? rb.n_part : nullptr, k ? n_norm_parts : 0, rb.scal, s, N, 0);
      exchange_scal_b(h, s);
      {
        ProfScope ps(h, 11);
        pba::launch_lm_energy(h->ctl, h->lmopt, h->fparams, N, ro.scal, Hm, bm,
                              k == 0 ? pba::LM_ENERGY_INITIAL : pba::LM_ENERGY_TRIAL, s, nullptr, 0, nullptr, 0, 0, 0);
      }
      if (k == h->dbg_freeze) {
        pba::stamps_off_async(s);
        pba::peer_stamps_off_async(s);
      }
      if (more && (rc = stream_edge(h, s2, s))) return rc;  // the summed system before the LM step
      if (k > 0) {
        if ((rc = stream_edge(h, s, s2))) return rc;
        ProfScope ps(h, 11, s2);
        pba::launch_accept(w, 0, nullptr, s2, h->ctl, 1);
      }
      if (!more) {
        if (k > 0 && (rc = stream_edge(h, s2, s))) return rc;
        break;
      }
      {
        ProfScope ps(h, 7);
        pba::launch_lm_step(h->ctl, h->lmopt, h->fparams, h->fixed_dev, N, ro, Hm, bm, h->step_dev, s, 0);
      }
      if (k > 0 && (rc = stream_edge(h, s2, s))) return rc;
      if ((rc = stream_edge(h, s, s2))) return rc;
      {
        ProfScope ps(h, 6, s2);
        pba::launch_pair_setup(h->fparams, N, h->pairs, h->pasm, s2);
      }
      {
        ProfScope ps(h, 5);
        pba::launch_back_substitute(w, h->step_dev, 0.0, s, h->ctl, rb.n_part);
      }
      if ((rc = stream_edge(h, s2, s))) return rc;
    }
    for (int k = 0; !split && k <= od.max_it; ++k) {
      const bool more = k < od.max_it;
      {
        ProfScope ps(h, 0);
        shape = pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl, 1);
      }
      if (more) {
        if ((rc = stream_edge(h, s, s2))) return rc;
        ProfScope ps(h, 10, s2);
        pba::launch_finish_fused(w, rb, shape, s2, h->ctl, fx);
      }
      {
        ProfScope ps(h, 8);
        pba::launch_core_reduce(w, rb, shape, s, h->ctl);
      }
      pba::launch_reduce_scal(h->ctl, 1, rb.core, N * (N - 1), k ? rb.n_part : nullptr, k ? n_norm_parts : 0, rb.scal, s, N, fx);
      if (more) {
        {
          ProfScope ps(h, 9);
          pba::launch_assemble_blocks(w, fej, rb, shape, s, h->ctl, fx);
        }
        if ((rc = stream_edge(h, s2, s))) return rc;
      }
      if (!fx && (rc = exchange(h, more ? EX_ALL : EX_SCAL))) return rc;
      {
        ProfScope ps(h, 11);
        pba::launch_lm_energy(h->ctl, h->lmopt, h->fparams, N, ro.scal, Hm, bm,
                              k == 0 ? pba::LM_ENERGY_INITIAL : pba::LM_ENERGY_TRIAL, s, nullptr, 0, nullptr, 0, 0,
                              fx ? (more ? 1 : 2) : 0);
      }
      if (k == h->dbg_freeze) {
        pba::stamps_off_async(s);
        pba::peer_stamps_off_async(s);
      }
      if (k > 0) {
        if ((rc = stream_edge(h, s, s2))) return rc;
        ProfScope ps(h, 11, s2);
        pba::launch_accept(w, 0, nullptr, s2, h->ctl, 1);
      }
      if (!more) {
        if (k > 0 && (rc = stream_edge(h, s2, s))) return rc;
        break;
      }
      {
        ProfScope ps(h, 7);
        pba::launch_lm_step(h->ctl, h->lmopt, h->fparams, h->fixed_dev, N, ro, Hm, bm, h->step_dev, s, fx ? sys_ctas : 0);
      }
      if (k > 0 && (rc = stream_edge(h, s2, s))) return rc;
      if ((rc = stream_edge(h, s, s2))) return rc;
      {
        ProfScope ps(h, 6, s2);
        pba::launch_pair_setup(h->fparams, N, h->pairs, h->pasm, s2);
      }
      {
        ProfScope ps(h, 5);
        pba::launch_back_substitute(w, h->step_dev, 0.0, s, h->ctl, rb.n_part);
      }
      if ((rc = stream_edge(h, s2, s))) return rc;
    }
    pairs();
    {
      ProfScope ps(h, 0);
      pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl, 3);
    }
    // The closing calculateEnergy() (levenberg_marquardt_algorithm.hpp:119,126) is a no-op here: the last sweep that ran
    // -- the trial evaluation of the last accepted step, or the re-linearisation at the restored state after a rejected
    // one -- already left every residual's energy and candidate at the final state (same per-residual code, same state), and
    // the accept / reject kernel committed the statuses.  Option "final_sweep" brings the redundant sweep back.
    if (h->final_sweep && (rc = energy_eval(0, 0, pba::LM_ENERGY_FINAL))) return rc;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->ctl_h, h->ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->fparams_h, h->fparams, sizeof(FrameParams) * N, cudaMemcpyDeviceToHost, s));
    return 0;
  }
  // result.energy = problem.calculateEnergy()
  if ((rc = energy_eval(0, 0, pba::LM_ENERGY_INITIAL))) return rc;
  for (int it = 0; it < od.max_it; ++it) {
    // linearize().  The per-pair constants are already those of the current state: an accepted trial state IS the
    // new state (eps + step, bit for bit), and after a rejection the loop either ends (force_accept) or keeps the
    // previous linear system and skips the sweep.
    FusedShape shape;
    {
      ProfScope ps(h, 0);
      shape = pba::launch_linearize_fused(w, sigma, 1, fej, 0, rb, s, h->ctl);
    }
    // the two second-stage reductions are independent: H_pp / b_p (core reduce -> block assembly) on the main
    // branch, the Schur partial sum on the side branch
    if ((rc = stream_edge(h, s, s2))) return rc;
    {
      ProfScope ps(h, 10, s2);
      pba::launch_finish_fused(w, rb, shape, s2, h->ctl);
    }
    {
      ProfScope ps(h, 8);
      pba::launch_core_reduce(w, rb, shape, s, h->ctl);
    }
    {
      ProfScope ps(h, 9);
      pba::launch_assemble_blocks(w, fej, rb, shape, s, h->ctl);
    }
    if ((rc = stream_edge(h, s2, s))) return rc;
    if ((rc = exchange(h, EX_SYSTEM))) return rc;
    // calculateStep(lambda)
    {
      ProfScope ps(h, 7);
      pba::launch_lm_step(h->ctl, h->lmopt, h->fparams, h->fixed_dev, N, ro, Hm, bm, h->step_dev, s);
    }
    // back-substitution (landmarks) and the per-pair constants of the trial state (frames) are independent
    if ((rc = stream_edge(h, s, s2))) return rc;
    {
      ProfScope ps(h, 6, s2);
      pba::launch_pair_setup(h->fparams, N, h->pairs, h->pasm, s2);
    }
    {
      ProfScope ps(h, 5);
      pba::launch_back_substitute(w, h->step_dev, 0.0, s, h->ctl, rb.n_part);
    }
    if ((rc = stream_edge(h, s2, s))) return rc;
    // calculateEnergy() at state + step
    if ((rc = energy_eval(1, n_norm_parts, pba::LM_ENERGY_TRIAL))) return rc;
    // acceptStep() / rejectStep() incl. changeResidualStatuses
    {
      ProfScope ps(h, 11);
      pba::launch_accept(w, 0, nullptr, s, h->ctl, 1);
    }
  }
  // the trailing problem.calculateEnergy() of both exits (lm.hpp:119,126)
  pairs();
  if ((rc = energy_eval(0, 0, pba::LM_ENERGY_FINAL))) return rc;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->ctl_h, h->ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h->fparams_h, h->fparams, sizeof(FrameParams) * N, cudaMemcpyDeviceToHost, s));
  return 0;
}

int dpba_solve_lm(dpba_handle* h, const dpba_lm_options* o, const double* H_marg, const double* b_marg,
                  double energy_marg, dpba_lm_result* result) {
  REQUIRE(h, "null handle");
  REQUIRE(o, "null options");
  REQUIRE(h->n_frames >= 2, "need at least two frames");
  REQUIRE((H_marg == nullptr) == (b_marg == nullptr), "H_marg and b_marg must be given together");
  REQUIRE(o->max_num_iterations >= 0 && o->max_num_iterations <= 64, "max_num_iterations out of range");
  h->rb_valid = false;  // a graph replay does not pass through make_window()
  const int N = h->n_frames, D = 8 * N;
  LmOptionsDev& od = *h->lmopt_h;
  od.max_it = o->max_num_iterations;
  od.min_it = o->min_num_iterations;
  od.force_accept = o->force_accept;
  od.fej = o->first_estimate_jacobians;
  od.huber = 1;
  od.lambda0 = o->initial_levenberg_marquardt_regularizer;
  od.ftol = o->function_tolerance;
  od.ptol = o->parameter_tolerance;
  od.dec = o->levenberg_marquardt_regularizer_decrease_on_accept;
  od.inc = o->levenberg_marquardt_regularizer_increase_on_reject;
  od.sigma = o->sigma_huber_loss;
  od.ab_reg[0] = o->affine_brightness_regularizer[0];
  od.ab_reg[1] = o->affine_brightness_regularizer[1];
  od.fixed_reg = o->fixed_state_regularizer;
  od.energy_marg = energy_marg;
  for (int f = 0; f < PBA_MAXF; ++f) h->fixed_h[f] = f < N ? h->fr[f].fixed : 0;
  if (H_marg) {
    memcpy(h->marg_h, H_marg, (size_t)D * D * sizeof(double));
    memcpy(h->marg_h + (size_t)D * D, b_marg, D * sizeof(double));
  }
  fill_frame_params(h);
  h->linearized = true;
  h->lin_frames = N;
  wait_images(h);  // before any capture starts: an event recorded outside a capture cannot be waited on inside it

  // The launch sequence only depends on the window shape and a few options: capture it once into a CUDA graph and
  // replay it (one cudaGraphLaunch instead of ~110 launches per solve); all inputs travel through pinned buffers.
  std::vector<long long> key = {N, od.max_it, od.fej, H_marg != nullptr, h->world, (long long)h->profiling,
                                (long long)(h->speculative && od.force_accept), (long long)h->speculative_multi,
                                (long long)llround(od.sigma * 1e6), (long long)h->peer_on + 2 * (long long)h->peer_fused,
                                (long long)h->merged_tail + 2 * (long long)h->three_branch + 4 * (long long)h->final_sweep,
                                (long long)pba::fused_version()};
  for (int f = 0; f < N; ++f) {  // every per-frame field of WindowDev (the captured kernels hold it BY VALUE)
    key.push_back(h->fr[f].n_lm);
    key.push_back(h->fr[f].phys);
    key.push_back(h->fr[f].mask_all | (h->fr[f].fixed << 1) | (h->fr[f].is_marg << 2));
  }
  int rc = 0;
  if (h->use_graph) {
    if (h->lm_graph_key.empty()) {  // an option changed (or first solve): every cached graph is stale
      for (auto& g : h->lm_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
      h->lm_graphs.clear();
    }
    dpba_handle::LmGraph* cur = nullptr;
    for (auto& g : h->lm_graphs)
      if (g.key == key) cur = &g;
    if (!cur) {
      if (h->lm_graphs.size() >= 16) {  // oldest out
        if (h->lm_graphs.front().exec) cudaGraphExecDestroy(h->lm_graphs.front().exec);
        h->lm_graphs.erase(h->lm_graphs.begin());
      }
      dpba_handle::LmGraph fresh;
      profile_collect(h);
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
      rc = lm_enqueue(h, H_marg != nullptr);
      cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      if (ce != cudaSuccess) return fail(h, DPBA_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
      {  // kernels per replay, for dpba_launch_count
        size_t nn = 0;
        fresh.kernels = 0;
        if (cudaGraphGetNodes(graph, nullptr, &nn) == cudaSuccess && nn) {
          std::vector<cudaGraphNode_t> nodes(nn);
          cudaGraphGetNodes(graph, nodes.data(), &nn);
          for (auto nd : nodes) {
            cudaGraphNodeType ty;
            if (cudaGraphNodeGetType(nd, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) ++fresh.kernels;
          }
        }
      }
      ce = cudaGraphInstantiate(&fresh.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) return fail(h, DPBA_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
      fresh.key = key;
      fresh.events = h->ev_used;
      h->lm_graph_key = key;
      h->ev_used = 0;
      h->lm_graph_fresh = true;
      h->lm_graphs.push_back(fresh);
      cur = &h->lm_graphs.back();
    }
    CK(cudaGraphLaunch(cur->exec, h->stream));
    if (!h->lm_graph_fresh) pba::add_launches(cur->kernels);  // the capture pass already counted once
    h->lm_graph_fresh = false;
    if (h->profiling) h->ev_used = cur->events;
  } else {
    if ((rc = lm_enqueue(h, H_marg != nullptr))) return rc;
  }
  CK(cudaStreamSynchronize(h->stream));
  if (peer_check(h)) return DPBA_E_COMM;
  if (h->profiling) profile_collect(h);
  for (int f = 0; f < N; ++f)
    for (int k = 0; k < 8; ++k) {
      h->fr[f].eps[k] = h->fparams_h[f].eps[k];
      h->fr[f].step[k] = 0;
    }
  if (result) {
    result->energy = h->ctl_h->energy;
    result->number_of_valid_residuals = h->ctl_h->n_valid;
    result->converged = h->ctl_h->converged;
    result->iterations = h->ctl_h->iterations_executed;
  }
  return DPBA_SUCCESS;
}

int dpba_create_reference_depth_maps(dpba_handle* h, int32_t n_levels, double idepth_variance, float* const* idepth_sum,
                                     float* const* weight) {
  REQUIRE(h, "null handle");
  REQUIRE(h->n_frames >= 1, "no frame");
  REQUIRE(n_levels >= 1 && n_levels <= 8, "1..8 pyramid levels");
  REQUIRE((h->cfg.width >> (n_levels - 1)) >= 1 && (h->cfg.height >> (n_levels - 1)) >= 1, "too many levels for this image");
  const int W = h->cfg.width, H = h->cfg.height;
  if (!h->dm_buf || h->dm_levels < n_levels) {
    if (h->dm_buf) {
      CK(cudaStreamSynchronize(h->stream));
      cudaFree(h->dm_buf);
      h->dm_buf = nullptr;
    }
    CK(cudaMalloc(&h->dm_buf, pba::dm_level_offset(W, H, n_levels) * sizeof(float)));
    h->dm_levels = n_levels;
  }
  int rc = sync_pairs(h);  // T_target_reference of every older keyframe at the accepted state (tWorldAgent, :29)
  if (rc) return rc;
  const bool was_valid = h->rb_valid;
  const WindowDev w = make_window(h);
  h->rb_valid = was_valid;  // read-only on the window: the host mirror stays valid
  pba::launch_reference_depth_maps(w, n_levels, (float)idepth_variance, h->dm_buf, h->stream);
  CK(cudaGetLastError());
  for (int l = 0; l < n_levels; ++l) {
    const size_t nl = (size_t)(W >> l) * (size_t)(H >> l);
    const float* base = h->dm_buf + pba::dm_level_offset(W, H, l);
    if (idepth_sum && idepth_sum[l])
      CK(cudaMemcpyAsync(idepth_sum[l], base + 2 * nl, nl * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (weight && weight[l])
      CK(cudaMemcpyAsync(weight[l], base + 3 * nl, nl * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return DPBA_SUCCESS;
}

int dpba_set_option(dpba_handle* h, const char* name, int64_t value) {
  REQUIRE(h && name, "null argument");
  if (!strcmp(name, "cuda_graph")) {
    h->use_graph = value != 0;
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "speculative_multi_gpu")) {
    h->speculative_multi = value != 0;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "device_quantile")) {  // updatePointStatuses: radix select on the device instead of host nth_element
    h->device_quantile = value != 0;
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "peer_exchange")) {  // our NVLink mailbox all-reduce instead of ncclAllReduce
    if (value && !h->peer_attached) return fail(h, DPBA_E_STATE, "peer_exchange needs dpba_peer_attach first");
    h->peer_on = value != 0;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "speculative_linearize")) {
    h->speculative = value != 0;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "schur_tensor_cores")) {  // process-wide
    pba::set_schur_mma(value != 0);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "peer_fence_all")) {  // A/B: 1 = system-scope fence in every pushing thread of the mailbox exchange (round-1 form)
    pba::set_peer_fence_all((int)value);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "split_exchange")) {  // sharded LM with the mailbox kernel: scalars and system in two concurrent exchanges
    h->split_exchange = value != 0;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "fused2_min_blocks")) {  // process-wide A/B switch: 2 (default) or 3 resident CTAs per SM for the fused sweep
    pba::set_fused2_min_blocks((int)value);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "fused_reduce_once")) {  // process-wide A/B switch: the sweep's per-group warp reductions folded into one per CTA
    pba::set_fused_reduce_once((int)value);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "fused_lpb_max")) {  // process-wide tuning switch: cap of the fused sweep's landmarks per CTA (32..256)
    pba::set_fused_lpb_max((int)value);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "fused_epilogue")) {  // process-wide A/B switch: 1 = second-generation epilogue of the fused sweep
    pba::set_fused_epilogue((int)value);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "pixelinfo_tma")) {  // process-wide A/B switch: {I,dx,dy} packing with the tile staged by TMA
    pba::set_pixelinfo_tma(value != 0);
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "fused_prefetch")) {  // process-wide A/B switch
    pba::set_fused_prefetch(value != 0);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "debug_freeze_stamps")) {  // diagnostics: see dpba_debug_kernel_times
    h->dbg_freeze = (int)value;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "final_sweep")) {
    h->final_sweep = value != 0;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "three_branch")) {
    h->three_branch = value != 0;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "peer_fused")) {
    h->peer_fused = value != 0;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "pdl")) {  // process-wide: programmatic dependent launch between the kernels of the device LM
    pba::set_pdl(value != 0);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "merged_tail")) {
    h->merged_tail = value != 0;
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "fused_version")) {  // process-wide: 2 = one thread per patch-residual (default), 1 = 8 lanes per patch
    pba::set_fused_version((int)value);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  if (!strcmp(name, "fused_min_blocks")) {  // process-wide: 4 (64 registers per thread) or 3 (96)
    pba::set_fused_min_blocks((int)value);
    h->lm_graph_key.clear();
    return DPBA_SUCCESS;
  }
  return fail(h, DPBA_E_INVALID, std::string("unknown option ") + name);
}

int64_t dpba_launch_count(void) { return (int64_t)pba::launch_count(); }

int dpba_debug_pixelinfo_ab(int32_t W, int32_t H, int32_t reps, double ms_per_launch[2], int64_t* mismatching_words) {
  if (W < 8 || H < 8 || reps < 1 || !ms_per_launch || !mismatching_words) return DPBA_E_INVALID;
  const size_t n = (size_t)W * H;
  std::vector<float> host(n);
  uint32_t st = 12345u;
  for (size_t i = 0; i < n; ++i) {
    st = st * 1664525u + 1013904223u;
    host[i] = (float)(st >> 8) * (255.f / 16777216.f);
  }
  float* I = nullptr;
  float4 *a = nullptr, *b = nullptr;
  cudaStream_t s;
  cudaEvent_t e0, e1;
  if (cudaMalloc(&I, n * 4) != cudaSuccess || cudaMalloc(&a, n * 32) != cudaSuccess || cudaMalloc(&b, n * 32) != cudaSuccess)
    return DPBA_E_CUDA;
  cudaStreamCreate(&s);
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaMemcpy(I, host.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemset(a, 0xff, n * 32);
  cudaMemset(b, 0x7f, n * 32);
  int rc = DPBA_SUCCESS;
  const bool tma_option = pba::get_pixelinfo_tma();  // the process-wide option is left as it was found
  for (int variant = 0; variant < 2 && rc == DPBA_SUCCESS; ++variant) {
    pba::set_pixelinfo_tma(false);
    for (int r = -3; r < reps; ++r) {  // three warm-up launches
      if (r == 0) cudaEventRecord(e0, s);
      if (variant == 0) pba::launch_pixelinfo(I, a, W, H, s);
      else if (!pba::launch_pixelinfo_tma(I, b, W, H, s)) rc = DPBA_E_STATE;
    }
    cudaEventRecord(e1, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = DPBA_E_CUDA;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms_per_launch[variant] = ms / reps;
  }
  if (rc == DPBA_SUCCESS) {
    std::vector<uint32_t> ha(n * 8), hb(n * 8);
    cudaMemcpy(ha.data(), a, n * 32, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb.data(), b, n * 32, cudaMemcpyDeviceToHost);
    int64_t bad = 0;
    for (size_t i = 0; i < n * 8; ++i) bad += ha[i] != hb[i];
    *mismatching_words = bad;
  }
  pba::set_pixelinfo_tma(tma_option);
  cudaFree(I);
  cudaFree(a);
  cudaFree(b);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(s);
  return rc;
}

int dpba_debug_kernel_times(int64_t out[32]) {
  cudaDeviceSynchronize();
  pba::debug_kernel_times(reinterpret_cast<long long*>(out));
  long long pk[4];
  pba::debug_peer_times(1, pk);  // slots 10, 11: the mailbox exchange kernel (entry / exit, wait for the peers begin / end)
  out[20] = pk[0], out[21] = pk[1], out[22] = pk[2], out[23] = pk[3];
  return DPBA_SUCCESS;
}

int dpba_debug_cta_times(int64_t* out, int32_t n) {
  cudaDeviceSynchronize();
  pba::debug_cta_times(reinterpret_cast<long long*>(out), n);
  return DPBA_SUCCESS;
}

int dpba_debug_stamps(int32_t enable, int64_t out[64]) {
  long long tmp[64];
  cudaDeviceSynchronize();
  pba::debug_peer_times(enable, nullptr);
  pba::debug_stamps(enable, tmp);
  if (out)
    for (int i = 0; i < 64; ++i) out[i] = tmp[i];
  return DPBA_SUCCESS;
}

int dpba_profile_enable(dpba_handle* h, int32_t on) {
  REQUIRE(h, "null handle");
  profile_collect(h);
  h->profiling = on != 0;
  for (int k = 0; k < DPBA_PROFILE_KINDS; ++k) {
    h->prof_ms[k] = 0;
    h->prof_n[k] = 0;
  }
  return DPBA_SUCCESS;
}

int dpba_profile_read(dpba_handle* h, double ms[DPBA_PROFILE_KINDS], int32_t launches[DPBA_PROFILE_KINDS]) {
  REQUIRE(h, "null handle");
  profile_collect(h);
  for (int k = 0; k < DPBA_PROFILE_KINDS; ++k) {
    if (ms) ms[k] = h->prof_ms[k];
    if (launches) launches[k] = h->prof_n[k];
  }
  return DPBA_SUCCESS;
}

int dpba_comm_unique_id(uint8_t id[128]) {
  if (!id) return DPBA_E_INVALID;
  ncclUniqueId uid;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (!nccl_api().ok || nccl_api().GetUniqueId(&uid) != ncclSuccess) return DPBA_E_COMM;
  memcpy(id, &uid, 128);
  return DPBA_SUCCESS;
}

int dpba_comm_init(dpba_handle* h, const uint8_t id[128], int32_t rank, int32_t world) {
  REQUIRE(h, "null handle");
  REQUIRE(id && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
  CK(cudaSetDevice(h->cfg.device));
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  NcclApi& nc = nccl_api();
  if (!nc.ok) return fail(h, DPBA_E_COMM, "libnccl.so.2 could not be loaded");
  ncclResult_t r = nc.CommInitRank(&h->comm, world, uid, rank);
  if (r != ncclSuccess) return fail(h, DPBA_E_COMM, std::string("ncclCommInitRank: ") + nc.GetErrorString(r));
  h->world = world;
  h->rank = rank;
  return DPBA_SUCCESS;
}

int dpba_peer_export(dpba_handle* h, uint8_t handle[64]) {
  REQUIRE(h && handle, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  CK(cudaSetDevice(h->cfg.device));
  if (!h->peer_box) {
    CK(cudaMalloc(&h->peer_box, PEER_BOX_BYTES));
    CK(cudaMalloc(&h->peer_ctr, 8 * sizeof(unsigned)));  // A: seq, done, error, pad; B: seq, done, pad, pad
    CK(cudaHostAlloc(&h->peer_err_h, sizeof(int), cudaHostAllocMapped));
    *h->peer_err_h = 0;
  }
  CK(cudaMemset(h->peer_box, 0, PEER_BOX_BYTES));
  CK(cudaMemset(h->peer_ctr, 0, 8 * sizeof(unsigned)));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t ipc;
  CK(cudaIpcGetMemHandle(&ipc, h->peer_box));
  memcpy(handle, &ipc, 64);
  return DPBA_SUCCESS;
}

int dpba_peer_barrier(dpba_handle* h) {
  REQUIRE(h, "null handle");
  REQUIRE(h->peer_attached, "dpba_peer_attach first");
  // a mailbox exchange of the 8 scalar slots into the scratch half of the reduction buffer: it completes on every rank
  // within a flag's flight time of the last rank's arrival, which is all a rendezvous needs
  const RedLayout L = red_layout(h->n_frames >= 2 ? h->n_frames : 2);
  pba::launch_peer_allreduce(h->peer, h->red, h->red2, L.scal, 8, h->stream);
  CK(cudaGetLastError());
  return DPBA_SUCCESS;
}

int dpba_peer_attach(dpba_handle* h, const uint8_t* handles, int32_t rank, int32_t world) {
  REQUIRE(h && handles, "null argument");
  REQUIRE(world >= 2 && world <= pba::PEER_MAXW && rank >= 0 && rank < world, "peer exchange: 2..8 ranks of one node");
  REQUIRE(h->peer_box, "dpba_peer_export first");
  REQUIRE(!h->peer_attached, "peers already attached");
  REQUIRE(!h->comm || (h->world == world && h->rank == rank), "rank / world differ from dpba_comm_init");
  CK(cudaSetDevice(h->cfg.device));
  pba::PeerDev pd{};
  for (int r = 0; r < world; ++r) {
    void* base = h->peer_box;
    if (r != rank) {
      cudaIpcMemHandle_t ipc;
      memcpy(&ipc, handles + (size_t)r * 64, 64);
      cudaError_t e = cudaIpcOpenMemHandle(&base, ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return fail(h, DPBA_E_COMM, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
      h->peer_open[r] = base;
    }
    pd.data[r] = static_cast<double*>(base);
    pd.flag[r] = reinterpret_cast<unsigned*>(static_cast<double*>(base) + PEER_BOX_DATA);
    pd.cnt[r] = pd.flag[r] + PEER_BOX_FLAGS;
  }
  pd.red_base = h->red;
  pd.seq = h->peer_ctr;
  pd.done = h->peer_ctr + 1;
  int* err_dev = nullptr;
  CK(cudaHostGetDevicePointer(&err_dev, h->peer_err_h, 0));
  pd.error = reinterpret_cast<int*>(h->peer_ctr + 2);
  pd.error_host = err_dev;
  pd.rank = rank;
  pd.world = world;
  pd.slot = N_EXCHANGE;
  h->peer = pd;
  {  // channel B: same ranks, same error word, its own data / flags / epoch counters
    pba::PeerDev pb = pd;
    for (int r = 0; r < world; ++r) {
      char* base = reinterpret_cast<char*>(pd.data[r]) + PEER_BOX_B_OFFSET;
      pb.data[r] = reinterpret_cast<double*>(base);
      pb.flag[r] = reinterpret_cast<unsigned*>(reinterpret_cast<double*>(base) + PEER_BOX_B_DATA);
      pb.cnt[r] = nullptr;
    }
    pb.seq = h->peer_ctr + 4;
    pb.done = h->peer_ctr + 5;
    pb.slot = PEER_B_SLOT;
    h->peer_b = pb;
  }
  pba::set_peer_context(pd);  // process-wide __device__ copy for the fused exchange (one sharded handle per process)
  CK(cudaGetLastError());
  h->world = world;
  h->rank = rank;
  h->peer_attached = true;
  return DPBA_SUCCESS;
}

}  // extern "C"
#pragma GCC visibility pop
