// extern "C" boundary of libapg_b200.so (declared in include/apg_b200.h).  Plain pointers and sizes only.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/apg_b200.h"
#include "kernels.h"
#include "pack_tables.h"
#include <mutex>

using namespace apg;

namespace {

// Per-device state, keyed by cudaGetDevice() under one mutex: the SM count (grid and workspace sizes follow it) and
// the cached buffer / stream of the host-buffer entry point.  A process may use several devices from several threads.
constexpr int APG_MAX_DEVICES = 64;
std::mutex g_dev_mutex;
int g_sm_counts[APG_MAX_DEVICES];
bool g_sm_known[APG_MAX_DEVICES];

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= APG_MAX_DEVICES) return -1;
  return dev;
}

int sm_count() {
  const int dev = current_device();
  if (dev < 0) return -1;
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  if (!g_sm_known[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    g_sm_counts[dev] = n;
    g_sm_known[dev] = true;
  }
  return g_sm_counts[dev];
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t up256(size_t x) { return (x + 255) & ~size_t(255); }

int state_dim(int system) { return system == SYS_CARTPOLE ? 4 : 12; }
int action_dim(int system) { return system == SYS_CARTPOLE ? 1 : 4; }

bool is_hutter(const apg_config* c) { return c->net == NET_HUTTER_CONV || c->net == NET_HUTTER_LIN; }

int check_config(const apg_config* c) {
  if (!c) return APG_ERR_BAD_CONFIG;
  if (c->n_drones <= 0 || c->horizon <= 0 || c->horizon > 64) return APG_ERR_BAD_CONFIG;
  if (c->system < 0 || c->system > 2 || c->mode < 0 || c->mode > 2) return APG_ERR_BAD_CONFIG;
  if (is_hutter(c) && c->mode == MODE_AUTOREGRESSIVE) {
    // Net(15, h, 9, 4) evaluated every step on features(state) and the next-h reference rows
    if (c->net != NET_HUTTER_CONV || c->system != SYS_QUAD) return APG_ERR_UNSUPPORTED;
    if (c->out_dim != 4 || c->state_feat != 15 || c->ref_dim != 9 || c->ref_len != c->horizon) return APG_ERR_BAD_CONFIG;
    if (c->horizon < 3 || c->horizon > 10) return APG_ERR_BAD_CONFIG;
    if (c->window != WINDOW_CUMULATIVE && c->window != WINDOW_RELATIVE) return APG_ERR_BAD_CONFIG;
    return 0;
  }
  if (is_hutter(c)) {
    if (c->mode != MODE_CONCURRENT) return APG_ERR_UNSUPPORTED;
    if (c->out_dim != action_dim(c->system) * c->horizon) return APG_ERR_BAD_CONFIG;
    if (c->net == NET_HUTTER_CONV) {
      if (c->system != SYS_QUAD) return APG_ERR_UNSUPPORTED;
      if (c->ref_len < 3 || c->ref_len > 11 || 3 * c->ref_dim > 32) return APG_ERR_BAD_CONFIG;
    } else {
      if (c->system != SYS_WING) return APG_ERR_UNSUPPORTED;
      if (c->ref_len * c->ref_dim > 32) return APG_ERR_BAD_CONFIG;
    }
    if (c->state_feat < 1 || c->state_feat > 32) return APG_ERR_BAD_CONFIG;
    return 0;
  }
  if (c->net == NET_LSTM) {
    if (c->mode != MODE_LSTM || c->system != SYS_QUAD) return APG_ERR_UNSUPPORTED;
    if (c->out_dim != 4 || c->state_feat != 15 || c->ref_dim != 9 || c->ref_len != c->horizon) return APG_ERR_BAD_CONFIG;
    if (c->horizon < 3 || c->horizon > 10) return APG_ERR_BAD_CONFIG;
    if (c->window != WINDOW_CUMULATIVE && c->window != WINDOW_RELATIVE) return APG_ERR_BAD_CONFIG;
    return 0;
  }
  if (c->net == NET_SIMPLE) {
    if (c->mode != MODE_CONCURRENT || c->system != SYS_CARTPOLE) return APG_ERR_UNSUPPORTED;
    if (c->out_dim != c->horizon || c->state_feat != 4 || c->horizon < 2) return APG_ERR_BAD_CONFIG;
    return 0;
  }
  return APG_ERR_UNSUPPORTED;
}

SimpleLayout simple_layout(const apg_config* c) { return make_simple_layout(c->state_feat, c->out_dim); }
LstmLayout lstm_layout(const apg_config* c) { return make_lstm_layout(c->state_feat, c->ref_len, c->ref_dim, c->out_dim); }

// per-net sizes the workspace plan needs
struct NetInfo { int n_params, f_total, b_total, x1_rows, h_rows, act_rows, steps; };
bool is_recurrent(const apg_config* c) { return c->mode != MODE_CONCURRENT; }

NetInfo net_info(const apg_config* c);

HutterLayout hutter_layout(const apg_config* c) {
  return make_hutter_layout(c->state_feat, c->ref_len, c->ref_dim, c->out_dim, c->net == NET_HUTTER_CONV);
}

struct Plan {
  int grid, ntiles;
  size_t o_wf, o_wb, o_lossp, o_gradp, o_x1, o_h1, o_h2, o_h3, o_act, o_states, o_hdr, o_tq_w, o_tq_t, o_tq_f, o_tq_z,
      total;
  int tq_grid, tq_dyn_grid;
};

NetInfo net_info(const apg_config* c) {
  NetInfo n;
  if (is_hutter(c)) {
    const HutterLayout y = hutter_layout(c);
    n.n_params = y.n_params; n.f_total = y.f_total; n.b_total = y.b_total;
    n.x1_rows = y.K1; n.h_rows = HID; n.act_rows = y.Mo4;
    n.steps = is_recurrent(c) ? c->horizon : 1;
  } else if (c->net == NET_LSTM) {
    const LstmLayout y = lstm_layout(c);
    n.n_params = y.n_params; n.f_total = y.f_total; n.b_total = y.b_total;
    n.x1_rows = y.ROWS; n.h_rows = 0; n.act_rows = 0; n.steps = c->horizon;
  } else {
    const SimpleLayout y = simple_layout(c);
    n.n_params = y.n_params; n.f_total = y.f_total; n.b_total = y.b_total;
    n.x1_rows = y.rows_total; n.h_rows = 0; n.act_rows = 0; n.steps = 1;
  }
  return n;
}

Plan make_plan(const apg_config* c, const NetInfo& y) {
  Plan p;
  p.ntiles = (c->n_drones + TM - 1) / TM;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  p.grid = p.ntiles < sms ? p.ntiles : sms;
  const int S = state_dim(c->system);
  size_t o = 0;
  p.o_wf = o;     o += up256(sizeof(float) * y.f_total);
  p.o_wb = o;     o += up256(sizeof(float) * y.b_total);
  p.o_lossp = o;  o += up256(sizeof(float) * 1024);
  p.o_gradp = o;  o += up256(sizeof(float) * (size_t)sms * y.n_params);
  const size_t nst = (size_t)p.ntiles * y.steps;
  p.o_x1 = o;     o += up256(sizeof(float) * nst * y.x1_rows * TMP);
  p.o_h1 = o;     o += up256(sizeof(float) * nst * y.h_rows * TMP);
  p.o_h2 = o;     o += up256(sizeof(float) * nst * y.h_rows * TMP);
  p.o_h3 = o;     o += up256(sizeof(float) * nst * y.h_rows * TMP);
  p.o_act = o;    o += up256(sizeof(float) * nst * y.act_rows * TMP);
  p.o_states = o; o += up256(sizeof(float) * (size_t)p.ntiles * c->horizon * S * TMP);
  // header: which forward variant produced the stash / weight images of this workspace (read by backward)
  p.o_hdr = o;    o += 256;
  // tcgen05 path (quadrotor concurrent): forward / transposed weight images, X stash and dZ stash in operand-image
  // format (tq_layout.cuh)
  const bool tq_cfg = is_hutter(c) && !is_recurrent(c) && c->system == SYS_QUAD && tq_supported(hutter_layout(c), c->horizon);
  p.o_tq_w = o;   o += tq_cfg ? up256(tq_blob_bytes()) : 0;
  p.o_tq_t = o;   o += tq_cfg ? up256(tq_tblob_bytes()) : 0;
  p.o_tq_f = o;   o += tq_cfg ? up256(tq_fstash_bytes(c->n_drones)) + 1024 : 0;
  p.o_tq_z = o;   o += tq_cfg ? up256(tq_zstash_bytes(c->n_drones)) + 1024 : 0;
  p.tq_grid = tq_cfg ? tq_grid(c->n_drones, sms) : 0;
  p.tq_dyn_grid = tq_cfg ? tq_dyn_grid(c->n_drones, sms) : 0;
  p.total = o;
  return p;
}

RolloutArgs make_args(const apg_config* c, const Plan& p, const float* in_state, const float* cur, const float* in_ref,
                      const float* ref, const float* h0c0, void* workspace) {
  RolloutArgs a;
  memset(&a, 0, sizeof(a));
  char* w = static_cast<char*>(workspace);
  a.in_state = in_state; a.cur = cur; a.in_ref = in_ref; a.ref = ref; a.h0c0 = h0c0;
  a.N = c->n_drones; a.h = c->horizon; a.ref_rows = is_recurrent(c) ? 2 * c->horizon : c->horizon;
  a.window = c->window; a.dt = c->dt;
  a.raw_inputs = (!is_recurrent(c) && is_hutter(c) && !in_state && !in_ref) ? 1 : 0;
  memcpy(a.pc.v, c->phys, sizeof(float) * MAX_PHYS);
  a.wf = reinterpret_cast<float*>(w + p.o_wf);
  a.wb = reinterpret_cast<float*>(w + p.o_wb);
  a.loss_partials = reinterpret_cast<float*>(w + p.o_lossp);
  a.grad_partials = reinterpret_cast<float*>(w + p.o_gradp);
  a.st_x1 = reinterpret_cast<float*>(w + p.o_x1);
  a.st_h1 = reinterpret_cast<float*>(w + p.o_h1);
  a.st_h2 = reinterpret_cast<float*>(w + p.o_h2);
  a.st_h3 = reinterpret_cast<float*>(w + p.o_h3);
  a.st_act = reinterpret_cast<float*>(w + p.o_act);
  a.st_states = reinterpret_cast<float*>(w + p.o_states);
  return a;
}

bool env_flag(const char* name);
bool use_tq(const apg_config* c, const HutterLayout& y);

int check_ptrs(const apg_config* c, const float* params, const float* in_state, const float* cur, const float* in_ref,
               const float* ref, const float* h0c0, void* workspace) {
  if (!params || !cur || !workspace) return APG_ERR_BAD_CONFIG;
  // the LSTM policy starts from the caller's (h0, c0) in the forward AND in the adjoint
  if (c->net == NET_LSTM && !h0c0) return APG_ERR_BAD_CONFIG;
  if (!aligned16(h0c0)) return APG_ERR_ALIGNMENT;
  // recurrent modes featurise `cur` in-kernel; the tcgen05 path takes RAW samples when in_state and in_ref are both
  // NULL (cur = raw states, ref = raw reference rows: QuadDataset.prepare_data runs in the kernels' prologue)
  const bool raw_ok = is_hutter(c) && !is_recurrent(c) && use_tq(c, hutter_layout(c)) && !in_state && !in_ref && ref;
  if (!in_state && !is_recurrent(c) && !raw_ok) return APG_ERR_BAD_CONFIG;
  if (c->system != SYS_CARTPOLE && ((!in_ref && !raw_ok) || !ref)) return APG_ERR_BAD_CONFIG;
  if (!aligned16(params) || !aligned16(in_state) || !aligned16(cur) || !aligned16(in_ref) || !aligned16(ref) ||
      (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return APG_ERR_ALIGNMENT;
  return 0;
}

bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1';
}

// The tcgen05 path (tq_kernels.cu / tq_dw_kernels.cu) is THE path of the quadrotor concurrent configuration it is
// written for (Net(15,10,9,40,conv), h = 10); APG_LEGACY_MMA=1 selects the mma.sync kernels instead (kept for the other
// configurations and as a cross-check).
bool use_tq(const apg_config* c, const HutterLayout& y) {
  if (env_flag("APG_LEGACY_MMA")) return false;
  return c->system == SYS_QUAD && c->mode == MODE_CONCURRENT && tq_supported(y, c->horizon);
}
inline unsigned char* align1024(unsigned char* p) {
  return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
enum FwdVariant { FWD_LEGACY = 1, FWD_TQ = 3 };

// Optional per-kernel device timing of the tcgen05 path (apg_debug_timing / apg_debug_kernel_times): CUDA events on
// the launching stream between the launches, so that a caller (bench.py) can attribute the step time to kernels
// without a profiler.  Off by default; not part of any timed measurement.
#ifndef APG_SIM
constexpr int N_TMARK = 9;                 // [pack | fwd | dyn | sum_loss] [dx | dw | reduce]
bool g_timing = false;
cudaEvent_t g_tev[N_TMARK] = {};
void tmark(int i, cudaStream_t st) {
  if (!g_timing) return;
  if (!g_tev[i]) cudaEventCreate(&g_tev[i]);
  cudaEventRecord(g_tev[i], st);
}
#else
void tmark(int, cudaStream_t) {}
#endif

// cached device buffer and stream of the host-buffer entry point, one per device; a call holds its device's entry
// locked from the first copy to the final synchronize (the entry point is synchronous anyway)
struct HostCache {
  void* buf = nullptr;
  size_t cap = 0;
  cudaStream_t stream = nullptr;
  std::mutex busy;
} g_caches[APG_MAX_DEVICES];

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int apg_version(void) { return 1; }
__attribute__((visibility("default"))) int apg_sm_count(void) { return sm_count(); }

__attribute__((visibility("default"))) const char* apg_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case APG_ERR_BAD_CONFIG: return "apg: bad configuration or null pointer";
    case APG_ERR_UNSUPPORTED: return "apg: unsupported system/net/mode combination";
    case APG_ERR_ALIGNMENT: return "apg: pointer not 16-byte aligned (workspace: 256)";
    case APG_ERR_NO_DEVICE: return "apg: no CUDA device";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "apg: unknown error";
  }
}

// per-kernel device times of the LAST forward + backward on the tcgen05 path (see tmark above): ms_out[7] =
// pack, forward chain, dynamics + reverse sweep, loss sum, dX chain, dW GEMM, gradient reduce.  Synchronises.
__attribute__((visibility("default"))) int apg_debug_timing(int enable) {
#ifndef APG_SIM
  g_timing = enable != 0;
#endif
  return 0;
}
__attribute__((visibility("default"))) int apg_debug_kernel_times(float* ms_out) {
#ifndef APG_SIM
  if (!ms_out) return APG_ERR_BAD_CONFIG;
  for (int i = 0; i < N_TMARK; ++i) if (!g_tev[i]) return APG_ERR_BAD_CONFIG;
  cudaError_t ce = cudaEventSynchronize(g_tev[N_TMARK - 1]);
  if (ce) return (int)ce;
  const int pairs[7][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {5, 6}, {6, 7}, {7, 8}};
  for (int k = 0; k < 7; ++k)
    if ((ce = cudaEventElapsedTime(&ms_out[k], g_tev[pairs[k][0]], g_tev[pairs[k][1]]))) return (int)ce;
#endif
  return 0;
}

// which kernels apg_rollout_forward / backward launch for this configuration (and environment): 1 = the tcgen05 /
// TMEM path (tq_*_kernel), 0 = the mma.sync / FFMA tile-engine kernels
__attribute__((visibility("default"))) int apg_rollout_kernel_path(const apg_config* cfg) {
  const int e = check_config(cfg);
  if (e) return e;
  return (is_hutter(cfg) && !is_recurrent(cfg) && use_tq(cfg, hutter_layout(cfg))) ? 1 : 0;
}

__attribute__((visibility("default"))) int apg_num_params(const apg_config* cfg) {
  const int e = check_config(cfg);
  if (e) return e;
  return net_info(cfg).n_params;
}

__attribute__((visibility("default"))) size_t apg_workspace_bytes(const apg_config* cfg) {
  if (check_config(cfg)) return 0;
  return make_plan(cfg, net_info(cfg)).total;
}

}  // extern "C"

namespace {
int rollout_forward_impl(const apg_config* cfg, const float* params, const float* learnt, const float* in_state,
                         const float* cur, const float* in_ref, const float* ref, const float* h0c0, void* workspace,
                         float* loss, float* states_out, float* actions_out, void* stream) {
  int e = check_config(cfg);
  if (e) return e;
  if ((e = check_ptrs(cfg, params, in_state, cur, in_ref, ref, h0c0, workspace))) return e;
  if (sm_count() <= 0) return APG_ERR_NO_DEVICE;
  // learnt dynamics inside the horizon: the dynamics kernel of the tcgen05 path has the variant, nothing else does
  if (learnt && !(is_hutter(cfg) && !is_recurrent(cfg) && use_tq(cfg, hutter_layout(cfg)))) return APG_ERR_UNSUPPORTED;
  if (!aligned16(learnt)) return APG_ERR_ALIGNMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(cfg, net_info(cfg));
  RolloutArgs a = make_args(cfg, p, in_state, cur, in_ref, ref, h0c0, workspace);
  a.states_out = states_out;
  a.actions_out = actions_out;
  a.learnt = learnt;
  cudaError_t ce;
  if (is_hutter(cfg)) {
    const HutterLayout y = hutter_layout(cfg);
    // the mma.sync kernels' packed weights; not needed when forward AND both adjoint halves run on the tcgen05 images
    const bool tq = !is_recurrent(cfg) && use_tq(cfg, y);
    const int variant = tq ? FWD_TQ : FWD_LEGACY;
    // stamp the workspace with the variant that fills its stash (one byte value, capturable memset node); the adjoint
    // kernels of the tcgen05 path check it on the device and poison the gradient on a mismatch
    if ((ce = cudaMemsetAsync(static_cast<char*>(workspace) + p.o_hdr, variant, 16, st))) return (int)ce;
    // the mma.sync kernels' packed weights (not needed on the tcgen05 path, which packs its own images)
    if (!tq &&
        (ce = launch_pack(hutter_pack_table(y), params, const_cast<float*>(a.wf), const_cast<float*>(a.wb), st)))
      return (int)ce;
    if (is_recurrent(cfg)) { if ((ce = launch_rec_fwd(y, a, p.grid, st))) return (int)ce; }
    else if (tq) {
      unsigned char* w = static_cast<unsigned char*>(workspace);
      unsigned char* fs = align1024(w + p.o_tq_f);
      unsigned char* zs = align1024(w + p.o_tq_z);
      tmark(0, st);
      if ((ce = launch_tq_pack(y, params, w + p.o_tq_w, w + p.o_tq_t, st))) return (int)ce;
      tmark(1, st);
      if ((ce = launch_tq_fwd(w + p.o_tq_w, a, fs, p.tq_grid, st))) return (int)ce;
      tmark(2, st);
      // the loss sum rides on the dynamics kernel (last block; the ticket word is part of the stamp set above)
      if ((ce = launch_tq_dyn(a, fs, zs, loss, reinterpret_cast<unsigned*>(w + p.o_hdr + 8), 0x01010101u * FWD_TQ,
                              p.tq_dyn_grid, st)))
        return (int)ce;
      tmark(3, st);
      tmark(4, st);
      return 0;
    }
    else if ((ce = launch_hutter_fwd(cfg->system, y, a, p.grid, st))) return (int)ce;
  } else if (cfg->net == NET_LSTM) {
    if (!h0c0) return APG_ERR_BAD_CONFIG;
    const LstmLayout y = lstm_layout(cfg);
    if ((ce = launch_pack(lstm_pack_table(y), params, const_cast<float*>(a.wf), const_cast<float*>(a.wb), st)))
      return (int)ce;
    if ((ce = launch_lstm_fwd(y, a, p.grid, st))) return (int)ce;
  } else {
    const SimpleLayout y = simple_layout(cfg);
    if ((ce = launch_pack(simple_pack_table(y), params, const_cast<float*>(a.wf), const_cast<float*>(a.wb), st)))
      return (int)ce;
    if ((ce = launch_simple_fwd(y, a, p.grid, st))) return (int)ce;
  }
  if (loss && (ce = launch_sum_loss(a.loss_partials, p.grid, loss, st))) return (int)ce;
  return 0;
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) int apg_rollout_forward(const apg_config* cfg, const float* params, const float* in_state, const float* cur,
                        const float* in_ref, const float* ref, const float* h0c0, void* workspace, float* loss,
                        float* states_out, float* actions_out, void* stream) {
  return rollout_forward_impl(cfg, params, nullptr, in_state, cur, in_ref, ref, h0c0, workspace, loss, states_out,
                              actions_out, stream);
}

// apg_rollout_forward with the h dynamics steps taken by the LEARNT residual model (LearntDynamics.forward,
// neural_control/dynamics/quad_dynamics_trained.py:58-69) instead of the analytic one: the controller-training phase of
// run_dynamics (scripts/train_base.py:334-375 -> train_drone.py:175-199 with train_dynamics = LearntDynamics) as one
// fused rollout.  `learnt_params`: the 1891 floats of apg_learnt_num_params(APG_SYS_QUAD), named_parameters() order;
// cfg->phys carries the construction-time constants of the learnt object.  The reverse sweep (d loss / d logits through
// the learnt steps) is part of this call, so the adjoint of it is the ordinary apg_rollout_backward / _sgd / _p2p.
// Configurations served by the tcgen05 path only (apg_rollout_kernel_path(cfg) == 1), APG_ERR_UNSUPPORTED otherwise.
__attribute__((visibility("default"))) int apg_rollout_forward_learnt(const apg_config* cfg, const float* params,
                        const float* learnt_params, const float* in_state, const float* cur, const float* in_ref,
                        const float* ref, void* workspace, float* loss, float* states_out, float* actions_out,
                        void* stream) {
  if (!learnt_params) return APG_ERR_BAD_CONFIG;
  return rollout_forward_impl(cfg, params, learnt_params, in_state, cur, in_ref, ref, nullptr, workspace, loss,
                              states_out, actions_out, stream);
}

}  // extern "C"

namespace {
// gradient reduction that ends the adjoint pass: into `grad_params` (default), or - `comm` given - straight into
// every rank's receive slot over peer memory (p2p_kernels.cu)
cudaError_t finish_gradient(const float* partials, int ncta, int n, float scale, float* grad_params,
                            const apg_grad_comm* comm, int pm_off, int pm_k1, int pm_npos, bool sliced,
                            cudaStream_t st) {
  if (comm)
    return launch_reduce_scatter_p2p(partials, ncta, n, scale, pm_off, pm_k1, pm_npos,
                                     static_cast<float* const*>(comm->slot_ptrs),
                                     static_cast<unsigned* const*>(comm->flag_ptrs), comm->rank, comm->world,
                                     comm->epoch, static_cast<unsigned*>(comm->ticket), st);
  if (sliced) return launch_reduce_grad4(partials, ncta, n, scale, grad_params, st);
  return launch_reduce_grad(partials, ncta, n, scale, grad_params, st, pm_off, pm_k1, pm_npos);
}

// optimizer step fused into the gradient reduction (tcgen05 path only)
struct SgdFuse { float* params; float* momentum_buf; float lr, momentum; };

int rollout_backward_impl(const apg_config* cfg, const float* params, const float* in_state, const float* cur,
                          const float* in_ref, const float* ref, const float* h0c0, void* workspace, float grad_loss,
                          float* grad_params, const apg_grad_comm* comm, void* stream, const SgdFuse* sgd = nullptr) {
  int e = check_config(cfg);
  if (e) return e;
  if ((e = check_ptrs(cfg, params, in_state, cur, in_ref, ref, h0c0, workspace))) return e;
  if (!grad_params && !comm && !sgd) return APG_ERR_BAD_CONFIG;
  if (comm && (!comm->slot_ptrs || !comm->flag_ptrs || !comm->ticket || comm->world < 1 || comm->rank < 0 ||
               comm->rank >= comm->world))
    return APG_ERR_BAD_CONFIG;
  if (sm_count() <= 0) return APG_ERR_NO_DEVICE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const NetInfo ni = net_info(cfg);
  const Plan p = make_plan(cfg, ni);
  RolloutArgs a = make_args(cfg, p, in_state, cur, in_ref, ref, h0c0, workspace);
  cudaError_t ce;
  if (is_hutter(cfg)) {
    if (is_recurrent(cfg)) { if ((ce = launch_rec_adj(hutter_layout(cfg), a, p.grid, st))) return (int)ce; }
    else if (use_tq(cfg, hutter_layout(cfg))) {
      // tcgen05 path: dX chain (reverse sweep + dZ stash), then the streaming weight-gradient GEMM on the two stashes
      const HutterLayout y = hutter_layout(cfg);
      unsigned char* w = static_cast<unsigned char*>(workspace);
      unsigned char* fs = align1024(w + p.o_tq_f);
      unsigned char* zs = align1024(w + p.o_tq_z);
      tmark(5, st);
      if ((ce = launch_tq_dx(w + p.o_tq_t, a, fs, zs, w + p.o_hdr, FWD_TQ, p.tq_grid, st))) return (int)ce;
      tmark(6, st);
      if ((ce = launch_tq_dw(y, a, fs, zs, p.tq_grid, st))) return (int)ce;
      tmark(7, st);
      if (sgd) {
        if ((ce = launch_reduce_grad4_sgd(a.grad_partials, p.tq_grid, ni.n_params, grad_loss, grad_params, sgd->params,
                                          sgd->momentum_buf, sgd->lr, sgd->momentum, st)))
          return (int)ce;
      } else if ((ce = finish_gradient(a.grad_partials, p.tq_grid, ni.n_params, grad_loss, grad_params, comm, 0, 0, 0,
                                       true, st)))
        return (int)ce;
      tmark(8, st);
      return 0;
    }
    else if (sgd) return APG_ERR_UNSUPPORTED;
    else if ((ce = launch_hutter_adj(cfg->system, hutter_layout(cfg), a, p.grid, st))) return (int)ce;
  } else if (sgd) {
    return APG_ERR_UNSUPPORTED;
  } else if (cfg->net == NET_LSTM) {
    if ((ce = launch_lstm_adj(lstm_layout(cfg), a, p.grid, st))) return (int)ce;
  } else {
    if ((ce = launch_simple_adj(simple_layout(cfg), a, p.grid, st))) return (int)ce;
  }
  int pm_off = 0, pm_k1 = 0, pm_npos = 0;
  if (is_hutter(cfg)) {
    const HutterLayout y = hutter_layout(cfg);
    pm_off = y.t_w1; pm_k1 = y.K1; pm_npos = y.perm_npos;
  }
  if ((ce = finish_gradient(a.grad_partials, p.grid, ni.n_params, grad_loss, grad_params, comm, pm_off, pm_k1, pm_npos,
                            false, st)))
    return (int)ce;
  return 0;
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) int apg_rollout_backward(const apg_config* cfg, const float* params, const float* in_state, const float* cur,
                         const float* in_ref, const float* ref, const float* h0c0, void* workspace, float grad_loss,
                         float* grad_params, void* stream) {
  if (!grad_params) return APG_ERR_BAD_CONFIG;
  return rollout_backward_impl(cfg, params, in_state, cur, in_ref, ref, h0c0, workspace, grad_loss, grad_params,
                               nullptr, stream);
}

// apg_rollout_backward + the reference's optimizer step (optim.SGD(momentum): buf = momentum * buf + g; p -= lr * buf,
// train_base.py:139-143) fused into the gradient reduction: `params_rw` (the vector the forward read) and
// `momentum_buf` are updated in place; `grad_params` may be NULL.  Configurations served by the tcgen05 path only
// (apg_rollout_kernel_path(cfg) == 1), APG_ERR_UNSUPPORTED otherwise.
__attribute__((visibility("default"))) int apg_rollout_backward_sgd(const apg_config* cfg, float* params_rw, const float* in_state, const float* cur,
                             const float* in_ref, const float* ref, const float* h0c0, void* workspace,
                             float grad_loss, float* grad_params, float* momentum_buf, float lr, float momentum,
                             void* stream) {
  if (!params_rw || !momentum_buf) return APG_ERR_BAD_CONFIG;
  if (!aligned16(momentum_buf) || (grad_params && !aligned16(grad_params))) return APG_ERR_ALIGNMENT;
  const SgdFuse sgd{params_rw, momentum_buf, lr, momentum};
  return rollout_backward_impl(cfg, params_rw, in_state, cur, in_ref, ref, h0c0, workspace, grad_loss, grad_params,
                               nullptr, stream, &sgd);
}

// ---- the gradient all-reduce as this library's own kernels over NVLink peer memory (p2p_kernels.cu, p2p_math.cuh)
__attribute__((visibility("default"))) size_t apg_grad_comm_bytes(int world, int n_params) {
  if (world < 1 || n_params < 1) return 0;
  GradCommLayout L{world, n_params};
  return sizeof(float) * L.total_floats();
}

__attribute__((visibility("default"))) int apg_grad_comm_offsets(int world, int n_params, int set, size_t* slots_offset_bytes, size_t* flags_offset_bytes) {
  if (world < 1 || n_params < 1 || set < 0 || set > 1 || !slots_offset_bytes || !flags_offset_bytes)
    return APG_ERR_BAD_CONFIG;
  GradCommLayout L{world, n_params};
  *slots_offset_bytes = sizeof(float) * L.slot_off(set, 0);
  *flags_offset_bytes = sizeof(float) * L.flag_off(set, 0);
  return 0;
}

__attribute__((visibility("default"))) int apg_rollout_backward_p2p(const apg_config* cfg, const float* params, const float* in_state, const float* cur,
                             const float* in_ref, const float* ref, const float* h0c0, void* workspace,
                             float grad_loss, const apg_grad_comm* comm, void* stream) {
  if (!comm) return APG_ERR_BAD_CONFIG;
  return rollout_backward_impl(cfg, params, in_state, cur, in_ref, ref, h0c0, workspace, grad_loss, nullptr, comm,
                               stream);
}

__attribute__((visibility("default"))) int apg_grad_gather_sgd_p2p(const apg_grad_comm* comm, const void* local_set, int n_params, float* grad_out,
                            float* params, float* momentum_buf, float lr, float momentum, void* stream) {
  if (!comm || !local_set || n_params < 1 || comm->world < 1) return APG_ERR_BAD_CONFIG;
  if (params && !momentum_buf) return APG_ERR_BAD_CONFIG;
  if (sm_count() <= 0) return APG_ERR_NO_DEVICE;
  const float* slots = static_cast<const float*>(local_set);
  const unsigned* flags = reinterpret_cast<const unsigned*>(slots + (size_t)comm->world * n_params);
  return (int)launch_gather_sgd_p2p(slots, flags, comm->world, n_params, comm->epoch, grad_out, params, momentum_buf,
                                    lr, momentum, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int apg_rollout_value_and_grad_host(const apg_config* cfg, const float* params_host, const float* in_state_host,
                                    const float* cur_host, const float* in_ref_host, const float* ref_host,
                                    const float* h0c0_host, float* loss_host, float* grad_params_host) {
  int e = check_config(cfg);
  if (e) return e;
  if (sm_count() <= 0) return APG_ERR_NO_DEVICE;
  const NetInfo y = net_info(cfg);
  const int N = cfg->n_drones, h = cfg->horizon, S = state_dim(cfg->system);
  const int refw = cfg->system == SYS_QUAD ? 9 : (cfg->system == SYS_WING ? 3 : 0);
  const size_t b_params = up256(sizeof(float) * y.n_params);
  const size_t b_ins = up256(sizeof(float) * (size_t)N * cfg->state_feat);
  const size_t b_cur = up256(sizeof(float) * (size_t)N * S);
  const int rec = is_recurrent(cfg) ? 2 : 1;        // recurrent modes carry 2h reference rows
  const size_t n_inr = (size_t)N * cfg->ref_len * cfg->ref_dim * rec, n_ref = (size_t)N * h * refw * rec;
  const size_t b_inr = up256(sizeof(float) * n_inr + 16);
  const size_t b_ref = up256(sizeof(float) * n_ref + 16);
  const size_t n_hc = cfg->net == NET_LSTM ? (size_t)2 * N * LSTM_HS : 0;
  const size_t b_hc = up256(sizeof(float) * n_hc + 16);
  const size_t b_ws = apg_workspace_bytes(cfg);
  const size_t need = 2 * b_params + b_ins + b_cur + b_inr + b_ref + b_hc + 256 + b_ws;
  cudaError_t ce;
  const int dev = current_device();
  if (dev < 0) return APG_ERR_BAD_CONFIG;
  HostCache& g_cache = g_caches[dev];
  std::lock_guard<std::mutex> cache_lock(g_cache.busy);
  if (!g_cache.stream && (ce = cudaStreamCreateWithFlags(&g_cache.stream, cudaStreamNonBlocking))) return (int)ce;
  if (g_cache.cap < need) {
    if (g_cache.buf) cudaFree(g_cache.buf);
    g_cache.buf = nullptr;
    g_cache.cap = 0;
    if ((ce = cudaMalloc(&g_cache.buf, need))) return (int)ce;
    g_cache.cap = need;
  }
  char* b = static_cast<char*>(g_cache.buf);
  float* d_params = reinterpret_cast<float*>(b); b += b_params;
  float* d_grad = reinterpret_cast<float*>(b);   b += b_params;
  float* d_ins = reinterpret_cast<float*>(b);    b += b_ins;
  float* d_cur = reinterpret_cast<float*>(b);    b += b_cur;
  float* d_inr = reinterpret_cast<float*>(b);    b += b_inr;
  float* d_ref = reinterpret_cast<float*>(b);    b += b_ref;
  float* d_hc = reinterpret_cast<float*>(b);     b += b_hc;
  float* d_loss = reinterpret_cast<float*>(b);   b += 256;
  void* d_ws = b;
  cudaStream_t st = g_cache.stream;
#define APG_H2D(dst, src, bytes) \
  if ((src) && (bytes) && (ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st))) return (int)ce;
  APG_H2D(d_params, params_host, sizeof(float) * y.n_params)
  APG_H2D(d_ins, in_state_host, sizeof(float) * (size_t)N * cfg->state_feat)
  APG_H2D(d_cur, cur_host, sizeof(float) * (size_t)N * S)
  APG_H2D(d_inr, in_ref_host, sizeof(float) * n_inr)
  APG_H2D(d_ref, ref_host, sizeof(float) * n_ref)
  APG_H2D(d_hc, h0c0_host, sizeof(float) * n_hc)
#undef APG_H2D
  const float* p_hc = (h0c0_host && n_hc) ? d_hc : nullptr;
  const float* p_ins = in_state_host ? d_ins : nullptr;
  const float* p_inr = in_ref_host ? d_inr : nullptr;
  const float* p_ref = ref_host ? d_ref : nullptr;
  if ((e = apg_rollout_forward(cfg, d_params, p_ins, d_cur, p_inr, p_ref, p_hc, d_ws, d_loss, nullptr, nullptr, st)))
    return e;
  if ((e = apg_rollout_backward(cfg, d_params, p_ins, d_cur, p_inr, p_ref, p_hc, d_ws, 1.0f, d_grad, st))) return e;
  if (loss_host && (ce = cudaMemcpyAsync(loss_host, d_loss, sizeof(float), cudaMemcpyDeviceToHost, st))) return (int)ce;
  if (grad_params_host &&
      (ce = cudaMemcpyAsync(grad_params_host, d_grad, sizeof(float) * y.n_params, cudaMemcpyDeviceToHost, st)))
    return (int)ce;
  if ((ce = cudaStreamSynchronize(st))) return (int)ce;
  return 0;
}

__attribute__((visibility("default"))) int apg_dynamics_step(int system, const float* phys, const float* state, const float* action, float dt, int n,
                      float* out, void* stream) {
  if (!phys || !state || !action || !out || n < 0) return APG_ERR_BAD_CONFIG;
  PhysConsts pc;
  memcpy(pc.v, phys, sizeof(float) * MAX_PHYS);
  return (int)launch_step(system, pc, state, action, dt, n, out, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int apg_dynamics_step_adjoint(int system, const float* phys, const float* state, const float* action, float dt, int n,
                              const float* grad_out, float* grad_state, float* grad_action, void* stream) {
  if (!phys || !state || !action || !grad_out || !grad_state || !grad_action || n < 0) return APG_ERR_BAD_CONFIG;
  PhysConsts pc;
  memcpy(pc.v, phys, sizeof(float) * MAX_PHYS);
  return (int)launch_step_adj(system, pc, state, action, dt, n, grad_out, grad_state, grad_action,
                              static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int apg_quad_features(const float* state, int n, float* feat, void* stream) {
  if (!state || !feat || n < 0) return APG_ERR_BAD_CONFIG;
  return (int)launch_features(state, n, feat, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int apg_quad_features_adjoint(const float* state, const float* grad_feat, int n, float* grad_state, void* stream) {
  if (!state || !grad_feat || !grad_state || n < 0) return APG_ERR_BAD_CONFIG;
  return (int)launch_features_adj(state, grad_feat, n, grad_state, static_cast<cudaStream_t>(stream));
}

// Closed-loop evaluation on table references (eval_kernels.cu).
__attribute__((visibility("default"))) int apg_eval_rollout(const apg_config* cfg, const float* params, const float* tables, const int* table_index,
                     int n_tables, int table_rows, const float* init_states, int steps, float thresh_div,
                     float thresh_stable, int test_time, void* workspace, float* states_out, float* div_out,
                     float* actions_out, int* n_steps_out, void* stream) {
  int e = check_config(cfg);
  if (e) return e;
  if (!is_hutter(cfg) || cfg->net != NET_HUTTER_CONV || cfg->system != SYS_QUAD) return APG_ERR_UNSUPPORTED;
  if (cfg->state_feat != 15 || cfg->ref_dim != 9 || cfg->ref_len != cfg->horizon) return APG_ERR_BAD_CONFIG;
  if (!params || !tables || !init_states || !workspace) return APG_ERR_BAD_CONFIG;
  if (n_tables < 1 || table_rows < 1 || steps < 1) return APG_ERR_BAD_CONFIG;
  if (!table_index && n_tables < cfg->n_drones) return APG_ERR_BAD_CONFIG;
  if (!aligned16(params) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return APG_ERR_ALIGNMENT;
  if (sm_count() <= 0) return APG_ERR_NO_DEVICE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(cfg, net_info(cfg));
  char* w = static_cast<char*>(workspace);
  float* wf = reinterpret_cast<float*>(w + p.o_wf);
  float* wb = reinterpret_cast<float*>(w + p.o_wb);
  const HutterLayout y = hutter_layout(cfg);
  cudaError_t ce;
  if ((ce = launch_pack(hutter_pack_table(y), params, wf, wb, st))) return (int)ce;
  PhysConsts pc;
  memcpy(pc.v, cfg->phys, sizeof(float) * MAX_PHYS);
  EvalParams ev;
  ev.steps = steps; ev.table_rows = table_rows; ev.test_time = test_time ? 1 : 0;
  ev.thresh_div = thresh_div; ev.thresh_stable = thresh_stable;
  if ((ce = launch_eval_rollout(y, wf, tables, table_index, init_states, cfg->n_drones, cfg->dt, pc, ev, states_out,
                                div_out, actions_out, n_steps_out, p.grid, st)))
    return (int)ce;
  return 0;
}

// apg_eval_rollout for the LSTM policy (train_mode "LSTM", models/rnn.py LSTM_NEW): h0c0 [2][N][8] is every drone's
// hidden / cell state before its first policy call, hc_out (optional) the state after its last one.
__attribute__((visibility("default"))) int apg_eval_rollout_lstm(const apg_config* cfg, const float* params, const float* h0c0, const float* tables,
                          const int* table_index, int n_tables, int table_rows, const float* init_states, int steps,
                          float thresh_div, float thresh_stable, int test_time, void* workspace, float* states_out,
                          float* div_out, float* actions_out, int* n_steps_out, float* hc_out, void* stream) {
  int e = check_config(cfg);
  if (e) return e;
  if (cfg->net != NET_LSTM || cfg->system != SYS_QUAD) return APG_ERR_UNSUPPORTED;
  if (cfg->state_feat != 15 || cfg->ref_dim != 9 || cfg->ref_len != cfg->horizon || cfg->out_dim != 4) return APG_ERR_BAD_CONFIG;
  if (!params || !h0c0 || !tables || !init_states || !workspace) return APG_ERR_BAD_CONFIG;
  if (n_tables < 1 || table_rows < 1 || steps < 1) return APG_ERR_BAD_CONFIG;
  if (!table_index && n_tables < cfg->n_drones) return APG_ERR_BAD_CONFIG;
  if (!aligned16(params) || !aligned16(h0c0) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return APG_ERR_ALIGNMENT;
  if (sm_count() <= 0) return APG_ERR_NO_DEVICE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(cfg, net_info(cfg));
  char* w = static_cast<char*>(workspace);
  float* wf = reinterpret_cast<float*>(w + p.o_wf);
  float* wb = reinterpret_cast<float*>(w + p.o_wb);
  const LstmLayout y = lstm_layout(cfg);
  cudaError_t ce;
  if ((ce = launch_pack(lstm_pack_table(y), params, wf, wb, st))) return (int)ce;
  PhysConsts pc;
  memcpy(pc.v, cfg->phys, sizeof(float) * MAX_PHYS);
  EvalParams ev;
  ev.steps = steps; ev.table_rows = table_rows; ev.test_time = test_time ? 1 : 0;
  ev.thresh_div = thresh_div; ev.thresh_stable = thresh_stable;
  if ((ce = launch_eval_rollout_lstm(y, wf, h0c0, tables, table_index, init_states, cfg->n_drones, cfg->dt, pc, ev,
                                     states_out, div_out, actions_out, n_steps_out, hc_out, p.grid, st)))
    return (int)ce;
  return 0;
}

// Closed-loop evaluation of the fixed wing towards target points (eval_kernels.cu).
__attribute__((visibility("default"))) int apg_eval_fly_to_points(const apg_config* cfg, const float* params, const float* targets, int n_targets,
                           const float* init_states, const float* mean_host, const float* std_host, float dt_data,
                           int steps, float thresh_div, float thresh_stable, int test_time, void* workspace,
                           float* states_out, float* div_linear_out, float* actions_out, int* n_steps_out,
                           float* div_target_sum_out, float* div_target_cnt_out, void* stream) {
  int e = check_config(cfg);
  if (e) return e;
  if (!is_hutter(cfg) || cfg->net != NET_HUTTER_LIN || cfg->system != SYS_WING || cfg->mode != MODE_CONCURRENT)
    return APG_ERR_UNSUPPORTED;
  if (cfg->state_feat != 9 || cfg->ref_dim != 3 || cfg->ref_len != 1) return APG_ERR_BAD_CONFIG;
  if (!params || !targets || !init_states || !mean_host || !std_host || !workspace) return APG_ERR_BAD_CONFIG;
  if (n_targets < 1 || steps < 1) return APG_ERR_BAD_CONFIG;
  if (!aligned16(params) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return APG_ERR_ALIGNMENT;
  if (sm_count() <= 0) return APG_ERR_NO_DEVICE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(cfg, net_info(cfg));
  char* w = static_cast<char*>(workspace);
  float* wf = reinterpret_cast<float*>(w + p.o_wf);
  float* wb = reinterpret_cast<float*>(w + p.o_wb);
  const HutterLayout y = hutter_layout(cfg);
  cudaError_t ce;
  if ((ce = launch_pack(hutter_pack_table(y), params, wf, wb, st))) return (int)ce;
  PhysConsts pc;
  memcpy(pc.v, cfg->phys, sizeof(float) * MAX_PHYS);
  WingEvalParams ev;
  ev.steps = steps; ev.n_targets = n_targets; ev.test_time = test_time ? 1 : 0; ev.h = cfg->horizon;
  ev.thresh_div = thresh_div; ev.thresh_stable = thresh_stable;
  ev.vlen = (float)(12.0 * (double)dt_data); ev.des_speed = 11.5f;
  if ((ce = launch_eval_wing(y, wf, targets, init_states, cfg->n_drones, cfg->dt, pc, mean_host, std_host, ev,
                             states_out, div_linear_out, actions_out, n_steps_out, div_target_sum_out,
                             div_target_cnt_out, p.grid, st)))
    return (int)ce;
  return 0;
}

// Closed-loop balancing evaluation of the cartpole (eval_kernels.cu).
__attribute__((visibility("default"))) int apg_eval_cartpole(const apg_config* cfg, const float* params, const float* init_states, int steps,
                      float thresh_div, int burn_in_steps, void* workspace, float* states_out, float* actions_out,
                      int* n_steps_out, float* angle_sum_out, float* angle_cnt_out, float* vel_sum_out, void* stream) {
  int e = check_config(cfg);
  if (e) return e;
  if (cfg->net != NET_SIMPLE || cfg->system != SYS_CARTPOLE || cfg->mode != MODE_CONCURRENT) return APG_ERR_UNSUPPORTED;
  if (cfg->state_feat != 4) return APG_ERR_BAD_CONFIG;
  if (!params || !init_states || !workspace || steps < 1 || burn_in_steps < 0) return APG_ERR_BAD_CONFIG;
  if (!aligned16(params) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return APG_ERR_ALIGNMENT;
  if (sm_count() <= 0) return APG_ERR_NO_DEVICE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(cfg, net_info(cfg));
  char* w = static_cast<char*>(workspace);
  float* wf = reinterpret_cast<float*>(w + p.o_wf);
  float* wb = reinterpret_cast<float*>(w + p.o_wb);
  const SimpleLayout y = simple_layout(cfg);
  cudaError_t ce;
  if ((ce = launch_pack(simple_pack_table(y), params, wf, wb, st))) return (int)ce;
  PhysConsts pc;
  memcpy(pc.v, cfg->phys, sizeof(float) * MAX_PHYS);
  CartpoleEvalParams ev;
  ev.steps = steps; ev.burn_in = burn_in_steps; ev.thresh_div = thresh_div;
  if ((ce = launch_eval_cartpole(y, wf, init_states, cfg->n_drones, cfg->dt, pc, ev, states_out, actions_out,
                                 n_steps_out, angle_sum_out, angle_cnt_out, vel_sum_out, p.grid, st)))
    return (int)ce;
  return 0;
}

}  // extern "C"
