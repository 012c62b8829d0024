// Engine state behind rbslam_ctx (one context = one GPU = one shard of particles).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/rbslam.h"
#include "models.cuh"

enum { RB_PH_RESAMPLE = 0, RB_PH_PROPAGATE, RB_PH_MEAS, RB_PH_KALMAN, RB_PH_NORMALIZE,
       RB_PH_ANCESTOR, RB_PH_INFO, RB_PH_COUNT };

struct rbslam_ctx {
  rbslam_config cfg{};
  rb::ModelConsts mc{};
  int N = 0, T = 0, M = 0, d = 0, n = 0, nz = 0, nw = 0, n_odo = 0, ld = 0, ldh = 0;
  size_t slab = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int num_sms = 148;
  size_t smem_optin = 0, smem_small_max = 0, smem_resample_max = 0;

  // model constants on device
  int *d_NN = nullptr;

  // particle state
  double *d_P = nullptr;         // [N][M cols][ld rows] covariance slabs
  double *d_Imat = nullptr;      // information-form slabs (same slot map)
  double *d_xl[2] = {nullptr, nullptr};     // [N][M] logical order, ping-pong
  double *d_ivec[2] = {nullptr, nullptr};
  double *d_hld[2] = {nullptr, nullptr};    // halfLogDetP
  int *d_slot[2] = {nullptr, nullptr};      // logical -> physical slab
  int *d_src_slot = nullptr, *d_first_child = nullptr, *d_free_list = nullptr;
  int *d_listA = nullptr, *d_listB = nullptr, *d_counts = nullptr;
  double *d_H = nullptr, *d_yhat = nullptr;
  double *d_PHpart = nullptr, *d_G = nullptr, *d_KS = nullptr;   // legacy 3-kernel path
  // streaming path (kalman_stream.cuh): pending downdates and partial P H'
  double *d_G4[2] = {nullptr, nullptr}, *d_KS4[2] = {nullptr, nullptr}, *d_PHp = nullptr;
  int cg = 0;            // d_G4[cg]/d_KS4[cg]: pending downdate written by the last step
  bool pending = false;  // a deferred downdate is outstanding
  int kpath = 0;         // 0 shared-memory single pass, 1 streaming (lazy), 2 legacy 3-kernel
  int nsplit = 1, cw = 0;
  // sibling fusion (family_kernels.cuh): scratch [7][N], main/surplus family lists [2][5][N], counters
  int *d_fam = nullptr;
  bool use_fam = true;
  // kalman_variant 7: packed symmetric tile slabs (packed_kernels.cuh)
  bool pt = false;
  int layout = 0;            // RB_LAYOUT_FULL / _SYM / _PT: how element (r,c) of a slab is stored
  int pt_ts = 120, pt_ns = 2; // tiles per stage, ring slots (RBSLAM_PT_CFG="TS,NS")
  int pt_nw = 15;            // consumer warps per CTA (+ the service warp = 16 warps x 128 registers)
  int pt_psplit[9] = {0};    // panel ranges of the nsplit items per family
  const int *item_group = nullptr;   // sharded engine: work group of each local item (see stream_groups)
  // sharded engine, fused migration: source keys >= st_nloc name migrants whose ancestor state is read in
  // place from a peer (kalman_stream.cuh, SrcTab); d_peer_tab [kind][1 + world] base pointers
  int st_nloc = 0;
  const int *st_fetch = nullptr;
  const double **d_peer_tab = nullptr;
  int fam_slabs = 0;                 // source-key space of k_build_families (0: N)
  int stream_ctas_per_sm = 0, stream_hints = 0;   // tuning knobs (env)
  size_t hs_p = 0; int hs_a = 0, hs_c = 1;   // layout of d_H: H_i(a,c) at i*hs_p + a*hs_a + c*hs_c
  double *d_logw = nullptr, *d_w = nullptr, *d_wc = nullptr;
  double *d_Xhist = nullptr;     // [T_hist][N][n]
  int *d_Ahist = nullptr;        // [T_hist][N]
  int T_hist = 2;
  double *d_traj_max = nullptr, *d_traj_mean = nullptr;   // [T][n]
  int *d_iwmax = nullptr;        // [T]
  double *d_yhattraj = nullptr;  // [T][d] predicted measurement of the highest-weight particle (makePlots tap)
  double *d_logw_hist = nullptr, *d_w_hist = nullptr;     // [T][N] optional taps
  rb::DevStatus *d_status = nullptr;
  double *d_scratch = nullptr;   // misc outputs (xl_mean, P_mean, ...)
  size_t scratch_doubles = 0;

  // run inputs on device
  double *d_odo = nullptr;       // [T-1][n_odo] row-major
  double *d_y = nullptr;         // [T][d] row-major
  double *d_Q = nullptr;         // [pages][nw*nw]
  double *d_R = nullptr;         // [d*d]
  double *d_P0 = nullptr;        // [M*M]
  double *d_x0lin = nullptr;     // [M x cols]
  double *d_U = nullptr, *d_Z = nullptr;   // injected streams
  int *d_forced = nullptr;
  std::vector<double> h_dt, h_Uend;
  std::vector<int> h_forced_ak;
  std::vector<double> h_x0n, h_y, h_R;
  int Q_pages = 1, x0_cols = 1, run_T = 0, run_K = 1;
  bool have_U = false, have_Z = false, have_forced = false;
  double jitter = 1e-3;

  // progress
  int t = 0, sweep = 0;
  int cs = 0;   // d_slot[cs]: current logical->physical slab map
  int cx = 0;   // d_xl[cx] (and d_ivec/d_hld): current linear-state means
  bool running = false;

  // callback / counters / timing
  rbslam_step_fn step_fn = nullptr;
  void *step_user = nullptr;
  int64_t launches = 0, h2d = 0, d2h = 0;
  bool phase_timing = false;
  std::vector<cudaEvent_t> ph_events;   // pairs per recorded phase
  std::vector<int> ph_ids;
  double phase_ms[RB_PH_COUNT] = {0};
  cudaEvent_t user_events[16] = {nullptr};

  void *smoother_ws = nullptr;   // SmootherWs (smoother.cu)
  double *d_normws = nullptr;    // partials of the chunked normalisation (large N)
  int *d_chol_fail = nullptr;    // [chol_fail_cap] not-PD flags of the batched panel Cholesky
  int chol_fail_cap = 0;
  void *shard_ws = nullptr;      // ShardWs (sharded.cu); N is the LOCAL particle count when set
  std::vector<rbslam_ctx *> group;   // leader of a single-process group (rbslam_create_group): every shard, itself first
  bool group_member = false;         // a non-leader shard of such a group (destroyed with its leader)
  void *replica_group = nullptr;     // ReplicaGroup (smoother.cu): full replicas on several GPUs, smoothers
  int replica_rank = 0;
  const int *anc_override = nullptr;   // sharded engine: thin arrays are slot-indexed
  // streaming pass in groups: group g uses listA/listB + group_off[g][phase], counts d_counts[2g+phase];
  // group_hook(ctx, g) runs before group g>0 (sharded engine: wait for migrants + peer barrier)
  int stream_groups = 1;
  int group_off[2][2] = {{0, 0}, {0, 0}};
  const int *group_off_dev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // device-side offsets (planner on GPU)
  int (*group_hook)(rbslam_ctx *, int) = nullptr;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fetch = nullptr, ev_plan = nullptr;


  void fail_cuda(cudaError_t e, const char *what, const char *file, int line);
  int fail(int code, const std::string &msg) { err = msg; return code; }
};
