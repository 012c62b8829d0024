/*
 * rbslam.h -- C ABI of librbslam.so: the B200 (sm_100a) implementation of the
 * Rao-Blackwellized particle filter / smoother hot path of
 * manonkok/Rao-Blackwellized-SLAM-smoothing.
 *
 * This header is the drop-in boundary.  The reference has no FFI of its own (it
 * is interpreted MATLAB); the entry points below are what a MEX gateway binds so
 * that the three reference functions keep their MATLAB signatures:
 *
 *   particleFilter(...)                  src/particleFilter.m:1-3
 *   particleSmoother(...)                src/particleSmoother.m:1-2
 *   particleSmootherInformationForm(...) src/particleSmootherInformationForm.m:1-2
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the boundary;
 *   - every matrix is fp64, column-major, exactly as MATLAB stores it (mxGetDoubles);
 *   - the caller owns all host buffers; the library never keeps a host pointer
 *     after a call returns and owns all device memory inside the context;
 *   - indices returned to the caller are 0-based int32 (MATLAB shims add 1);
 *   - every function returns an rbslam_status; rbslam_last_error() gives text;
 *   - a context is not thread-safe (MATLAB calls mexFunction on one thread).
 *
 * The MATLAB function handles dynModel / measModel / dynResNorm cannot be called
 * from CUDA.  The caller names one of the model families below and passes the
 * constants its closures captured (rbslam_config); an unknown family is
 * RBSLAM_EMODEL -- there is no CPU fallback.
 */
#ifndef RBSLAM_H
#define RBSLAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBSLAM_VERSION 1

#if defined(__GNUC__)
#define RBSLAM_API __attribute__((visibility("default")))
#else
#define RBSLAM_API
#endif

typedef struct rbslam_ctx rbslam_ctx;

typedef enum {
  RBSLAM_OK = 0,
  RBSLAM_EARG = 1,    /* bad argument / shape / unsupported size */
  RBSLAM_ECUDA = 2,   /* CUDA or NCCL failure */
  RBSLAM_ENOTPD = 3,  /* innovation covariance not PD even after jitter
                         (MATLAB would throw from chol, src/particleFilter.m:147) */
  RBSLAM_EMODEL = 4   /* unsupported model family */
} rbslam_status;

/* Model families = the closures the reference's runners pass as handles. */
typedef enum {
  /* examples/slam-dense-mag/run_dense3D_magfield.m:265-279,301-308,202-203
     xn=[pos(3);quat(4)], odometry [dpos(3) dquat(4)], d=3, M=m_basis+3, 6 normals */
  RBSLAM_MODEL_DENSE_MAG3D = 1,
  /* examples/slam-dense-radio/run_dense2D_withHeading.m:75-77,168
     xn=[pos(2);heading], d=1, M=m_basis, 1 normal */
  RBSLAM_MODEL_DENSE_RADIO2D = 2,
  /* examples/slam-sparse-visual/pfslam.m:81-82, measurement.m:32-84
     xn=[pos(2);heading], d=m_basis landmarks, M=2*m_basis, 3 normals, dynResNorm=[] */
  RBSLAM_MODEL_SPARSE_VISUAL2D = 3
} rbslam_model;

typedef enum {
  RBSLAM_RNG_INJECTED = 0, /* caller supplies U,Z (compat mode: MATLAB pre-draws them) */
  RBSLAM_RNG_PHILOX = 1    /* device Philox4x32-10, counter=(particle,step,block,sweep) */
} rbslam_rng;

typedef struct {
  int32_t struct_size;     /* sizeof(rbslam_config), for ABI checks */
  int32_t device;          /* CUDA device ordinal */
  int32_t model;           /* rbslam_model */
  int32_t N;               /* N_P particles (global count) */
  int32_t T;               /* N_T time steps (capacity; inputs may use fewer) */
  int32_t m_basis;         /* eigenfunctions (dense) or landmarks (sparse) */
  const int32_t *NN;       /* [m_basis x dim] column-major eigenfunction indices (dense) */
  const double *L;         /* [dim] domain half-widths (dense) */
  double cam_f, cam_fp, cam_fw; /* pinhole constants (sparse), measurement.m:32 */
  int32_t rng_mode;        /* rbslam_rng */
  uint64_t seed;           /* Philox key */
  int32_t ld;              /* leading dimension of covariance slabs, 0 = auto */
  int32_t information_form;/* also carry ivec/Imat/halfLogDetP (information-form smoother) */
  int32_t keep_history;    /* 1: keep xn history for xn_traj / traj_sample outputs */
  int32_t rank, world;     /* particle sharding: this context owns N/world particles */
  int32_t kalman_variant;  /* Kalman-update kernel.
                              0 = auto, valid for every entry point: shared-memory single pass when the
                                  slab fits one CTA, else the streaming pass over full [ld x M] slabs with
                                  sibling fusion;
                             -1 = auto for contexts that only run the FILTER entry points (what the
                                  particleFilter drop-ins pass): as 0, but large slabs are stored as packed
                                  symmetric tiles (variant 7);
                              2 = force the streaming pass over full slabs;
                              3 = legacy three-kernel path (A/B baseline);
                              7 = packed symmetric tile slabs streamed on the fp64 tensor cores: half the
                                  memory and half the HBM traffic of 2; filter only, covariance form,
                                  d <= 4, M <= 1080 (csrc/packed_kernels.cuh) */
} rbslam_config;

typedef struct {
  int32_t T;               /* N_T = rows of y */
  const double *odometry;  /* [odo_rows x n_odo]; row t-1 is used at step t */
  int32_t odo_rows;        /* leading dimension of odometry (>= T-1) */
  const double *y;         /* [T x d]; NaN = unobserved (sparse family) */
  const double *x0_nonLin; /* [n] */
  const double *x0_lin;    /* [M x x0_lin_cols] */
  int32_t x0_lin_cols;     /* 1 or N (src/particleFilter.m:60-64) */
  const double *P0_lin;    /* [M x M] */
  const double *Q;         /* [nw x nw x Q_pages] */
  int32_t Q_pages;         /* 1 or >= T-1 (src/particleFilter.m:75-77) */
  const double *R;         /* [d x d] */
  const double *dt;        /* [dt_len] */
  int32_t dt_len;          /* 1 or >= T-1 (src/particleFilter.m:80-82) */
  /* injected streams (RBSLAM_RNG_INJECTED), MATLAB layouts: */
  const double *U;         /* [N x T x K] uniforms of sample(); column t=0 unused */
  const double *Z;         /* [nz x N x T x K] normals of dynModel */
  const double *Uend;      /* [K] sweep-end uniforms (smoothers) */
  /* teacher forcing for parity tests (optional, 0-based) */
  const int32_t *forced_ancestors; /* [N x T x K] */
  const int32_t *forced_ak;        /* [K] */
} rbslam_inputs;

typedef struct {          /* any pointer may be NULL = not wanted */
  double *traj_max;        /* [n x T] */
  double *traj_mean;       /* [n x T] */
  double *xl_max;          /* [M] */
  double *xl_mean;         /* [M] */
  double *P_max;           /* [M x M] */
  double *P_mean;          /* [M x M] (reference quirk: last particle's term only) */
  double *traj_sample_iwmax; /* [n x T] */
  double *xn_traj;         /* [n x N x T] */
  /* taps for parity tests */
  double *logw_hist;       /* [N x T] unnormalised log-weights */
  double *w_hist;          /* [N x T] normalised weights */
  int32_t *ancestors;      /* [N x T] 0-based, column 0 unused */
} rbslam_filter_outputs;

typedef struct {
  double *XNK;             /* [n x T x N_K] */
  double *XLK;             /* [M x N_K] */
  double *PK;              /* [M x M x N_K] */
  double *AI;              /* [N x T x N_K] ancestor probabilities of the reference
                              particle (the reference's debugging matrix AI(:,t)) */
  int32_t *ak;             /* [N_K] sampled trajectory index per sweep, 0-based */
} rbslam_smoother_outputs;

/* called after every time step t = 0..T-1 (the filter's makePlots hook, src/particleFilter.m:215-217) and,
   by rbslam_smoother_run, once more per sweep with t == T when the sweep's outputs XNK(:,:,sweep),
   XLK(:,sweep), PK(:,:,sweep) have been written (the smoother's makePlots hook and progress line,
   src/particleSmoother.m:359-365) */
typedef void (*rbslam_step_fn)(void *user, int32_t sweep, int32_t t);

/* ---- life cycle --------------------------------------------------------- */
RBSLAM_API int rbslam_version(void);
RBSLAM_API int rbslam_device_count(void);
RBSLAM_API int rbslam_create(rbslam_ctx **out, const rbslam_config *cfg);
/* ONE filter whose particles are sharded over several GPUs of the box, driven from ONE host process --
   what a MATLAB caller (a single process, src/particleFilter.m:1-3 called from one interpreter thread)
   needs to use the whole machine.  devices[n_devices]: CUDA ordinals (they may repeat: shards then
   share a GPU); cfg->N is the global particle count (divisible by n_devices); cfg->device/rank/world
   are ignored; rng_mode PHILOX, keep_history = 1, dense model on the streaming path.  The returned
   leader context drives every shard through rbslam_filter_begin/step/end/run, rbslam_sync,
   rbslam_counters; rbslam_destroy(leader) destroys the group.  The shards are the contexts of the
   one-process-per-GPU mode below with their peer tables wired by plain device pointers
   (cudaDeviceEnablePeerAccess); the step never synchronises with the host. */
RBSLAM_API int rbslam_create_group(rbslam_ctx **out, const rbslam_config *cfg, const int32_t *devices,
                        int32_t n_devices);
/* The smoothers on several GPUs from one process.  A sweep is sequential in time and over sweeps; what
   parallelises is the ancestor-weight evaluation of the reference trajectory (src/particleSmoother.m:171-233,
   src/particleSmootherInformationForm.m:203-239: N independent dense factorisations per time step, the
   FP64-bound bulk of a sweep).  Every device holds a full replica of the particle state and runs the
   filter part of the sweep redundantly (bit-identical on identical GPUs); replica r evaluates the ancestor
   weights of its block of N / n_devices particles and stores them into every replica's array over peer
   memory -- the all-gather of weights BASELINE.json's north_star names -- then all replicas normalise and
   draw the same ancestor.  rbslam_smoother_run on the returned leader drives the group (one host thread
   per replica); rbslam_destroy(leader) destroys it.  cfg->device/rank/world are ignored. */
RBSLAM_API int rbslam_create_replicas(rbslam_ctx **out, const rbslam_config *cfg, const int32_t *devices,
                           int32_t n_devices);
RBSLAM_API void rbslam_destroy(rbslam_ctx *ctx);
RBSLAM_API const char *rbslam_last_error(const rbslam_ctx *ctx); /* ctx may be NULL: last create error */
/* derived sizes: n, d, M, nz, nw, n_odo, ld (in that order) */
RBSLAM_API int rbslam_dims(const rbslam_ctx *ctx, int32_t out7[7]);

/* ---- whole-run entry points (host buffers in, host buffers out) --------- */
/* replaces the body of src/particleFilter.m:52-233 */
RBSLAM_API int rbslam_filter_run(rbslam_ctx *ctx, const rbslam_inputs *in, rbslam_filter_outputs *out);
/* replaces src/particleSmoother.m:48-366 (form=0) and
   src/particleSmootherInformationForm.m:54-361 (form=1) */
RBSLAM_API int rbslam_smoother_run(rbslam_ctx *ctx, const rbslam_inputs *in, int32_t N_K, int32_t form,
                        rbslam_smoother_outputs *out);
/* per-step callback (the makePlots hook, src/particleFilter.m:215-217); the
   callback may call rbslam_read_particles */
RBSLAM_API int rbslam_step_callback(rbslam_ctx *ctx, rbslam_step_fn fn, void *user);

/* ---- stepwise filter (tests, benchmarks, makePlots-style taps) ---------- */
RBSLAM_API int rbslam_filter_begin(rbslam_ctx *ctx, const rbslam_inputs *in);
RBSLAM_API int rbslam_filter_step(rbslam_ctx *ctx);              /* advances one time step, asynchronous */
RBSLAM_API int rbslam_filter_end(rbslam_ctx *ctx, rbslam_filter_outputs *out);
RBSLAM_API int rbslam_sync(rbslam_ctx *ctx);                     /* wait for enqueued work, report errors */
/* logical-order copies of the particle state (any pointer may be NULL):
   xn [n x N], xl [M x N], P [M x M x N], logw [N], w [N], ai [N] */
RBSLAM_API int rbslam_read_particles(rbslam_ctx *ctx, double *xn, double *xl, double *P, double *logw,
                          double *w, int32_t *ai);
/* the trajectory outputs as the reference holds them when it calls makePlots at step t
   (src/particleFilter.m:92-97,215-217): traj_max, traj_mean [n x T] and yhattraj [d x T] (predicted
   measurement of the highest-weight particle; recorded while a step callback is registered) with NaN in
   the columns of steps not yet run, xn_traj [n x N x T] (genealogy of the current particles) with zero
   pages.  Any pointer may be NULL. */
RBSLAM_API int rbslam_read_trajectories(rbslam_ctx *ctx, double *traj_max, double *traj_mean, double *yhattraj,
                             double *xn_traj);
/* information-form extras: ivec [M x N], Imat [M x M x N], halfLogDetP [N] */
RBSLAM_API int rbslam_read_information(rbslam_ctx *ctx, double *ivec, double *Imat, double *halfLogDetP);
/* counters since context creation: kernels launched by this library, bytes copied */
RBSLAM_API int rbslam_counters(rbslam_ctx *ctx, int64_t *kernel_launches, int64_t *h2d_bytes,
                    int64_t *d2h_bytes);
/* diagnostic counters of the device status word since the last rbslam_filter_begin / smoother sweep / op call:
   chol retries with jitter (src/particleFilter.m:145-148), draws u > wc(end) that were clamped (tools/sample.m
   would raise an index error), and how many resampling steps needed the exact sequential scan after the
   parallel fast path could not prove a draw independent of the rounding order.  Synchronises the stream. */
RBSLAM_API int rbslam_status_counters(rbslam_ctx *ctx, int32_t *used_jitter, int32_t *clamped_draws,
                           int32_t *exact_scan_runs);
/* CUDA-event timing on the context's own stream (torch.cuda.Event cannot see it).
   rbslam_event_record marks slot 0..15; rbslam_event_elapsed syncs and returns ms. */
RBSLAM_API int rbslam_event_record(rbslam_ctx *ctx, int32_t slot);
RBSLAM_API int rbslam_event_elapsed(rbslam_ctx *ctx, int32_t slot_a, int32_t slot_b, float *ms);
/* per-phase device time: enable, run steps, then read accumulated ms per phase:
   [0] resample+plan [1] propagate [2] measurement Jacobian [3] Kalman update
   (gather + log-weight + downdate) [4] normalise [5] ancestor weights
   [6] information-form update [7] reserved.  Reading syncs and resets. */
RBSLAM_API int rbslam_phase_timing(rbslam_ctx *ctx, int32_t enable);
RBSLAM_API int rbslam_phase_times(rbslam_ctx *ctx, double ms8[8]);
/* raw stream handle (cudaStream_t) for hosts that enqueue their own collectives */
RBSLAM_API void *rbslam_stream(rbslam_ctx *ctx);

/* ---- kernel-level entry points (host buffers; parity tests) ------------- */
/* tools/sample.m:30-32 for a vector of uniforms: ai[j] = sum(cumsum(w) < u[j]) (0-based,
   clamped to N-1) */
RBSLAM_API int rbslam_op_resample(rbslam_ctx *ctx, int32_t N, const double *w, int32_t n_draws,
                       const double *u, int32_t *ai);
/* src/particleFilter.m:153-159: w = exp(logw - lse), iw_max = first argmax */
RBSLAM_API int rbslam_op_normalize(rbslam_ctx *ctx, int32_t N, const double *logw, double *w,
                        int32_t *iw_max);
/* dynModel for N particles: xn_out(:,i) = dynModel(xn_in(:,ai(i)), dx, dt, Q) with
   normals Z [nz x N] */
RBSLAM_API int rbslam_op_propagate(rbslam_ctx *ctx, int32_t N, const double *xn_in, const int32_t *ai,
                        const double *dx, double dt, const double *Q, const double *Z,
                        double *xn_out);
/* measModel: dy [N x d x M] (MATLAB layout, particle index fastest); for the
   sparse family xl [M x N] is required and yhat [d x N] is also returned */
RBSLAM_API int rbslam_op_meas_jacobian(rbslam_ctx *ctx, int32_t N, const double *xn, const double *xl,
                            double *dy, double *yhat);
/* JacobianPhi3D (tools/JacobianPhi3D.m:1, called from run_dense3D_magfield.m:292-294):
   Hessians of the context's m basis functions at N points.  x [3 x N]; lo = [xl yl zl],
   hi = [xu yu zu] (the reference's six scalar bounds); out J [3 x 3 x m x N].
   Dense-mag (3-D basis) contexts only, RBSLAM_EMODEL otherwise. */
RBSLAM_API int rbslam_op_jacobian_phi3d(rbslam_ctx *ctx, int32_t N, const double *x, const double *lo,
                             const double *hi, double *J);
/* fused log-weight + Kalman update (src/particleFilter.m:126-151,164-204):
   in/out xl [M x N], P [M x M x N]; H [N x d x M] MATLAB layout (NULL = evaluate the
   model's measModel at xn); out logw [N] */
RBSLAM_API int rbslam_op_kalman_update(rbslam_ctx *ctx, int32_t N, const double *xn, const double *H,
                            const double *y_t, const double *R, double jitter, double *xl,
                            double *P, double *logw);
/* dynResNorm log-density: logwDyn[i] = -0.5*||dynResNorm(xnk_t, xn(:,i), dx, dt, Q)||^2 */
RBSLAM_API int rbslam_op_dyn_logweight(rbslam_ctx *ctx, int32_t N, const double *xnk_t, const double *xn,
                            const double *dx, double dt, const double *Q, int32_t use_default,
                            double *logwDyn);

/* Measurement part of the ancestor weights of the reference trajectory (logwMeas) for N particles,
   through the kernels the smoother sweeps use (K6: batched fp64 GEMM + Cholesky; K7: batched M x M
   Cholesky), at any size:
     form 0, covariance form (src/particleSmoother.m:187-229, dense families):
        A = P [M x M x N], v = xl [M x N], S = D [ne x M] stacked future Jacobians (time-major, d rows per
        step), r = stacked future measurements [ne], R [d x d];  q2, hld unused (may be NULL)
        logwMeas_i = -sum(log(diag(cS))) - 1/2 v'v - ne/2 log(2 pi),  cS = chol(D P_i D' + kron(I, R)),
        v = cS \ (r - D xl_i); jitter as src/particleSmoother.m:70, 222-225
     form 1, information form (src/particleSmootherInformationForm.m:225-236):
        A = Imat [M x M x N], v = ivec [M x N], S = ImatAddt [M x M], r = ivecAddt [M], q2 [N] = ivec_i' P_i ivec_i,
        hld [N] = halfLogDetP;  ne, R unused;  jitter < 0 reproduces quirk Q7 (no retry: failure is an error)
        logwMeas_i = -1/2 q2_i - hld_i - sum(log(diag(cI))) + 1/2 |cI \ (ivec_i + ivecAddt)|^2,
        cI = chol(Imat_i + ImatAddt) */
RBSLAM_API int rbslam_op_ancestor_weights(rbslam_ctx *ctx, int32_t form, int32_t N, int32_t ne, const double *A,
                               const double *v, const double *S, const double *r, const double *R,
                               const double *q2, const double *hld, double jitter, double *logwMeas);

/* ---- EKF baseline of the dense magnetic-field example ---------------------- */
/* examples/slam-dense-mag/ekf_dense.m:37-102 with the closures dynModel_ekf / measModel_ekf of
   run_dense3D_magfield.m:281-316, run on the device for a dense-mag context (its basis NN, L is used;
   N and T of the context are irrelevant).  State x = [pos(3); orientation deviation(3); map(M)],
   ns = M + 6; x0 [ns], q0 [4] (linearisation point), P0 [ns x ns]; odometry [odo_rows x 7]; y [T x 3];
   Q [6 x 6 x Q_pages]; R [3 x 3]; dt [dt_len]; LL [2 x 3] = the domain bounds measModel_ekf hands to
   JacobianPhi3D.  Outputs (any may be NULL): xf_traj [ns x T], qnb_traj [4 x T], Pf_last [ns x ns]
   (= Pf_traj(:,:,T)), Pf_traj [ns x ns x T] (the reference keeps all of them: 8.5 MB each at M = 1027). */
RBSLAM_API int rbslam_ekf_run(rbslam_ctx *ctx, int32_t T, const double *odometry, int32_t odo_rows, const double *y,
                   const double *x0, const double *q0, const double *P0, const double *Q, int32_t Q_pages,
                   const double *R, const double *dt, int32_t dt_len, const double *LL, double *xf_traj,
                   double *qnb_traj, double *Pf_last, double *Pf_traj);

/* ---- localisation-only particle filter (fixed map) --------------------------- */
/* examples/mag-localization-mapping/particleFilterLocalization.m:50-132 with the closures dynModel / measModel
   of run_localization.m:241-280, for a dense-mag context (its basis NN, L).  Bootstrap particle filter of the
   7-state pose against a FIXED reduced-rank GP map: map_mean [M] (the example's `foo`), var_rows [N x 3]
   (row i = the predictive variance the reference reads for particle i, dVarft(i,:)), sigma2.
   odometry [odo_rows x 7]; y [T x 3]; x0 [7 x x0_cols], x0_cols 1 or N; Q [6 x 6 x Q_pages]; dt [dt_len].
   U [N x T], Z [6 x N x T]: injected uniforms / normals in the reference's order (column 0 unused), or both
   NULL for the device Philox stream.  Weights are w = measModel(yt, xn) ./ sum (plain sum, the reference's
   quirk of SUMMING the three component densities kept).  Outputs (any may be NULL): traj_max, traj_mean
   [7 x T]; xn_traj [7 x N x T]; ancestors [N x T] 0-based; w_hist [N x T]; n_diverged = steps with
   sum(w) <= 1e-12 (the reference prints a message and carries on). */
RBSLAM_API int rbslam_localization_run(rbslam_ctx *ctx, int32_t N, int32_t T, const double *odometry, int32_t odo_rows,
                            const double *y, const double *x0, int32_t x0_cols, const double *Q, int32_t Q_pages,
                            const double *dt, int32_t dt_len, const double *map_mean, const double *var_rows,
                            double sigma2, const double *U, const double *Z, double *traj_max, double *traj_mean,
                            double *xn_traj, int32_t *ancestors, double *w_hist, int32_t *n_diverged);

/* ---- multi-GPU sharding (one process per GPU) --------------------------- */
/* Host-only planner: given the ancestors of all N new particles and the owner
   rank of every old particle, assign new particles to ranks so that offspring
   stay on their ancestor's rank up to the per-rank capacity N/world; the rest
   migrates.  Deterministic, identical on every rank.  Outputs: new_owner [N],
   n_migrate (total). */
RBSLAM_API int rbslam_plan_migration(int32_t N, int32_t world, const int32_t *ai, const int32_t *old_owner,
                          int32_t *new_owner, int32_t *n_migrate);
/* Full per-step plan of the sharded engine (host only, deterministic, identical on every
   rank): owner and local slab slot of every new particle.  First offspring staying on the
   ancestor's rank keep the ancestor's slab; migrants are placed in DEAD slabs (no offspring
   anywhere) because they are fetched before the peer barrier. */
RBSLAM_API int rbslam_plan_shard(int32_t N, int32_t world, const int32_t *ai, const int32_t *owner_old,
                                 const int32_t *lslot_old, int32_t *owner_new, int32_t *lslot_new,
                                 int32_t *n_migrate);
/* The same plan computed ON THE DEVICE (what the sharded step uses; no host round trip),
   plus rank `rank`'s work: src_slot/glob [N/world], listA/listB [N/world] (safe group first,
   then the group deferred behind the peer barrier), fetch [N/world][4] and counts8 =
   {nA0, nB0, nA1, nB1, nFetch, nMigrantsTotal, -, -}.  Kernel-level entry point for tests. */
RBSLAM_API int rbslam_op_plan_shard(int32_t device, int32_t N, int32_t world, int32_t rank, const int32_t *ai,
                                    const int32_t *owner_old, const int32_t *lslot_old, int32_t *owner_new,
                                    int32_t *lslot_new, int32_t *src_slot, int32_t *glob, int32_t *listA,
                                    int32_t *listB, int32_t *fetch, int32_t *counts8);
/* A context created with world > 1 owns N/world slabs and shares rbslam_ipc_count() device
   buffers with its peers through CUDA IPC (64-byte handles exchanged by the host): slabs,
   pending (G,KS) ping-pong, xl ping-pong, the replicated log-weight array and the barrier
   flags.  After every peer's handles are imported, rbslam_filter_begin/step/end run the
   SHARDED filter: migrants' state is read straight from the exporter's HBM over NVLink,
   log-weights are all-gathered by peer stores, barriers are flags in peer memory. */
RBSLAM_API int rbslam_ipc_count(void);
RBSLAM_API int rbslam_ipc_export(rbslam_ctx *ctx, int32_t which, void *handle64);
RBSLAM_API int rbslam_ipc_import(rbslam_ctx *ctx, int32_t peer_rank, int32_t which, const void *handle64);
#ifdef __cplusplus
}
#endif
#endif /* RBSLAM_H */
