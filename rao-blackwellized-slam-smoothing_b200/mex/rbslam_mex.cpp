// rbslam_mex.cpp -- thin MEX gateway from MATLAB to the C ABI of librbslam.so.
//
//   [traj_max,traj_mean,xl_max,xl_mean,P_max,P_mean,traj_sample_iwmax,xn_traj] = ...
//       rbslam_mex('filter',   model, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, dt, opts)
//   [XNK,XLK,PK] = ...
//       rbslam_mex('smoother', model, form, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, N_K, dt, opts)
//   J = rbslam_mex('jacobianphi3d', x, N_m, xl, xu, yl, yu, zl, zu, Indices)      (tools/JacobianPhi3D.m:1)
//   [xf_traj,qnb_traj,Pf_traj] = rbslam_mex('ekf', model, odometry, y, x0, q0, P0, Q, R, dt, LL)   (ekf_dense.m:1-2)
//   [traj_max,traj_mean] = rbslam_mex('localization', model, odometry, y, x0_nonLin, Q, N_P, dt, foo, dVarft, sigma2, opts)
//   rbslam_mex('release')
//
// `model` is the descriptor struct made by matlab/rbslam_model.m (family, NN, L, camera);
// `opts` carries device / devices, rng ('philox' | 'compat'), seed, in compat mode the pre-drawn
// U / Z / Uend arrays, and makePlots (the caller's function handle or []).
// The gateway only marshals: every mxArray is checked against the sizes the context derives
// from the model (rbslam_dims) -- a wrong shape raises rbslam:badArgument exactly where the
// reference would raise a MATLAB dimension error, it never becomes an out-of-bounds read --
// and mxGetDoubles pointers are then handed to the library unchanged (MATLAB arrays are
// already column-major fp64); outputs are allocated with mxCreate* and filled in place.
// makePlots is forwarded from the library's step callback with mexCallMATLAB (feval), with
// the reference's own argument lists (src/particleFilter.m:215-217, src/particleSmoother.m:359-361).
// Build (on a machine with MATLAB):
//   mex -R2018a rbslam_mex.cpp -I../../include -L../lib -lrbslam
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "mex.h"
#include "rbslam.h"

namespace {

rbslam_ctx *g_ctx = nullptr;
std::vector<int32_t> g_NN;
bool g_locked = false;

void at_exit() {
  if (g_ctx) { rbslam_destroy(g_ctx); g_ctx = nullptr; }
}
void release_all() {
  at_exit();
  if (g_locked) { mexUnlock(); g_locked = false; }   // balanced: the MEX file can be cleared again
}

void fail(rbslam_ctx *ctx, int rc) {
  static const char *ids[] = {"rbslam:ok", "rbslam:badArgument", "rbslam:cuda", "rbslam:notPositiveDefinite",
                              "rbslam:unsupportedModel"};
  std::string msg = rbslam_last_error(ctx);
  if (ctx == g_ctx) at_exit();
  mexErrMsgIdAndTxt(ids[rc >= 0 && rc <= 4 ? rc : 1], "%s", msg.c_str());
}
void bad(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  at_exit();
  mexErrMsgIdAndTxt("rbslam:badArgument", "%s", buf);
}

const mxArray *field(const mxArray *s, const char *name) {
  return (s && mxIsStruct(s)) ? mxGetField(s, 0, name) : nullptr;
}
double scalar_field(const mxArray *s, const char *name, double dflt) {
  const mxArray *f = field(s, name);
  return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}
const double *doubles_or_null(const mxArray *a) { return (a && !mxIsEmpty(a)) ? mxGetDoubles(a) : nullptr; }

// ---- shape checks ------------------------------------------------------------------------
size_t dim_of(const mxArray *a, int k) {
  return (mwSize)k < mxGetNumberOfDimensions(a) ? (size_t)mxGetDimensions(a)[k] : 1;
}
void need_double(const mxArray *a, const char *name) {
  if (!a || !mxIsDouble(a) || mxIsEmpty(a)) bad("%s must be a non-empty real double array", name);
}
void need_matrix(const mxArray *a, const char *name, size_t rows, size_t cols) {
  need_double(a, name);
  if (mxGetNumberOfDimensions(a) > 2 || mxGetM(a) != rows || mxGetN(a) != cols)
    bad("%s must be %zu x %zu (got %zu x %zu)", name, rows, cols, mxGetM(a), mxGetN(a));
}
void need_vector(const mxArray *a, const char *name, size_t len) {
  need_double(a, name);
  if (mxGetNumberOfElements(a) != len) bad("%s must have %zu elements (got %zu)", name, len, mxGetNumberOfElements(a));
}

int family_id(const mxArray *model) {
  const mxArray *f = field(model, "family");
  if (!f || !mxIsChar(f)) return -1;
  char *s = mxArrayToString(f);
  int id = -1;
  if (!strcmp(s, "denseMag3D")) id = RBSLAM_MODEL_DENSE_MAG3D;
  else if (!strcmp(s, "denseRadio2D")) id = RBSLAM_MODEL_DENSE_RADIO2D;
  else if (!strcmp(s, "sparseVisual2D")) id = RBSLAM_MODEL_SPARSE_VISUAL2D;
  mxFree(s);
  return id;
}

// (re)create the context for this problem size; the context (and CUDA init) is kept
// across calls (mexLock once + mexAtExit); 'release' destroys it and unlocks.
void make_context(const mxArray *model, const mxArray *opts, int N, int T, int info_form, int filter_only) {
  at_exit();
  if (N < 1 || T < 1) bad("N_P and the number of rows of y must be >= 1");
  rbslam_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = (int32_t)sizeof cfg;
  cfg.device = (int32_t)scalar_field(opts, "device", 0);
  cfg.model = family_id(model);
  if (cfg.model < 0)
    mexErrMsgIdAndTxt("rbslam:unsupportedModel",
                      "model handle is not one of the registered families (there is no CPU fallback)");
  cfg.N = N; cfg.T = T;
  const mxArray *NN = field(model, "NN"), *L = field(model, "L"), *cam = field(model, "camera");
  if (cfg.model == RBSLAM_MODEL_SPARSE_VISUAL2D) {
    cfg.m_basis = (int32_t)scalar_field(model, "nLandmarks", 0);
    need_vector(cam, "model.camera", 3);
    const double *c = mxGetDoubles(cam);
    cfg.cam_f = c[0]; cfg.cam_fp = c[1]; cfg.cam_fw = c[2];
  } else {
    const size_t dim = cfg.model == RBSLAM_MODEL_DENSE_MAG3D ? 3 : 2;
    need_double(NN, "model.NN");
    if (mxGetN(NN) != dim) bad("model.NN must have %zu columns", dim);
    need_vector(L, "model.L", dim);
    cfg.m_basis = (int32_t)mxGetM(NN);
    const double *nn = mxGetDoubles(NN);
    g_NN.assign(nn, nn + mxGetNumberOfElements(NN));   // MATLAB doubles -> int32, same layout
    cfg.NN = g_NN.data();
    cfg.L = mxGetDoubles(L);
  }
  const mxArray *rng = field(opts, "rng");
  char *rs = (rng && mxIsChar(rng)) ? mxArrayToString(rng) : nullptr;
  cfg.rng_mode = (rs && !strcmp(rs, "compat")) ? RBSLAM_RNG_INJECTED : RBSLAM_RNG_PHILOX;
  if (rs) mxFree(rs);
  cfg.seed = (uint64_t)scalar_field(opts, "seed", 0);
  cfg.information_form = info_form;
  cfg.keep_history = 1;
  cfg.rank = 0; cfg.world = 1;
  cfg.kalman_variant = filter_only ? -1 : 0;   // filter: packed symmetric slabs where they apply
  // opts.devices = [0 1 ...]: ONE filter sharded over several GPUs of the box, driven from this
  // (MATLAB's) process (rbslam_create_group); filter with the device RNG only
  const mxArray *devs = field(opts, "devices");
  int rc;
  if (filter_only && devs && mxGetNumberOfElements(devs) > 1) {
    if (cfg.rng_mode != RBSLAM_RNG_PHILOX) bad("opts.devices needs the device RNG (setenv RBSLAM_RNG philox)");
    std::vector<int32_t> dv;
    const double *dd = mxGetDoubles(devs);
    for (size_t k = 0; k < mxGetNumberOfElements(devs); ++k) dv.push_back((int32_t)dd[k]);
    rc = rbslam_create_group(&g_ctx, &cfg, dv.data(), (int32_t)dv.size());
  } else {
    rc = rbslam_create(&g_ctx, &cfg);
  }
  if (rc) fail(nullptr, rc);
  if (!g_locked) { mexLock(); g_locked = true; mexAtExit(at_exit); }
}

// validate every array against the sizes the context derived from the model, then fill the struct
void fill_inputs(rbslam_inputs &in, const mxArray *odo, const mxArray *y, const mxArray *x0n,
                 const mxArray *x0l, const mxArray *P0, const mxArray *Q, const mxArray *R,
                 const mxArray *dt, const mxArray *opts, int N, int K, bool compat) {
  int32_t d7[7];
  rbslam_dims(g_ctx, d7);
  const size_t n = d7[0], d = d7[1], M = d7[2], nz = d7[3], nw = d7[4], n_odo = d7[5];
  memset(&in, 0, sizeof in);
  need_double(y, "y");
  if (mxGetNumberOfDimensions(y) > 2 || mxGetN(y) != d) bad("y must be N_T x %zu (got %zu columns)", d, mxGetN(y));
  const size_t T = mxGetM(y);
  in.T = (int32_t)T;
  if (T > 1) {
    need_double(odo, "odometry");
    if (mxGetNumberOfDimensions(odo) > 2 || mxGetN(odo) != n_odo || mxGetM(odo) < T - 1)
      bad("odometry must be (>= N_T-1) x %zu (got %zu x %zu)", n_odo, mxGetM(odo), mxGetN(odo));
  }
  in.odometry = doubles_or_null(odo); in.odo_rows = odo ? (int32_t)mxGetM(odo) : 0;
  in.y = mxGetDoubles(y);
  need_vector(x0n, "x0_nonLin", n);
  in.x0_nonLin = mxGetDoubles(x0n);
  need_double(x0l, "x0_lin");
  if (mxGetNumberOfDimensions(x0l) > 2 || mxGetM(x0l) != M || (mxGetN(x0l) != 1 && mxGetN(x0l) != (size_t)N))
    bad("x0_lin must be %zu x 1 or %zu x N_P (got %zu x %zu)", M, M, mxGetM(x0l), mxGetN(x0l));
  in.x0_lin = mxGetDoubles(x0l); in.x0_lin_cols = (int32_t)mxGetN(x0l);
  need_matrix(P0, "P0_lin", M, M);
  in.P0_lin = mxGetDoubles(P0);
  need_double(Q, "Q");
  if (mxGetNumberOfDimensions(Q) > 3 || dim_of(Q, 0) != nw || dim_of(Q, 1) != nw ||
      (dim_of(Q, 2) != 1 && dim_of(Q, 2) + 1 < T))
    bad("Q must be %zu x %zu or %zu x %zu x (>= N_T-1)", nw, nw, nw, nw);
  in.Q = mxGetDoubles(Q); in.Q_pages = (int32_t)dim_of(Q, 2);
  need_matrix(R, "R", d, d);
  in.R = mxGetDoubles(R);
  need_double(dt, "dt");
  if (mxGetNumberOfElements(dt) != 1 && mxGetNumberOfElements(dt) + 1 < T) bad("dt must be a scalar or have >= N_T-1 elements");
  in.dt = mxGetDoubles(dt); in.dt_len = (int32_t)mxGetNumberOfElements(dt);
  if (compat) {   // pre-drawn streams (matlab/rbslam_streams.m): U [N x T x K], Z [nz x N x T x K], Uend [K]
    const mxArray *U = field(opts, "U"), *Z = field(opts, "Z"), *Ue = field(opts, "Uend");
    need_double(U, "opts.U"); need_double(Z, "opts.Z");
    if (mxGetNumberOfElements(U) != (size_t)N * T * K) bad("opts.U must be N_P x N_T x N_K");
    if (mxGetNumberOfElements(Z) != nz * (size_t)N * T * K) bad("opts.Z must be %zu x N_P x N_T x N_K", nz);
    if (Ue && !mxIsEmpty(Ue) && mxGetNumberOfElements(Ue) != (size_t)K) bad("opts.Uend must have N_K elements");
    in.U = mxGetDoubles(U); in.Z = mxGetDoubles(Z); in.Uend = doubles_or_null(Ue);
  }
}

bool compat_mode(const mxArray *opts) {
  const mxArray *rng = field(opts, "rng");
  char *rs = (rng && mxIsChar(rng)) ? mxArrayToString(rng) : nullptr;
  const bool c = rs && !strcmp(rs, "compat");
  if (rs) mxFree(rs);
  return c;
}

// ---- makePlots forwarding ------------------------------------------------------------------
// The reference hands ALL linear states and covariances to makePlots at every step
// (src/particleFilter.m:216): N*M^2 doubles cross PCIe per step.  Supported at example scale,
// refused above this many bytes per step.
const double kMakePlotsMaxBytes = 2.0e9;

struct FilterCb {
  mxArray *handle;
  int N, T;
  size_t n, d, M;
  std::string error;
};
void filter_step_cb(void *user, int32_t /*sweep*/, int32_t /*t*/) {
  FilterCb *cb = static_cast<FilterCb *>(user);
  if (!cb->error.empty()) return;
  const mwSize dP[3] = {cb->M, cb->M, (mwSize)cb->N}, dX[3] = {cb->n, (mwSize)cb->N, (mwSize)cb->T};
  mxArray *xn = mxCreateDoubleMatrix(cb->n, cb->N, mxREAL), *xl = mxCreateDoubleMatrix(cb->M, cb->N, mxREAL);
  mxArray *P = mxCreateNumericArray(3, dP, mxDOUBLE_CLASS, mxREAL), *w = mxCreateDoubleMatrix(cb->N, 1, mxREAL);
  mxArray *tmax = mxCreateDoubleMatrix(cb->n, cb->T, mxREAL), *tmean = mxCreateDoubleMatrix(cb->n, cb->T, mxREAL);
  mxArray *yh = mxCreateDoubleMatrix(cb->d, cb->T, mxREAL), *xt = mxCreateNumericArray(3, dX, mxDOUBLE_CLASS, mxREAL);
  int rc = rbslam_read_particles(g_ctx, mxGetDoubles(xn), mxGetDoubles(xl), mxGetDoubles(P), nullptr, mxGetDoubles(w), nullptr);
  if (!rc) rc = rbslam_read_trajectories(g_ctx, mxGetDoubles(tmax), mxGetDoubles(tmean), mxGetDoubles(yh), mxGetDoubles(xt));
  if (rc) { cb->error = rbslam_last_error(g_ctx); return; }
  // iw_max = first index of the largest weight (src/particleFilter.m:158)
  const double *wv = mxGetDoubles(w);
  int im = 0;
  for (int i = 1; i < cb->N; ++i) if (wv[i] > wv[im]) im = i;
  mxArray *xlm = mxCreateDoubleMatrix(cb->M, 1, mxREAL), *Pm = mxCreateDoubleMatrix(cb->M, cb->M, mxREAL);
  memcpy(mxGetDoubles(xlm), mxGetDoubles(xl) + (size_t)im * cb->M, sizeof(double) * cb->M);
  memcpy(mxGetDoubles(Pm), mxGetDoubles(P) + (size_t)im * cb->M * cb->M, sizeof(double) * cb->M * cb->M);
  // makePlots(xn,xl(:,iw_max),P(:,:,iw_max),traj_max,yhattraj,xn_traj,traj_mean,xl,P)
  mxArray *rhs[10] = {cb->handle, xn, xlm, Pm, tmax, yh, xt, tmean, xl, P};
  if (mexCallMATLAB(0, nullptr, 10, rhs, "feval")) cb->error = "makePlots raised an error";
  for (int k = 1; k < 10; ++k) mxDestroyArray(rhs[k]);
  mxDestroyArray(w);
}

struct SmootherCb {
  mxArray *handle;   // may be null: progress line only
  int T, NK;
  size_t n, M;
  mxArray *XNK, *XLK, *PK;
  std::string error;
};
void smoother_step_cb(void *user, int32_t sweep, int32_t t) {
  SmootherCb *cb = static_cast<SmootherCb *>(user);
  if (t != cb->T || !cb->error.empty()) return;     // t == T: sweep complete, its outputs are written
  if (cb->handle) {
    // makePlots(xnk,xlk,k,XNK,XLK,PK)   (src/particleSmoother.m:359-361)
    mxArray *xnk = mxCreateDoubleMatrix(cb->n, cb->T, mxREAL), *xlk = mxCreateDoubleMatrix(cb->M, 1, mxREAL);
    memcpy(mxGetDoubles(xnk), mxGetDoubles(cb->XNK) + (size_t)sweep * cb->n * cb->T, sizeof(double) * cb->n * cb->T);
    memcpy(mxGetDoubles(xlk), mxGetDoubles(cb->XLK) + (size_t)sweep * cb->M, sizeof(double) * cb->M);
    mxArray *k = mxCreateDoubleMatrix(1, 1, mxREAL);
    mxGetDoubles(k)[0] = sweep + 1;
    mxArray *rhs[7] = {cb->handle, xnk, xlk, k, cb->XNK, cb->XLK, cb->PK};
    if (mexCallMATLAB(0, nullptr, 7, rhs, "feval")) cb->error = "makePlots raised an error";
    mxDestroyArray(xnk); mxDestroyArray(xlk); mxDestroyArray(k);
  }
  mexPrintf("Particle smoother iteration %i/%i done.\n", sweep + 1, cb->NK);   // src/particleSmoother.m:365
}

mxArray *plots_handle(const mxArray *opts) {
  const mxArray *h = field(opts, "makePlots");
  return (h && !mxIsEmpty(h) && mxIsClass(h, "function_handle")) ? const_cast<mxArray *>(h) : nullptr;
}

}  // namespace

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("rbslam:badArgument", "first argument: command string");
  char *cmd = mxArrayToString(prhs[0]);
  const std::string c(cmd);
  mxFree(cmd);
  if (c == "filter") {
    if (nrhs < 11) mexErrMsgIdAndTxt("rbslam:badArgument", "filter: 11 or 12 arguments expected");
    const mxArray *model = prhs[1], *opts = nrhs > 11 ? prhs[11] : nullptr;
    need_double(prhs[9], "N_P"); need_double(prhs[3], "y");
    const int N = (int)mxGetScalar(prhs[9]), T = (int)mxGetM(prhs[3]);
    make_context(model, opts, N, T, 0, 1);
    rbslam_inputs in;
    fill_inputs(in, prhs[2], prhs[3], prhs[4], prhs[5], prhs[6], prhs[7], prhs[8], prhs[10], opts, N, 1,
                compat_mode(opts));
    int32_t dims7[7];
    rbslam_dims(g_ctx, dims7);
    const size_t n = dims7[0], d = dims7[1], M = dims7[2];
    rbslam_filter_outputs out;
    memset(&out, 0, sizeof out);
    mxArray *o[8] = {nullptr};
    o[0] = mxCreateDoubleMatrix(n, T, mxREAL); out.traj_max = mxGetDoubles(o[0]);
    o[1] = mxCreateDoubleMatrix(n, T, mxREAL); out.traj_mean = mxGetDoubles(o[1]);
    o[2] = mxCreateDoubleMatrix(M, 1, mxREAL); out.xl_max = mxGetDoubles(o[2]);
    o[3] = mxCreateDoubleMatrix(M, 1, mxREAL); out.xl_mean = mxGetDoubles(o[3]);
    o[4] = mxCreateDoubleMatrix(M, M, mxREAL); out.P_max = mxGetDoubles(o[4]);
    o[5] = mxCreateDoubleMatrix(M, M, mxREAL); out.P_mean = mxGetDoubles(o[5]);
    o[6] = mxCreateDoubleMatrix(n, T, mxREAL); out.traj_sample_iwmax = mxGetDoubles(o[6]);
    if (nlhs >= 8) {   // n*N*T doubles (1.1 GB at C4): allocated only when asked for
      const mwSize d3[3] = {n, (mwSize)N, (mwSize)T};
      o[7] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
      out.xn_traj = mxGetDoubles(o[7]);
    }
    FilterCb cb{plots_handle(opts), N, T, n, d, M, std::string()};
    if (cb.handle) {
      if ((double)N * M * M * 8.0 > kMakePlotsMaxBytes)
        bad("makePlots receives all N_P covariances every step (%.1f GB here): not supported at this size, pass []",
            (double)N * M * M * 8.0 / 1e9);
      rbslam_step_callback(g_ctx, filter_step_cb, &cb);
    }
    int rc = rbslam_filter_run(g_ctx, &in, &out);
    rbslam_step_callback(g_ctx, nullptr, nullptr);
    if (rc) fail(g_ctx, rc);
    if (!cb.error.empty()) { at_exit(); mexErrMsgIdAndTxt("rbslam:makePlots", "%s", cb.error.c_str()); }
    for (int k = 0; k < 8 && k < (nlhs > 0 ? nlhs : 1); ++k) plhs[k] = o[k];
  } else if (c == "smoother") {
    if (nrhs < 13) mexErrMsgIdAndTxt("rbslam:badArgument", "smoother: 13 or 14 arguments expected");
    const mxArray *model = prhs[1], *opts = nrhs > 13 ? prhs[13] : nullptr;
    need_double(prhs[2], "form"); need_double(prhs[10], "N_P"); need_double(prhs[11], "N_K"); need_double(prhs[4], "y");
    const int form = (int)mxGetScalar(prhs[2]);
    const int N = (int)mxGetScalar(prhs[10]), NK = (int)mxGetScalar(prhs[11]), T = (int)mxGetM(prhs[4]);
    if ((form != 0 && form != 1) || NK < 1) bad("smoother: form must be 0 or 1 and N_K >= 1");
    make_context(model, opts, N, T, form == 1, 0);
    rbslam_inputs in;
    fill_inputs(in, prhs[3], prhs[4], prhs[5], prhs[6], prhs[7], prhs[8], prhs[9], prhs[12], opts, N, NK,
                compat_mode(opts));
    int32_t dims7[7];
    rbslam_dims(g_ctx, dims7);
    const size_t n = dims7[0], M = dims7[2];
    rbslam_smoother_outputs out;
    memset(&out, 0, sizeof out);
    const mwSize dx[3] = {n, (mwSize)T, (mwSize)NK}, dp[3] = {M, M, (mwSize)NK};
    mxArray *XNK = mxCreateNumericArray(3, dx, mxDOUBLE_CLASS, mxREAL);
    mxArray *XLK = mxCreateDoubleMatrix(M, NK, mxREAL);
    mxArray *PK = mxCreateNumericArray(3, dp, mxDOUBLE_CLASS, mxREAL);
    out.XNK = mxGetDoubles(XNK); out.XLK = mxGetDoubles(XLK); out.PK = mxGetDoubles(PK);
    // per sweep: makePlots (if any), then the reference's progress line
    SmootherCb cb{plots_handle(opts), T, NK, n, M, XNK, XLK, PK, std::string()};
    rbslam_step_callback(g_ctx, smoother_step_cb, &cb);
    int rc = rbslam_smoother_run(g_ctx, &in, NK, form, &out);
    rbslam_step_callback(g_ctx, nullptr, nullptr);
    if (rc) fail(g_ctx, rc);
    if (!cb.error.empty()) { at_exit(); mexErrMsgIdAndTxt("rbslam:makePlots", "%s", cb.error.c_str()); }
    plhs[0] = XNK;
    if (nlhs > 1) plhs[1] = XLK;
    if (nlhs > 2) plhs[2] = PK;
  } else if (c == "jacobianphi3d") {
    if (nrhs != 10) mexErrMsgIdAndTxt("rbslam:badArgument", "jacobianphi3d: 9 arguments expected");
    need_double(prhs[1], "x"); need_double(prhs[2], "N_m"); need_double(prhs[9], "Indices");
    const int Np = (int)mxGetN(prhs[1]), Nm = (int)mxGetScalar(prhs[2]);
    const mxArray *Ind = prhs[9];
    if (mxGetM(prhs[1]) != 3 || (int)mxGetM(Ind) < Nm || mxGetN(Ind) != 3 || Nm < 1 || Np < 1)
      mexErrMsgIdAndTxt("rbslam:badArgument", "jacobianphi3d: x must be 3 x N and Indices N_m x 3");
    const double lo[3] = {mxGetScalar(prhs[3]), mxGetScalar(prhs[5]), mxGetScalar(prhs[7])};
    const double hi[3] = {mxGetScalar(prhs[4]), mxGetScalar(prhs[6]), mxGetScalar(prhs[8])};
    const double Lh[3] = {0.5 * (hi[0] - lo[0]), 0.5 * (hi[1] - lo[1]), 0.5 * (hi[2] - lo[2])};
    const size_t rows = mxGetM(Ind);
    const double *ind = mxGetDoubles(Ind);
    std::vector<int32_t> nn((size_t)3 * Nm);
    for (int dcol = 0; dcol < 3; ++dcol)
      for (int b = 0; b < Nm; ++b) nn[b + (size_t)Nm * dcol] = (int32_t)ind[b + rows * dcol];
    rbslam_config cfg;   // a private one-particle context: the filter's context is left alone
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (int32_t)sizeof cfg;
    cfg.model = RBSLAM_MODEL_DENSE_MAG3D;
    cfg.N = 1; cfg.T = 1; cfg.m_basis = Nm; cfg.NN = nn.data(); cfg.L = Lh;
    cfg.rng_mode = RBSLAM_RNG_PHILOX; cfg.world = 1;
    rbslam_ctx *tmp = nullptr;
    int rc = rbslam_create(&tmp, &cfg);
    if (rc) fail(nullptr, rc);
    const mwSize d4[4] = {3, 3, (mwSize)Nm, (mwSize)Np};
    mxArray *J = mxCreateNumericArray(4, d4, mxDOUBLE_CLASS, mxREAL);
    rc = rbslam_op_jacobian_phi3d(tmp, Np, mxGetDoubles(prhs[1]), lo, hi, mxGetDoubles(J));
    if (rc) {
      const std::string msg = rbslam_last_error(tmp);
      rbslam_destroy(tmp);
      mexErrMsgIdAndTxt(rc == 4 ? "rbslam:unsupportedModel" : "rbslam:cuda", "%s", msg.c_str());
    }
    rbslam_destroy(tmp);
    plhs[0] = J;
  } else if (c == "ekf") {
    // [xf_traj, qnb_traj, Pf_traj] = rbslam_mex('ekf', model, odometry, y, x0, q0, P0, Q, R, dt, LL): the EKF
    // baseline examples/slam-dense-mag/ekf_dense.m:1-2 (LL = the bounds measModel_ekf gives JacobianPhi3D)
    if (nrhs != 11) mexErrMsgIdAndTxt("rbslam:badArgument", "ekf: 10 arguments expected");
    const mxArray *model = prhs[1];
    need_double(prhs[3], "y");
    const int T = (int)mxGetM(prhs[3]);
    make_context(model, nullptr, 1, T, 0, 0);
    int32_t d7[7];
    rbslam_dims(g_ctx, d7);
    const size_t M = d7[2], ns = M + 6;
    need_matrix(prhs[3], "y", T, 3);
    if (T > 1) {
      need_double(prhs[2], "odometry");
      if (mxGetN(prhs[2]) != 7 || (int)mxGetM(prhs[2]) < T - 1) bad("odometry must be (>= N_T-1) x 7");
    }
    need_vector(prhs[4], "x0", ns);
    need_vector(prhs[5], "q0", 4);
    need_matrix(prhs[6], "P0", ns, ns);
    need_double(prhs[7], "Q");
    if (dim_of(prhs[7], 0) != 6 || dim_of(prhs[7], 1) != 6 || (dim_of(prhs[7], 2) != 1 && (int)dim_of(prhs[7], 2) < T - 1))
      bad("Q must be 6 x 6 or 6 x 6 x (>= N_T-1)");
    need_matrix(prhs[8], "R", 3, 3);
    need_double(prhs[9], "dt");
    if (mxGetNumberOfElements(prhs[9]) != 1 && (int)mxGetNumberOfElements(prhs[9]) < T - 1) bad("dt must be a scalar or have >= N_T-1 elements");
    need_matrix(prhs[10], "LL", 2, 3);
    mxArray *X = mxCreateDoubleMatrix(ns, T, mxREAL), *Qn = mxCreateDoubleMatrix(4, T, mxREAL);
    const mwSize dP[3] = {ns, ns, (mwSize)T};
    mxArray *P = nlhs > 2 ? mxCreateNumericArray(3, dP, mxDOUBLE_CLASS, mxREAL) : nullptr;
    int rc = rbslam_ekf_run(g_ctx, T, doubles_or_null(prhs[2]), (int32_t)(T > 1 ? mxGetM(prhs[2]) : 0), mxGetDoubles(prhs[3]),
                            mxGetDoubles(prhs[4]), mxGetDoubles(prhs[5]), mxGetDoubles(prhs[6]), mxGetDoubles(prhs[7]),
                            (int32_t)dim_of(prhs[7], 2), mxGetDoubles(prhs[8]), mxGetDoubles(prhs[9]),
                            (int32_t)mxGetNumberOfElements(prhs[9]), mxGetDoubles(prhs[10]), mxGetDoubles(X),
                            mxGetDoubles(Qn), nullptr, P ? mxGetDoubles(P) : nullptr);
    if (rc) fail(g_ctx, rc);
    plhs[0] = X;
    if (nlhs > 1) plhs[1] = Qn;
    if (nlhs > 2) plhs[2] = P;
  } else if (c == "localization") {
    // [traj_max, traj_mean] = rbslam_mex('localization', model, odometry, y, x0_nonLin, Q, N_P, dt, foo, dVarft, sigma2, opts)
    // examples/mag-localization-mapping/particleFilterLocalization.m:1-2; foo / dVarft / sigma2 are what the
    // reference's measModel closure captures (run_localization.m:259-270)
    if (nrhs < 11) mexErrMsgIdAndTxt("rbslam:badArgument", "localization: 10 or 11 arguments expected");
    const mxArray *model = prhs[1], *opts = nrhs > 11 ? prhs[11] : nullptr;
    need_double(prhs[3], "y"); need_double(prhs[6], "N_P");
    const int T = (int)mxGetM(prhs[3]), N = (int)mxGetScalar(prhs[6]);
    make_context(model, opts, 1, 1, 0, 0);
    int32_t d7[7];
    rbslam_dims(g_ctx, d7);
    const size_t M = d7[2];
    need_matrix(prhs[3], "y", T, 3);
    if (T > 1) {
      need_double(prhs[2], "odometry");
      if (mxGetN(prhs[2]) != 7 || (int)mxGetM(prhs[2]) < T - 1) bad("odometry must be (>= N_T-1) x 7");
    }
    need_double(prhs[4], "x0_nonLin");
    if (mxGetM(prhs[4]) != 7 || (mxGetN(prhs[4]) != 1 && (int)mxGetN(prhs[4]) != N)) bad("x0_nonLin must be 7 x 1 or 7 x N_P");
    need_double(prhs[5], "Q");
    if (dim_of(prhs[5], 0) != 6 || dim_of(prhs[5], 1) != 6 || (dim_of(prhs[5], 2) != 1 && (int)dim_of(prhs[5], 2) < T - 1))
      bad("Q must be 6 x 6 or 6 x 6 x (>= N_T-1)");
    need_double(prhs[7], "dt");
    if (mxGetNumberOfElements(prhs[7]) != 1 && (int)mxGetNumberOfElements(prhs[7]) < T - 1) bad("dt must be a scalar or have >= N_T-1 elements");
    need_vector(prhs[8], "foo", M);
    need_double(prhs[9], "dVarft");
    if ((int)mxGetM(prhs[9]) < N || mxGetN(prhs[9]) != 3) bad("dVarft must have >= N_P rows and 3 columns");
    need_double(prhs[10], "sigma2");
    std::vector<double> var((size_t)N * 3);   // rows 1..N_P of dVarft, as the reference indexes them
    for (int a = 0; a < 3; ++a)
      for (int i = 0; i < N; ++i) var[i + (size_t)N * a] = mxGetDoubles(prhs[9])[i + mxGetM(prhs[9]) * a];
    const bool compat = compat_mode(opts);
    const mxArray *U = field(opts, "U"), *Z = field(opts, "Z");
    if (compat) {
      need_double(U, "opts.U"); need_double(Z, "opts.Z");
      if (mxGetNumberOfElements(U) != (size_t)N * T || mxGetNumberOfElements(Z) != (size_t)6 * N * T) bad("opts.U must be N_P x N_T and opts.Z 6 x N_P x N_T");
    }
    mxArray *tm = mxCreateDoubleMatrix(7, T, mxREAL), *tmean = mxCreateDoubleMatrix(7, T, mxREAL);
    int32_t ndiv = 0;
    int rc = rbslam_localization_run(g_ctx, N, T, doubles_or_null(prhs[2]), (int32_t)(T > 1 ? mxGetM(prhs[2]) : 0),
                                     mxGetDoubles(prhs[3]), mxGetDoubles(prhs[4]), (int32_t)mxGetN(prhs[4]), mxGetDoubles(prhs[5]),
                                     (int32_t)dim_of(prhs[5], 2), mxGetDoubles(prhs[7]), (int32_t)mxGetNumberOfElements(prhs[7]),
                                     mxGetDoubles(prhs[8]), var.data(), mxGetScalar(prhs[10]), compat ? mxGetDoubles(U) : nullptr,
                                     compat ? mxGetDoubles(Z) : nullptr, mxGetDoubles(tm), mxGetDoubles(tmean), nullptr, nullptr,
                                     nullptr, &ndiv);
    if (rc) fail(g_ctx, rc);
    if (ndiv > 0) mexPrintf("Weights filter close to zero at %d time step(s) !!!\n", ndiv);   // particleFilterLocalization.m:113-115
    plhs[0] = tm;
    if (nlhs > 1) plhs[1] = tmean;
  } else if (c == "release") {
    release_all();
  } else {
    mexErrMsgIdAndTxt("rbslam:badArgument", "unknown command '%s'", c.c_str());
  }
}
