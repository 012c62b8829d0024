// rbslam_mex.cpp -- thin MEX gateway from MATLAB to the C ABI of librbslam.so.
//
//   [traj_max,traj_mean,xl_max,xl_mean,P_max,P_mean,traj_sample_iwmax,xn_traj] = ...
//       rbslam_mex('filter',   model, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, dt, opts)
//   [XNK,XLK,PK] = ...
//       rbslam_mex('smoother', model, form, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, N_K, dt, opts)
//   J = rbslam_mex('jacobianphi3d', x, N_m, xl, xu, yl, yu, zl, zu, Indices)      (tools/JacobianPhi3D.m:1)
//
// `model` is the descriptor struct made by matlab/rbslam_model.m (family, NN, L, camera);
// `opts` carries device, rng ('philox' | 'compat'), seed and, in compat mode, the
// pre-drawn U / Z / Uend arrays.  The gateway only marshals: mxGetDoubles pointers are
// handed to the library unchanged (MATLAB arrays are already column-major fp64), outputs
// are allocated with mxCreate* and filled in place.  Build (on a machine with MATLAB):
//   mex -R2018a rbslam_mex.cpp -I../../include -L../lib -lrbslam
#include <cstring>
#include <string>
#include <vector>
#include "mex.h"
#include "rbslam.h"

namespace {

rbslam_ctx *g_ctx = nullptr;
std::vector<int32_t> g_NN;

void at_exit() {
  if (g_ctx) { rbslam_destroy(g_ctx); g_ctx = nullptr; }
}

void fail(rbslam_ctx *ctx, int rc) {
  static const char *ids[] = {"rbslam:ok", "rbslam:badArgument", "rbslam:cuda", "rbslam:notPositiveDefinite",
                              "rbslam:unsupportedModel"};
  std::string msg = rbslam_last_error(ctx);
  if (ctx == g_ctx) at_exit();
  mexErrMsgIdAndTxt(ids[rc >= 0 && rc <= 4 ? rc : 1], "%s", msg.c_str());
}

const mxArray *field(const mxArray *s, const char *name) {
  return (s && mxIsStruct(s)) ? mxGetField(s, 0, name) : nullptr;
}
double scalar_field(const mxArray *s, const char *name, double dflt) {
  const mxArray *f = field(s, name);
  return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}
const double *doubles_or_null(const mxArray *a) { return (a && !mxIsEmpty(a)) ? mxGetDoubles(a) : nullptr; }

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
// across calls with mexLock/mexAtExit.
void make_context(const mxArray *model, const mxArray *opts, int N, int T, int info_form) {
  at_exit();
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
    const double *c = doubles_or_null(cam);
    if (c) { cfg.cam_f = c[0]; cfg.cam_fp = c[1]; cfg.cam_fw = c[2]; }
  } else {
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
  int rc = rbslam_create(&g_ctx, &cfg);
  if (rc) fail(nullptr, rc);
  mexLock();
  mexAtExit(at_exit);
}

void fill_inputs(rbslam_inputs &in, const mxArray *odo, const mxArray *y, const mxArray *x0n,
                 const mxArray *x0l, const mxArray *P0, const mxArray *Q, const mxArray *R,
                 const mxArray *dt, const mxArray *opts) {
  memset(&in, 0, sizeof in);
  in.T = (int32_t)mxGetM(y);
  in.odometry = doubles_or_null(odo); in.odo_rows = (int32_t)mxGetM(odo);
  in.y = mxGetDoubles(y);
  in.x0_nonLin = mxGetDoubles(x0n);
  in.x0_lin = mxGetDoubles(x0l); in.x0_lin_cols = (int32_t)mxGetN(x0l);
  in.P0_lin = mxGetDoubles(P0);
  in.Q = mxGetDoubles(Q);
  in.Q_pages = mxGetNumberOfDimensions(Q) > 2 ? (int32_t)mxGetDimensions(Q)[2] : 1;
  in.R = mxGetDoubles(R);
  in.dt = mxGetDoubles(dt); in.dt_len = (int32_t)mxGetNumberOfElements(dt);
  in.U = doubles_or_null(field(opts, "U"));
  in.Z = doubles_or_null(field(opts, "Z"));
  in.Uend = doubles_or_null(field(opts, "Uend"));
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
    const int N = (int)mxGetScalar(prhs[9]), T = (int)mxGetM(prhs[3]);
    make_context(model, opts, N, T, 0);
    rbslam_inputs in;
    fill_inputs(in, prhs[2], prhs[3], prhs[4], prhs[5], prhs[6], prhs[7], prhs[8], prhs[10], opts);
    int32_t dims7[7];
    rbslam_dims(g_ctx, dims7);
    const size_t n = dims7[0], M = dims7[2];
    rbslam_filter_outputs out;
    memset(&out, 0, sizeof out);
    mxArray *o[8];
    o[0] = mxCreateDoubleMatrix(n, T, mxREAL); out.traj_max = mxGetDoubles(o[0]);
    o[1] = mxCreateDoubleMatrix(n, T, mxREAL); out.traj_mean = mxGetDoubles(o[1]);
    o[2] = mxCreateDoubleMatrix(M, 1, mxREAL); out.xl_max = mxGetDoubles(o[2]);
    o[3] = mxCreateDoubleMatrix(M, 1, mxREAL); out.xl_mean = mxGetDoubles(o[3]);
    o[4] = mxCreateDoubleMatrix(M, M, mxREAL); out.P_max = mxGetDoubles(o[4]);
    o[5] = mxCreateDoubleMatrix(M, M, mxREAL); out.P_mean = mxGetDoubles(o[5]);
    o[6] = mxCreateDoubleMatrix(n, T, mxREAL); out.traj_sample_iwmax = mxGetDoubles(o[6]);
    const mwSize d3[3] = {n, (mwSize)N, (mwSize)T};
    o[7] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
    if (nlhs >= 8) out.xn_traj = mxGetDoubles(o[7]);   // 1.1 GB at C4: only when asked for
    int rc = rbslam_filter_run(g_ctx, &in, &out);
    if (rc) fail(g_ctx, rc);
    for (int k = 0; k < 8 && k < (nlhs > 0 ? nlhs : 1); ++k) plhs[k] = o[k];
  } else if (c == "smoother") {
    if (nrhs < 13) mexErrMsgIdAndTxt("rbslam:badArgument", "smoother: 13 or 14 arguments expected");
    const mxArray *model = prhs[1], *opts = nrhs > 13 ? prhs[13] : nullptr;
    const int form = (int)mxGetScalar(prhs[2]);
    const int N = (int)mxGetScalar(prhs[10]), NK = (int)mxGetScalar(prhs[11]), T = (int)mxGetM(prhs[4]);
    make_context(model, opts, N, T, form == 1);
    rbslam_inputs in;
    fill_inputs(in, prhs[3], prhs[4], prhs[5], prhs[6], prhs[7], prhs[8], prhs[9], prhs[12], opts);
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
    int rc = rbslam_smoother_run(g_ctx, &in, NK, form, &out);
    if (rc) fail(g_ctx, rc);
    for (int k = 1; k <= NK; ++k)   // same progress line as src/particleSmoother.m:365
      mexPrintf("Particle smoother iteration %i/%i done.\n", k, NK);
    plhs[0] = XNK;
    if (nlhs > 1) plhs[1] = XLK;
    if (nlhs > 2) plhs[2] = PK;
  } else if (c == "jacobianphi3d") {
    if (nrhs != 10) mexErrMsgIdAndTxt("rbslam:badArgument", "jacobianphi3d: 9 arguments expected");
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
  } else if (c == "release") {
    at_exit();
  } else {
    mexErrMsgIdAndTxt("rbslam:badArgument", "unknown command '%s'", c.c_str());
  }
}
