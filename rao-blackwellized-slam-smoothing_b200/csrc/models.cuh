// Device restatement of the three model-closure families (dynModel, measModel,
// dynResNorm) the reference passes into its engines as MATLAB function handles.
#pragma once
#include "common.cuh"

namespace rb {

enum { FAM_DENSE_MAG3D = 1, FAM_DENSE_RADIO2D = 2, FAM_SPARSE_VISUAL2D = 3 };

struct ModelConsts {
  int family;
  int n;       // nNonLin
  int d;       // measurements per step
  int M;       // nLin
  int nz;      // normals per dynModel call
  int nw;      // process-noise dimension (Q is nw x nw)
  int n_odo;   // odometry columns
  int dim;     // spatial dimension of the eigenbasis
  int m;       // basis functions / landmarks
  int maxn[3]; // max eigenfunction index per dimension
  double L[3];
  const int *NN;  // device, [m x dim] column-major
  double cam_f, cam_fp, cam_fw;
};

// 3x3 lower Cholesky of dt*Q(o:o+2,o:o+2) times z  (chol(dt*Q(..),'lower')*randn(3,1))
__device__ __forceinline__ void chol3_mul(const double *Q, int ldq, int o, double dt,
                                          const double z[3], double out[3]) {
  const double a00 = dt * Q[(o + 0) + (o + 0) * ldq];
  const double a10 = dt * Q[(o + 1) + (o + 0) * ldq];
  const double a20 = dt * Q[(o + 2) + (o + 0) * ldq];
  const double a11 = dt * Q[(o + 1) + (o + 1) * ldq];
  const double a21 = dt * Q[(o + 2) + (o + 1) * ldq];
  const double a22 = dt * Q[(o + 2) + (o + 2) * ldq];
  const double l00 = sqrt(a00);
  const double l10 = a10 / l00, l20 = a20 / l00;
  const double l11 = sqrt(a11 - l10 * l10);
  const double l21 = (a21 - l20 * l10) / l11;
  const double l22 = sqrt(a22 - l20 * l20 - l21 * l21);
  out[0] = l00 * z[0];
  out[1] = l10 * z[0] + l11 * z[1];
  out[2] = l20 * z[0] + l21 * z[1] + l22 * z[2];
}

// dynModel of one particle.  xin: ancestor state, dx: odometry row, Q: nw x nw page.
__device__ __forceinline__ void dyn_model(const ModelConsts &mc, const double *xin,
                                          const double *dx, double dt, const double *Q,
                                          const double *z, double *xout) {
  if (mc.family == FAM_DENSE_MAG3D) {
    // examples/slam-dense-mag/run_dense3D_magfield.m:301-308
    double np3[3], nq3[3], eq[4], dq[4], qo[4];
    chol3_mul(Q, 6, 0, dt, z, np3);
    chol3_mul(Q, 6, 3, dt, z + 3, nq3);
    xout[0] = xin[0] + dx[0] + np3[0];
    xout[1] = xin[1] + dx[1] + np3[1];
    xout[2] = xin[2] + dx[2] + np3[2];
    expq(nq3, eq);
    qmul(dx + 3, eq, dq);    // dQuat = qLeft(dx(4:7)') * expq(..)
    qmul(xin + 3, dq, qo);   // q+ = qLeft(q) * dQuat, no renormalisation
    xout[3] = qo[0]; xout[4] = qo[1]; xout[5] = qo[2]; xout[6] = qo[3];
  } else if (mc.family == FAM_DENSE_RADIO2D) {
    // examples/slam-dense-radio/run_dense2D_withHeading.m:75-76
    double s, c;
    sincos(xin[2], &s, &c);
    xout[0] = xin[0] + (c * dx[0] + s * dx[1]);
    xout[1] = xin[1] + (-s * dx[0] + c * dx[1]);
    xout[2] = xin[2] + dx[2] + sqrt(dt * Q[0]) * z[0];
  } else {
    // examples/slam-sparse-visual/pfslam.m:81: xn + dx' + sqrt(dt*Q)*randn(3,1)
    for (int r = 0; r < 3; ++r) {
      double acc = 0.0;
      for (int c2 = 0; c2 < 3; ++c2) acc += sqrt(dt * Q[r + c2 * 3]) * z[c2];
      xout[r] = xin[r] + dx[r] + acc;
    }
  }
}

// -0.5*||dynResNorm(xnk, xni, dx, dt, Q)||^2  (src/particleSmoother.m:175-182).
// use_default: the reference's form for dynResNorm == [] (:175-177).
__device__ __forceinline__ double dyn_logweight(const ModelConsts &mc, const double *xnk,
                                                const double *xni, const double *dx, double dt,
                                                const double *Q, bool use_default) {
  double r[8];
  int nr;
  if (use_default || mc.family == FAM_SPARSE_VISUAL2D) {
    nr = mc.n;
    for (int j = 0; j < nr; ++j) r[j] = xnk[j] - xni[j] - dx[j];
  } else if (mc.family == FAM_DENSE_MAG3D) {
    // run_dense3D_magfield.m:202-203
    nr = 6;
    for (int j = 0; j < 3; ++j) r[j] = xnk[j] - xni[j] - dx[j];
    double dqi[4] = {dx[3], -dx[4], -dx[5], -dx[6]};
    double qii[4] = {xni[3], -xni[4], -xni[5], -xni[6]};
    double t1[4], t2[4];
    qmul(dqi, qii, t1);
    qmul(t1, xnk + 3, t2);
    logq(t2, r + 3);
  } else {
    // run_dense2D_withHeading.m:77: heading residual only
    const double e = (xnk[2] - xni[2] - dx[2]) / sqrt(dt * Q[0]);
    return -0.5 * e * e;
  }
  // row / chol(dt*Q,'lower'):  x L = r  ->  L' x' = r'  (back substitution)
  double Lc[36];
  const int nw = mc.nw;
  for (int c = 0; c < nw; ++c)
    for (int rr = 0; rr < nw; ++rr) Lc[rr + c * nw] = dt * Q[rr + c * nw];
  chol_small(Lc, nw, nw);
  double x[8];
  double ss = 0.0;
  for (int j = nr - 1; j >= 0; --j) {
    double v = r[j];
    for (int k = j + 1; k < nr; ++k) v -= Lc[k + j * nw] * x[k];
    x[j] = v / Lc[j + j * nw];
    ss += x[j] * x[j];
  }
  return -0.5 * ss;
}

}  // namespace rb
