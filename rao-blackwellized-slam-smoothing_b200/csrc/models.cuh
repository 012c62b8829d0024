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

// ||r' / chol(dt*Q,'lower')||^2 for a fixed size: x L = r  ->  L' x' = r' (back substitution)
template <int NW>
__device__ __forceinline__ double whitened_sq(const double *r, double dt, const double *Q) {
  double L[NW * NW];
#pragma unroll
  for (int q = 0; q < NW * NW; ++q) L[q] = dt * Q[q];
#pragma unroll
  for (int j = 0; j < NW; ++j) {
    double s = L[j + j * NW];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[j + k * NW] * L[j + k * NW];
    const double ljj = sqrt(s);
    L[j + j * NW] = ljj;
#pragma unroll
    for (int i = j + 1; i < NW; ++i) {
      double v = L[i + j * NW];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i + k * NW] * L[j + k * NW];
      L[i + j * NW] = v / ljj;
    }
  }
  double x[NW];
  double ss = 0.0;
#pragma unroll
  for (int j = NW - 1; j >= 0; --j) {
    double v = r[j];
#pragma unroll
    for (int k = j + 1; k < NW; ++k) v -= L[k + j * NW] * x[k];
    x[j] = v / L[j + j * NW];
    ss += x[j] * x[j];
  }
  return ss;
}

// -0.5*||dynResNorm(xnk, xni, dx, dt, Q)||^2  (src/particleSmoother.m:175-182).
// use_default: the reference's form for dynResNorm == [] (:175-177).
__device__ __forceinline__ double dyn_logweight(const ModelConsts &mc, const double *xnk,
                                                const double *xni, const double *dx, double dt,
                                                const double *Q, bool use_default) {
  if (mc.family == FAM_DENSE_RADIO2D && !use_default) {
    // run_dense2D_withHeading.m:77: heading residual only
    const double r0 = xnk[2] - xni[2] - dx[2];
    return -0.5 * whitened_sq<1>(&r0, dt, Q);
  }
  if (mc.family == FAM_DENSE_MAG3D && !use_default) {
    // run_dense3D_magfield.m:202-203
    double r[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) r[j] = xnk[j] - xni[j] - dx[j];
    const double dqi[4] = {dx[3], -dx[4], -dx[5], -dx[6]};
    const double qii[4] = {xni[3], -xni[4], -xni[5], -xni[6]};
    const double qk[4] = {xnk[3], xnk[4], xnk[5], xnk[6]};
    double t1[4], t2[4], lq[3];
    qmul(dqi, qii, t1);
    qmul(t1, qk, t2);
    logq(t2, lq);
    r[3] = lq[0]; r[4] = lq[1]; r[5] = lq[2];
    return -0.5 * whitened_sq<6>(r, dt, Q);
  }
  // default form (src/particleSmoother.m:175-177): (x'_t - x_i - odo')' / chol(dt*Q); it
  // needs n == nw, which only the 3-state families satisfy (MATLAB would raise otherwise)
  if (mc.nw == 3) {
    double r[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) r[j] = xnk[j] - xni[j] - dx[j];
    return -0.5 * whitened_sq<3>(r, dt, Q);
  }
  return nan("");
}

}  // namespace rb
