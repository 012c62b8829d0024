// librbslam: the EKF baseline of the dense magnetic-field example on the device
// (examples/slam-dense-mag/ekf_dense.m:37-102 with the closures dynModel_ekf / measModel_ekf of
// run_dense3D_magfield.m:281-316).  One Gaussian with ns = M + 6 states [pos; orientation
// deviation; map] and the linearisation point q_nb next to it: the same rank-3 Kalman update
// as the particle filter's map update (src/particleFilter.m:184-198), N = 1, plus the
// prediction of the 6 x 6 pose block, the symmetrisation (:91) and the relinearisation (:94-95).
// Nothing here is throughput-critical (one 8.5 MB covariance at M = 1027); four small kernels
// per time step, no host round trip inside the recursion.
#include <vector>
#include "engine_internal.h"
#include "step_kernels.cuh"

using namespace rb;

namespace {

struct EkfArgs {
  ModelConsts mc;
  int ns, ld, t;
  double *x;        // [ns] state mean (in/out)
  double *q;        // [4] linearisation point q_nb (in/out)
  double *P;        // [ld x ns] covariance, column-major
  const double *dx; // [7] odometry row t-1 (t > 0)
  const double *Qp; // [6 x 6] process noise of step t-1
  double dt;
  const double *y_t, *R;
  double lo[3], hi[3];   // domain bounds handed to JacobianPhi3D (run_dense3D_magfield.m:292-294)
  double *dy4;      // [ns][4] measurement Jacobian, dy(a, c) at [c][a]
  double *yhat;     // [3]
  double *PH4;      // [ns][4]  Pp dy'
  double *K4, *KS4; // [ns][4]  gain, gain * SS
  double *xf_traj, *qnb_traj;   // [ns x T], [4 x T]
  DevStatus *status;
  double jitter;
};

__device__ __forceinline__ double block_sum_128(double v, double *s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  return (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
}

// prediction of the pose block (ekf_dense.m:68-73, dynModel_ekf :310-316) and the measurement
// model at the predicted state (measModel_ekf :281-299).  One CTA of 128 threads.
__global__ void __launch_bounds__(128) k_ekf_predict_meas(EkfArgs a) {
  __shared__ double s_sin[3][RB_MAXTAB], s_cos[3][RB_MAXTAB];     // eigenfun_dx tables, domain [-L, L]
  __shared__ double s_sj[3][RB_MAXTAB], s_cj[3][RB_MAXTAB];       // JacobianPhi3D tables, domain [lo, hi]
  __shared__ double s_red[4];
  __shared__ double s_R[3][3], s_v[3], s_J[9];
  const ModelConsts &mc = a.mc;
  const int M = mc.M, m = mc.m, tid = threadIdx.x;
  if (tid == 0) {
    double q[4] = {a.q[0], a.q[1], a.q[2], a.q[3]};
    if (a.t > 0) {
      // xpred(iPos) = x(iPos) + dx(iPos)';  qpred = qLeft(q) * dx(iQuat)'
      for (int j = 0; j < 3; ++j) a.x[j] += a.dx[j];
      const double dq[4] = {a.dx[3], a.dx[4], a.dx[5], a.dx[6]};
      double qp[4];
      qmul(q, dq, qp);
      for (int j = 0; j < 4; ++j) { q[j] = qp[j]; a.q[j] = qp[j]; }
      // Pp = F Pf F' + G Qt G',  F = I, G = [blkdiag(I3, quat2rmat(qpred)); 0]: only the 6 x 6 pose block changes
      double Rq[3][3];
      quat2rmat(q, Rq);
      double G6[6][6] = {{0}};
      for (int j = 0; j < 3; ++j) G6[j][j] = 1.0;
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) G6[3 + r][3 + c] = Rq[r][c];
      double GQ[6][6];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += G6[r][k] * (a.dt * a.Qp[k + 6 * c]);
          GQ[r][c] = s;
        }
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += GQ[r][k] * G6[c][k];
          a.P[r + (size_t)c * a.ld] += s;
        }
    }
    double Rq[3][3];
    quat2rmat(q, Rq);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) s_R[r][c] = Rq[r][c];
  }
  __syncthreads();
  const double pos[3] = {a.x[0], a.x[1], a.x[2]};
  for (int idx = tid; idx < 3 * RB_MAXTAB; idx += blockDim.x) {
    const int j = idx / RB_MAXTAB, nn = idx % RB_MAXTAB;
    if (nn >= 1 && nn <= mc.maxn[j]) {
      double sv, cv;
      sincos((RB_PI * (double)nn) * (pos[j] + mc.L[j]) / (2.0 * mc.L[j]), &sv, &cv);   // tools/domain_cartesian_dx.m:88-91
      s_sin[j][nn] = sv; s_cos[j][nn] = cv;
      const double mult = 1.0 / sqrt(0.5 * (a.hi[j] - a.lo[j]));                         // tools/JacobianPhi3D.m:51-56
      sincos((RB_PI * (double)nn) * (pos[j] - a.lo[j]) / (a.hi[j] - a.lo[j]), &sv, &cv);
      s_sj[j][nn] = sv * mult; s_cj[j][nn] = cv * mult;
    }
  }
  __syncthreads();
  // gradient columns g_c = dPhi(:, c) (3-vector), v = dPhi * x(7:end), J = sum_b Hess_b * x(9 + b)
  const double rL[3] = {sqrt(mc.L[0]), sqrt(mc.L[1]), sqrt(mc.L[2])};
  double pv[3] = {0, 0, 0}, pJ[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = tid; c < M; c += blockDim.x) {
    double gc[3];
    const double xm = a.x[6 + c];
    if (c < 3) {
      gc[0] = c == 0; gc[1] = c == 1; gc[2] = c == 2;
    } else {
      const int b = c - 3;
      const int nn[3] = {mc.NN[b], mc.NN[b + m], mc.NN[b + 2 * m]};
#pragma unroll
      for (int di = 0; di < 3; ++di) {
        double v = 1.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j == di) v = v * RB_PI * (double)nn[j] / (2.0 * mc.L[j] * rL[j]) * s_cos[j][nn[j]];
          else v = v * 1.0 / rL[j] * s_sin[j][nn[j]];
        }
        gc[di] = v;
      }
      double f[3], sn[3], cs[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        f[j] = (RB_PI * (double)nn[j]) / (a.hi[j] - a.lo[j]);
        sn[j] = s_sj[j][nn[j]]; cs[j] = s_cj[j][nn[j]];
      }
      // J(r, c) column-major at pJ[r + 3 c]   (tools/JacobianPhi3D.m:58-66)
      pJ[0] += -(f[0] * f[0]) * sn[0] * sn[1] * sn[2] * xm;
      pJ[1] += f[1] * f[0] * cs[0] * cs[1] * sn[2] * xm;
      pJ[2] += f[2] * f[0] * cs[0] * sn[1] * cs[2] * xm;
      pJ[3] += f[0] * f[1] * cs[0] * cs[1] * sn[2] * xm;
      pJ[4] += -(f[1] * f[1]) * sn[0] * sn[1] * sn[2] * xm;
      pJ[5] += f[2] * f[1] * sn[0] * cs[1] * cs[2] * xm;
      pJ[6] += f[0] * f[2] * cs[0] * sn[1] * cs[2] * xm;
      pJ[7] += f[1] * f[2] * sn[0] * cs[1] * cs[2] * xm;
      pJ[8] += -(f[2] * f[2]) * sn[0] * sn[1] * sn[2] * xm;
    }
#pragma unroll
    for (int di = 0; di < 3; ++di) pv[di] = fma(gc[di], xm, pv[di]);
    // dy(:, 7:end) = Rnb' * dPhi
    double o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int aa = 0; aa < 3; ++aa) o[aa] = s_R[0][aa] * gc[0] + s_R[1][aa] * gc[1] + s_R[2][aa] * gc[2];
    *reinterpret_cast<double4 *>(a.dy4 + (size_t)(6 + c) * 4) = make_double4(o[0], o[1], o[2], o[3]);
  }
  for (int k = 0; k < 3; ++k) { const double s = block_sum_128(pv[k], s_red); if (tid == 0) s_v[k] = s; }
  for (int k = 0; k < 9; ++k) { const double s = block_sum_128(pJ[k], s_red); if (tid == 0) s_J[k] = s; }
  __syncthreads();
  if (tid < 6) {
    // dy(:,1:3) = Rnb' * J;  dy(:,4:6) = Rnb' * mcross(dPhi * x(7:end))
    const double vx[3][3] = {{0, -s_v[2], s_v[1]}, {s_v[2], 0, -s_v[0]}, {-s_v[1], s_v[0], 0}};   // tools/mcross.m:33-36
    double o[4] = {0, 0, 0, 0};
    for (int aa = 0; aa < 3; ++aa) {
      double s = 0.0;
      for (int r = 0; r < 3; ++r) s += s_R[r][aa] * (tid < 3 ? s_J[r + 3 * tid] : vx[r][tid - 3]);
      o[aa] = s;
    }
    *reinterpret_cast<double4 *>(a.dy4 + (size_t)tid * 4) = make_double4(o[0], o[1], o[2], o[3]);
  }
  if (tid < 3) a.yhat[tid] = s_R[0][tid] * s_v[0] + s_R[1][tid] * s_v[1] + s_R[2][tid] * s_v[2];   // yhat = Rnb' * dPhi * x(7:end)
  for (int r = a.ns + tid; r < a.ld; r += blockDim.x)
    *reinterpret_cast<double4 *>(a.dy4 + (size_t)r * 4) = make_double4(0, 0, 0, 0);
}

// PH = Pp * dy'  (one thread per row; P is column-major, so a column step is coalesced)
__global__ void __launch_bounds__(128) k_ekf_ph(EkfArgs a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.ld) return;
  double acc[3] = {0, 0, 0};
  if (r < a.ns) {
    for (int c = 0; c < a.ns; ++c) {
      const double p = a.P[r + (size_t)c * a.ld];
      const double4 h = *reinterpret_cast<const double4 *>(a.dy4 + (size_t)c * 4);
      acc[0] = fma(p, h.x, acc[0]); acc[1] = fma(p, h.y, acc[1]); acc[2] = fma(p, h.z, acc[2]);
    }
  }
  *reinterpret_cast<double4 *>(a.PH4 + (size_t)r * 4) = make_double4(acc[0], acc[1], acc[2], 0.0);
}

// SS, Cholesky (+ jitter), gain, state update, relinearisation (ekf_dense.m:78-100).  One CTA.
__global__ void __launch_bounds__(128) k_ekf_gain(EkfArgs a) {
  __shared__ double s_red[4];
  __shared__ double s_L[9], s_SS[9], s_e[3];
  const int tid = threadIdx.x, ns = a.ns;
  double part[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = tid; c < ns; c += blockDim.x) {
    const double4 h = *reinterpret_cast<const double4 *>(a.dy4 + (size_t)c * 4);
    const double4 p = *reinterpret_cast<const double4 *>(a.PH4 + (size_t)c * 4);
    const double hv[3] = {h.x, h.y, h.z}, ph[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int aa = 0; aa < 3; ++aa)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) part[aa + 3 * bb] = fma(hv[aa], ph[bb], part[aa + 3 * bb]);
  }
  double S[9];
  for (int k = 0; k < 9; ++k) S[k] = block_sum_128(part[k], s_red);
  if (tid == 0) {
    double Lc[9];
    for (int k = 0; k < 9; ++k) { S[k] += a.R[k]; Lc[k] = S[k]; s_SS[k] = S[k]; }
    int flag = chol_small(Lc, 3, 3);
    if (flag) {   // ekf_dense.m:83-85
      for (int k = 0; k < 9; ++k) Lc[k] = S[k] + ((k % 3) == (k / 3) ? a.jitter : 0.0);
      atomicAdd(&a.status->used_jitter, 1);
      flag = chol_small(Lc, 3, 3);
      if (flag && atomicCAS(&a.status->not_pd, 0, 1) == 0) { a.status->not_pd_step = a.t; a.status->not_pd_particle = 0; }
    }
    for (int k = 0; k < 9; ++k) s_L[k] = Lc[k];
    for (int k = 0; k < 3; ++k) s_e[k] = a.y_t[k] - a.yhat[k];
  }
  __syncthreads();
  for (int r = tid; r < a.ld; r += blockDim.x) {
    double g[4] = {0, 0, 0, 0}, ks[4] = {0, 0, 0, 0};
    if (r < ns) {
      const double4 p = *reinterpret_cast<const double4 *>(a.PH4 + (size_t)r * 4);
      const double ph[3] = {p.x, p.y, p.z};
      // K(r, :) = PH(r, :) / cS' / cS
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        double s = ph[q];
#pragma unroll
        for (int k = 0; k < 3; ++k) if (k < q) s -= s_L[q + 3 * k] * g[k];
        g[q] = s / s_L[q + 3 * q];
      }
#pragma unroll
      for (int q = 2; q >= 0; --q) {
        double s = g[q];
#pragma unroll
        for (int k = 0; k < 3; ++k) if (k > q) s -= s_L[k + 3 * q] * g[k];
        g[q] = s / s_L[q + 3 * q];
      }
      double ge = 0.0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        ge = fma(g[q], s_e[q], ge);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s = fma(g[k], s_SS[k + 3 * q], s);
        ks[q] = s;
      }
      a.x[r] += ge;   // xf = xp + K*e
    }
    *reinterpret_cast<double4 *>(a.K4 + (size_t)r * 4) = make_double4(g[0], g[1], g[2], 0.0);
    *reinterpret_cast<double4 *>(a.KS4 + (size_t)r * 4) = make_double4(ks[0], ks[1], ks[2], 0.0);
  }
  __syncthreads();
  if (tid == 0) {
    // q_nb = qLeft(expq(xf(iOri)/2)) * q_nb;  xf(iOri) = 0   (ekf_dense.m:94-95)
    const double phi[3] = {a.x[3] / 2, a.x[4] / 2, a.x[5] / 2};
    double eq[4], qn[4];
    const double q[4] = {a.q[0], a.q[1], a.q[2], a.q[3]};
    expq(phi, eq);
    qmul(eq, q, qn);
    for (int j = 0; j < 4; ++j) { a.q[j] = qn[j]; a.qnb_traj[j + 4 * (size_t)a.t] = qn[j]; }
    a.x[3] = a.x[4] = a.x[5] = 0.0;
  }
  __syncthreads();
  for (int r = tid; r < ns; r += blockDim.x) a.xf_traj[r + (size_t)ns * a.t] = a.x[r];
}

// Pf = Pp - K*SS*K';  Pf = 0.5*(Pf + Pf')   (ekf_dense.m:90-91).  32 x 32 tiles of the lower triangle.
__global__ void __launch_bounds__(256) k_ekf_downdate(EkfArgs a) {
  const int tr = blockIdx.y, tc = blockIdx.x;
  if (tc > tr) return;
  const int ns = a.ns;
  for (int idx = threadIdx.x; idx < 32 * 32; idx += blockDim.x) {
    const int r = 32 * tr + (idx & 31), c = 32 * tc + (idx >> 5);
    if (r >= ns || c >= ns || c > r) continue;
    const double4 kr = *reinterpret_cast<const double4 *>(a.K4 + (size_t)r * 4);
    const double4 kc = *reinterpret_cast<const double4 *>(a.K4 + (size_t)c * 4);
    const double4 sr = *reinterpret_cast<const double4 *>(a.KS4 + (size_t)r * 4);
    const double4 sc = *reinterpret_cast<const double4 *>(a.KS4 + (size_t)c * 4);
    // (K SS K')(r, c) = sum_a KS(r, a) K(c, a)
    const double lo = a.P[r + (size_t)c * a.ld] - (sr.x * kc.x + sr.y * kc.y + sr.z * kc.z);
    const double up = a.P[c + (size_t)r * a.ld] - (sc.x * kr.x + sc.y * kr.y + sc.z * kr.z);
    const double v = 0.5 * (lo + up);
    a.P[r + (size_t)c * a.ld] = v;
    a.P[c + (size_t)r * a.ld] = v;
  }
}

}  // namespace

extern "C" int rbslam_ekf_run(rbslam_ctx *ctx, int32_t T, const double *odometry, int32_t odo_rows, const double *y,
                              const double *x0, const double *q0, const double *P0, const double *Q, int32_t Q_pages,
                              const double *R, const double *dt, int32_t dt_len, const double *LL, double *xf_traj,
                              double *qnb_traj, double *Pf_last, double *Pf_traj) {
  if (!ctx || !y || !x0 || !q0 || !P0 || !Q || !R || !dt || !LL || T < 1) return RBSLAM_EARG;
  if (ctx->mc.family != FAM_DENSE_MAG3D) return ctx->fail(RBSLAM_EMODEL, "the EKF baseline is defined for the dense magnetic-field model only");
  if (T > 1 && (!odometry || odo_rows < T - 1)) return ctx->fail(RBSLAM_EARG, "odometry needs >= T-1 rows");
  if (Q_pages != 1 && Q_pages < T - 1) return ctx->fail(RBSLAM_EARG, "Q needs 1 or >= T-1 pages");
  if (dt_len != 1 && dt_len < T - 1) return ctx->fail(RBSLAM_EARG, "dt needs 1 or >= T-1 entries");
  for (int k = 0; k < 3; ++k)
    if (!(LL[1 + 2 * k] > LL[2 * k])) return ctx->fail(RBSLAM_EARG, "LL: upper bound must exceed lower bound");
  CK(cudaSetDevice(ctx->cfg.device));
  const int M = ctx->M, ns = M + 6, ld = (ns + 1) & ~1;
  struct Buf { void *p = nullptr; ~Buf() { if (p) cudaFree(p); } };
  Buf bx, bq, bP, bodo, by, bQ, bR, bdy, byh, bPH, bK, bKS, bxf, bqn, bPt;
  auto alloc = [&](Buf &b, size_t bytes) -> int {
    if (cudaMalloc(&b.p, bytes ? bytes : 8) != cudaSuccess) { cudaGetLastError(); return ctx->fail(RBSLAM_ECUDA, "ekf: out of device memory"); }
    return RBSLAM_OK;
  };
  int rc;
  std::vector<double> odo((size_t)std::max(T - 1, 1) * 7, 0.0), yy((size_t)T * 3), Pp((size_t)ld * ns, 0.0);
  for (int t = 0; t + 1 < T; ++t) for (int j = 0; j < 7; ++j) odo[(size_t)t * 7 + j] = odometry[t + (size_t)j * odo_rows];
  for (int t = 0; t < T; ++t) for (int j = 0; j < 3; ++j) yy[(size_t)t * 3 + j] = y[t + (size_t)j * T];
  for (int c = 0; c < ns; ++c) for (int r = 0; r < ns; ++r) Pp[r + (size_t)c * ld] = P0[r + (size_t)c * ns];
  if ((rc = alloc(bx, 8 * (size_t)ld)) || (rc = alloc(bq, 32)) || (rc = alloc(bP, 8 * Pp.size())) ||
      (rc = alloc(bodo, 8 * odo.size())) || (rc = alloc(by, 8 * yy.size())) || (rc = alloc(bQ, 8 * (size_t)36 * Q_pages)) ||
      (rc = alloc(bR, 72)) || (rc = alloc(bdy, 32 * (size_t)ld)) || (rc = alloc(byh, 32)) || (rc = alloc(bPH, 32 * (size_t)ld)) ||
      (rc = alloc(bK, 32 * (size_t)ld)) || (rc = alloc(bKS, 32 * (size_t)ld)) || (rc = alloc(bxf, 8 * (size_t)ns * T)) ||
      (rc = alloc(bqn, 32 * (size_t)T)))
    return rc;
  if (Pf_traj && (rc = alloc(bPt, 8 * (size_t)ns * ns))) return rc;
  if ((rc = rb_h2d(ctx, bx.p, x0, 8 * (size_t)ns)) || (rc = rb_h2d(ctx, bq.p, q0, 32)) ||
      (rc = rb_h2d(ctx, bP.p, Pp.data(), 8 * Pp.size())) || (rc = rb_h2d(ctx, bodo.p, odo.data(), 8 * odo.size())) ||
      (rc = rb_h2d(ctx, by.p, yy.data(), 8 * yy.size())) || (rc = rb_h2d(ctx, bQ.p, Q, 8 * (size_t)36 * Q_pages)) ||
      (rc = rb_h2d(ctx, bR.p, R, 72)))
    return rc;
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(DevStatus), ctx->stream));
  EkfArgs a;
  a.mc = ctx->mc; a.ns = ns; a.ld = ld;
  a.x = (double *)bx.p; a.q = (double *)bq.p; a.P = (double *)bP.p; a.R = (const double *)bR.p;
  for (int k = 0; k < 3; ++k) { a.lo[k] = LL[2 * k]; a.hi[k] = LL[1 + 2 * k]; }
  a.dy4 = (double *)bdy.p; a.yhat = (double *)byh.p; a.PH4 = (double *)bPH.p; a.K4 = (double *)bK.p; a.KS4 = (double *)bKS.p;
  a.xf_traj = (double *)bxf.p; a.qnb_traj = (double *)bqn.p; a.status = ctx->d_status; a.jitter = 1e-3;   // ekf_dense.m:56
  const int nt = (ns + 31) / 32;
  for (int t = 0; t < T; ++t) {
    a.t = t;
    a.dx = (const double *)bodo.p + (size_t)std::max(t - 1, 0) * 7;
    a.Qp = (const double *)bQ.p + (Q_pages > 1 ? (size_t)std::max(t - 1, 0) * 36 : 0);
    a.dt = t > 0 ? (dt_len > 1 ? dt[t - 1] : dt[0]) : 0.0;
    a.y_t = (const double *)by.p + (size_t)t * 3;
    k_ekf_predict_meas<<<1, 128, 0, ctx->stream>>>(a);
    k_ekf_ph<<<(ld + 127) / 128, 128, 0, ctx->stream>>>(a);
    k_ekf_gain<<<1, 128, 0, ctx->stream>>>(a);
    k_ekf_downdate<<<dim3(nt, nt), 256, 0, ctx->stream>>>(a);
    ctx->launches += 4;
    if (Pf_traj) {   // the reference keeps every filtered covariance (ekf_dense.m:99): optional here
      CK(cudaMemcpy2DAsync(bPt.p, 8 * (size_t)ns, a.P, 8 * (size_t)ld, 8 * (size_t)ns, ns, cudaMemcpyDeviceToDevice, ctx->stream));
      if ((rc = rb_d2h(ctx, Pf_traj + (size_t)t * ns * ns, bPt.p, 8 * (size_t)ns * ns))) return rc;
    }
  }
  CK(cudaGetLastError());
  if (xf_traj && (rc = rb_d2h(ctx, xf_traj, bxf.p, 8 * (size_t)ns * T))) return rc;
  if (qnb_traj && (rc = rb_d2h(ctx, qnb_traj, bqn.p, 32 * (size_t)T))) return rc;
  if (Pf_last) {
    CK(cudaMemcpy2DAsync(Pp.data(), 8 * (size_t)ns, a.P, 8 * (size_t)ld, 8 * (size_t)ns, ns, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->d2h += 8 * (int64_t)ns * ns;
    memcpy(Pf_last, Pp.data(), 8 * (size_t)ns * ns);
  }
  return rb_check_status(ctx);
}
