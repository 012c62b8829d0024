// librbslam: the localisation-only particle filter of the mag-localization-mapping example
// (examples/mag-localization-mapping/particleFilterLocalization.m:50-132 with the closures
// dynModel / measModel of run_localization.m:241-280): a bootstrap particle filter against a FIXED
// reduced-rank GP map -- no per-particle covariances, so it is the pose part of the hot path only
// (multinomial resampling, pose propagation, basis evaluation, weight normalisation).
#include <vector>
#include "engine_internal.h"
#include "step_kernels.cuh"

using namespace rb;

namespace {

// run_localization.m:274-280: pos + dx(iPos)' + sqrt(dt*Q(iPos,iPos))*randn(3,1) (element-wise sqrt of the
// 3 x 3 block); quat = qLeft(qRight(q)*dx(iQuat)') * expq(sqrt(dt*Q(4:6,4:6))*randn(3,1))
__global__ void k_loc_propagate(int N, const double *__restrict__ xn_old, const int *__restrict__ ai,
                                const double *__restrict__ dx, double dt, const double *__restrict__ Q, NormalSrc nsrc,
                                double *__restrict__ xn_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double z[6];
  if (nsrc.Z) {
    for (int j = 0; j < 6; ++j) z[j] = nsrc.Z[j + (size_t)i * 6];
  } else {
    for (int p = 0; p < 3; ++p) philox_normal_pair(nsrc.seed, nsrc.sweep, nsrc.t, i, p, z[2 * p], z[2 * p + 1]);
  }
  const double *xi = xn_old + (size_t)ai[i] * 7;
  double *xo = xn_new + (size_t)i * 7;
  double nq[3];
  for (int r = 0; r < 3; ++r) {
    double ap = 0.0, aq = 0.0;
    for (int c = 0; c < 3; ++c) {
      ap += sqrt(dt * Q[r + 6 * c]) * z[c];
      aq += sqrt(dt * Q[(3 + r) + 6 * (3 + c)]) * z[3 + c];
    }
    xo[r] = xi[r] + dx[r] + ap;
    nq[r] = aq;
  }
  const double q[4] = {xi[3], xi[4], xi[5], xi[6]}, dq[4] = {dx[3], dx[4], dx[5], dx[6]};
  double t1[4], eq[4], qo[4];
  qmul(dq, q, t1);          // qRight(q) * dq = dq (x) q
  expq(nq, eq);
  qmul(t1, eq, qo);
  xo[3] = qo[0]; xo[4] = qo[1]; xo[5] = qo[2]; xo[6] = qo[3];
}

// run_localization.m:241-272: w_i = sum_a normpdf(y_a, (Rnb_i' * dPhi(pos_i) * foo)_a, sqrt(dVarft(i,a) + sigma2))
// One CTA per particle: separable sin / cos tables, then one pass over the basis.
__global__ void __launch_bounds__(128)
k_loc_weights(ModelConsts mc, int N, const double *__restrict__ xn, const double *__restrict__ foo,
              const double *__restrict__ var_rows, double sigma2, const double *__restrict__ y_t,
              double *__restrict__ w) {
  __shared__ double s_sin[3][RB_MAXTAB], s_cos[3][RB_MAXTAB];
  __shared__ double s_red[3][4];
  const int i = blockIdx.x, tid = threadIdx.x;
  const double *x = xn + (size_t)i * 7;
  for (int idx = tid; idx < 3 * RB_MAXTAB; idx += blockDim.x) {
    const int j = idx / RB_MAXTAB, nn = idx % RB_MAXTAB;
    if (nn >= 1 && nn <= mc.maxn[j]) {
      double s, c;
      sincos((RB_PI * (double)nn) * (x[j] + mc.L[j]) / (2.0 * mc.L[j]), &s, &c);
      s_sin[j][nn] = s; s_cos[j][nn] = c;
    }
  }
  __syncthreads();
  const double rL[3] = {sqrt(mc.L[0]), sqrt(mc.L[1]), sqrt(mc.L[2])};
  double g[3] = {0, 0, 0};
  for (int c = tid; c < mc.M; c += blockDim.x) {
    const double f = foo[c];
    if (c < 3) { g[c] += f; continue; }
    const int b = c - 3;
    const int nn[3] = {mc.NN[b], mc.NN[b + mc.m], mc.NN[b + 2 * mc.m]};
#pragma unroll
    for (int di = 0; di < 3; ++di) {
      double v = 1.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j == di) v = v * RB_PI * (double)nn[j] / (2.0 * mc.L[j] * rL[j]) * s_cos[j][nn[j]];
        else v = v * 1.0 / rL[j] * s_sin[j][nn[j]];
      }
      g[di] = fma(v, f, g[di]);
    }
  }
  for (int k = 0; k < 3; ++k) {
    const double s = warp_sum(g[k]);
    if ((tid & 31) == 0) s_red[k][tid >> 5] = s;
  }
  __syncthreads();
  if (tid == 0) {
    double dE[3];
    for (int k = 0; k < 3; ++k) dE[k] = (s_red[k][0] + s_red[k][1]) + (s_red[k][2] + s_red[k][3]);
    double R[3][3];
    quat2rmat(x + 3, R);
    double acc = 0.0;
    for (int a = 0; a < 3; ++a) {
      const double mu = R[0][a] * dE[0] + R[1][a] * dE[1] + R[2][a] * dE[2];   // (Rnb' * dEft')(a)
      const double sd = sqrt(var_rows[i + (size_t)N * a] + sigma2);
      const double u = (y_t[a] - mu) / sd;
      acc += exp(-0.5 * u * u) / (sd * sqrt(2.0 * RB_PI));                     // normpdf
    }
    w[i] = acc;
  }
}

// w = w ./ sum(w); [~, iw_max] = max(w); traj_max, traj_mean  (particleFilterLocalization.m:111-122)
__global__ void __launch_bounds__(1024)
k_loc_normalize(int N, double *__restrict__ w, const double *__restrict__ xn, double *__restrict__ traj_max_t,
                double *__restrict__ traj_mean_t, double *__restrict__ w_hist_t, int *__restrict__ n_diverged) {
  __shared__ double s_red[32];
  __shared__ int s_idx[32];
  __shared__ double s_sum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += w[i];
  s = warp_sum(s);
  if (lane == 0) s_red[wid] = s;
  __syncthreads();
  if (wid == 0) {
    double v = lane < nwarp ? s_red[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) { s_sum = v; if (v <= 1e-12) atomicAdd(n_diverged, 1); }   // "Weights filter close to zero"
  }
  __syncthreads();
  const double tot = s_sum;
  double best = -1.0;
  int bidx = 0x7fffffff;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double wi = w[i] / tot;
    w[i] = wi;
    if (w_hist_t) w_hist_t[i] = wi;
    if (wi > best) { best = wi; bidx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  __syncthreads();
  if (lane == 0) { s_red[wid] = best; s_idx[wid] = bidx; }
  __syncthreads();
  if (wid == 0) {
    best = lane < nwarp ? s_red[lane] : -1.0;
    bidx = lane < nwarp ? s_idx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (bidx < 0 || bidx >= N) bidx = 0;     // all-NaN weights: MATLAB's max returns index 1
    if (lane == 0) s_idx[0] = bidx;
  }
  __syncthreads();
  const int imax = s_idx[0];
  if (threadIdx.x < 7) traj_max_t[threadIdx.x] = xn[threadIdx.x + (size_t)imax * 7];
  for (int j = wid; j < 7; j += nwarp) {
    double acc = 0.0;
    for (int i = lane; i < N; i += 32) acc += xn[j + (size_t)i * 7] * w[i];
    acc = warp_sum(acc);
    if (lane == 0) traj_mean_t[j] = acc;
  }
}

__global__ void k_loc_init(int N, int cols, const double *__restrict__ x0, double *__restrict__ xn, double *__restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int j = 0; j < 7; ++j) xn[j + (size_t)i * 7] = x0[j + (cols > 1 ? (size_t)i * 7 : 0)];
  w[i] = 1.0 / N;
}
__global__ void k_loc_trace(int N, int T, const double *__restrict__ X, const int *__restrict__ A, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int b = i;
  for (int s = T - 1; s >= 0; --s) {
    for (int j = 0; j < 7; ++j) out[((size_t)s * N + i) * 7 + j] = X[((size_t)s * N + b) * 7 + j];
    if (s > 0) b = A[(size_t)s * N + b];
  }
}

}  // namespace

extern "C" int rbslam_localization_run(rbslam_ctx *ctx, int32_t N, int32_t T, const double *odometry, int32_t odo_rows,
                                       const double *y, const double *x0, int32_t x0_cols, const double *Q,
                                       int32_t Q_pages, const double *dt, int32_t dt_len, const double *map_mean,
                                       const double *var_rows, double sigma2, const double *U, const double *Z,
                                       double *traj_max, double *traj_mean, double *xn_traj, int32_t *ancestors,
                                       double *w_hist, int32_t *n_diverged) {
  if (!ctx || !y || !x0 || !Q || !dt || !map_mean || !var_rows || N < 1 || T < 1) return RBSLAM_EARG;
  if (ctx->mc.family != FAM_DENSE_MAG3D) return ctx->fail(RBSLAM_EMODEL, "the localisation filter is defined for the dense magnetic-field model only");
  if (T > 1 && (!odometry || odo_rows < T - 1)) return ctx->fail(RBSLAM_EARG, "odometry needs >= T-1 rows");
  if (x0_cols != 1 && x0_cols != N) return ctx->fail(RBSLAM_EARG, "x0_nonLin must have 1 or N columns");
  if (Q_pages != 1 && Q_pages < T - 1) return ctx->fail(RBSLAM_EARG, "Q needs 1 or >= T-1 pages");
  if (dt_len != 1 && dt_len < T - 1) return ctx->fail(RBSLAM_EARG, "dt needs 1 or >= T-1 entries");
  if ((U == nullptr) != (Z == nullptr)) return ctx->fail(RBSLAM_EARG, "pass both U and Z (injected streams) or neither (device Philox)");
  CK(cudaSetDevice(ctx->cfg.device));
  const int M = ctx->M;
  struct Buf { void *p = nullptr; ~Buf() { if (p) cudaFree(p); } };
  Buf bodo, by, bx0, bQ, bfoo, bvar, bU, bZ, bX, bA, bw, bwc, btm, btmean, bwh, bdiv, btr;
  auto alloc = [&](Buf &b, size_t bytes) -> int {
    if (cudaMalloc(&b.p, bytes ? bytes : 8) != cudaSuccess) { cudaGetLastError(); return ctx->fail(RBSLAM_ECUDA, "localization: out of device memory"); }
    return RBSLAM_OK;
  };
  int rc;
  std::vector<double> odo((size_t)std::max(T - 1, 1) * 7, 0.0), yy((size_t)T * 3);
  for (int t = 0; t + 1 < T; ++t) for (int j = 0; j < 7; ++j) odo[(size_t)t * 7 + j] = odometry[t + (size_t)j * odo_rows];
  for (int t = 0; t < T; ++t) for (int j = 0; j < 3; ++j) yy[(size_t)t * 3 + j] = y[t + (size_t)j * T];
  if ((rc = alloc(bodo, 8 * odo.size())) || (rc = alloc(by, 8 * yy.size())) || (rc = alloc(bx0, 56 * (size_t)x0_cols)) ||
      (rc = alloc(bQ, 8 * (size_t)36 * Q_pages)) || (rc = alloc(bfoo, 8 * (size_t)M)) || (rc = alloc(bvar, 24 * (size_t)N)) ||
      (rc = alloc(bX, 56 * (size_t)N * T)) || (rc = alloc(bA, 4 * (size_t)N * T)) || (rc = alloc(bw, 8 * (size_t)N)) ||
      (rc = alloc(bwc, 8 * (size_t)N)) || (rc = alloc(btm, 56 * (size_t)T)) || (rc = alloc(btmean, 56 * (size_t)T)) ||
      (rc = alloc(bdiv, 4)))
    return rc;
  if (w_hist && (rc = alloc(bwh, 8 * (size_t)N * T))) return rc;
  if ((rc = rb_h2d(ctx, bodo.p, odo.data(), 8 * odo.size())) || (rc = rb_h2d(ctx, by.p, yy.data(), 8 * yy.size())) ||
      (rc = rb_h2d(ctx, bx0.p, x0, 56 * (size_t)x0_cols)) || (rc = rb_h2d(ctx, bQ.p, Q, 8 * (size_t)36 * Q_pages)) ||
      (rc = rb_h2d(ctx, bfoo.p, map_mean, 8 * (size_t)M)) || (rc = rb_h2d(ctx, bvar.p, var_rows, 24 * (size_t)N)))
    return rc;
  if (U) {
    if ((rc = alloc(bU, 8 * (size_t)N * T)) || (rc = alloc(bZ, 48 * (size_t)N * T))) return rc;
    if ((rc = rb_h2d(ctx, bU.p, U, 8 * (size_t)N * T)) || (rc = rb_h2d(ctx, bZ.p, Z, 48 * (size_t)N * T))) return rc;
  }
  CK(cudaMemsetAsync(bdiv.p, 0, 4, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(DevStatus), ctx->stream));
  double *X = (double *)bX.p, *w = (double *)bw.p, *wc = (double *)bwc.p;
  int *A = (int *)bA.p;
  k_loc_init<<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, x0_cols, (const double *)bx0.p, X, w);
  ctx->launches += 1;
  size_t smem = std::min(sizeof(double) * (size_t)N, std::min(ctx->smem_resample_max, (size_t)(96 << 10)));
  for (int t = 0; t < T; ++t) {
    double *xt = X + (size_t)t * N * 7;
    if (t > 0) {
      RngSrc rs;
      rs.U = U ? (const double *)bU.p + (size_t)t * N : nullptr; rs.seed = ctx->cfg.seed; rs.sweep = 0; rs.t = t;
      NormalSrc ns;
      ns.Z = Z ? (const double *)bZ.p + (size_t)t * N * 6 : nullptr; ns.seed = ctx->cfg.seed; ns.sweep = 0; ns.t = t;
      int *ai = A + (size_t)t * N;
      k_resample<<<1, 1024, smem, ctx->stream>>>(N, 0, N, w, wc, rs, nullptr, ai, ctx->d_status);
      k_loc_propagate<<<(N + 127) / 128, 128, 0, ctx->stream>>>(
          N, X + (size_t)(t - 1) * N * 7, ai, (const double *)bodo.p + (size_t)(t - 1) * 7, dt_len > 1 ? dt[t - 1] : dt[0],
          (const double *)bQ.p + (Q_pages > 1 ? (size_t)(t - 1) * 36 : 0), ns, xt);
      ctx->launches += 2;
    }
    k_loc_weights<<<N, 128, 0, ctx->stream>>>(ctx->mc, N, xt, (const double *)bfoo.p, (const double *)bvar.p, sigma2,
                                              (const double *)by.p + (size_t)t * 3, w);
    k_loc_normalize<<<1, 1024, 0, ctx->stream>>>(N, w, xt, (double *)btm.p + (size_t)t * 7, (double *)btmean.p + (size_t)t * 7,
                                                 w_hist ? (double *)bwh.p + (size_t)t * N : nullptr, (int *)bdiv.p);
    ctx->launches += 2;
  }
  CK(cudaGetLastError());
  if (traj_max && (rc = rb_d2h(ctx, traj_max, btm.p, 56 * (size_t)T))) return rc;
  if (traj_mean && (rc = rb_d2h(ctx, traj_mean, btmean.p, 56 * (size_t)T))) return rc;
  if (ancestors && (rc = rb_d2h(ctx, ancestors, A, 4 * (size_t)N * T))) return rc;
  if (w_hist && (rc = rb_d2h(ctx, w_hist, bwh.p, 8 * (size_t)N * T))) return rc;
  if (n_diverged && (rc = rb_d2h(ctx, n_diverged, bdiv.p, 4))) return rc;
  if (xn_traj) {   // xn_traj(:,:,1:t-1) = xn_traj(:,ai,1:t-1) every step == one backward trace at the end
    if ((rc = alloc(btr, 56 * (size_t)N * T))) return rc;
    k_loc_trace<<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, T, X, A, (double *)btr.p);
    ctx->launches += 1;
    if ((rc = rb_d2h(ctx, xn_traj, btr.p, 56 * (size_t)N * T))) return rc;
  }
  return rb_check_status(ctx);
}
