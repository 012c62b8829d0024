// Kernel-level entry points with host buffers (parity tests drive each kernel on
// its own through these) and the host-only multi-GPU migration planner.
#include <vector>
#include <algorithm>
#include <cstring>
#include "engine_internal.h"
#include "step_kernels.cuh"
#include "packed_kernels.cuh"

using namespace rb;

namespace {
struct TmpBuf {   // RAII device scratch
  void *p = nullptr;
  ~TmpBuf() { if (p) cudaFree(p); }
};
}  // namespace

#define TMP_ALLOC(buf, type, count)                                               \
  TmpBuf buf##_holder;                                                            \
  type *buf = nullptr;                                                            \
  do {                                                                            \
    int rc__ = dev_alloc(ctx, &buf, (size_t)(count));                             \
    if (rc__) return rc__;                                                        \
    buf##_holder.p = buf;                                                         \
  } while (0)

extern "C" int rbslam_op_resample(rbslam_ctx *ctx, int32_t N, const double *w, int32_t n_draws,
                                  const double *u, int32_t *ai) {
  if (!ctx || !w || !u || !ai || N < 1 || n_draws < 0) return RBSLAM_EARG;
  CK(cudaSetDevice(ctx->cfg.device));
  TMP_ALLOC(d_w, double, N);
  TMP_ALLOC(d_wc, double, N);
  TMP_ALLOC(d_u, double, n_draws);
  TMP_ALLOC(d_ai, int, n_draws);
  int rc;
  if ((rc = rb_h2d(ctx, d_w, w, sizeof(double) * N))) return rc;
  if ((rc = rb_h2d(ctx, d_u, u, sizeof(double) * n_draws))) return rc;
  RngSrc rs;
  rs.U = d_u; rs.seed = 0; rs.sweep = 0; rs.t = 0;
  size_t smem = sizeof(double) * (size_t)N;
  smem = std::min(smem, std::min(ctx->smem_resample_max, (size_t)(96 << 10)));
  if (N >= 4096 && n_draws > 1024) {   // the launch sequence of rb_resample_phase for large populations
    const int fast = rb_fast_scan() ? 1 : 0;
    if (fast) {
      k_scan_approx<<<1, 1024, 0, ctx->stream>>>(N, d_w, d_wc, ctx->d_status);
      k_search_checked<<<(n_draws + 255) / 256, 256, 0, ctx->stream>>>(N, 0, n_draws, d_wc, rs, nullptr, d_ai, ctx->d_status);
    }
    k_resample<<<1, 1024, smem, ctx->stream>>>(N, 0, -1, d_w, d_wc, rs, nullptr, d_ai, ctx->d_status, fast);
    k_resample_search<<<(n_draws + 255) / 256, 256, 0, ctx->stream>>>(N, 0, n_draws, d_wc, rs, nullptr, d_ai, ctx->d_status, fast);
    ctx->launches += 2 + 2 * fast;
  } else {
    k_resample<<<1, 1024, smem, ctx->stream>>>(N, 0, n_draws, d_w, d_wc, rs, nullptr, d_ai, ctx->d_status);
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  return rb_d2h(ctx, ai, d_ai, sizeof(int) * n_draws);
}

extern "C" int rbslam_op_normalize(rbslam_ctx *ctx, int32_t N, const double *logw, double *w,
                                   int32_t *iw_max) {
  if (!ctx || !logw || !w || N < 1) return RBSLAM_EARG;
  CK(cudaSetDevice(ctx->cfg.device));
  TMP_ALLOC(d_lw, double, N);
  TMP_ALLOC(d_w, double, N);
  TMP_ALLOC(d_i, int, 1);
  int rc;
  if ((rc = rb_h2d(ctx, d_lw, logw, sizeof(double) * N))) return rc;
  if ((rc = rb_normalize(ctx, N, 0, d_lw, d_w, nullptr, nullptr, nullptr, d_i, nullptr, nullptr))) return rc;
  if ((rc = rb_d2h(ctx, w, d_w, sizeof(double) * N))) return rc;
  if (iw_max && (rc = rb_d2h(ctx, iw_max, d_i, sizeof(int)))) return rc;
  return RBSLAM_OK;
}

extern "C" int rbslam_op_propagate(rbslam_ctx *ctx, int32_t N, const double *xn_in, const int32_t *ai,
                                   const double *dx, double dt, const double *Q, const double *Z,
                                   double *xn_out) {
  if (!ctx || !xn_in || !ai || !dx || !Q || !xn_out || N < 1) return RBSLAM_EARG;
  CK(cudaSetDevice(ctx->cfg.device));
  const int n = ctx->n;
  TMP_ALLOC(d_in, double, (size_t)n * N);
  TMP_ALLOC(d_out, double, (size_t)n * N);
  TMP_ALLOC(d_ai, int, N);
  TMP_ALLOC(d_dx, double, ctx->n_odo);
  TMP_ALLOC(d_Q, double, ctx->nw * ctx->nw);
  TMP_ALLOC(d_Z, double, (size_t)ctx->nz * N);
  int rc;
  if ((rc = rb_h2d(ctx, d_in, xn_in, sizeof(double) * n * N))) return rc;
  if ((rc = rb_h2d(ctx, d_ai, ai, sizeof(int) * N))) return rc;
  if ((rc = rb_h2d(ctx, d_dx, dx, sizeof(double) * ctx->n_odo))) return rc;
  if ((rc = rb_h2d(ctx, d_Q, Q, sizeof(double) * ctx->nw * ctx->nw))) return rc;
  NormalSrc ns;
  ns.Z = nullptr; ns.seed = ctx->cfg.seed; ns.sweep = 0; ns.t = 0;
  if (Z) {
    if ((rc = rb_h2d(ctx, d_Z, Z, sizeof(double) * ctx->nz * N))) return rc;
    ns.Z = d_Z;
  }
  k_propagate<<<(N + 127) / 128, 128, 0, ctx->stream>>>(ctx->mc, N, N, d_in, d_ai, d_dx, dt, d_Q, ns, d_out);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return rb_d2h(ctx, xn_out, d_out, sizeof(double) * n * N);
}

extern "C" int rbslam_op_dyn_logweight(rbslam_ctx *ctx, int32_t N, const double *xnk_t, const double *xn,
                                       const double *dx, double dt, const double *Q, int32_t use_default,
                                       double *logwDyn) {
  if (!ctx || !xnk_t || !xn || !dx || !Q || !logwDyn || N < 1) return RBSLAM_EARG;
  CK(cudaSetDevice(ctx->cfg.device));
  const int n = ctx->n;
  TMP_ALLOC(d_xk, double, n);
  TMP_ALLOC(d_xn, double, (size_t)n * N);
  TMP_ALLOC(d_dx, double, ctx->n_odo);
  TMP_ALLOC(d_Q, double, ctx->nw * ctx->nw);
  TMP_ALLOC(d_o, double, N);
  int rc;
  if ((rc = rb_h2d(ctx, d_xk, xnk_t, sizeof(double) * n))) return rc;
  if ((rc = rb_h2d(ctx, d_xn, xn, sizeof(double) * n * N))) return rc;
  if ((rc = rb_h2d(ctx, d_dx, dx, sizeof(double) * ctx->n_odo))) return rc;
  if ((rc = rb_h2d(ctx, d_Q, Q, sizeof(double) * ctx->nw * ctx->nw))) return rc;
  k_dyn_logweight<<<(N + 127) / 128, 128, 0, ctx->stream>>>(ctx->mc, N, d_xk, d_xn, d_dx, dt, d_Q, use_default, d_o);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return rb_d2h(ctx, logwDyn, d_o, sizeof(double) * N);
}

// dy [N x d x M] MATLAB layout (particle index fastest)
extern "C" int rbslam_op_meas_jacobian(rbslam_ctx *ctx, int32_t N, const double *xn, const double *xl,
                                       double *dy, double *yhat) {
  if (!ctx || !xn || !dy || N < 1) return RBSLAM_EARG;
  if (N > ctx->N) return ctx->fail(RBSLAM_EARG, "op N exceeds context N");
  const bool sparse = ctx->mc.family == FAM_SPARSE_VISUAL2D;
  if (sparse && !xl) return ctx->fail(RBSLAM_EARG, "sparse measModel needs xl");
  CK(cudaSetDevice(ctx->cfg.device));
  const int n = ctx->n, M = ctx->M, d = ctx->d;
  TMP_ALLOC(d_xn, double, (size_t)n * N);
  TMP_ALLOC(d_xl, double, (size_t)M * N);
  int rc;
  if ((rc = rb_h2d(ctx, d_xn, xn, sizeof(double) * n * N))) return rc;
  if (sparse && (rc = rb_h2d(ctx, d_xl, xl, sizeof(double) * (size_t)M * N))) return rc;
  k_meas<<<N, 128, 0, ctx->stream>>>(ctx->mc, N, d_xn, d_xl, M, nullptr, ctx->d_H, ctx->hs_p, ctx->hs_a,
                                     ctx->hs_c, ctx->ld, ctx->d_yhat);
  ctx->launches += 1;
  CK(cudaGetLastError());
  std::vector<double> h((size_t)N * ctx->hs_p);
  if ((rc = rb_d2h(ctx, h.data(), ctx->d_H, h.size() * 8))) return rc;
  for (int i = 0; i < N; ++i)
    for (int a = 0; a < d; ++a)
      for (int c = 0; c < M; ++c)
        dy[i + (size_t)N * (a + (size_t)d * c)] = h[(size_t)i * ctx->hs_p + (size_t)a * ctx->hs_a + (size_t)c * ctx->hs_c];
  if (sparse && yhat && (rc = rb_d2h(ctx, yhat, ctx->d_yhat, sizeof(double) * d * N))) return rc;
  return RBSLAM_OK;
}

extern "C" int rbslam_op_jacobian_phi3d(rbslam_ctx *ctx, int32_t N, const double *x, const double *lo,
                                        const double *hi, double *J) {
  if (!ctx || !x || !lo || !hi || !J || N < 1) return RBSLAM_EARG;
  if (ctx->mc.family != FAM_DENSE_MAG3D) return ctx->fail(RBSLAM_EMODEL, "JacobianPhi3D needs the 3-D dense basis");
  for (int k = 0; k < 3; ++k)
    if (!(hi[k] > lo[k])) return ctx->fail(RBSLAM_EARG, "JacobianPhi3D: upper bound must exceed lower bound");
  CK(cudaSetDevice(ctx->cfg.device));
  const int m = ctx->mc.m;
  TMP_ALLOC(d_x, double, (size_t)3 * N);
  TMP_ALLOC(d_J, double, (size_t)9 * m * N);
  int rc;
  if ((rc = rb_h2d(ctx, d_x, x, sizeof(double) * 3 * N))) return rc;
  k_jacobian_phi3d<<<N, 128, 0, ctx->stream>>>(ctx->mc, N, d_x, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], d_J);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return rb_d2h(ctx, J, d_J, sizeof(double) * 9 * (size_t)m * N);
}

__global__ void k_op_unpack(int M, int ld, size_t slab, double *__restrict__ P, const double *__restrict__ in) {
  const int c = blockIdx.x, i = blockIdx.y;
  double *dst = P + (size_t)i * slab + (size_t)c * ld;
  const double *src = in + ((size_t)i * M + c) * M;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) dst[r] = r < M ? src[r] : 0.0;
}
__global__ void k_op_iota(int *slot, int *src, int *listB, int *counts, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { slot[i] = i; src[i] = i; listB[i] = i; }
  if (i == 0) { counts[0] = 0; counts[1] = N; }
}

extern "C" int rbslam_op_kalman_update(rbslam_ctx *ctx, int32_t N, const double *xn, const double *H,
                                       const double *y_t, const double *R, double jitter, double *xl,
                                       double *P, double *logw) {
  if (!ctx || !y_t || !R || !xl || !P || !logw) return RBSLAM_EARG;
  if (ctx->shard_ws) return ctx->fail(RBSLAM_EARG, "op_kalman_update is not available on a sharded context");
  if (N != ctx->N) return ctx->fail(RBSLAM_EARG, "op_kalman_update: N must equal the context's N");
  if (!H && !xn) return ctx->fail(RBSLAM_EARG, "op_kalman_update: need H or xn");
  CK(cudaSetDevice(ctx->cfg.device));
  const int M = ctx->M, d = ctx->d, n = ctx->n;
  const bool sparse = ctx->mc.family == FAM_SPARSE_VISUAL2D;
  int rc;
  ctx->cs = 0; ctx->cx = 0; ctx->t = 0;
  ctx->running = false;
  if (ctx->kpath == 1) {
    ctx->cg = 0; ctx->pending = false;
    CK(cudaMemsetAsync(ctx->d_G4[0], 0, sizeof(double) * (size_t)N * ctx->ld * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_KS4[0], 0, sizeof(double) * (size_t)N * ctx->ld * 4, ctx->stream));
  }
  k_op_iota<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_slot[0], ctx->d_src_slot, ctx->d_listB, ctx->d_counts, N);
  if ((rc = rb_h2d(ctx, ctx->d_xl[0], xl, sizeof(double) * (size_t)M * N))) return rc;
  {  // P -> slabs in chunks
    const size_t per = (size_t)M * M;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>(N, (64u << 20) / (per * 8) + 1));
    TMP_ALLOC(tmp, double, per * chunk);
    for (int i0 = 0; i0 < N; i0 += chunk) {
      const int cnt = std::min(chunk, N - i0);
      if ((rc = rb_h2d(ctx, tmp, P + (size_t)i0 * per, per * cnt * 8))) return rc;
      if (ctx->pt) k_unpack_slabs_pt<<<dim3(ctx->ld / 8, cnt), 256, 0, ctx->stream>>>(M, ctx->ld, ctx->slab, ctx->d_P + (size_t)i0 * ctx->slab, tmp);
      else k_op_unpack<<<dim3(M, cnt), 128, 0, ctx->stream>>>(M, ctx->ld, ctx->slab, ctx->d_P + (size_t)i0 * ctx->slab, tmp);
      CK(cudaStreamSynchronize(ctx->stream));
    }
  }
  TMP_ALLOC(d_y, double, d);
  TMP_ALLOC(d_R, double, d * d);
  if ((rc = rb_h2d(ctx, d_y, y_t, sizeof(double) * d))) return rc;
  if ((rc = rb_h2d(ctx, d_R, R, sizeof(double) * d * d))) return rc;
  if (H) {
    if (sparse) return ctx->fail(RBSLAM_EARG, "sparse family evaluates its own Jacobian (pass H=NULL)");
    std::vector<double> h((size_t)N * ctx->hs_p, 0.0);
    for (int i = 0; i < N; ++i)
      for (int a = 0; a < d; ++a)
        for (int c = 0; c < M; ++c)
          h[(size_t)i * ctx->hs_p + (size_t)a * ctx->hs_a + (size_t)c * ctx->hs_c] = H[i + (size_t)N * (a + (size_t)d * c)];
    if ((rc = rb_h2d(ctx, ctx->d_H, h.data(), h.size() * 8))) return rc;
  } else {
    TMP_ALLOC(d_xn, double, (size_t)n * N);
    if ((rc = rb_h2d(ctx, d_xn, xn, sizeof(double) * n * N))) return rc;
    k_meas<<<N, 128, 0, ctx->stream>>>(ctx->mc, N, d_xn, ctx->d_xl[0], M, nullptr, ctx->d_H, ctx->hs_p,
                                       ctx->hs_a, ctx->hs_c, ctx->ld, ctx->d_yhat);
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(DevStatus), ctx->stream));
  double *R_save = ctx->d_R;
  const double jit_save = ctx->jitter;
  ctx->d_R = d_R; ctx->jitter = jitter;
  rc = rb_kalman_phase(ctx, d_y, false);
  ctx->d_R = R_save; ctx->jitter = jit_save;
  if (rc) return rc;
  if ((rc = rb_check_status(ctx))) return rc;
  if ((rc = rb_d2h(ctx, xl, ctx->d_xl[ctx->cx], sizeof(double) * (size_t)M * N))) return rc;
  if ((rc = rb_read_slabs(ctx, ctx->d_P, P))) return rc;
  return rb_d2h(ctx, logw, ctx->d_logw, sizeof(double) * N);
}

// ---------------------------------------------------------------------------
// Host-only migration planner (multi-GPU).  Particles are owned by ranks in
// blocks; after resampling, offspring stay on their ancestor's rank up to the
// per-rank capacity; the surplus is shipped to ranks with a deficit, lowest
// particle index first, lowest rank first.  Pure integer logic: deterministic and
// identical on every rank.
// ---------------------------------------------------------------------------
extern "C" int rbslam_plan_migration(int32_t N, int32_t world, const int32_t *ai,
                                     const int32_t *old_owner, int32_t *new_owner,
                                     int32_t *n_migrate) {
  if (N < 1 || world < 1 || !ai || !old_owner || !new_owner) return RBSLAM_EARG;
  std::vector<int> cap(world), load(world, 0);
  for (int r = 0; r < world; ++r) cap[r] = N / world + (r < N % world ? 1 : 0);
  std::vector<int> surplus;
  for (int i = 0; i < N; ++i) {
    if (ai[i] < 0 || ai[i] >= N) return RBSLAM_EARG;
    const int r = old_owner[ai[i]];
    if (r < 0 || r >= world) return RBSLAM_EARG;
    if (load[r] < cap[r]) { new_owner[i] = r; ++load[r]; }
    else { new_owner[i] = -1; surplus.push_back(i); }
  }
  int r = 0;
  for (int i : surplus) {
    while (r < world && load[r] >= cap[r]) ++r;
    if (r >= world) return RBSLAM_EARG;
    new_owner[i] = r; ++load[r];
  }
  if (n_migrate) *n_migrate = (int)surplus.size();
  return RBSLAM_OK;
}
