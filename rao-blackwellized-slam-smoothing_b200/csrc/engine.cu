// librbslam: context life cycle and the particle-filter engine
// (replaces the body of src/particleFilter.m:52-233 of the reference).
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include "engine.h"
#include "step_kernels.cuh"
#include "kalman_kernels.cuh"
#include "kalman_stream.cuh"
#include "family_kernels.cuh"
#include "packed_kernels.cuh"
#include "engine_internal.h"

using namespace rb;

static std::string g_create_error;

void rbslam_ctx::fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file,
           line, what);
  err = buf;
}

// ---------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------
__global__ void k_fill_int_iota(int *p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// P[slot] = P0 for every slab (pads zero); grid (col chunk, particle)
__global__ void k_init_slabs(double *__restrict__ P, size_t slab, int ld, int M, int N,
                             const double *__restrict__ P0, int diag_inverse) {
  const int c = blockIdx.x;
  for (int i = blockIdx.y; i < N; i += gridDim.y) {
    double *dst = P + (size_t)i * slab + (size_t)c * ld;
    for (int r = threadIdx.x; r < ld; r += blockDim.x) {
      double v = 0.0;
      if (r < M) {
        if (diag_inverse) v = (r == c) ? 1.0 / P0[r + (size_t)r * M] : 0.0;  // Imat0 = diag(1./diag(P0))
        else v = P0[r + (size_t)c * M];
      }
      dst[r] = v;
    }
  }
}

__global__ void k_init_xl(double *__restrict__ xl, int M, int N, const double *__restrict__ x0,
                          int cols) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * N) return;
  const int r = idx % M;
  const size_t i = idx / M;
  xl[idx] = cols > 1 ? x0[r + i * M] : x0[r];
}

__global__ void k_init_xn(double *__restrict__ xn, int n, int N, const double *__restrict__ x0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n * N) xn[idx] = x0[idx % n];
}

__global__ void k_fill(double *p, size_t n, double v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// final map extraction (src/particleFilter.m:220-230):
//   out[0:M)           xl_max  = xl(:,iw_max)
//   out[M:2M)          xl_mean = sum(xl.*w,2)   (fixed i order)
__global__ void k_final_means(int N, int M, const double *__restrict__ xl,
                              const double *__restrict__ w, const int *__restrict__ iw_max,
                              double *__restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  out[r] = xl[(size_t)(*iw_max) * M + r];
  double acc = 0.0;
  for (int i = 0; i < N; ++i) acc = fma(xl[(size_t)i * M + r], w[i], acc);
  out[M + r] = acc;
}
// P_max = P(:,:,iw_max);  P_mean = w(N)*(P(:,:,N) + dx dx'), dx = xl_mean - xl(:,N)
// (quirk Q1: the reference assigns inside its loop, src/particleFilter.m:228-230)
__global__ void k_final_cov(int N, int M, int ld, size_t slab, const double *__restrict__ P,
                            const int *__restrict__ slot, const double *__restrict__ xl,
                            const double *__restrict__ w, const int *__restrict__ iw_max,
                            const double *__restrict__ means, double *__restrict__ Pmax,
                            double *__restrict__ Pmean, const double *__restrict__ G4,
                            const double *__restrict__ KS4, int layout) {
  const int c = blockIdx.x;
  const int im = *iw_max, il = N - 1;
  const double *Pm = P + (size_t)slot[im] * slab;
  const double *Pl = P + (size_t)slot[il] * slab;
  const double *xll = xl + (size_t)il * M;
  const double *xmean = means + M;
  const double wl = w[il];
  const double dc = xmean[c] - xll[c];
  for (int r = threadIdx.x; r < M; r += blockDim.x) {
    // packed symmetric slabs store (r,c) of an upper block as its mirror image
    int rr, cc;
    slab_elem_rc(layout, r, c, rr, cc);
    const size_t off = slab_elem(layout, ld, r, c);
    double pm = Pm[off], pl = Pl[off];
    if (G4) {   // deferred downdate of the streaming path: P(r,c) -= KS(r,b) G(c,b)
      for (int b = 0; b < 4; ++b) {
        pm = fma(-KS4[((size_t)im * ld + rr) * 4 + b], G4[((size_t)im * ld + cc) * 4 + b], pm);
        pl = fma(-KS4[((size_t)il * ld + rr) * 4 + b], G4[((size_t)il * ld + cc) * 4 + b], pl);
      }
    }
    Pmax[r + (size_t)c * M] = pm;
    Pmean[r + (size_t)c * M] = wl * (pl + (xmean[r] - xll[r]) * dc);
  }
}

// pack slabs into dense [M x M x cnt] logical order (read_particles)
__global__ void k_pack_slabs(int M, int ld, size_t slab, const double *__restrict__ P,
                             const int *__restrict__ slot, int i0, double *__restrict__ out, int layout) {
  const int c = blockIdx.x, i = blockIdx.y;
  const double *src = P + (size_t)slot[i0 + i] * slab;
  double *dst = out + ((size_t)i * M + c) * M;
  for (int r = threadIdx.x; r < M; r += blockDim.x) dst[r] = src[slab_elem(layout, ld, r, c)];
}
__global__ void k_unpack_slabs(int M, int ld, size_t slab, double *__restrict__ P,
                               const int *__restrict__ slot, int i0, const double *__restrict__ in) {
  const int c = blockIdx.x, i = blockIdx.y;
  double *dst = P + (size_t)slot[i0 + i] * slab + (size_t)c * ld;
  const double *src = in + ((size_t)i * M + c) * M;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) dst[r] = r < M ? src[r] : 0.0;
}

// genealogy: xn_traj(:,i,s) = X_s(:, b_s(i)),  b_T(i)=i, b_{s-1} = A_s(b_s)
// replaces the per-step reshuffle xn_traj(:,:,1:t-1)=xn_traj(:,ai,1:t-1) (src/particleFilter.m:118)
__global__ void k_trace(int N, int n, int T, const double *__restrict__ X, const int *__restrict__ A,
                        const int *__restrict__ only_particle, double *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_out = only_particle ? 1 : N;
  if (i >= n_out) return;
  int b = only_particle ? *only_particle : i;
  for (int s = T - 1; s >= 0; --s) {
    const double *x = X + ((size_t)s * N + b) * n;
    double *o = only_particle ? out + (size_t)s * n : out + ((size_t)s * N + i) * n;
    for (int j = 0; j < n; ++j) o[j] = x[j];
    if (s > 0) b = A[(size_t)s * N + b];
  }
}

// yhattraj(:,t) = predicted measurement of the highest-weight particle (src/particleFilter.m:200-203):
// dense families yhat = H_i * xl_i with the mean BEFORE the update, sparse family the model's yhat.
// Only evaluated when a step callback (the makePlots hook) is registered.
__global__ void k_yhat_max(int M, int d, const int *__restrict__ iw_max, const double *__restrict__ H, size_t hs_p,
                           int hs_a, int hs_c, const double *__restrict__ xl_old, const int *__restrict__ anc,
                           const double *__restrict__ yhat_model, double *__restrict__ out) {
  const int i = *iw_max;
  const int a = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (a >= d) return;
  if (yhat_model) { if (lane == 0) out[a] = yhat_model[a + (size_t)i * d]; return; }
  const double *x = xl_old + (size_t)(anc ? anc[i] : i) * M;
  const double *Hi = H + (size_t)i * hs_p + (size_t)a * hs_a;
  double acc = 0.0;
  for (int c = lane; c < M; c += 32) acc = fma(Hi[(size_t)c * hs_c], x[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[a] = acc;
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
int rb_h2d(rbslam_ctx *ctx, void *dst, const void *src, size_t bytes) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // caller's buffer may be pageable / reused
  ctx->h2d += (int64_t)bytes;
  return RBSLAM_OK;
}
int rb_d2h(rbslam_ctx *ctx, void *dst, const void *src, size_t bytes) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->d2h += (int64_t)bytes;
  return RBSLAM_OK;
}

void rb_phase_begin(rbslam_ctx *ctx, int id) {
  if (!ctx->phase_timing) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, ctx->stream);
  ctx->ph_events.push_back(e);
  ctx->ph_ids.push_back(id);
}
void rb_phase_end(rbslam_ctx *ctx) {
  if (!ctx->phase_timing) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, ctx->stream);
  ctx->ph_events.push_back(e);
}

bool rb_fast_scan() {
  static const bool on = [] { const char *e = getenv("RBSLAM_EXACT_SCAN"); return !(e && atoi(e)); }();
  return on;
}

int rb_check_status(rbslam_ctx *ctx) {
  DevStatus st;
  CK(cudaMemcpyAsync(&st, ctx->d_status, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  if (st.peer_timeout)
    return ctx->fail(RBSLAM_ECUDA, "sharded filter: a peer rank did not reach the barrier (it failed or exited)");
  if (st.not_pd) {
    char buf[256];
    if (st.not_pd_step < 0)
      snprintf(buf, sizeof buf, "innovation covariance not positive definite even with jitter (reported by peer rank %d)",
               -1 - st.not_pd_particle);
    else
      snprintf(buf, sizeof buf,
               "innovation covariance not positive definite even with jitter (step %d, particle %d)",
               st.not_pd_step, st.not_pd_particle);
    return ctx->fail(RBSLAM_ENOTPD, buf);
  }
  return RBSLAM_OK;
}

// ---------------------------------------------------------------------------
// life cycle
// ---------------------------------------------------------------------------
// CUDA loads kernels lazily by default, and the first launch of a kernel may have to synchronise the device.
// A single-process group whose shards SHARE a GPU enqueues shard 0's step -- including a peer barrier kernel that
// spins until shard 1 arrives -- before shard 1's; if one of shard 1's kernels is launched for the first time at
// that point, its load waits for the spinning kernel and the barrier only ends by its time-out.  (Seen as a
// 20 s "peer did not reach the barrier" when a group on one GPU was the first thing a process ran.)  Eager
// module loading removes the hazard; it is requested when the library is loaded, unless the user chose a mode.
__attribute__((constructor)) static void rb_request_eager_module_loading() { setenv("CUDA_MODULE_LOADING", "EAGER", 0); }

extern "C" int rbslam_version(void) { return RBSLAM_VERSION; }

extern "C" int rbslam_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const char *rbslam_last_error(const rbslam_ctx *ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

static int create_impl(rbslam_ctx *ctx, const rbslam_config *cfg) {
  if (cfg->struct_size != (int32_t)sizeof(rbslam_config))
    return ctx->fail(RBSLAM_EARG, "rbslam_config.struct_size mismatch (ABI)");
  ctx->cfg = *cfg;
  ModelConsts &mc = ctx->mc;
  memset(&mc, 0, sizeof mc);
  mc.family = cfg->model;
  mc.m = cfg->m_basis;
  if (cfg->N < 1 || cfg->T < 1 || cfg->m_basis < 1) return ctx->fail(RBSLAM_EARG, "N, T, m_basis must be >= 1");
  switch (cfg->model) {
    case RBSLAM_MODEL_DENSE_MAG3D:
      mc.n = 7; mc.d = 3; mc.M = cfg->m_basis + 3; mc.nz = 6; mc.nw = 6; mc.n_odo = 7; mc.dim = 3;
      break;
    case RBSLAM_MODEL_DENSE_RADIO2D:
      mc.n = 3; mc.d = 1; mc.M = cfg->m_basis; mc.nz = 1; mc.nw = 1; mc.n_odo = 3; mc.dim = 2;
      break;
    case RBSLAM_MODEL_SPARSE_VISUAL2D:
      mc.n = 3; mc.d = cfg->m_basis; mc.M = 2 * cfg->m_basis; mc.nz = 3; mc.nw = 3; mc.n_odo = 3;
      mc.dim = 0;
      mc.cam_f = cfg->cam_f; mc.cam_fp = cfg->cam_fp; mc.cam_fw = cfg->cam_fw;
      break;
    default:
      return ctx->fail(RBSLAM_EMODEL, "unsupported model family (no CPU fallback by design)");
  }
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world)
    return ctx->fail(RBSLAM_EARG, "bad rank/world");
  if (cfg->world > 1 && cfg->N % cfg->world) return ctx->fail(RBSLAM_EARG, "N must be divisible by world");
  ctx->N = cfg->N / cfg->world;   // local particle count (== N on one GPU)
  ctx->T = cfg->T; ctx->M = mc.M; ctx->d = mc.d; ctx->n = mc.n;
  ctx->nz = mc.nz; ctx->nw = mc.nw; ctx->n_odo = mc.n_odo;
  const int M = ctx->M, N = ctx->N, T = ctx->T, d = ctx->d, n = ctx->n;
  ctx->ld = cfg->ld > 0 ? cfg->ld : ((M + 7) / 8) * 8;
  if (ctx->ld < M || (ctx->ld & 1)) return ctx->fail(RBSLAM_EARG, "ld must be even and >= M");
  ctx->ldh = ctx->ld;
  ctx->slab = (size_t)ctx->ld * M;
  if (d > RB_DMAX) return ctx->fail(RBSLAM_EARG, "more than 32 measurements per step is not supported");

  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return ctx->fail(RBSLAM_EARG, "bad device ordinal");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) return ctx->fail(RBSLAM_ECUDA, "librbslam is built for sm_100a (B200) only");
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));

  // Kalman-update path: 0 = single pass in shared memory (M*M*8 B fits one CTA),
  // 1 = streaming pass with deferred downdate (kalman_stream.cuh), 2 = legacy 3-kernel path
  const bool small = kalman_small_smem(M, d) + 1024 <= ctx->smem_optin;
  const bool can_stream = d <= 4 && ctx->ld <= 4 * 2 * RB_STREAM_THREADS;
  if (cfg->information_form && (!can_stream || d > 3))
    return ctx->fail(RBSLAM_EARG, "information form needs d<=3 and M<=1536");
  // packed symmetric tile slabs (packed_kernels.cuh): explicit (7), or chosen by the filter-only
  // auto mode (-1) whenever the streaming path would be used
  const bool pt_ok = d <= 4 && (ctx->ld % 8) == 0 && ctx->ld / 8 <= 135 && !cfg->information_form;
  if (cfg->kalman_variant == 7 && !pt_ok)
    return ctx->fail(RBSLAM_EARG, "kalman_variant 7 needs d<=4, ld a multiple of 8, M<=1080 and the covariance form");
  if (cfg->kalman_variant >= 4 && cfg->kalman_variant <= 6)
    return ctx->fail(RBSLAM_EARG, "kalman_variant 4/5/6 were removed (superseded by 7, packed symmetric slabs)");
  if (cfg->kalman_variant == 7 || (cfg->kalman_variant == -1 && pt_ok && !small)) {
    ctx->kpath = 1;
    ctx->pt = true;
    ctx->layout = RB_LAYOUT_PT;
    ctx->slab = pt_slab_doubles(ctx->ld);
    if (const char *e = getenv("RBSLAM_PT_CFG")) { int ts = 48, ns = 4; if (sscanf(e, "%d,%d", &ts, &ns) == 2) { ctx->pt_ts = ts; ctx->pt_ns = ns; } }
    if (ctx->pt_ns < 2 || ctx->pt_ns > 8 || ctx->pt_ts < 1) return ctx->fail(RBSLAM_EARG, "RBSLAM_PT_CFG: need 2..8 slots");
    while (ctx->pt_ns > 2 && pt_smem_bytes(ctx->ld, ctx->pt_ts, ctx->pt_ns, ctx->pt_nw) + 2048 > ctx->smem_optin) --ctx->pt_ns;
    if (pt_smem_bytes(ctx->ld, ctx->pt_ts, ctx->pt_ns, ctx->pt_nw) + 2048 > ctx->smem_optin)
      return ctx->fail(RBSLAM_EARG, "packed streaming pass does not fit shared memory");
  } else if (small && cfg->kalman_variant != 2 && cfg->kalman_variant != 3 && !cfg->information_form) ctx->kpath = 0;
  else if (can_stream && cfg->kalman_variant != 3) ctx->kpath = 1;
  else if (d <= 4 && ctx->ld <= 2048) ctx->kpath = 2;
  else return ctx->fail(RBSLAM_EARG, "unsupported size: d>4 needs M*M*8 B to fit shared memory; "
                                     "d<=4 needs M <= 2048");
  if (ctx->kpath == 1) {
    ctx->hs_p = (size_t)ctx->ld * 4; ctx->hs_a = 1; ctx->hs_c = 4;
    int ns = N >= 4096 ? 2 : std::max(1, std::min(8, (4 * 148 + N - 1) / N));
    if (ctx->pt && N >= 5000) ns = 1;   // > 20 waves of families per launch: whole-slab items balance well enough
    if (const char *e = getenv("RBSLAM_NSPLIT")) ns = std::max(1, atoi(e));
    if (const char *e = getenv("RBSLAM_CTAS_PER_SM")) ctx->stream_ctas_per_sm = atoi(e);
    if (const char *e = getenv("RBSLAM_STREAM_HINTS")) ctx->stream_hints = atoi(e);
    int cw = ((M + ns - 1) / ns + 3) / 4 * 4;
    ctx->nsplit = (M + cw - 1) / cw; ctx->cw = cw;
    if (ctx->pt) {   // the nsplit items of a family stream panel ranges of (nearly) equal tile counts
      const int nb = ctx->ld / 8;
      ns = std::max(1, std::min(std::min(ns, RB_PT_MAXSPLIT), nb));
      const size_t tot = pt_panel_off(nb, nb);
      ctx->pt_psplit[0] = 0;
      for (int sp = 1; sp < ns; ++sp) {
        int p = ctx->pt_psplit[sp - 1] + 1;
        while (p < nb - (ns - sp) && pt_panel_off(nb, p) < tot * sp / ns) ++p;
        ctx->pt_psplit[sp] = p;
      }
      ctx->pt_psplit[ns] = nb;
      ctx->nsplit = ns;
    }
  } else {
    ctx->hs_p = (size_t)d * ctx->ldh; ctx->hs_a = ctx->ldh; ctx->hs_c = 1;
    ctx->nsplit = (M + 127) / 128; ctx->cw = (M + ctx->nsplit - 1) / ctx->nsplit;
  }

  if (mc.dim > 0) {
    if (!cfg->NN || !cfg->L) return ctx->fail(RBSLAM_EARG, "NN and L are required for dense models");
    for (int j = 0; j < mc.dim; ++j) {
      mc.L[j] = cfg->L[j];
      if (!(mc.L[j] > 0)) return ctx->fail(RBSLAM_EARG, "L must be positive");
      int mx = 0;
      for (int c = 0; c < mc.m; ++c) {
        const int v = cfg->NN[c + (size_t)j * mc.m];
        if (v < 1) return ctx->fail(RBSLAM_EARG, "NN indices must be >= 1");
        mx = std::max(mx, v);
      }
      if (mx >= RB_MAXTAB) return ctx->fail(RBSLAM_EARG, "eigenfunction index too large (>= 96)");
      mc.maxn[j] = mx;
    }
    RB_ALLOC(ctx->d_NN, (size_t)mc.m * mc.dim);
    int rc = rb_h2d(ctx, ctx->d_NN, cfg->NN, sizeof(int) * (size_t)mc.m * mc.dim);
    if (rc) return rc;
    mc.NN = ctx->d_NN;
  }

  RB_ALLOC(ctx->d_P, (size_t)N * ctx->slab);
  if (cfg->information_form) {
    RB_ALLOC(ctx->d_Imat, (size_t)N * ctx->slab);
    for (int b = 0; b < 2; ++b) { RB_ALLOC(ctx->d_ivec[b], (size_t)N * M); RB_ALLOC(ctx->d_hld[b], N); }
  }
  for (int b = 0; b < 2; ++b) { RB_ALLOC(ctx->d_xl[b], (size_t)N * M); RB_ALLOC(ctx->d_slot[b], N); }
  RB_ALLOC(ctx->d_src_slot, N); RB_ALLOC(ctx->d_first_child, N); RB_ALLOC(ctx->d_free_list, N);
  RB_ALLOC(ctx->d_listA, N); RB_ALLOC(ctx->d_listB, N); RB_ALLOC(ctx->d_counts, 8);
  RB_ALLOC(ctx->d_H, (size_t)N * ctx->hs_p);
  CK(cudaMemsetAsync(ctx->d_H, 0, sizeof(double) * (size_t)N * ctx->hs_p, ctx->stream));   // unused lanes of the 4-wide layout stay 0
  RB_ALLOC(ctx->d_yhat, (size_t)N * d);
  if (ctx->kpath == 2) {
    RB_ALLOC(ctx->d_PHpart, (size_t)N * ctx->nsplit * d * ctx->ld);
    RB_ALLOC(ctx->d_G, (size_t)N * d * ctx->ld);
    RB_ALLOC(ctx->d_KS, (size_t)N * d * ctx->ld);
  } else if (ctx->kpath == 1) {
    for (int b = 0; b < 2; ++b) {
      RB_ALLOC(ctx->d_G4[b], (size_t)N * ctx->ld * 4);
      RB_ALLOC(ctx->d_KS4[b], (size_t)N * ctx->ld * 4);
    }
    // the packed pass keeps one more partial slot (the column-side sums)
    RB_ALLOC(ctx->d_PHp, (size_t)N * (ctx->nsplit + (ctx->pt ? 1 : 0)) * ctx->ld * 4);
    // 7 scratch arrays over the source keys (N slabs; + N migrant keys in a sharded filter), two family
    // lists of 5 arrays over the items, counters: n_fb, n_fa, work counters [2]
    RB_ALLOC(ctx->d_fam, (size_t)(cfg->world > 1 ? 24 : 17) * N + 8);
    ctx->use_fam = getenv("RBSLAM_NO_FAM") == nullptr;
  }
  RB_ALLOC(ctx->d_logw, N); RB_ALLOC(ctx->d_w, N); RB_ALLOC(ctx->d_wc, N);
  ctx->T_hist = cfg->keep_history ? T : 2;
  RB_ALLOC(ctx->d_Xhist, (size_t)ctx->T_hist * N * n);
  RB_ALLOC(ctx->d_Ahist, (size_t)ctx->T_hist * N);
  RB_ALLOC(ctx->d_traj_max, (size_t)T * n); RB_ALLOC(ctx->d_traj_mean, (size_t)T * n);
  RB_ALLOC(ctx->d_iwmax, T);
  RB_ALLOC(ctx->d_status, 1);
  ctx->scratch_doubles = 2 * (size_t)M * M + 4 * (size_t)M + 64;
  RB_ALLOC(ctx->d_scratch, ctx->scratch_doubles);
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(DevStatus), ctx->stream));
  {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_kalman_small));
    ctx->smem_small_max = ctx->smem_optin - fa.sharedSizeBytes;
    CK(cudaFuncSetAttribute(k_kalman_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_small_max));
    CK(cudaFuncGetAttributes(&fa, k_resample));
    ctx->smem_resample_max = ctx->smem_optin - fa.sharedSizeBytes;
    CK(cudaFuncSetAttribute(k_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_resample_max));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if (cfg->world > 1) {
    int rcs = rb_shard_create(ctx, cfg->world, cfg->rank, cfg->N);
    if (rcs) return rcs;
  }
  return RBSLAM_OK;
}

extern "C" int rbslam_create(rbslam_ctx **out, const rbslam_config *cfg) {
  if (!out || !cfg) { g_create_error = "null argument"; return RBSLAM_EARG; }
  *out = nullptr;
  rbslam_ctx *ctx = new rbslam_ctx();
  int rc = create_impl(ctx, cfg);
  if (rc != RBSLAM_OK) {
    g_create_error = ctx->err;
    rbslam_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return RBSLAM_OK;
}

// ---------------------------------------------------------------------------
// ONE filter sharded over several GPUs, driven from ONE host process (SURVEY 8(b) "Threading": a
// MATLAB caller has a single process).  The shards are ordinary sharded contexts (sharded.cu);
// their peer tables hold plain device pointers instead of CUDA-IPC mappings, and the host thread
// enqueues every shard's step in turn -- the step itself never synchronises with the host, the
// shards meet in the peer-memory barriers on the devices.
// ---------------------------------------------------------------------------
extern "C" int rbslam_create_group(rbslam_ctx **out, const rbslam_config *cfg, const int32_t *devices, int32_t n_devices) {
  if (!out || !cfg || !devices) { g_create_error = "null argument"; return RBSLAM_EARG; }
  *out = nullptr;
  if (n_devices < 1 || n_devices > 8) { g_create_error = "1..8 devices"; return RBSLAM_EARG; }
  if (n_devices == 1) {
    rbslam_config c1 = *cfg;
    c1.device = devices[0]; c1.rank = 0; c1.world = 1;
    return rbslam_create(out, &c1);
  }
  if (cfg->rng_mode != RBSLAM_RNG_PHILOX || !cfg->keep_history) {
    g_create_error = "a device group runs the sharded filter: rng_mode PHILOX and keep_history=1";
    return RBSLAM_EARG;
  }
  std::vector<rbslam_ctx *> sh;
  int rc = RBSLAM_OK;
  for (int r = 0; r < n_devices && rc == RBSLAM_OK; ++r) {
    rbslam_config c = *cfg;
    c.device = devices[r]; c.rank = r; c.world = n_devices;
    rbslam_ctx *ctx = nullptr;
    rc = rbslam_create(&ctx, &c);
    if (rc == RBSLAM_OK) sh.push_back(ctx);
  }
  for (int r = 0; r < (int)sh.size() && rc == RBSLAM_OK; ++r)
    for (int p = 0; p < (int)sh.size() && rc == RBSLAM_OK; ++p) {
      if (p == r) continue;
      if (devices[p] != devices[r]) {
        cudaSetDevice(devices[r]);
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[r], devices[p]);
        cudaError_t e = can ? cudaDeviceEnablePeerAccess(devices[p], 0) : cudaErrorPeerAccessUnsupported;
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        if (e != cudaSuccess) { g_create_error = "devices of the group cannot access each other's memory"; rc = RBSLAM_ECUDA; break; }
      }
      rc = rb_shard_wire(sh[r], sh[p]);
    }
  if (rc != RBSLAM_OK) {
    for (rbslam_ctx *c : sh) rbslam_destroy(c);
    return rc;
  }
  for (size_t r = 1; r < sh.size(); ++r) sh[r]->group_member = true;
  sh[0]->group = sh;
  *out = sh[0];
  return RBSLAM_OK;
}

static void free_run_inputs(rbslam_ctx *ctx) {
  void **ptrs[] = {(void **)&ctx->d_odo, (void **)&ctx->d_y, (void **)&ctx->d_Q, (void **)&ctx->d_R,
                   (void **)&ctx->d_P0, (void **)&ctx->d_x0lin, (void **)&ctx->d_U, (void **)&ctx->d_Z,
                   (void **)&ctx->d_forced, (void **)&ctx->d_logw_hist, (void **)&ctx->d_w_hist};
  for (auto p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
}

extern "C" void rbslam_destroy(rbslam_ctx *ctx) {
  if (!ctx) return;
  if (!ctx->group.empty()) {   // group leader: the other shards go first
    std::vector<rbslam_ctx *> members(ctx->group.begin() + 1, ctx->group.end());
    ctx->group.clear();
    for (rbslam_ctx *m : members) { m->group_member = false; rbslam_destroy(m); }
  }
  if (ctx->replica_group && ctx->replica_rank == 0) rb_replicas_free(ctx);   // leader: the other replicas go first
  if (ctx->cfg.device >= 0) cudaSetDevice(ctx->cfg.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  free_run_inputs(ctx);
  rb_smoother_free(ctx);
  rb_shard_free(ctx);
  void *ptrs[] = {ctx->d_NN, ctx->d_P, ctx->d_Imat, ctx->d_xl[0], ctx->d_xl[1], ctx->d_ivec[0],
                  ctx->d_ivec[1], ctx->d_hld[0], ctx->d_hld[1], ctx->d_slot[0], ctx->d_slot[1],
                  ctx->d_src_slot, ctx->d_first_child, ctx->d_free_list, ctx->d_listA, ctx->d_listB,
                  ctx->d_counts, ctx->d_H, ctx->d_yhat, ctx->d_PHpart, ctx->d_G, ctx->d_KS,
                  ctx->d_G4[0], ctx->d_G4[1], ctx->d_KS4[0], ctx->d_KS4[1], ctx->d_PHp, ctx->d_fam,
                  ctx->d_logw, ctx->d_w, ctx->d_wc, ctx->d_Xhist, ctx->d_Ahist, ctx->d_traj_max,
                  ctx->d_traj_mean, ctx->d_iwmax, ctx->d_status, ctx->d_scratch, ctx->d_yhattraj, ctx->d_chol_fail, ctx->d_normws};
  for (void *p : ptrs) if (p) cudaFree(p);
  for (auto e : ctx->ph_events) cudaEventDestroy(e);
  for (auto &e : ctx->user_events) if (e) cudaEventDestroy(e);
  if (ctx->ev_fetch) cudaEventDestroy(ctx->ev_fetch);
  if (ctx->ev_plan) cudaEventDestroy(ctx->ev_plan);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int rbslam_dims(const rbslam_ctx *ctx, int32_t out7[7]) {
  if (!ctx || !out7) return RBSLAM_EARG;
  out7[0] = ctx->n; out7[1] = ctx->d; out7[2] = ctx->M; out7[3] = ctx->nz; out7[4] = ctx->nw;
  out7[5] = ctx->n_odo; out7[6] = ctx->ld;
  return RBSLAM_OK;
}

extern "C" int rbslam_step_callback(rbslam_ctx *ctx, rbslam_step_fn fn, void *user) {
  if (!ctx) return RBSLAM_EARG;
  ctx->step_fn = fn; ctx->step_user = user;
  return RBSLAM_OK;
}

extern "C" void *rbslam_stream(rbslam_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int rbslam_sync(rbslam_ctx *ctx) {
  if (!ctx) return RBSLAM_EARG;
  for (size_t r = 1; r < ctx->group.size(); ++r) {
    rbslam_ctx *m = ctx->group[r];
    cudaSetDevice(m->cfg.device);
    if (cudaStreamSynchronize(m->stream) != cudaSuccess) return ctx->fail(RBSLAM_ECUDA, "group shard failed");
    int rcm = rb_check_status(m);
    if (rcm) return ctx->fail(rcm, m->err);
  }
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaStreamSynchronize(ctx->stream));
  return rb_check_status(ctx);
}

extern "C" int rbslam_status_counters(rbslam_ctx *ctx, int32_t *used_jitter, int32_t *clamped_draws,
                                      int32_t *exact_scan_runs) {
  if (!ctx) return RBSLAM_EARG;
  CK(cudaSetDevice(ctx->cfg.device));
  DevStatus st;
  CK(cudaMemcpyAsync(&st, ctx->d_status, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (used_jitter) *used_jitter = st.used_jitter;
  if (clamped_draws) *clamped_draws = st.clamp_sample;
  if (exact_scan_runs) *exact_scan_runs = st.scan_fallbacks;
  return RBSLAM_OK;
}

extern "C" int rbslam_counters(rbslam_ctx *ctx, int64_t *kl, int64_t *h2d, int64_t *d2h) {
  if (!ctx) return RBSLAM_EARG;
  if (kl) *kl = ctx->launches;
  if (h2d) *h2d = ctx->h2d;
  if (d2h) *d2h = ctx->d2h;
  for (size_t r = 1; r < ctx->group.size(); ++r) {   // group leader: the whole group's work
    if (kl) *kl += ctx->group[r]->launches;
    if (h2d) *h2d += ctx->group[r]->h2d;
    if (d2h) *d2h += ctx->group[r]->d2h;
  }
  return RBSLAM_OK;
}

extern "C" int rbslam_event_record(rbslam_ctx *ctx, int32_t slot) {
  if (!ctx || slot < 0 || slot >= 16) return RBSLAM_EARG;
  CK(cudaSetDevice(ctx->cfg.device));
  if (!ctx->user_events[slot]) CK(cudaEventCreate(&ctx->user_events[slot]));
  CK(cudaEventRecord(ctx->user_events[slot], ctx->stream));
  return RBSLAM_OK;
}
extern "C" int rbslam_event_elapsed(rbslam_ctx *ctx, int32_t a, int32_t b, float *ms) {
  if (!ctx || !ms || a < 0 || b < 0 || a >= 16 || b >= 16 || !ctx->user_events[a] || !ctx->user_events[b])
    return RBSLAM_EARG;
  CK(cudaEventSynchronize(ctx->user_events[b]));
  CK(cudaEventElapsedTime(ms, ctx->user_events[a], ctx->user_events[b]));
  return RBSLAM_OK;
}
extern "C" int rbslam_phase_timing(rbslam_ctx *ctx, int32_t enable) {
  if (!ctx) return RBSLAM_EARG;
  ctx->phase_timing = enable != 0;
  return RBSLAM_OK;
}
extern "C" int rbslam_phase_times(rbslam_ctx *ctx, double ms8[8]) {
  if (!ctx || !ms8) return RBSLAM_EARG;
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t k = 0; k < ctx->ph_ids.size(); ++k) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ph_events[2 * k], ctx->ph_events[2 * k + 1]);
    ctx->phase_ms[ctx->ph_ids[k]] += ms;
  }
  for (auto e : ctx->ph_events) cudaEventDestroy(e);
  ctx->ph_events.clear(); ctx->ph_ids.clear();
  for (int k = 0; k < 8; ++k) { ms8[k] = k < RB_PH_COUNT ? ctx->phase_ms[k] : 0.0; }
  for (int k = 0; k < RB_PH_COUNT; ++k) ctx->phase_ms[k] = 0.0;
  return RBSLAM_OK;
}

// ---------------------------------------------------------------------------
// run inputs
// ---------------------------------------------------------------------------
int rb_upload_inputs(rbslam_ctx *ctx, const rbslam_inputs *in, int K) {
  const int N = ctx->N, M = ctx->M, d = ctx->d, n = ctx->n;
  if (!in || !in->y || !in->x0_nonLin || !in->x0_lin || !in->P0_lin || !in->Q || !in->R || !in->dt)
    return ctx->fail(RBSLAM_EARG, "missing input array");
  const int T = in->T;
  if (T < 1 || T > ctx->T) return ctx->fail(RBSLAM_EARG, "inputs.T exceeds the context's capacity");
  if (T > 1 && (!in->odometry || in->odo_rows < T - 1)) return ctx->fail(RBSLAM_EARG, "odometry needs >= T-1 rows");
  if (in->x0_lin_cols != 1 && in->x0_lin_cols != N) return ctx->fail(RBSLAM_EARG, "x0_lin must have 1 or N columns");
  if (in->Q_pages != 1 && in->Q_pages < T - 1) return ctx->fail(RBSLAM_EARG, "Q needs 1 or >= T-1 pages");
  if (in->dt_len != 1 && in->dt_len < T - 1) return ctx->fail(RBSLAM_EARG, "dt needs 1 or >= T-1 entries");
  if (ctx->cfg.rng_mode == RBSLAM_RNG_INJECTED && T > 1 && !in->forced_ancestors && !in->U)
    return ctx->fail(RBSLAM_EARG, "rng_mode INJECTED needs U (and Z)");
  if (ctx->cfg.rng_mode == RBSLAM_RNG_INJECTED && T > 1 && !in->Z)
    return ctx->fail(RBSLAM_EARG, "rng_mode INJECTED needs Z");
  CK(cudaSetDevice(ctx->cfg.device));
  free_run_inputs(ctx);
  ctx->run_T = T; ctx->run_K = K;
  ctx->Q_pages = in->Q_pages; ctx->x0_cols = in->x0_lin_cols;
  int rc;
  // odometry -> row-major [T-1][n_odo]
  {
    std::vector<double> odo((size_t)std::max(T - 1, 1) * ctx->n_odo, 0.0);
    for (int t = 0; t + 1 < T; ++t)
      for (int j = 0; j < ctx->n_odo; ++j) odo[(size_t)t * ctx->n_odo + j] = in->odometry[t + (size_t)j * in->odo_rows];
    RB_ALLOC(ctx->d_odo, odo.size());
    if ((rc = rb_h2d(ctx, ctx->d_odo, odo.data(), odo.size() * 8))) return rc;
  }
  {
    ctx->h_y.assign((size_t)T * d, 0.0);
    for (int t = 0; t < T; ++t)
      for (int j = 0; j < d; ++j) ctx->h_y[(size_t)t * d + j] = in->y[t + (size_t)j * T];
    RB_ALLOC(ctx->d_y, ctx->h_y.size());
    if ((rc = rb_h2d(ctx, ctx->d_y, ctx->h_y.data(), ctx->h_y.size() * 8))) return rc;
  }
  {
    const size_t qn = (size_t)ctx->nw * ctx->nw * in->Q_pages;
    RB_ALLOC(ctx->d_Q, qn);
    if ((rc = rb_h2d(ctx, ctx->d_Q, in->Q, qn * 8))) return rc;
  }
  RB_ALLOC(ctx->d_R, (size_t)d * d);
  if ((rc = rb_h2d(ctx, ctx->d_R, in->R, sizeof(double) * d * d))) return rc;
  ctx->h_R.assign(in->R, in->R + (size_t)d * d);
  RB_ALLOC(ctx->d_P0, (size_t)M * M);
  if ((rc = rb_h2d(ctx, ctx->d_P0, in->P0_lin, sizeof(double) * M * M))) return rc;
  RB_ALLOC(ctx->d_x0lin, (size_t)M * in->x0_lin_cols);
  if ((rc = rb_h2d(ctx, ctx->d_x0lin, in->x0_lin, sizeof(double) * M * in->x0_lin_cols))) return rc;
  ctx->h_x0n.assign(in->x0_nonLin, in->x0_nonLin + n);
  ctx->h_dt.assign(std::max(T - 1, 1), in->dt[0]);
  if (in->dt_len > 1) for (int t = 0; t + 1 < T; ++t) ctx->h_dt[t] = in->dt[t];
  ctx->have_U = in->U != nullptr; ctx->have_Z = in->Z != nullptr;
  ctx->have_forced = in->forced_ancestors != nullptr;
  if (ctx->cfg.rng_mode == RBSLAM_RNG_PHILOX) { ctx->have_U = false; ctx->have_Z = false; }
  if (ctx->have_U) {
    RB_ALLOC(ctx->d_U, (size_t)N * T * K);
    if ((rc = rb_h2d(ctx, ctx->d_U, in->U, sizeof(double) * N * T * K))) return rc;
  }
  if (ctx->have_Z) {
    RB_ALLOC(ctx->d_Z, (size_t)ctx->nz * N * T * K);
    if ((rc = rb_h2d(ctx, ctx->d_Z, in->Z, sizeof(double) * ctx->nz * N * T * K))) return rc;
  }
  if (ctx->have_forced) {
    for (size_t q = 0; q < (size_t)N * T * K; ++q) {   // used as indices on the device: check them here
      if (q % ((size_t)N * T) < (size_t)N) continue;    // column t = 0 is unused
      if (in->forced_ancestors[q] < 0 || in->forced_ancestors[q] >= N)
        return ctx->fail(RBSLAM_EARG, "forced_ancestors out of range [0, N)");
    }
    RB_ALLOC(ctx->d_forced, (size_t)N * T * K);
    if ((rc = rb_h2d(ctx, ctx->d_forced, in->forced_ancestors, sizeof(int) * (size_t)N * T * K))) return rc;
  }
  ctx->h_Uend.assign(K, 0.5);
  if (in->Uend) ctx->h_Uend.assign(in->Uend, in->Uend + K);
  ctx->h_forced_ak.clear();
  if (in->forced_ak) {
    for (int k = 0; k < K; ++k)
      if (in->forced_ak[k] < 0 || in->forced_ak[k] >= N) return ctx->fail(RBSLAM_EARG, "forced_ak out of range [0, N)");
    ctx->h_forced_ak.assign(in->forced_ak, in->forced_ak + K);
  }
  return RBSLAM_OK;
}

// (re-)initialise the particle state of one sweep (src/particleFilter.m:55-67)
int rb_init_state(rbslam_ctx *ctx, bool info_form) {
  const int N = ctx->N, M = ctx->M, n = ctx->n;
  CK(cudaSetDevice(ctx->cfg.device));
  ctx->cs = 0; ctx->cx = 0; ctx->t = 0;
  if (ctx->kpath == 1) {
    ctx->cg = 0; ctx->pending = false;
    CK(cudaMemsetAsync(ctx->d_G4[0], 0, sizeof(double) * (size_t)N * ctx->ld * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_KS4[0], 0, sizeof(double) * (size_t)N * ctx->ld * 4, ctx->stream));
  }
  k_fill_int_iota<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_slot[0], N);
  dim3 g(M, std::min(N, 4 * ctx->num_sms));
  if (ctx->pt) k_init_slabs_pt<<<dim3(256, std::min(N, 4 * ctx->num_sms)), 64, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ctx->ld, M, N, ctx->d_P0);
  else k_init_slabs<<<g, 128, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ctx->ld, M, N, ctx->d_P0, 0);
  const size_t tot = (size_t)M * N;
  k_init_xl<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_xl[0], M, N, ctx->d_x0lin, ctx->x0_cols);
  // xn(:, i) = x0_nonLin for all i -> history slot 0
  double *x0d = ctx->d_scratch;
  CK(cudaMemcpyAsync(x0d, ctx->h_x0n.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  k_init_xn<<<(n * N + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_Xhist, n, N, x0d);
  k_fill<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_w, (size_t)N, 1.0 / N);
  ctx->launches += 5;
  if (info_form) {
    if (!ctx->d_Imat) return ctx->fail(RBSLAM_EARG, "context was created without information_form");
    int rc = rb_info_init(ctx);
    if (rc) return rc;
  }
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

// ---------------------------------------------------------------------------
// one time step of the filter recursion (src/particleFilter.m:100-204)
// ---------------------------------------------------------------------------
static void pick_large_launch(int ld, int &threads, int &R2) {
  const int npairs = ld / 2;
  double best = -1.0;
  threads = 256; R2 = 4;
  for (int r2 = 1; r2 <= 4; ++r2) {
    int th = ((npairs + r2 - 1) / r2 + 31) / 32 * 32;
    if (th > 256 || th < 64) continue;
    const double eff = (double)npairs / ((double)th * r2);
    if (eff > best + 1e-9 || (std::fabs(eff - best) < 1e-9 && th > threads)) { best = eff; threads = th; R2 = r2; }
  }
}

template <int D>
static int launch_large(rbslam_ctx *ctx, const KalmanArgs &a) {
  const int M = ctx->M, N = ctx->N, ld = ctx->ld;
  const int nsplit = ctx->nsplit, cw = ctx->cw;
  int threads, R2;
  pick_large_launch(ld, threads, R2);
  dim3 g(N, nsplit);
#define RB_LAUNCH_PH(R2V)                                                                        \
  k_ph<D, R2V><<<g, threads, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ld, M, ctx->d_src_slot, a.H,  \
                                               a.ldh, nsplit, cw, ctx->d_PHpart)
  switch (R2) { case 1: RB_LAUNCH_PH(1); break; case 2: RB_LAUNCH_PH(2); break;
                case 3: RB_LAUNCH_PH(3); break; default: RB_LAUNCH_PH(4); break; }
  k_innov<D><<<N, 128, sizeof(double) * D * ld, ctx->stream>>>(a, nsplit, ctx->d_PHpart, ctx->d_G, ctx->d_KS);
#define RB_LAUNCH_DD(R2V, LIST, CNT)                                                              \
  k_downdate<D, R2V><<<g, threads, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ld, M, ctx->d_src_slot, \
                                                     a.dst_slot, LIST, CNT, ctx->d_G, ctx->d_KS, cw)
  for (int phase = 0; phase < 2; ++phase) {
    const int *list = phase == 0 ? ctx->d_listA : ctx->d_listB;
    const int *cnt = ctx->d_counts + phase;
    if (phase == 0 && ctx->t == 0) continue;  // nothing copies at the first step
    switch (R2) { case 1: RB_LAUNCH_DD(1, list, cnt); break; case 2: RB_LAUNCH_DD(2, list, cnt); break;
                  case 3: RB_LAUNCH_DD(3, list, cnt); break; default: RB_LAUNCH_DD(4, list, cnt); break; }
    ctx->launches += 1;
  }
  ctx->launches += 2;
  return RBSLAM_OK;
}

// where family sources live (single GPU: all local; sharded with fused migration: see SrcTab)
static SrcTab make_src_tab(const rbslam_ctx *ctx) {
  SrcTab t;
  if (ctx->st_nloc > 0) {
    const double *const *tab = ctx->d_peer_tab;
    t.nloc = ctx->st_nloc; t.fetch = ctx->st_fetch;
    t.P = tab + SH_P * (RB_MAXW + 1);
    t.G4 = tab + (ctx->cg ? SH_G4B : SH_G4A) * (RB_MAXW + 1);
    t.KS4 = tab + (ctx->cg ? SH_KS4B : SH_KS4A) * (RB_MAXW + 1);
    t.xl = tab + (ctx->cx ? SH_XLB : SH_XLA) * (RB_MAXW + 1);
  }
  return t;
}
// carve the family-construction scratch out of d_fam: 7 arrays over the source-key space, the two
// family lists (5 arrays over the items each), counters
static void fam_layout(rbslam_ctx *ctx, const KalmanArgs &a, int cb, FamBuildArgs &fb, int *&la, int *&lb, int *&cnts) {
  const int N = ctx->N;
  const size_t NS = ctx->fam_slabs > 0 ? ctx->fam_slabs : N;
  int *fm = ctx->d_fam;
  fb.n_items = N; fb.n_slabs = (int)NS; fb.n_items_dev = nullptr;
  fb.src_slot = a.src_slot; fb.dst_slot = a.dst_slot; fb.anc = a.ai;
  fb.s_cnt = fm; fb.s_keeper = fm + NS; fb.s_cursor = fm + 2 * NS; fb.s_first = fm + 3 * NS;
  fb.s_fid = fm + 4 * NS; fb.s_xoff = fm + 5 * NS; fb.s_xfam = fm + 6 * NS;
  lb = fm + 7 * NS; la = lb + 5 * (size_t)N; cnts = la + 5 * (size_t)N;
  fb.fb_src = lb; fb.fb_anc = lb + N; fb.fb_first = lb + 2 * (size_t)N; fb.fb_cnt = lb + 3 * (size_t)N;
  fb.fb_child = lb + 4 * (size_t)N; fb.n_fb = cnts;
  fb.fa_src = la; fb.fa_anc = la + N; fb.fa_first = la + 2 * (size_t)N; fb.fa_cnt = la + 3 * (size_t)N;
  fb.fa_child = la + 4 * (size_t)N; fb.n_fa = cnts + 1;
  fb.work_ctr = cnts + 2;
  fb.cb = cb; fb.kf = 2 * cb;
}

// streaming path: one pass per slab with the deferred downdate (kalman_stream.cuh)
// (KC columns per stage, S stages) is a tuning knob: RBSLAM_STREAM_CFG="KC,S" (d=3 only)
template <int D, int R2, int KC, int S>
static int launch_stream_cfg(rbslam_ctx *ctx, const KalmanArgs &a, bool resampled) {
  const int N = ctx->N, ld = ctx->ld;
  StreamArgs sa;
  sa.hints = ctx->stream_hints; sa.gk_by_particle = 0;
  sa.M = ctx->M; sa.ld = ld; sa.cw = ctx->cw; sa.nsplit = ctx->nsplit; sa.slab = ctx->slab;
  sa.P = ctx->d_P; sa.src_slot = a.src_slot; sa.dst_slot = a.dst_slot; sa.anc = a.ai;
  sa.G4prev = ctx->d_G4[ctx->cg]; sa.KS4prev = ctx->d_KS4[ctx->cg]; sa.H4 = a.H; sa.PHp = ctx->d_PHp;
  const size_t smem = sizeof(double) * (size_t)S * ((size_t)KC * ld + 8 * KC);
  auto kern = k_stream_pass<D, D, R2, KC, S>;
  RB_OPTIN_SMEM(kern, ctx->smem_optin - 1024);
  if (smem > ctx->smem_optin - 1024) return ctx->fail(RBSLAM_EARG, "streaming stage ring does not fit shared memory");
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (ctx->smem_optin) / (smem + 2048)));
  if (ctx->stream_ctas_per_sm > 0) per_sm = ctx->stream_ctas_per_sm;
  const int grid = std::min(N * ctx->nsplit, per_sm * ctx->num_sms);
  if (ctx->use_fam) {
    // sibling fusion: families of offspring share one read of the ancestor slab
    FamBuildArgs fb;
    int *la, *lb, *cnts;
    fam_layout(ctx, a, RB_CB, fb, la, lb, cnts);
    sa.st = make_src_tab(ctx);
    auto fkern = k_stream_fam<D, R2, KC, S, RB_CB>;
    const size_t fsmem = sizeof(double) * (size_t)S * ((size_t)KC * ld + 4 * KC * (1 + RB_CB));
    RB_OPTIN_SMEM(fkern, ctx->smem_optin - 1024);
    if (fsmem > ctx->smem_optin - 1024) return ctx->fail(RBSLAM_EARG, "streaming stage ring does not fit shared memory");
    const int fgrid = std::min(N * ctx->nsplit, ctx->num_sms);
    const int ngrp = ctx->item_group ? ctx->stream_groups : 1;   // no group tags: one pass over all families
    for (int grp = 0; grp < ngrp; ++grp) {
      if (grp > 0 && ctx->group_hook) {   // sharded engine: migrants landed, peers done reading
        int rch = ctx->group_hook(ctx, grp);
        if (rch) return rch;
      }
      fb.item_group = ngrp > 1 ? ctx->item_group : nullptr;
      fb.group = grp;
      k_build_families<<<1, 1024, 0, ctx->stream>>>(fb);
      ctx->launches += 1;
      for (int phase = 0; phase < 2; ++phase) {
        if (phase == 0 && !resampled) continue;   // surplus families exist only after resampling
        FamLists fl;
        const int *base = phase == 0 ? la : lb;
        fl.n_fam = cnts + (phase == 0 ? 1 : 0);
        fl.work_counter = cnts + 2 + phase;
        fl.src = base; fl.anc = base + N; fl.first = base + 2 * (size_t)N; fl.cnt = base + 3 * (size_t)N;
        fl.child = base + 4 * (size_t)N;
        fkern<<<fgrid, RB_STREAM_THREADS, fsmem, ctx->stream>>>(sa, fl);
        ctx->launches += 1;
      }
    }
  } else {
  for (int grp = 0; grp < ctx->stream_groups; ++grp) {
    if (grp > 0 && ctx->group_hook) {
      int rch = ctx->group_hook(ctx, grp);
      if (rch) return rch;
    }
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 0 && !resampled) continue;
      kern<<<grid, RB_STREAM_THREADS, smem, ctx->stream>>>(
          sa, (phase == 0 ? ctx->d_listA : ctx->d_listB) + ctx->group_off[grp][phase],
          ctx->d_counts + 2 * grp + phase, ctx->group_off_dev[grp][phase]);
      ctx->launches += 1;
    }
  }
  }
  Innov4Args ia;
  ia.st = make_src_tab(ctx);
  ia.N = N; ia.M = ctx->M; ia.ld = ld; ia.nsplit = ctx->nsplit; ia.PHp = ctx->d_PHp; ia.H4 = a.H;
  ia.xl_old = a.xl_old; ia.anc = a.ai; ia.xl_new = a.xl_new;
  ia.G4new = ctx->d_G4[1 - ctx->cg]; ia.KS4new = ctx->d_KS4[1 - ctx->cg];
  ia.y_t = a.y_t; ia.R = a.R; ia.jitter = a.jitter; ia.logw = a.logw; ia.status = a.status; ia.t = a.t;
  k_innov4<D><<<N, 128, sizeof(double) * 4 * ld, ctx->stream>>>(ia);
  ctx->launches += 1;
  ctx->cg ^= 1;
  ctx->pending = true;
  return RBSLAM_OK;
}
template <int D, int R2>
static int launch_stream_r(rbslam_ctx *ctx, const KalmanArgs &a, bool resampled) {
  // KC = 8 columns per stage, S = 2 stages measured best on C4 (profiles/tuning_r1.md section 2)
  return launch_stream_cfg<D, R2, 8, 2>(ctx, a, resampled);
}
template <int D>
static int launch_stream(rbslam_ctx *ctx, const KalmanArgs &a, bool resampled) {
  const int npairs = ctx->ld / 2;
  const int R2 = (npairs + RB_STREAM_THREADS - 1) / RB_STREAM_THREADS;
  switch (R2) {
    case 1: return launch_stream_r<D, 1>(ctx, a, resampled);
    case 2: return launch_stream_r<D, 2>(ctx, a, resampled);
    case 3: return launch_stream_r<D, 3>(ctx, a, resampled);
    default: return launch_stream_r<D, 4>(ctx, a, resampled);
  }
}

// packed symmetric slabs (kalman_variant 7): k_stream_fam_pt on the fp64 tensor cores
template <int NW, int MAXQ>
static int launch_pt_q(rbslam_ctx *ctx, const KalmanArgs &a, bool resampled) {
  constexpr int CB = RB_PT_CB;
  const int N = ctx->N, ld = ctx->ld;
  PtArgs pa;
  pa.M = ctx->M; pa.ld = ld; pa.nb = ld / 8; pa.nsplit = ctx->nsplit; pa.ts = ctx->pt_ts; pa.ns = ctx->pt_ns;
  for (int q = 0; q <= RB_PT_MAXSPLIT; ++q) pa.psplit[q] = ctx->pt_psplit[std::min(q, ctx->nsplit)];
  pa.slab = ctx->slab; pa.P = ctx->d_P; pa.dst_slot = a.dst_slot;
  pa.G4prev = ctx->d_G4[ctx->cg]; pa.KS4prev = ctx->d_KS4[ctx->cg]; pa.H4 = a.H; pa.PHp = ctx->d_PHp;
  FamBuildArgs fb;
  int *la, *lb, *cnts;
  fam_layout(ctx, a, CB, fb, la, lb, cnts);
  pa.st = make_src_tab(ctx);
  {
    // sharded filter, fused migration: RBSLAM_MIGRANTS_FIRST=1 starts with the families whose slab comes from a
    // peer (highest source keys).  Measured SLOWER on 4 GPUs (weak 15.31 vs 15.11 ms/step, strong 4.57 vs 4.45:
    // every rank would open its pass with NVLink reads), so the default keeps them last (profiles/tuning_r2.md 5)
    static const int mig_first = [] { const char *e = getenv("RBSLAM_MIGRANTS_FIRST"); return (e && atoi(e)) ? 1 : 0; }();
    pa.rev = (ctx->st_nloc > 0 && mig_first) ? 1 : 0;
  }
  auto fkern = k_stream_fam_pt<NW, MAXQ>;
  const size_t fsmem = pt_smem_bytes(ld, ctx->pt_ts, ctx->pt_ns, NW);
  RB_OPTIN_SMEM(fkern, ctx->smem_optin - 1024);
  const int fgrid = std::min(N * ctx->nsplit, ctx->num_sms);
  const int ngrp = ctx->item_group ? ctx->stream_groups : 1;
  for (int grp = 0; grp < ngrp; ++grp) {
    if (grp > 0 && ctx->group_hook) {
      int rch = ctx->group_hook(ctx, grp);
      if (rch) return rch;
    }
    fb.item_group = ngrp > 1 ? ctx->item_group : nullptr;
    fb.group = grp;
    k_build_families<<<1, 1024, 0, ctx->stream>>>(fb);
    ctx->launches += 1;
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 0 && !resampled) continue;   // surplus families exist only after resampling
      FamLists fl;
      const int *base = phase == 0 ? la : lb;
      fl.n_fam = cnts + (phase == 0 ? 1 : 0);
      fl.work_counter = cnts + 2 + phase;
      fl.src = base; fl.anc = base + N; fl.first = base + 2 * (size_t)N; fl.cnt = base + 3 * (size_t)N;
      fl.child = base + 4 * (size_t)N;
      fkern<<<fgrid, 32 * (NW + 1), fsmem, ctx->stream>>>(pa, fl);
      ctx->launches += 1;
    }
  }
  Innov4Args ia;
  ia.st = make_src_tab(ctx);
  ia.N = N; ia.M = ctx->M; ia.ld = ld; ia.nsplit = ctx->nsplit + 1; ia.PHp = ctx->d_PHp; ia.H4 = a.H;
  ia.xl_old = a.xl_old; ia.anc = a.ai; ia.xl_new = a.xl_new;
  ia.G4new = ctx->d_G4[1 - ctx->cg]; ia.KS4new = ctx->d_KS4[1 - ctx->cg];
  ia.y_t = a.y_t; ia.R = a.R; ia.jitter = a.jitter; ia.logw = a.logw; ia.status = a.status; ia.t = a.t;
  ia.G4prev = ctx->d_G4[ctx->cg]; ia.KS4prev = ctx->d_KS4[ctx->cg];   // the products used the slab before its downdate
  const size_t ism = sizeof(double) * 4 * ld;
  switch (ctx->d) {
    case 1: k_innov4<1><<<N, 128, ism, ctx->stream>>>(ia); break;
    case 2: k_innov4<2><<<N, 128, ism, ctx->stream>>>(ia); break;
    case 3: k_innov4<3><<<N, 128, ism, ctx->stream>>>(ia); break;
    default: k_innov4<4><<<N, 128, ism, ctx->stream>>>(ia); break;
  }
  ctx->launches += 1;
  ctx->cg ^= 1;
  ctx->pending = true;
  return RBSLAM_OK;
}
static int launch_pt(rbslam_ctx *ctx, const KalmanArgs &a, bool resampled) {
  // MAXQ = row blocks per consumer warp: NW warps share ld / 8 blocks
  const int nb = ctx->ld / 8;
  return nb <= 75 ? launch_pt_q<15, 5>(ctx, a, resampled) : launch_pt_q<15, 9>(ctx, a, resampled);
}

// apply the deferred downdate to every slab (before the state is read out as a whole)
int rb_flush_pending(rbslam_ctx *ctx) {
  if (ctx->kpath != 1 || !ctx->pending) return RBSLAM_OK;
  const int N = ctx->N, M = ctx->M, ld = ctx->ld;
  dim3 g(std::min(M, 64), N);
  const double *G4 = ctx->d_G4[ctx->cg], *KS4 = ctx->d_KS4[ctx->cg];
  if (ctx->pt) {
    k_apply_pending_pt<<<dim3(std::min(ld / 8, 64), N), 256, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ld, ctx->d_slot[ctx->cs], G4, KS4);
  } else
  switch (ctx->d) {
    case 1: k_apply_pending<1><<<g, 256, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ld, M, ctx->d_slot[ctx->cs], G4, KS4); break;
    case 2: k_apply_pending<2><<<g, 256, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ld, M, ctx->d_slot[ctx->cs], G4, KS4); break;
    case 3: k_apply_pending<3><<<g, 256, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ld, M, ctx->d_slot[ctx->cs], G4, KS4); break;
    default: k_apply_pending<4><<<g, 256, 0, ctx->stream>>>(ctx->d_P, ctx->slab, ld, M, ctx->d_slot[ctx->cs], G4, KS4); break;
  }
  CK(cudaMemsetAsync(ctx->d_G4[ctx->cg], 0, sizeof(double) * (size_t)N * ld * 4, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_KS4[ctx->cg], 0, sizeof(double) * (size_t)N * ld * 4, ctx->stream));
  ctx->launches += 1;
  ctx->pending = false;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

// Kalman phase for the particles in the current plan; xl: cur -> 1-cur
int rb_kalman_phase(rbslam_ctx *ctx, const double *y_t_dev, bool resampled) {
  const int N = ctx->N, M = ctx->M, d = ctx->d;
  KalmanArgs a;
  a.N = N; a.M = M; a.d = d; a.ld = ctx->ld; a.ldh = ctx->ldh; a.slab = ctx->slab;
  a.P = ctx->d_P;
  a.src_slot = ctx->d_src_slot;
  a.dst_slot = ctx->d_slot[ctx->cs];
  a.xl_old = ctx->d_xl[ctx->cx];
  a.ai = ctx->anc_override ? ctx->anc_override
                           : (resampled ? ctx->d_Ahist + (size_t)(ctx->t % ctx->T_hist) * N : nullptr);
  a.xl_new = ctx->d_xl[1 - ctx->cx];
  a.H = ctx->d_H;
  a.yhat = ctx->mc.family == FAM_SPARSE_VISUAL2D ? ctx->d_yhat : nullptr;
  a.y_t = y_t_dev;
  a.R = ctx->d_R;
  a.jitter = ctx->jitter;
  a.logw = ctx->d_logw;
  a.status = ctx->d_status;
  a.t = ctx->t;
  if (ctx->kpath == 0) {
    const size_t smem = kalman_small_smem(M, d);
    const int grid = std::min(N, 8 * ctx->num_sms);
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 0 && !resampled) continue;
      k_kalman_small<<<grid, 256, smem, ctx->stream>>>(a, phase == 0 ? ctx->d_listA : ctx->d_listB,
                                                       ctx->d_counts + phase);
      ctx->launches += 1;
    }
  } else if (ctx->kpath == 1 && ctx->pt) {
    int rc = launch_pt(ctx, a, resampled);
    if (rc) return rc;
  } else if (ctx->kpath == 1) {
    int rc;
    switch (d) {
      case 1: rc = launch_stream<1>(ctx, a, resampled); break;
      case 2: rc = launch_stream<2>(ctx, a, resampled); break;
      case 3: rc = launch_stream<3>(ctx, a, resampled); break;
      default: rc = launch_stream<4>(ctx, a, resampled); break;
    }
    if (rc) return rc;
  } else {
    switch (d) {
      case 1: launch_large<1>(ctx, a); break;
      case 2: launch_large<2>(ctx, a); break;
      case 3: launch_large<3>(ctx, a); break;
      case 4: launch_large<4>(ctx, a); break;
      default: return ctx->fail(RBSLAM_EARG, "large-M Kalman path supports d<=4");
    }
  }
  CK(cudaGetLastError());
  ctx->cx ^= 1;   // updated means now live in d_xl[cx]
  return RBSLAM_OK;
}

// resample + plan + propagate for step t>=1.  n_draws < N leaves the last particle
// to the caller (smoother reference particle).
int rb_resample_phase(rbslam_ctx *ctx, int n_draws) {
  const int N = ctx->N, t = ctx->t;
  const int tb = t % ctx->T_hist;
  int *ai = ctx->d_Ahist + (size_t)tb * N;
  const size_t soff = ((size_t)ctx->sweep * ctx->run_T + t) * N;
  RngSrc rs;
  rs.U = ctx->have_U ? ctx->d_U + soff : nullptr;
  rs.seed = ctx->cfg.seed; rs.sweep = ctx->sweep; rs.t = t;
  const int *forced = ctx->have_forced ? ctx->d_forced + soff : nullptr;
  size_t smem = sizeof(double) * (size_t)N;
  smem = std::min(smem, std::min(ctx->smem_resample_max, (size_t)(96 << 10)));
  if (N >= 4096 && n_draws > 1024) {   // scan in one CTA, draws over the grid (N = 10^4: 0.06 ms against 0.17 ms in one CTA)
    // fast path first: parallel prefix sum + draws that prove themselves independent of the rounding order; the
    // exact pair runs only if a draw could not (step_kernels.cuh).  RBSLAM_EXACT_SCAN=1: exact pair always.
    const int fast = rb_fast_scan() ? 1 : 0;
    if (fast) {
      k_scan_approx<<<1, 1024, 0, ctx->stream>>>(N, ctx->d_w, ctx->d_wc, ctx->d_status);
      k_search_checked<<<(n_draws + 255) / 256, 256, 0, ctx->stream>>>(N, 0, n_draws, ctx->d_wc, rs, forced, ai, ctx->d_status);
      ctx->launches += 2;
    }
    k_resample<<<1, 1024, smem, ctx->stream>>>(N, 0, -1, ctx->d_w, ctx->d_wc, rs, forced, ai, ctx->d_status, fast);
    k_resample_search<<<(n_draws + 255) / 256, 256, 0, ctx->stream>>>(N, 0, n_draws, ctx->d_wc, rs, forced, ai, ctx->d_status, fast);
    ctx->launches += 2;
  } else {
    k_resample<<<1, 1024, smem, ctx->stream>>>(N, 0, n_draws, ctx->d_w, ctx->d_wc, rs, forced, ai, ctx->d_status);
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

int rb_plan_phase(rbslam_ctx *ctx) {
  const int N = ctx->N, tb = ctx->t % ctx->T_hist;
  const int *ai = ctx->d_Ahist + (size_t)tb * N;
  k_plan_slots<<<1, 1024, 0, ctx->stream>>>(N, ai, ctx->d_slot[ctx->cs], ctx->d_slot[1 - ctx->cs],
                                            ctx->d_src_slot, ctx->d_first_child, ctx->d_free_list,
                                            ctx->d_listA, ctx->d_listB, ctx->d_counts);
  ctx->cs ^= 1;   // d_slot[cs] is now the new logical->physical map
  ctx->launches += 1;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

int rb_propagate_phase(rbslam_ctx *ctx, int n_prop) {
  const int N = ctx->N, t = ctx->t, n = ctx->n;
  const int tb = t % ctx->T_hist, tp = (t - 1) % ctx->T_hist;
  const int *ai = ctx->d_Ahist + (size_t)tb * N;
  NormalSrc ns;
  ns.Z = ctx->have_Z ? ctx->d_Z + ((size_t)ctx->sweep * ctx->run_T + t) * N * ctx->nz : nullptr;
  ns.seed = ctx->cfg.seed; ns.sweep = ctx->sweep; ns.t = t;
  const double *Qp = ctx->d_Q + (ctx->Q_pages > 1 ? (size_t)(t - 1) * ctx->nw * ctx->nw : 0);
  k_propagate<<<(n_prop + 127) / 128, 128, 0, ctx->stream>>>(
      ctx->mc, N, n_prop, ctx->d_Xhist + (size_t)tp * N * n, ai, ctx->d_odo + (size_t)(t - 1) * ctx->n_odo,
      ctx->h_dt[t - 1], Qp, ns, ctx->d_Xhist + (size_t)tb * N * n);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

int rb_meas_phase(rbslam_ctx *ctx, bool resampled) {
  const int N = ctx->N, tb = ctx->t % ctx->T_hist;
  const double *xn = ctx->d_Xhist + (size_t)tb * N * ctx->n;
  const int *ai = resampled ? ctx->d_Ahist + (size_t)tb * N : nullptr;
  k_meas<<<N, 128, 0, ctx->stream>>>(ctx->mc, N, xn, ctx->d_xl[ctx->cx], ctx->M, ai, ctx->d_H,
                                     ctx->hs_p, ctx->hs_a, ctx->hs_c, ctx->ld, ctx->d_yhat);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

// w = exp(logw - lse), first arg-max, traj_max / traj_mean: one CTA with a fixed tree up to 32 767 weights,
// fixed chunks of 4096 (one CTA each, combined in chunk order) above -- both deterministic for a given N
int rb_normalize(rbslam_ctx *ctx, int N, int n, const double *logw, double *w, const double *xn, double *traj_max_t,
                 double *traj_mean_t, int *iw_max, double *logw_hist_t, double *w_hist_t) {
  const int nchunk = (N + RB_NCHUNK - 1) / RB_NCHUNK;
  if (N < 32768 || nchunk > RB_NCHUNK_MAX || n > 8) {
    k_normalize<<<1, 1024, 0, ctx->stream>>>(N, n, logw, w, xn, traj_max_t, traj_mean_t, iw_max, logw_hist_t, w_hist_t);
    ctx->launches += 1;
    CK(cudaGetLastError());
    return RBSLAM_OK;
  }
  if (!ctx->d_normws) {
    RB_ALLOC(ctx->d_normws, (size_t)RB_NCHUNK_MAX * 12);
  }
  NormWs ws;
  ws.pmax = ctx->d_normws; ws.psum = ws.pmax + RB_NCHUNK_MAX; ws.pbest = ws.psum + RB_NCHUNK_MAX;
  ws.pmean = ws.pbest + RB_NCHUNK_MAX;
  ws.pidx = reinterpret_cast<int *>(ws.pmean + (size_t)RB_NCHUNK_MAX * 8);
  k_norm_max<<<nchunk, 1024, 0, ctx->stream>>>(N, logw, ws);
  k_norm_sum<<<nchunk, 1024, 0, ctx->stream>>>(N, logw, ws);
  k_norm_write<<<nchunk, 1024, 0, ctx->stream>>>(N, n, logw, w, xn, ws, logw_hist_t, w_hist_t);
  k_norm_final<<<1, 32, 0, ctx->stream>>>(N, n, nchunk, xn, ws, traj_max_t, traj_mean_t, iw_max);
  ctx->launches += 4;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

int rb_normalize_phase(rbslam_ctx *ctx) {
  const int N = ctx->N, t = ctx->t, n = ctx->n, tb = t % ctx->T_hist;
  return rb_normalize(ctx, N, n, ctx->d_logw, ctx->d_w, ctx->d_Xhist + (size_t)tb * N * n, ctx->d_traj_max + (size_t)t * n,
                      ctx->d_traj_mean + (size_t)t * n, ctx->d_iwmax + t,
                      ctx->d_logw_hist ? ctx->d_logw_hist + (size_t)t * N : nullptr,
                      ctx->d_w_hist ? ctx->d_w_hist + (size_t)t * N : nullptr);
}

static int filter_step_impl(rbslam_ctx *ctx) {
  if (!ctx->running) return ctx->fail(RBSLAM_EARG, "filter_step without filter_begin");
  if (ctx->t >= ctx->run_T) return ctx->fail(RBSLAM_EARG, "all time steps already processed");
  CK(cudaSetDevice(ctx->cfg.device));
  const int N = ctx->N, t = ctx->t;
  int rc;
  const bool resampled = t > 0;
  if (resampled) {
    rb_phase_begin(ctx, RB_PH_RESAMPLE);
    if ((rc = rb_resample_phase(ctx, N))) return rc;
    if ((rc = rb_plan_phase(ctx))) return rc;
    rb_phase_end(ctx);
    rb_phase_begin(ctx, RB_PH_PROPAGATE);
    if ((rc = rb_propagate_phase(ctx, N))) return rc;
    rb_phase_end(ctx);
  } else {
    k_plan_identity<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, ctx->d_slot[ctx->cs], ctx->d_src_slot,
                                                             ctx->d_listB, ctx->d_counts);
    ctx->launches += 1;
  }
  rb_phase_begin(ctx, RB_PH_MEAS);
  if ((rc = rb_meas_phase(ctx, resampled))) return rc;
  rb_phase_end(ctx);
  rb_phase_begin(ctx, RB_PH_KALMAN);
  rc = rb_kalman_phase(ctx, ctx->d_y + (size_t)t * ctx->d, resampled);
  rb_phase_end(ctx);
  if (rc) return rc;
  rb_phase_begin(ctx, RB_PH_NORMALIZE);
  if ((rc = rb_normalize_phase(ctx))) return rc;
  rb_phase_end(ctx);
  ctx->t += 1;
  if (ctx->step_fn) {
    if (!ctx->d_yhattraj) {
      RB_ALLOC(ctx->d_yhattraj, (size_t)ctx->T * ctx->d);
    }
    const bool sparse = ctx->mc.family == FAM_SPARSE_VISUAL2D;
    k_yhat_max<<<1, 32 * ctx->d, 0, ctx->stream>>>(
        ctx->M, ctx->d, ctx->d_iwmax + t, ctx->d_H, ctx->hs_p, ctx->hs_a, ctx->hs_c, ctx->d_xl[1 - ctx->cx],
        resampled ? ctx->d_Ahist + (size_t)(t % ctx->T_hist) * N : nullptr, sparse ? ctx->d_yhat : nullptr,
        ctx->d_yhattraj + (size_t)t * ctx->d);
    ctx->launches += 1;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->step_fn(ctx->step_user, ctx->sweep, t);
  }
  return RBSLAM_OK;
}

// group leader: run `fn` on every shard (its own device current), first error wins
template <typename F>
static int group_each(rbslam_ctx *ctx, F fn) {
  for (rbslam_ctx *m : ctx->group) {
    if (cudaSetDevice(m->cfg.device) != cudaSuccess) return ctx->fail(RBSLAM_ECUDA, "cudaSetDevice failed");
    int rc = fn(m);
    if (rc) { if (m != ctx) ctx->err = m->err; return rc; }
  }
  return RBSLAM_OK;
}

extern "C" int rbslam_filter_begin(rbslam_ctx *ctx, const rbslam_inputs *in) {
  if (!ctx) return RBSLAM_EARG;
  if (!ctx->group.empty()) {
    int rc = group_each(ctx, [&](rbslam_ctx *m) { return rb_shard_begin(m, in, 1); });
    if (rc) return rc;
    return group_each(ctx, [&](rbslam_ctx *m) { return rb_shard_begin(m, in, 2); });
  }
  if (ctx->shard_ws) return rb_shard_begin(ctx, in);
  int rc = rb_upload_inputs(ctx, in, 1);
  if (rc) return rc;
  ctx->jitter = 1e-3;   // src/particleFilter.m:89
  ctx->sweep = 0;
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(DevStatus), ctx->stream));
  if ((rc = rb_init_state(ctx, false))) return rc;
  ctx->running = true;
  return RBSLAM_OK;
}

extern "C" int rbslam_filter_step(rbslam_ctx *ctx) {
  if (!ctx) return RBSLAM_EARG;
  if (!ctx->group.empty()) return group_each(ctx, [&](rbslam_ctx *m) { return rb_shard_step(m); });
  if (ctx->shard_ws) return rb_shard_step(ctx);
  return filter_step_impl(ctx);
}

int rb_enable_taps(rbslam_ctx *ctx, bool logw, bool w) {
  const size_t cnt = (size_t)ctx->N * ctx->run_T;
  if (logw && !ctx->d_logw_hist) RB_ALLOC(ctx->d_logw_hist, cnt);
  if (w && !ctx->d_w_hist) RB_ALLOC(ctx->d_w_hist, cnt);
  return RBSLAM_OK;
}

extern "C" int rbslam_filter_end(rbslam_ctx *ctx, rbslam_filter_outputs *out) {
  if (!ctx || !out) return RBSLAM_EARG;
  if (!ctx->group.empty()) {
    int rc;
    for (int phase = 1; phase <= 3; ++phase)
      if ((rc = group_each(ctx, [&](rbslam_ctx *m) { return rb_shard_end(m, out, phase); }))) return rc;
    return RBSLAM_OK;
  }
  if (ctx->shard_ws) return rb_shard_end(ctx, out);
  if (!ctx->running) return ctx->fail(RBSLAM_EARG, "filter_end without filter_begin");
  CK(cudaSetDevice(ctx->cfg.device));
  const int N = ctx->N, M = ctx->M, n = ctx->n, T = ctx->t;
  if (T < 1) return ctx->fail(RBSLAM_EARG, "no step was run");
  int rc = rb_check_status(ctx);
  if (rc) return rc;
  const int *iw = ctx->d_iwmax + (T - 1);
  double *means = ctx->d_scratch;                 // [2M]
  double *Pmax = ctx->d_scratch + 2 * (size_t)M + 64;
  double *Pmean = Pmax + (size_t)M * M;
  const double *xl = ctx->d_xl[ctx->cx];
  k_final_means<<<(M + 127) / 128, 128, 0, ctx->stream>>>(N, M, xl, ctx->d_w, iw, means);
  ctx->launches += 1;
  if (out->P_max || out->P_mean) {
    k_final_cov<<<M, 128, 0, ctx->stream>>>(N, M, ctx->ld, ctx->slab, ctx->d_P, ctx->d_slot[ctx->cs], xl,
                                            ctx->d_w, iw, means, Pmax, Pmean,
                                            (ctx->kpath == 1 && ctx->pending) ? ctx->d_G4[ctx->cg] : nullptr,
                                            (ctx->kpath == 1 && ctx->pending) ? ctx->d_KS4[ctx->cg] : nullptr,
                                            ctx->layout);
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  if (out->traj_max && (rc = rb_d2h(ctx, out->traj_max, ctx->d_traj_max, sizeof(double) * n * T))) return rc;
  if (out->traj_mean && (rc = rb_d2h(ctx, out->traj_mean, ctx->d_traj_mean, sizeof(double) * n * T))) return rc;
  if (out->xl_max && (rc = rb_d2h(ctx, out->xl_max, means, sizeof(double) * M))) return rc;
  if (out->xl_mean && (rc = rb_d2h(ctx, out->xl_mean, means + M, sizeof(double) * M))) return rc;
  if (out->P_max && (rc = rb_d2h(ctx, out->P_max, Pmax, sizeof(double) * M * M))) return rc;
  if (out->P_mean && (rc = rb_d2h(ctx, out->P_mean, Pmean, sizeof(double) * M * M))) return rc;
  if (out->traj_sample_iwmax || out->xn_traj) {
    if (!ctx->cfg.keep_history)
      return ctx->fail(RBSLAM_EARG, "xn_traj / traj_sample_iwmax need keep_history=1");
    if (out->traj_sample_iwmax) {
      double *tmp = nullptr;
      RB_ALLOC(tmp, (size_t)n * T);
      k_trace<<<1, 32, 0, ctx->stream>>>(N, n, T, ctx->d_Xhist, ctx->d_Ahist, iw, tmp);
      ctx->launches += 1;
      rc = rb_d2h(ctx, out->traj_sample_iwmax, tmp, sizeof(double) * n * T);
      cudaFree(tmp);
      if (rc) return rc;
    }
    if (out->xn_traj) {
      double *tmp = nullptr;
      RB_ALLOC(tmp, (size_t)n * N * T);
      k_trace<<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, n, T, ctx->d_Xhist, ctx->d_Ahist, nullptr, tmp);
      ctx->launches += 1;
      rc = rb_d2h(ctx, out->xn_traj, tmp, sizeof(double) * n * N * T);
      cudaFree(tmp);
      if (rc) return rc;
    }
  }
  if (out->ancestors) {
    if (!ctx->cfg.keep_history) return ctx->fail(RBSLAM_EARG, "ancestors need keep_history=1");
    if ((rc = rb_d2h(ctx, out->ancestors, ctx->d_Ahist, sizeof(int) * (size_t)N * T))) return rc;
  }
  if (out->logw_hist) {
    if (!ctx->d_logw_hist) return ctx->fail(RBSLAM_EARG, "logw_hist tap was not enabled before the run");
    if ((rc = rb_d2h(ctx, out->logw_hist, ctx->d_logw_hist, sizeof(double) * (size_t)N * T))) return rc;
  }
  if (out->w_hist) {
    if (!ctx->d_w_hist) return ctx->fail(RBSLAM_EARG, "w_hist tap was not enabled before the run");
    if ((rc = rb_d2h(ctx, out->w_hist, ctx->d_w_hist, sizeof(double) * (size_t)N * T))) return rc;
  }
  ctx->running = false;
  return rb_check_status(ctx);
}

extern "C" int rbslam_filter_run(rbslam_ctx *ctx, const rbslam_inputs *in, rbslam_filter_outputs *out) {
  if (!ctx || !in || !out) return RBSLAM_EARG;
  int rc = rbslam_filter_begin(ctx, in);
  if (rc) return rc;
  if (!ctx->shard_ws && (rc = rb_enable_taps(ctx, out->logw_hist != nullptr, out->w_hist != nullptr))) return rc;
  for (int t = 0; t < in->T; ++t)
    if ((rc = rbslam_filter_step(ctx))) return rc;
  return rbslam_filter_end(ctx, out);
}

// ---------------------------------------------------------------------------
// state read-back (makePlots taps, parity tests)
// ---------------------------------------------------------------------------
int rb_read_slabs(rbslam_ctx *ctx, const double *slabs, double *host) {
  const int N = ctx->N, M = ctx->M;
  if (slabs == ctx->d_P) {
    int rcf = rb_flush_pending(ctx);
    if (rcf) return rcf;
  }
  const size_t per = (size_t)M * M;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(N, (64u << 20) / (per * 8) + 1));
  double *tmp = nullptr;
  RB_ALLOC(tmp, per * chunk);
  int rc = RBSLAM_OK;
  for (int i0 = 0; i0 < N && !rc; i0 += chunk) {
    const int cnt = std::min(chunk, N - i0);
    k_pack_slabs<<<dim3(M, cnt), 128, 0, ctx->stream>>>(M, ctx->ld, ctx->slab, slabs, ctx->d_slot[ctx->cs], i0, tmp,
                                                        slabs == ctx->d_P ? ctx->layout : RB_LAYOUT_FULL);
    ctx->launches += 1;
    rc = rb_d2h(ctx, host + (size_t)i0 * per, tmp, per * cnt * 8);
  }
  cudaFree(tmp);
  return rc;
}

extern "C" int rbslam_read_particles(rbslam_ctx *ctx, double *xn, double *xl, double *P, double *logw,
                                     double *w, int32_t *ai) {
  if (!ctx) return RBSLAM_EARG;
  if (ctx->shard_ws) return ctx->fail(RBSLAM_EARG, "read_particles is not available on a sharded context (state lives on several GPUs)");
  CK(cudaSetDevice(ctx->cfg.device));
  const int N = ctx->N, M = ctx->M, n = ctx->n;
  const int tl = std::max(ctx->t - 1, 0) % ctx->T_hist;   // last completed step
  int rc;
  if (xn && (rc = rb_d2h(ctx, xn, ctx->d_Xhist + (size_t)tl * N * n, sizeof(double) * n * N))) return rc;
  if (xl && (rc = rb_d2h(ctx, xl, ctx->d_xl[ctx->cx], sizeof(double) * (size_t)M * N))) return rc;
  if (P && (rc = rb_read_slabs(ctx, ctx->d_P, P))) return rc;
  if (logw && (rc = rb_d2h(ctx, logw, ctx->d_logw, sizeof(double) * N))) return rc;
  if (w && (rc = rb_d2h(ctx, w, ctx->d_w, sizeof(double) * N))) return rc;
  if (ai && (rc = rb_d2h(ctx, ai, ctx->d_Ahist + (size_t)tl * N, sizeof(int) * N))) return rc;
  return RBSLAM_OK;
}

// the trajectory outputs as the reference holds them when it calls makePlots at step t
// (src/particleFilter.m:92-97, 215-217): columns / pages of steps not yet run are NaN (traj_max,
// traj_mean, yhattraj) or zero (xn_traj)
extern "C" int rbslam_read_trajectories(rbslam_ctx *ctx, double *traj_max, double *traj_mean, double *yhattraj,
                                        double *xn_traj) {
  if (!ctx) return RBSLAM_EARG;
  if (ctx->shard_ws) return ctx->fail(RBSLAM_EARG, "read_trajectories is not available on a sharded context");
  CK(cudaSetDevice(ctx->cfg.device));
  const int N = ctx->N, n = ctx->n, d = ctx->d, T = ctx->run_T, t = ctx->t;   // t steps are complete
  if (T < 1 || t < 1 || t > T) return ctx->fail(RBSLAM_EARG, "read_trajectories: no completed step");
  int rc;
  const double nanv = nan("");
  if (traj_max) {
    for (size_t q = (size_t)n * t; q < (size_t)n * T; ++q) traj_max[q] = nanv;
    if ((rc = rb_d2h(ctx, traj_max, ctx->d_traj_max, sizeof(double) * n * t))) return rc;
  }
  if (traj_mean) {
    for (size_t q = (size_t)n * t; q < (size_t)n * T; ++q) traj_mean[q] = nanv;
    if ((rc = rb_d2h(ctx, traj_mean, ctx->d_traj_mean, sizeof(double) * n * t))) return rc;
  }
  if (yhattraj) {
    if (!ctx->d_yhattraj) return ctx->fail(RBSLAM_EARG, "yhattraj is recorded only while a step callback is registered");
    for (size_t q = (size_t)d * t; q < (size_t)d * T; ++q) yhattraj[q] = nanv;
    if ((rc = rb_d2h(ctx, yhattraj, ctx->d_yhattraj, sizeof(double) * d * t))) return rc;
  }
  if (xn_traj) {
    if (!ctx->cfg.keep_history) return ctx->fail(RBSLAM_EARG, "xn_traj needs keep_history=1");
    for (size_t q = (size_t)n * N * t; q < (size_t)n * N * T; ++q) xn_traj[q] = 0.0;
    double *tmp = nullptr;
    RB_ALLOC(tmp, (size_t)n * N * t);
    k_trace<<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, n, t, ctx->d_Xhist, ctx->d_Ahist, nullptr, tmp);
    ctx->launches += 1;
    rc = rb_d2h(ctx, xn_traj, tmp, sizeof(double) * n * N * t);
    cudaFree(tmp);
    if (rc) return rc;
  }
  return RBSLAM_OK;
}

extern "C" int rbslam_read_information(rbslam_ctx *ctx, double *ivec, double *Imat, double *hld) {
  if (!ctx) return RBSLAM_EARG;
  if (!ctx->d_Imat) return ctx->fail(RBSLAM_EARG, "context was created without information_form");
  CK(cudaSetDevice(ctx->cfg.device));
  const int N = ctx->N, M = ctx->M;
  int rc;
  if (ivec && (rc = rb_d2h(ctx, ivec, ctx->d_ivec[ctx->cx], sizeof(double) * (size_t)M * N))) return rc;
  if (Imat && (rc = rb_read_slabs(ctx, ctx->d_Imat, Imat))) return rc;
  if (hld && (rc = rb_d2h(ctx, hld, ctx->d_hld[ctx->cx], sizeof(double) * N))) return rc;
  return RBSLAM_OK;
}
