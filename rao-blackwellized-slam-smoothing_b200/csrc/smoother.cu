// librbslam: conditional particle filter with ancestor sampling (CPF-AS), N_K sweeps.
//   form 0: covariance form      -- src/particleSmoother.m:48-366
//   form 1: information form     -- src/particleSmootherInformationForm.m:54-361
// The filter body of every sweep is the engine of engine.cu; this file adds the
// reference particle, its ancestor weights (K6 covariance form, K7 information
// form) and the information-form state update (K8).
#include <vector>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include "engine_internal.h"
#include "step_kernels.cuh"
#include "kalman_stream.cuh"
#include "dense_kernels.cuh"

using namespace rb;

struct SmootherWs {
  double *xnk = nullptr;        // [T][n] reference trajectory
  double *Hk = nullptr;         // [T][d][ldh] Jacobians along the reference trajectory (dense)
  double *lwdyn = nullptr, *paNtLog = nullptr, *paNt = nullptr, *sumlog = nullptr, *vtv = nullptr;
  double *W = nullptr, *SS = nullptr, *Lw = nullptr, *e = nullptr, *Dti = nullptr;
  double *Uend = nullptr;
  int *ak = nullptr;
  double *AI = nullptr;         // [K][T][N] tap
  int *obs_t = nullptr, *obs_l = nullptr;   // sparse: observed (ti, landmark) rows, time-major
  std::vector<int> obs_off;     // first row of step t in the obs list
  int n_obs = 0;
  size_t batch = 0, ntau_max = 0;
  std::vector<double> Rinv_h;   // host copy of inv(R): the per-step R^-1 y needs no device round trip
  // information form
  double *ImatAddt = nullptr, *ivecAddt = nullptr, *NHR4 = nullptr, *HRy = nullptr, *q2 = nullptr;
  double *Rinv = nullptr;
  double Riy_const = 0.0, half_logdetR = 0.0;
};

static SmootherWs *ws_of(rbslam_ctx *ctx) { return static_cast<SmootherWs *>(ctx->smoother_ws); }

// ---------------------------------------------------------------------------
// Replica group (rbslam_create_replicas): the smoothers on several GPUs from one process.
// A sweep is a Markov chain over time and over sweeps; what parallelises is the ancestor-weight
// evaluation of the reference trajectory -- N independent dense factorisations per time step, the
// FP64-bound part that dominates a sweep (src/particleSmoother.m:171-233, ...InformationForm.m:203-239).
// Every device holds a full replica of the particle state and runs the (HBM-bound, far cheaper)
// filter part of the sweep redundantly and bit-identically; replica r evaluates the ancestor weights
// of its block of particles only and stores them into EVERY replica's paNtLog array over peer memory
// (NVLink): the all-gather.  One host thread per replica; per step one cross-device event wait.
// ---------------------------------------------------------------------------
struct ReplicaGroup {
  int world = 1;
  std::vector<rbslam_ctx *> ctx;             // ctx[0] = leader
  cudaEvent_t ev[8][2] = {{nullptr}};        // replica r's block is stored everywhere (step parity)
  double *paNtLog[8] = {nullptr};
  std::atomic<int> arrived{0}, generation{0}, failed{0};
  // host barrier between the replicas' threads; false if a replica failed meanwhile
  bool barrier() {
    const int gen = generation.load();
    if (arrived.fetch_add(1) + 1 == world) { arrived.store(0); generation.fetch_add(1); return !failed.load(); }
    while (generation.load() == gen) {
      if (failed.load()) return false;
      std::this_thread::yield();
    }
    return !failed.load();
  }
};
static ReplicaGroup *rep_of(rbslam_ctx *ctx) { return static_cast<ReplicaGroup *>(ctx->replica_group); }

void rb_replicas_free(rbslam_ctx *leader) {
  ReplicaGroup *g = rep_of(leader);
  if (!g) return;
  for (size_t r = 1; r < g->ctx.size(); ++r) { g->ctx[r]->replica_group = nullptr; rbslam_destroy(g->ctx[r]); }
  for (auto &e : g->ev) for (auto &x : e) if (x) cudaEventDestroy(x);
  leader->replica_group = nullptr;
  delete g;
}

extern "C" int rbslam_create_replicas(rbslam_ctx **out, const rbslam_config *cfg, const int32_t *devices, int32_t n_devices) {
  if (!out || !cfg || !devices || n_devices < 1 || n_devices > 8) return RBSLAM_EARG;
  *out = nullptr;
  ReplicaGroup *g = new ReplicaGroup();
  g->world = n_devices;
  int rc = RBSLAM_OK;
  for (int r = 0; r < n_devices && rc == RBSLAM_OK; ++r) {
    rbslam_config c = *cfg;
    c.device = devices[r]; c.rank = 0; c.world = 1;
    rbslam_ctx *ctx = nullptr;
    rc = rbslam_create(&ctx, &c);
    if (rc == RBSLAM_OK) { ctx->replica_group = g; ctx->replica_rank = r; g->ctx.push_back(ctx); }
  }
  for (int r = 0; r < (int)g->ctx.size() && rc == RBSLAM_OK; ++r) {
    cudaSetDevice(devices[r]);
    for (int q = 0; q < 2; ++q)
      if (cudaEventCreateWithFlags(&g->ev[r][q], cudaEventDisableTiming) != cudaSuccess) rc = RBSLAM_ECUDA;
    for (int p = 0; p < (int)g->ctx.size() && rc == RBSLAM_OK; ++p) {
      if (devices[p] == devices[r]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[r], devices[p]);
      cudaError_t e = can ? cudaDeviceEnablePeerAccess(devices[p], 0) : cudaErrorPeerAccessUnsupported;
      if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
      if (e != cudaSuccess) rc = RBSLAM_ECUDA;
    }
  }
  if (rc != RBSLAM_OK) {
    for (rbslam_ctx *c : g->ctx) { c->replica_group = nullptr; rbslam_destroy(c); }
    for (auto &e : g->ev) for (auto &x : e) if (x) cudaEventDestroy(x);
    delete g;
    return rc;
  }
  *out = g->ctx[0];
  return RBSLAM_OK;
}

// this replica's block of particles [lo, hi) for the ancestor weights
static void rep_block(rbslam_ctx *ctx, int N, int &lo, int &hi) {
  ReplicaGroup *g = rep_of(ctx);
  lo = 0; hi = N;
  if (!g || g->world == 1) return;
  const int blk = (N + g->world - 1) / g->world;
  lo = std::min(N, ctx->replica_rank * blk); hi = std::min(N, lo + blk);
}
struct OutPtrs { double *p[8]; int n; };
static OutPtrs rep_outs(rbslam_ctx *ctx, double *own) {
  OutPtrs o; o.n = 1; o.p[0] = own;
  ReplicaGroup *g = rep_of(ctx);
  if (g && g->world > 1) { o.n = g->world; for (int r = 0; r < g->world; ++r) o.p[r] = g->paNtLog[r]; }
  return o;
}
// all-gather point: every replica's block has been stored into every paNtLog array
static int rep_exchange(rbslam_ctx *ctx, int t) {
  ReplicaGroup *g = rep_of(ctx);
  if (!g || g->world == 1) return RBSLAM_OK;
  const int r = ctx->replica_rank, q = t & 1;
  CK(cudaEventRecord(g->ev[r][q], ctx->stream));
  if (!g->barrier()) return ctx->fail(RBSLAM_ECUDA, "a smoother replica failed");
  for (int p = 0; p < g->world; ++p)
    if (p != r) CK(cudaStreamWaitEvent(ctx->stream, g->ev[p][q], 0));
  return RBSLAM_OK;
}

void rb_smoother_free(rbslam_ctx *ctx) {
  SmootherWs *w = ws_of(ctx);
  if (!w) return;
  void *ptrs[] = {w->xnk, w->Hk, w->lwdyn, w->paNtLog, w->paNt, w->sumlog, w->vtv, w->W, w->SS, w->Lw,
                  w->e, w->Dti, w->Uend, w->ak, w->AI, w->obs_t, w->obs_l, w->ImatAddt, w->ivecAddt,
                  w->NHR4, w->HRy, w->q2, w->Rinv};
  for (void *p : ptrs) if (p) cudaFree(p);
  delete w;
  ctx->smoother_ws = nullptr;
}

// ---------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------
__global__ void k_set_particle(double *__restrict__ xn, int n, int i, const double *__restrict__ src) {
  if ((int)threadIdx.x < n) xn[threadIdx.x + (size_t)i * n] = src[threadIdx.x];
}

// e[b][j] = yfut[j] - D(j,:) xl(:,i)   (src/particleSmoother.m:192-193); Dt col-major [M x ntau]
__global__ void k_future_resid(int M, int ntau, const double *__restrict__ Dt, int ldd,
                               const double *__restrict__ yfut, const double *__restrict__ xl, int i0,
                               double *__restrict__ e, size_t stride_e) {
  const int b = blockIdx.y, j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= ntau) return;
  const int lane = threadIdx.x & 31;
  const double *x = xl + (size_t)(i0 + b) * M;
  double acc = 0.0;
  for (int c = lane; c < M; c += 32) acc = fma(Dt[c + (size_t)j * ldd], x[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) e[(size_t)b * stride_e + j] = yfut[j] - acc;
}

// sparse family: stacked future Jacobian and residual of one particle, re-linearised at
// the particle's map (src/particleSmoother.m:197-216).  Dti col-major [M x nobs].
__global__ void k_sparse_future(ModelConsts mc, int nobs, const int *__restrict__ obs_t,
                                const int *__restrict__ obs_l, const double *__restrict__ xnk,
                                const double *__restrict__ y, const double *__restrict__ xl, int i0,
                                double *__restrict__ Dti, int ldd, size_t strideD, double *__restrict__ e,
                                size_t stride_e) {
  const int b = blockIdx.y;
  const int M = mc.M;
  const double *xli = xl + (size_t)(i0 + b) * M;
  double *D = Dti + (size_t)b * strideD;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nobs; j += gridDim.x * blockDim.x) {
    const int ti = obs_t[j], l = obs_l[j];
    const double *x = xnk + (size_t)ti * mc.n;
    double s, c;
    sincos(x[2], &s, &c);
    const double m1 = xli[2 * l], m2 = xli[2 * l + 1], p1 = x[0], p2 = x[1];
    const double t1 = -(c * p1 + s * p2), t2 = -(-s * p1 + c * p2);
    const double lx = c * m1 + s * m2 + t1, ly = -s * m1 + c * m2 + t2;
    const double yhat = (mc.cam_f * lx + mc.cam_fp * ly) / ly;
    const double dv = m2 * c - p2 * c - m1 * s + p1 * s;
    const double div = dv * dv;
    for (int r = 0; r < M; ++r) D[r + (size_t)j * ldd] = 0.0;
    D[2 * l + (size_t)j * ldd] = (mc.cam_f * (m2 - p2)) / div;
    D[2 * l + 1 + (size_t)j * ldd] = -(mc.cam_f * (m1 - p1)) / div;
    e[(size_t)b * stride_e + j] = y[(size_t)ti * mc.d + l] - yhat;
  }
}
// SS += blkdiag over time of R(ind,ind)  (src/particleSmoother.m:210,214)
__global__ void k_add_RS(int nobs, const int *__restrict__ obs_t, const int *__restrict__ obs_l,
                         const double *__restrict__ R, int d, double *__restrict__ SS, int ldss,
                         size_t strideSS) {
  const int b = blockIdx.y;
  double *S = SS + (size_t)b * strideSS;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nobs; j += gridDim.x * blockDim.x) {
    const int ti = obs_t[j];
    for (int j2 = j; j2 >= 0 && obs_t[j2] == ti; --j2) S[j + (size_t)j2 * ldss] += R[obs_l[j] + obs_l[j2] * d];
    for (int j2 = j + 1; j2 < nobs && obs_t[j2] == ti; ++j2) S[j + (size_t)j2 * ldss] += R[obs_l[j] + obs_l[j2] * d];
  }
}

// paNtLog = log w + logwDyn + logwMeas  (src/particleSmoother.m:229-232 / ...InformationForm.m:234-239)
__global__ void k_combine_cov(int N, int i0, int cnt, const double *__restrict__ w,
                              const double *__restrict__ lwdyn, const double *__restrict__ sumlog,
                              const double *__restrict__ vtv, double ne, OutPtrs out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cnt) return;
  const int i = i0 + b;
  const double lwm = -sumlog[b] - 0.5 * vtv[b] - ne / 2.0 * RB_LOG2PI;
  const double v = log(w[i]) + lwdyn[i] + lwm;
  for (int r = 0; r < out.n; ++r) out.p[r][i] = v;   // every replica's array (peer stores)
}
__global__ void k_combine_info(int N, int i0, int cnt, const double *__restrict__ w,
                               const double *__restrict__ lwdyn, const double *__restrict__ sumlog,
                               const double *__restrict__ vtv, const double *__restrict__ q2,
                               const double *__restrict__ hld, OutPtrs out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cnt) return;
  const int i = i0 + b;
  const double lwm = -0.5 * q2[i] - hld[i] - sumlog[b] + 0.5 * vtv[b];
  const double v = log(w[i]) + lwdyn[i] + lwm;
  for (int r = 0; r < out.n; ++r) out.p[r][i] = v;
}

// ---- information form -------------------------------------------------------------
// ivec0 = diag(1./diag(P0))*x0,  halfLogDetP = sum(log(sqrt(diag(P0))))   (...InformationForm.m:110-115)
__global__ void k_info_init_vec(int N, int M, const double *__restrict__ P0, const double *__restrict__ x0,
                                double *__restrict__ ivec, double *__restrict__ hld, double *__restrict__ q2) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)N * M) {
    const int r = idx % M;
    ivec[idx] = (1.0 / P0[r + (size_t)r * M]) * x0[r];
  }
  if (idx < (size_t)N) {
    double s = 0.0, q = 0.0;
    for (int r = 0; r < M; ++r) {
      const double p = P0[r + (size_t)r * M];
      s += log(sqrt(p));
      const double iv = (1.0 / p) * x0[r];
      q += iv * p * iv;
    }
    hld[idx] = s;
    q2[idx] = q;
  }
}

// suffix sums over the reference trajectory, accumulated in the reference's order
// (...InformationForm.m:132-146): ImatAddt = sum_jj H_jj'/R*H_jj, ivecAddt = sum_jj H_jj'/R*y_jj
__global__ void k_addt(int M, int d, int T0, int T1, double sign, const double *__restrict__ Hk,
                       int ldh, const double *__restrict__ Rinv, const double *__restrict__ y,
                       double *__restrict__ ImatAddt, int lda, double *__restrict__ ivecAddt, int init) {
  const int c = blockIdx.x;
  for (int r = threadIdx.x; r < M; r += blockDim.x) {
    double acc = init ? 0.0 : ImatAddt[r + (size_t)c * lda];
    double av = (init || c != 0) ? 0.0 : ivecAddt[r];
    for (int jj = T0; jj < T1; ++jj) {
      const double *H = Hk + (size_t)jj * d * ldh;
      double term = 0.0, tv = 0.0;
      for (int b = 0; b < d; ++b) {
        double hr = 0.0;   // (H'/R)(r,b)
        for (int a = 0; a < d; ++a) hr = fma(H[(size_t)a * ldh + r], Rinv[a + b * d], hr);
        term = fma(hr, H[(size_t)b * ldh + c], term);
        tv = fma(hr, y[(size_t)jj * d + b], tv);
      }
      acc += sign * term;
      av += sign * tv;
    }
    ImatAddt[r + (size_t)c * lda] = acc;
    if (c == 0) ivecAddt[r] = av;
  }
}

// per particle: NHR4(r,b) = -(H'/R)(r,b), HRy(r) = (H'/R*y)(r), and H4(.,d) = ivec of the ancestor
__global__ void k_info_prep(int M, int ld, int d, double *H4,
                            const double *__restrict__ Rinv, const double *__restrict__ y_t,
                            const double *__restrict__ ivec_old, const int *__restrict__ anc,
                            double *__restrict__ NHR4, double *__restrict__ HRy) {
  const int i = blockIdx.x;
  const double *iv = ivec_old + (size_t)(anc ? anc[i] : i) * M;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) {
    double h[4] = {0, 0, 0, 0}, o[4] = {0, 0, 0, 0};
    double hy = 0.0;
    if (r < M) {
      for (int a = 0; a < d; ++a) h[a] = H4[((size_t)i * ld + r) * 4 + a];
      for (int b = 0; b < d; ++b) {
        double hr = 0.0;
        for (int a = 0; a < d; ++a) hr = fma(h[a], Rinv[a + b * d], hr);
        o[b] = -hr;
        hy = fma(hr, y_t[b], hy);
      }
    }
    *reinterpret_cast<double4 *>(NHR4 + ((size_t)i * ld + r) * 4) = make_double4(o[0], o[1], o[2], o[3]);
    HRy[(size_t)i * ld + r] = hy;
    H4[((size_t)i * ld + r) * 4 + d] = r < M ? iv[r] : 0.0;   // column d of the DA = d+1 multiply
  }
}

struct InnovInfoArgs {
  Innov4Args base;
  const double *HRy;       // [N][ld]
  const double *ivec_old;  // [M x N]
  double *ivec_new;
  const double *hld_old;
  double *hld_new;
  double *q2;              // [N] ivec' P ivec after the update
  double Riy[4];           // R^-1 y, by value (a kernel argument: no copy, no synchronisation)
  double yRy_const;        // -1/2*y/R*y' - 1/2*log((2*pi)^d*det(R))
  double half_logdetR;
};

// information-form weights and update (...InformationForm.m:279-335) on top of k_innov4
template <int D>
__global__ void __launch_bounds__(128) k_innov4_info(InnovInfoArgs ia) {
  extern __shared__ double sm[];
  const Innov4Args &a = ia.base;
  const int M = a.M, ld = a.ld;
  double *sPH = sm;              // [ld][4]: P H' (0..D-1) and P ivec (D)
  __shared__ double s_red[4][D * D + D + 1];
  __shared__ double s_red2[4][2 * D + 1];
  __shared__ double s_L[D * D], s_SS[D * D], s_e[D];
  __shared__ double s_sumlog;
  const int i = blockIdx.x;
  const int an = a.anc ? a.anc[i] : i;
  const double *Hi = a.H4 + (size_t)i * ld * 4;
  const double *xls = a.xl_old + (size_t)an * M;
  const double *ivs = ia.ivec_old + (size_t)an * M;
  double part[D * D + D + 1];
#pragma unroll
  for (int q = 0; q < D * D + D + 1; ++q) part[q] = 0.0;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) {
    double ph[4] = {0, 0, 0, 0};
    for (int sp = 0; sp < a.nsplit; ++sp) {
      const double4 v = *reinterpret_cast<const double4 *>(a.PHp + (((size_t)i * a.nsplit + sp) * ld + r) * 4);
      ph[0] += v.x; ph[1] += v.y; ph[2] += v.z; ph[3] += v.w;
    }
    if (r >= M) { ph[0] = ph[1] = ph[2] = ph[3] = 0.0; }
    *reinterpret_cast<double4 *>(sPH + (size_t)r * 4) = make_double4(ph[0], ph[1], ph[2], ph[3]);
    if (r < M) {
      const double4 hv = *reinterpret_cast<const double4 *>(Hi + (size_t)r * 4);
      const double h[4] = {hv.x, hv.y, hv.z, hv.w};
      const double x = xls[r];
#pragma unroll
      for (int aa = 0; aa < D; ++aa) {
#pragma unroll
        for (int bb = 0; bb < D; ++bb) part[aa + bb * D] = fma(h[aa], ph[bb], part[aa + bb * D]);
        part[D * D + aa] = fma(h[aa], x, part[D * D + aa]);
      }
      part[D * D + D] = fma(ivs[r], ph[D], part[D * D + D]);   // q1 = ivec' P ivec
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < D * D + D + 1; ++q) {
    const double v = warp_sum(part[q]);
    if (lane == 0) s_red[wid][q] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S[D * D], e[D];
#pragma unroll
    for (int q = 0; q < D * D; ++q)
      S[q] = ((s_red[0][q] + s_red[1][q]) + (s_red[2][q] + s_red[3][q])) + a.R[q];
#pragma unroll
    for (int aa = 0; aa < D; ++aa)
      e[aa] = a.y_t[aa] - ((s_red[0][D * D + aa] + s_red[1][D * D + aa]) +
                           (s_red[2][D * D + aa] + s_red[3][D * D + aa]));
    double Lc[D * D];
#pragma unroll
    for (int q = 0; q < D * D; ++q) { Lc[q] = S[q]; s_SS[q] = S[q]; }
    int flag = chol_small(Lc, D, D);
    if (flag) {
#pragma unroll
      for (int q = 0; q < D * D; ++q) Lc[q] = S[q] + ((q % D) == (q / D) ? a.jitter : 0.0);
      atomicAdd(&a.status->used_jitter, 1);
      flag = chol_small(Lc, D, D);
      if (flag && atomicCAS(&a.status->not_pd, 0, 1) == 0) {
        a.status->not_pd_step = a.t;
        a.status->not_pd_particle = i;
      }
    }
    double sl = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r) sl += log(Lc[r + r * D]);
    s_sumlog = sl;
#pragma unroll
    for (int q = 0; q < D * D; ++q) s_L[q] = Lc[q];
#pragma unroll
    for (int aa = 0; aa < D; ++aa) s_e[aa] = e[aa];
  }
  __syncthreads();
  double *Gi = a.G4new + (size_t)i * ld * 4;
  double *KSi = a.KS4new + (size_t)i * ld * 4;
  double p2[2 * D + 1];
#pragma unroll
  for (int q = 0; q < 2 * D + 1; ++q) p2[q] = 0.0;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) {
    double g[4] = {0, 0, 0, 0}, ksv[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < D; ++q) {
      double s = sPH[(size_t)r * 4 + q];
#pragma unroll
      for (int k = 0; k < D; ++k) if (k < q) s -= s_L[q + k * D] * g[k];
      g[q] = s / s_L[q + q * D];
    }
#pragma unroll
    for (int q = D - 1; q >= 0; --q) {
      double s = g[q];
#pragma unroll
      for (int k = 0; k < D; ++k) if (k > q) s -= s_L[k + q * D] * g[k];
      g[q] = s / s_L[q + q * D];
    }
    double ge = 0.0;
#pragma unroll
    for (int q = 0; q < D; ++q) {
      ge = fma(g[q], s_e[q], ge);
      double ks = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) ks = fma(g[k], s_SS[k + q * D], ks);
      ksv[q] = ks;
    }
    *reinterpret_cast<double4 *>(Gi + (size_t)r * 4) = make_double4(g[0], g[1], g[2], g[3]);
    *reinterpret_cast<double4 *>(KSi + (size_t)r * 4) = make_double4(ksv[0], ksv[1], ksv[2], ksv[3]);
    if (r < M) {
      a.xl_new[(size_t)i * M + r] = xls[r] + ge;
      // ivecPlus = ivec + H'/R*y  (...InformationForm.m:292, 333)
      const double ivp = ivs[r] + ia.HRy[(size_t)i * ld + r];
      ia.ivec_new[(size_t)i * M + r] = ivp;
      // u = P*ivecPlus = P*ivec + PH*(R\y)
      double u = sPH[(size_t)r * 4 + D];
#pragma unroll
      for (int b = 0; b < D; ++b) u = fma(sPH[(size_t)r * 4 + b], ia.Riy[b], u);
      p2[2 * D] = fma(ivp, u, p2[2 * D]);
#pragma unroll
      for (int b = 0; b < D; ++b) {
        p2[b] = fma(ivp, ksv[b], p2[b]);          // ivecPlus' KS
        p2[D + b] = fma(g[b], ivp, p2[D + b]);    // G' ivecPlus
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2 * D + 1; ++q) {
    const double v = warp_sum(p2[q]);
    if (lane == 0) s_red2[wid][q] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[2 * D + 1];
#pragma unroll
    for (int q = 0; q < 2 * D + 1; ++q)
      t[q] = (s_red2[0][q] + s_red2[1][q]) + (s_red2[2][q] + s_red2[3][q]);
    double q2 = t[2 * D];   // ivecPlus' Pplus ivecPlus, Pplus = P - KS G'
#pragma unroll
    for (int b = 0; b < D; ++b) q2 -= t[b] * t[D + b];
    const double q1 = ((s_red[0][D * D + D] + s_red[1][D * D + D]) + (s_red[2][D * D + D] + s_red[3][D * D + D]));
    const double hld = ia.hld_old[an];
    const double hldp = -s_sumlog + ia.half_logdetR + hld;      // :298
    ia.hld_new[i] = hldp;
    ia.q2[i] = q2;
    a.logw[i] = -0.5 * q1 - hld + hldp + 0.5 * q2 + ia.yRy_const;   // :301-304
  }
}

// ---------------------------------------------------------------------------
// information-form hooks used by engine.cu
// ---------------------------------------------------------------------------
int rb_info_init(rbslam_ctx *ctx) {
  const int N = ctx->N, M = ctx->M;
  SmootherWs *w = ws_of(ctx);
  if (!w) return ctx->fail(RBSLAM_EARG, "information form state needs a smoother run");
  dim3 g(M, std::min(N, 4 * ctx->num_sms));
  k_init_slabs<<<g, 128, 0, ctx->stream>>>(ctx->d_Imat, ctx->slab, ctx->ld, M, N, ctx->d_P0, 1);
  const size_t tot = (size_t)M * N;
  k_info_init_vec<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(N, M, ctx->d_P0, ctx->d_x0lin,
                                                                          ctx->d_ivec[0], ctx->d_hld[0], w->q2);
  ctx->launches += 2;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

// information-form part of the Kalman phase: Imat gather + rank-d update (K8), then the
// covariance pass with the extra column P*ivec, then weights/update in information form.
template <int D>
static int info_kalman_phase_d(rbslam_ctx *ctx, const double *y_t_dev, const double *y_t_host, bool resampled) {
  SmootherWs *w = ws_of(ctx);
  const int N = ctx->N, M = ctx->M, ld = ctx->ld, d = ctx->d;
  const int *anc = resampled ? ctx->d_Ahist + (size_t)(ctx->t % ctx->T_hist) * N : nullptr;
  // R^-1 y and the constant of the weight (:303-304)
  double Riy[4] = {0, 0, 0, 0}, yRy = 0.0;
  if ((int)w->Rinv_h.size() != d * d) return ctx->fail(RBSLAM_EARG, "information form: inv(R) is set by the smoother run");
  const std::vector<double> &Rinv_h = w->Rinv_h;   // no stream synchronisation inside the step loop
  for (int a = 0; a < d; ++a) {
    for (int b = 0; b < d; ++b) Riy[a] += Rinv_h[a + b * d] * y_t_host[b];
  }
  for (int a = 0; a < d; ++a) yRy += y_t_host[a] * Riy[a];
  k_info_prep<<<N, 128, 0, ctx->stream>>>(M, ld, d, ctx->d_H, w->Rinv, y_t_dev, ctx->d_ivec[ctx->cx],
                                          anc, w->NHR4, w->HRy);
  ctx->launches += 1;
  const int npairs = ld / 2;
  const int R2 = (npairs + RB_STREAM_THREADS - 1) / RB_STREAM_THREADS;
  if (R2 > 4) return ctx->fail(RBSLAM_EARG, "information form: M too large for the streaming kernel");
  const size_t smem = sizeof(double) * 2 * ((size_t)8 * ld + 64);
  const int grid = std::min(N * ctx->nsplit, ctx->num_sms);
  // (1) Imat_dst = Imat_src + H'/R*H  (gather + rank-d update, :170,334)
  {
    StreamArgs sa;
    sa.hints = 0; sa.gk_by_particle = 1;
    sa.M = M; sa.ld = ld; sa.cw = ctx->cw; sa.nsplit = ctx->nsplit; sa.slab = ctx->slab;
    sa.P = ctx->d_Imat; sa.src_slot = ctx->d_src_slot; sa.dst_slot = ctx->d_slot[ctx->cs]; sa.anc = anc;
    sa.G4prev = ctx->d_H; sa.KS4prev = w->NHR4; sa.H4 = ctx->d_H; sa.PHp = ctx->d_PHp;
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 0 && !resampled) continue;
      const int *list = phase == 0 ? ctx->d_listA : ctx->d_listB;
      const int *cnt = ctx->d_counts + phase;
#define RB_INFO_LAUNCH(R2V, DA)                                                                         \
      {                                                                                                \
        auto kern = k_stream_pass<D, DA, R2V, 8, 2>;                                                   \
        RB_OPTIN_SMEM(kern, ctx->smem_optin - 1024);                                                   \
        kern<<<grid, RB_STREAM_THREADS, smem, ctx->stream>>>(sa, list, cnt, nullptr);                       \
      }
      switch (R2) { case 1: RB_INFO_LAUNCH(1, D) break; case 2: RB_INFO_LAUNCH(2, D) break;
                    case 3: RB_INFO_LAUNCH(3, D) break; default: RB_INFO_LAUNCH(4, D) break; }
      ctx->launches += 1;
    }
  }
  // (2) covariance pass: deferred downdate + P*[H' ivec]
  {
    StreamArgs sa;
    sa.hints = 0; sa.gk_by_particle = 0;
    sa.M = M; sa.ld = ld; sa.cw = ctx->cw; sa.nsplit = ctx->nsplit; sa.slab = ctx->slab;
    sa.P = ctx->d_P; sa.src_slot = ctx->d_src_slot; sa.dst_slot = ctx->d_slot[ctx->cs]; sa.anc = anc;
    sa.G4prev = ctx->d_G4[ctx->cg]; sa.KS4prev = ctx->d_KS4[ctx->cg]; sa.H4 = ctx->d_H; sa.PHp = ctx->d_PHp;
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 0 && !resampled) continue;
      const int *list = phase == 0 ? ctx->d_listA : ctx->d_listB;
      const int *cnt = ctx->d_counts + phase;
      switch (R2) { case 1: RB_INFO_LAUNCH(1, D + 1) break; case 2: RB_INFO_LAUNCH(2, D + 1) break;
                    case 3: RB_INFO_LAUNCH(3, D + 1) break; default: RB_INFO_LAUNCH(4, D + 1) break; }
      ctx->launches += 1;
    }
  }
  // (3) weights + update in information form
  InnovInfoArgs ia;
  ia.base.N = N; ia.base.M = M; ia.base.ld = ld; ia.base.nsplit = ctx->nsplit; ia.base.PHp = ctx->d_PHp;
  ia.base.H4 = ctx->d_H; ia.base.xl_old = ctx->d_xl[ctx->cx]; ia.base.anc = anc;
  ia.base.xl_new = ctx->d_xl[1 - ctx->cx];
  ia.base.G4new = ctx->d_G4[1 - ctx->cg]; ia.base.KS4new = ctx->d_KS4[1 - ctx->cg];
  ia.base.y_t = y_t_dev; ia.base.R = ctx->d_R; ia.base.jitter = ctx->jitter; ia.base.logw = ctx->d_logw;
  ia.base.status = ctx->d_status; ia.base.t = ctx->t;
  ia.HRy = w->HRy; ia.ivec_old = ctx->d_ivec[ctx->cx]; ia.ivec_new = ctx->d_ivec[1 - ctx->cx];
  ia.hld_old = ctx->d_hld[ctx->cx]; ia.hld_new = ctx->d_hld[1 - ctx->cx]; ia.q2 = w->q2;
  for (int a = 0; a < 4; ++a) ia.Riy[a] = Riy[a];
  ia.half_logdetR = w->half_logdetR;
  ia.yRy_const = -0.5 * yRy - 0.5 * (d * RB_LOG2PI + 2.0 * w->half_logdetR);
  k_innov4_info<D><<<N, 128, sizeof(double) * 8 * ld, ctx->stream>>>(ia);
  ctx->launches += 1;
  ctx->cg ^= 1;
  ctx->pending = true;
  ctx->cx ^= 1;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

int rb_info_kalman_phase(rbslam_ctx *ctx, const double *y_t_dev, const double *y_t_host, bool resampled) {
  switch (ctx->d) {
    case 1: return info_kalman_phase_d<1>(ctx, y_t_dev, y_t_host, resampled);
    case 2: return info_kalman_phase_d<2>(ctx, y_t_dev, y_t_host, resampled);
    case 3: return info_kalman_phase_d<3>(ctx, y_t_dev, y_t_host, resampled);
    default: return ctx->fail(RBSLAM_EARG, "information form supports d<=3");
  }
}

// ---------------------------------------------------------------------------
// ancestor weights of the reference particle
// ---------------------------------------------------------------------------
static int launch_gemm(rbslam_ctx *ctx, bool ta, const GemmArgs &g, int batch) {
  dim3 grid((g.m + 127) / 128, (g.n + 63) / 64, batch);
  // cp.async-pipelined kernel when every 16-byte copy is aligned (the sweeps allocate their operands that way);
  // otherwise the synchronous-staging kernel
  static const int force_sync = [] { const char *e = getenv("RBSLAM_GEMM_SYNC"); return (e && atoi(e)) ? 1 : 0; }();
  const bool aligned = !force_sync && (g.lda % 2 == 0) && (g.ldb % 2 == 0) && (g.strideA % 2 == 0) && (g.strideB % 2 == 0) &&
                       ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.B)) & 15) == 0;
  if (aligned) {
    RB_OPTIN_SMEM(k_dgemm_pipe<true>, dgemm_pipe_smem(true));
    RB_OPTIN_SMEM(k_dgemm_pipe<false>, dgemm_pipe_smem(false));
    if (ta) k_dgemm_pipe<true><<<grid, 256, dgemm_pipe_smem(true), ctx->stream>>>(g);
    else k_dgemm_pipe<false><<<grid, 256, dgemm_pipe_smem(false), ctx->stream>>>(g);
  } else {
    const size_t smem = sizeof(double) * 32 * (RB_LDA + RB_LDB);
    RB_OPTIN_SMEM(k_dgemm<true>, smem);
    RB_OPTIN_SMEM(k_dgemm<false>, smem);
    if (ta) k_dgemm<true><<<grid, 256, smem, ctx->stream>>>(g);
    else k_dgemm<false><<<grid, 256, smem, ctx->stream>>>(g);
  }
  ctx->launches += 1;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

// tuning switches of the factorisation, read once (thread-safe: the replica group calls this from one host
// thread per device at the same time)
struct CholTuning { int nt, panel_min, cfg; };
static const CholTuning &chol_tuning() {
  static const CholTuning t = [] {
    CholTuning v;
    v.nt = 64;   // 64: k_chol_inv (default); 128 / 256: k_chol_solve with that many threads (RBSLAM_CHOL_KERNEL=solve)
    if (const char *e = getenv("RBSLAM_CHOL_KERNEL"))
      if (!strcmp(e, "solve")) v.nt = 128;
    if (const char *e = getenv("RBSLAM_CHOL_THREADS")) v.nt = atoi(e) == 256 ? 256 : (atoi(e) == 128 ? 128 : v.nt);
    // batches at least this large go panel by panel across the batch.  Measured at C5 (N = 4096, M = 515):
    // 21.6 ms per step against 19.8 ms for k_chol_solve and 10.6 ms for k_chol_inv (profiles/tuning_r2.md), so the
    // path is opt-in (RBSLAM_CHOL_PANEL_MIN=<batch size>)
    v.panel_min = 1 << 30;
    if (const char *e = getenv("RBSLAM_CHOL_PANEL_MIN")) v.panel_min = std::max(1, atoi(e));
    // operand ring of k_chol_inv: columns per stage x stages, CTAs per SM (profiles/tuning_r2.md section 3)
    v.cfg = 0;
    if (const char *e = getenv("RBSLAM_CHOL_CFG")) v.cfg = atoi(e) & 1;
    return v;
  }();
  return t;
}

static int launch_chol(rbslam_ctx *ctx, const CholArgs &c, int batch) {
  const int nt = chol_tuning().nt, panel_min = chol_tuning().panel_min;
  RB_OPTIN_SMEM(k_chol_solve<128>, chol_solve_smem(128));
  RB_OPTIN_SMEM(k_chol_solve<256>, chol_solve_smem(256));
  RB_OPTIN_SMEM((k_chol_inv<16, 3, 4, 128>), chol_inv_smem(16, 3));
  RB_OPTIN_SMEM((k_chol_inv<32, 2, 3, 128>), chol_inv_smem(32, 2));
  RB_OPTIN_SMEM((k_chol_inv<16, 3, 1, 512>), chol_inv_smem(16, 3, 512));
  if (batch >= panel_min && c.n > RB_CP_NB) {
    // panel by panel across the batch (dense_kernels.cuh): thousands of independent tensor-core tiles
    // per launch instead of 3-4 latency-bound matrices per SM
    if (ctx->chol_fail_cap < batch) {
      if (ctx->d_chol_fail) cudaFree(ctx->d_chol_fail);
      ctx->d_chol_fail = nullptr; ctx->chol_fail_cap = 0;
      RB_ALLOC(ctx->d_chol_fail, batch);
      ctx->chol_fail_cap = batch;
    }
    CK(cudaMemsetAsync(ctx->d_chol_fail, 0, sizeof(int) * batch, ctx->stream));
    const size_t gsm = sizeof(double) * 2 * 32 * (RB_LDA + RB_LDB);
    RB_OPTIN_SMEM(k_chol_panel_gemm, gsm);
    CholPanelArgs p;
    p.c = c; p.fail = ctx->d_chol_fail;
    const int nr = c.n + 1;
    for (int jb = 0; jb < c.n; jb += RB_CP_NB) {
      p.jb = jb;
      k_chol_panel_gemm<<<dim3((nr - jb + 127) / 128, 1, batch), 256, gsm, ctx->stream>>>(p);
      k_chol_panel_factor<<<batch, 128, 0, ctx->stream>>>(p);
      ctx->launches += 2;
    }
    k_chol_finalize<<<batch, 128, 0, ctx->stream>>>(c, ctx->d_chol_fail);
    // matrices that were not positive definite: the one-CTA-per-matrix kernel redoes them with the
    // reference's retry (jitter) / error semantics; it returns at once for all the others
    CholArgs c2 = c;
    c2.only_failed = ctx->d_chol_fail;
    k_chol_solve<128><<<batch, 128, chol_solve_smem(128), ctx->stream>>>(c2);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return RBSLAM_OK;
  }
  if (nt == 256) k_chol_solve<256><<<batch, 256, chol_solve_smem(256), ctx->stream>>>(c);
  else if (nt == 64) {
    const int cfg = chol_tuning().cfg;
    // small batches (fewer matrices than SMs x 1.1: the C1 example) get 512 threads per matrix
    int wide_max = 160;   // read per call: the tests switch between the two shapes inside one process
    if (const char *e = getenv("RBSLAM_CHOL_WIDE_MAX")) wide_max = atoi(e);
    if (cfg == 0 && batch <= wide_max) k_chol_inv<16, 3, 1, 512><<<batch, 512, chol_inv_smem(16, 3, 512), ctx->stream>>>(c);
    else if (cfg == 0) k_chol_inv<16, 3, 4, 128><<<batch, 128, chol_inv_smem(16, 3), ctx->stream>>>(c);
    else k_chol_inv<32, 2, 3, 128><<<batch, 128, chol_inv_smem(32, 2), ctx->stream>>>(c);
  }
  else k_chol_solve<128><<<batch, 128, chol_solve_smem(128), ctx->stream>>>(c);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

// covariance form (K6): src/particleSmoother.m:156-245
static int ancestor_weights_cov(rbslam_ctx *ctx, int t, bool use_default_dyn) {
  SmootherWs *w = ws_of(ctx);
  const int N = ctx->N, M = ctx->M, d = ctx->d, n = ctx->n, T = ctx->run_T;
  const bool sparse = ctx->mc.family == FAM_SPARSE_VISUAL2D;
  int rc;
  if ((rc = rb_flush_pending(ctx))) return rc;
  const double *xn_old = ctx->d_Xhist + (size_t)((t - 1) % ctx->T_hist) * N * n;
  const double *Qp = ctx->d_Q + (ctx->Q_pages > 1 ? (size_t)(t - 1) * ctx->nw * ctx->nw : 0);
  k_dyn_logweight<<<(N + 127) / 128, 128, 0, ctx->stream>>>(ctx->mc, N, w->xnk + (size_t)t * n, xn_old,
                                                            ctx->d_odo + (size_t)(t - 1) * ctx->n_odo,
                                                            ctx->h_dt[t - 1], Qp, use_default_dyn ? 1 : 0, w->lwdyn);
  ctx->launches += 1;
  const int ne = sparse ? (w->n_obs - w->obs_off[t]) : d * (T - t);
  const double *xl_old = ctx->d_xl[ctx->cx];
  const int *slot_old = ctx->d_slot[ctx->cs];
  const int ldw = ctx->ld;   // W and Dti: leading dimension padded like the slabs (16-byte aligned columns)
  const size_t sW = (size_t)ldw * w->ntau_max, sS = w->ntau_max * w->ntau_max;
  const size_t sLw = (size_t)chol_ldl((int)w->ntau_max) * w->ntau_max;
  int blo, bhi;
  rep_block(ctx, N, blo, bhi);
  const OutPtrs outs = rep_outs(ctx, w->paNtLog);
  for (int b0 = blo; b0 < bhi; b0 += (int)w->batch) {
    const int cnt = std::min<int>((int)w->batch, bhi - b0);
    if (ne > 0) {
      GemmArgs g1{};   // W = P_i * D'
      g1.m = M; g1.n = ne; g1.k = M;
      g1.A = ctx->d_P; g1.lda = ctx->ld; g1.strideA = ctx->slab; g1.slotA = slot_old + b0;
      g1.C = w->W; g1.ldc = ldw; g1.strideC = sW; g1.Rblk = nullptr; g1.d = 1;
      GemmArgs g2{};   // SS = D * W (+ kron(I,R))
      g2.m = ne; g2.n = ne; g2.k = M; g2.slotA = nullptr;
      g2.B = w->W; g2.ldb = ldw; g2.strideB = sW;
      g2.C = w->SS; g2.ldc = ne; g2.strideC = sS;
      g2.lower = 1;    // S = D P D' + R is symmetric and the factorisation reads its lower triangle only
      if (sparse) {
        const int o0 = w->obs_off[t];
        k_sparse_future<<<dim3((ne + 127) / 128, cnt), 128, 0, ctx->stream>>>(
            ctx->mc, ne, w->obs_t + o0, w->obs_l + o0, w->xnk, ctx->d_y, xl_old, b0, w->Dti, ldw, sW, w->e, w->ntau_max);
        ctx->launches += 1;
        g1.B = w->Dti; g1.ldb = ldw; g1.strideB = sW;
        g2.A = w->Dti; g2.lda = ldw; g2.strideA = sW; g2.Rblk = nullptr; g2.d = 1;
      } else {
        const double *Dt = w->Hk + (size_t)t * d * ctx->ldh;
        g1.B = Dt; g1.ldb = ctx->ldh; g1.strideB = 0;
        g2.A = Dt; g2.lda = ctx->ldh; g2.strideA = 0; g2.Rblk = ctx->d_R; g2.d = d;
        k_future_resid<<<dim3((ne + 3) / 4, cnt), 128, 0, ctx->stream>>>(M, ne, Dt, ctx->ldh, ctx->d_y + (size_t)t * d,
                                                                         xl_old, b0, w->e, w->ntau_max);
        ctx->launches += 1;
      }
      if ((rc = launch_gemm(ctx, false, g1, cnt))) return rc;
      if ((rc = launch_gemm(ctx, true, g2, cnt))) return rc;
      if (sparse) {
        const int o0 = w->obs_off[t];
        k_add_RS<<<dim3((ne + 127) / 128, cnt), 128, 0, ctx->stream>>>(ne, w->obs_t + o0, w->obs_l + o0, ctx->d_R, d,
                                                                       w->SS, ne, sS);
        ctx->launches += 1;
      }
      CholArgs c{};
      c.n = ne; c.A1 = w->SS; c.lda1 = ne; c.strideA1 = sS; c.slot1 = nullptr; c.A2 = nullptr; c.lda2 = 0;
      c.L = w->Lw; c.ldl = chol_ldl(ne); c.strideL = sLw; c.rhs = w->e; c.stride_rhs = w->ntau_max; c.rhs2 = nullptr;
      c.jitter = ctx->jitter; c.sum_log_diag = w->sumlog; c.vtv = w->vtv; c.status = ctx->d_status; c.t = t;
      if ((rc = launch_chol(ctx, c, cnt))) return rc;
    } else {
      CK(cudaMemsetAsync(w->sumlog, 0, sizeof(double) * cnt, ctx->stream));
      CK(cudaMemsetAsync(w->vtv, 0, sizeof(double) * cnt, ctx->stream));
    }
    k_combine_cov<<<(cnt + 127) / 128, 128, 0, ctx->stream>>>(N, b0, cnt, ctx->d_w, w->lwdyn, w->sumlog, w->vtv,
                                                              (double)ne, outs);
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  return rep_exchange(ctx, t);
}

// information form (K7): ...InformationForm.m:187-254
static int ancestor_weights_info(rbslam_ctx *ctx, int t, bool use_default_dyn) {
  SmootherWs *w = ws_of(ctx);
  const int N = ctx->N, M = ctx->M, d = ctx->d, n = ctx->n;
  int rc;
  // the suffix sums lose the term of step t-1 (:192-201)
  k_addt<<<M, 128, 0, ctx->stream>>>(M, d, t - 1, t, -1.0, w->Hk, ctx->ldh, w->Rinv, ctx->d_y, w->ImatAddt, ctx->ld,
                                     w->ivecAddt, 0);
  ctx->launches += 1;
  const double *xn_old = ctx->d_Xhist + (size_t)((t - 1) % ctx->T_hist) * N * n;
  const double *Qp = ctx->d_Q + (ctx->Q_pages > 1 ? (size_t)(t - 1) * ctx->nw * ctx->nw : 0);
  k_dyn_logweight<<<(N + 127) / 128, 128, 0, ctx->stream>>>(ctx->mc, N, w->xnk + (size_t)t * n, xn_old,
                                                            ctx->d_odo + (size_t)(t - 1) * ctx->n_odo,
                                                            ctx->h_dt[t - 1], Qp, use_default_dyn ? 1 : 0, w->lwdyn);
  ctx->launches += 1;
  const size_t sL = (size_t)chol_ldl(M) * M;
  int blo, bhi;
  rep_block(ctx, N, blo, bhi);
  const OutPtrs outs = rep_outs(ctx, w->paNtLog);
  for (int b0 = blo; b0 < bhi; b0 += (int)w->batch) {
    const int cnt = std::min<int>((int)w->batch, bhi - b0);
    CholArgs c{};
    c.n = M; c.A1 = ctx->d_Imat; c.lda1 = ctx->ld; c.strideA1 = ctx->slab; c.slot1 = ctx->d_slot[ctx->cs] + b0;
    c.A2 = w->ImatAddt; c.lda2 = ctx->ld; c.L = w->Lw; c.ldl = chol_ldl(M); c.strideL = sL;
    c.rhs = ctx->d_ivec[ctx->cx] + (size_t)b0 * M; c.stride_rhs = M; c.rhs2 = w->ivecAddt;
    c.jitter = -1.0;   // quirk Q7: the reference's retry branch is broken and would raise
    c.sum_log_diag = w->sumlog; c.vtv = w->vtv; c.status = ctx->d_status; c.t = t;
    if ((rc = launch_chol(ctx, c, cnt))) return rc;
    k_combine_info<<<(cnt + 127) / 128, 128, 0, ctx->stream>>>(N, b0, cnt, ctx->d_w, w->lwdyn, w->sumlog, w->vtv,
                                                               w->q2, ctx->d_hld[ctx->cx], outs);
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  return rep_exchange(ctx, t);
}

// ---------------------------------------------------------------------------
// the sweep loop
// ---------------------------------------------------------------------------
static int smoother_run_impl(rbslam_ctx *ctx, const rbslam_inputs *in, int32_t N_K, int32_t form,
                             rbslam_smoother_outputs *out);

extern "C" int rbslam_smoother_run(rbslam_ctx *ctx, const rbslam_inputs *in, int32_t N_K, int32_t form,
                                   rbslam_smoother_outputs *out) {
  if (!ctx || !in || !out || N_K < 1 || (form != 0 && form != 1)) return RBSLAM_EARG;
  ReplicaGroup *g = rep_of(ctx);
  if (!g || g->world == 1) return smoother_run_impl(ctx, in, N_K, form, out);
  if (ctx->replica_rank != 0) return ctx->fail(RBSLAM_EARG, "call the smoother on the leader of the replica group");
  // one host thread per replica: every replica runs the same sweeps on the same inputs and streams
  g->arrived.store(0); g->failed.store(0);
  std::vector<int> rcs(g->world, RBSLAM_OK);
  std::vector<std::thread> th;
  rbslam_smoother_outputs none;
  memset(&none, 0, sizeof none);
  for (int r = 1; r < g->world; ++r)
    th.emplace_back([&, r]() {
      rcs[r] = smoother_run_impl(g->ctx[r], in, N_K, form, &none);
      if (rcs[r]) g->failed.store(1);
    });
  rcs[0] = smoother_run_impl(ctx, in, N_K, form, out);
  if (rcs[0]) g->failed.store(1);
  for (auto &t : th) t.join();
  cudaSetDevice(ctx->cfg.device);
  for (int r = 0; r < g->world; ++r)
    if (rcs[r]) { if (r) ctx->err = "replica " + std::to_string(r) + ": " + g->ctx[r]->err; return rcs[r]; }
  return RBSLAM_OK;
}

static int smoother_run_impl(rbslam_ctx *ctx, const rbslam_inputs *in, int32_t N_K, int32_t form,
                             rbslam_smoother_outputs *out) {
  if (ctx->shard_ws) return ctx->fail(RBSLAM_EARG, "rbslam_smoother_run: this context is a filter shard (world > 1)");
  if (!ctx->cfg.keep_history) return ctx->fail(RBSLAM_EARG, "the smoother needs keep_history=1");
  if (ctx->pt) return ctx->fail(RBSLAM_EARG, "packed symmetric slabs (kalman_variant 7 / -1) are filter-only: create the context with kalman_variant 0");
  const bool sparse = ctx->mc.family == FAM_SPARSE_VISUAL2D;
  if (form == 1 && sparse)
    return ctx->fail(RBSLAM_EARG, "This code has only been implemented for dense features");   // :77-80
  if (form == 1 && (!ctx->d_Imat || ctx->kpath != 1))
    return ctx->fail(RBSLAM_EARG, "information form needs a context created with information_form=1");
  if (form == 1 && in->x0_lin_cols != 1)
    return ctx->fail(RBSLAM_EARG, "information form takes a single x0_lin column (reference quirk Q8)");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc = rb_upload_inputs(ctx, in, N_K);
  if (rc) return rc;
  ctx->jitter = 1e-2;   // src/particleSmoother.m:70
  const int N = ctx->N, M = ctx->M, d = ctx->d, n = ctx->n, T = in->T;
  const bool use_default_dyn = sparse;   // dynResNorm = [] for the sparse family (psslam.m:118)
  rb_smoother_free(ctx);
  SmootherWs *w = new SmootherWs();
  ctx->smoother_ws = w;
  RB_ALLOC(w->xnk, (size_t)T * n);
  RB_ALLOC(w->lwdyn, N); RB_ALLOC(w->paNtLog, N); RB_ALLOC(w->paNt, N);
  if (ReplicaGroup *g = rep_of(ctx)) {   // publish this replica's array, wait for everybody's
    g->paNtLog[ctx->replica_rank] = w->paNtLog;
    if (g->world > 1 && !g->barrier()) return ctx->fail(RBSLAM_ECUDA, "a smoother replica failed");
  }
  RB_ALLOC(w->Uend, N_K); RB_ALLOC(w->ak, N_K);
  if ((rc = rb_h2d(ctx, w->Uend, ctx->h_Uend.data(), sizeof(double) * N_K))) return rc;
  if (out->AI) {
    RB_ALLOC(w->AI, (size_t)N_K * T * N);
    std::vector<double> nanv((size_t)N_K * T * N, NAN);
    if ((rc = rb_h2d(ctx, w->AI, nanv.data(), nanv.size() * 8))) return rc;
  }
  if (!sparse) RB_ALLOC(w->Hk, (size_t)T * d * ctx->ldh);
  if (N_K > 1) {
    if (form == 0) {
      if (sparse) {
        std::vector<int> ot, ol;
        w->obs_off.assign(T + 1, 0);
        for (int ti = 0; ti < T; ++ti) {
          w->obs_off[ti] = (int)ot.size();
          for (int l = 0; l < d; ++l)
            if (!std::isnan(ctx->h_y[(size_t)ti * d + l])) { ot.push_back(ti); ol.push_back(l); }
        }
        w->obs_off[T] = (int)ot.size();
        w->n_obs = (int)ot.size();
        RB_ALLOC(w->obs_t, ot.size()); RB_ALLOC(w->obs_l, ol.size());
        if (!ot.empty()) {
          if ((rc = rb_h2d(ctx, w->obs_t, ot.data(), sizeof(int) * ot.size()))) return rc;
          if ((rc = rb_h2d(ctx, w->obs_l, ol.data(), sizeof(int) * ol.size()))) return rc;
        }
        w->ntau_max = std::max(1, w->n_obs - (T > 1 ? w->obs_off[1] : 0));
      } else {
        w->ntau_max = (size_t)std::max(1, d * (T - 1));
      }
      const size_t per = ((size_t)M * w->ntau_max * (sparse ? 2 : 1) + 2 * w->ntau_max * w->ntau_max + w->ntau_max) * 8;
      w->batch = std::max<size_t>(1, std::min<size_t>(N, ((size_t)6 << 30) / per));
      RB_ALLOC(w->W, w->batch * ctx->ld * w->ntau_max);
      if (sparse) RB_ALLOC(w->Dti, w->batch * ctx->ld * w->ntau_max);
      RB_ALLOC(w->SS, w->batch * w->ntau_max * w->ntau_max);
      RB_ALLOC(w->Lw, w->batch * (size_t)chol_ldl((int)w->ntau_max) * w->ntau_max);
      RB_ALLOC(w->e, w->batch * w->ntau_max);
    } else {
      const size_t per = (size_t)M * M * 8;
      w->batch = std::max<size_t>(1, std::min<size_t>(N, ((size_t)12 << 30) / per));
      RB_ALLOC(w->Lw, w->batch * (size_t)chol_ldl(M) * M);
    }
    RB_ALLOC(w->sumlog, w->batch); RB_ALLOC(w->vtv, w->batch);
  }
  if (form == 1) {
    RB_ALLOC(w->ImatAddt, (size_t)ctx->ld * M); RB_ALLOC(w->ivecAddt, M);   // padded leading dimension: 16-byte loads in K7
    RB_ALLOC(w->NHR4, (size_t)N * ctx->ld * 4); RB_ALLOC(w->HRy, (size_t)N * ctx->ld);
    RB_ALLOC(w->q2, N); RB_ALLOC(w->Rinv, (size_t)d * d);
    // R^-1 and log det R on the host (d <= 3)
    std::vector<double> A(ctx->h_R), Ri((size_t)d * d, 0.0);
    double det = 1.0;
    {  // Gauss-Jordan with partial pivoting
      std::vector<double> aug((size_t)d * 2 * d, 0.0);
      for (int r = 0; r < d; ++r) { for (int c = 0; c < d; ++c) aug[r * 2 * d + c] = A[r + c * d]; aug[r * 2 * d + d + r] = 1.0; }
      for (int c = 0; c < d; ++c) {
        int p = c;
        for (int r = c + 1; r < d; ++r) if (std::fabs(aug[r * 2 * d + c]) > std::fabs(aug[p * 2 * d + c])) p = r;
        if (p != c) { for (int k = 0; k < 2 * d; ++k) std::swap(aug[p * 2 * d + k], aug[c * 2 * d + k]); det = -det; }
        const double pv = aug[c * 2 * d + c];
        det *= pv;
        for (int k = 0; k < 2 * d; ++k) aug[c * 2 * d + k] /= pv;
        for (int r = 0; r < d; ++r) if (r != c) {
          const double f = aug[r * 2 * d + c];
          for (int k = 0; k < 2 * d; ++k) aug[r * 2 * d + k] -= f * aug[c * 2 * d + k];
        }
      }
      for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) Ri[r + c * d] = aug[r * 2 * d + d + c];
    }
    if (!(det > 0)) return ctx->fail(RBSLAM_EARG, "R must be positive definite");
    w->half_logdetR = 0.5 * std::log(det);
    if ((rc = rb_h2d(ctx, w->Rinv, Ri.data(), sizeof(double) * d * d))) return rc;
    w->Rinv_h = Ri;
  }

  std::vector<int> h_ak(N_K, 0);
  for (int k = 0; k < N_K; ++k) {
    ctx->sweep = k;
    CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(DevStatus), ctx->stream));
    if ((rc = rb_init_state(ctx, form == 1))) return rc;
    if (k > 0) {
      // pin particle N_P to the reference trajectory (src/particleSmoother.m:93-96,111-113)
      k_set_particle<<<1, 32, 0, ctx->stream>>>(ctx->d_Xhist, n, N - 1, w->xnk);
      ctx->launches += 1;
      if (!sparse) {
        // dy_xnk = measModel(xnk)  (:119-121), row layout [T][d][ldh]
        k_meas<<<T, 128, 0, ctx->stream>>>(ctx->mc, T, w->xnk, nullptr, M, nullptr, w->Hk, (size_t)d * ctx->ldh,
                                           ctx->ldh, 1, ctx->ld, nullptr);
        ctx->launches += 1;
      }
      if (form == 1) {
        k_addt<<<M, 128, 0, ctx->stream>>>(M, d, 0, T, 1.0, w->Hk, ctx->ldh, w->Rinv, ctx->d_y, w->ImatAddt, ctx->ld,
                                           w->ivecAddt, 1);
        ctx->launches += 1;
      }
    }
    ctx->running = true;
    for (int t = 0; t < T; ++t) {
      ctx->t = t;
      const bool resampled = t > 0;
      if (resampled) {
        rb_phase_begin(ctx, RB_PH_RESAMPLE);
        if ((rc = rb_resample_phase(ctx, k == 0 ? N : N - 1))) return rc;
        rb_phase_end(ctx);
        if (k > 0) {
          rb_phase_begin(ctx, RB_PH_ANCESTOR);
          rc = form == 0 ? ancestor_weights_cov(ctx, t, use_default_dyn) : ancestor_weights_info(ctx, t, use_default_dyn);
          if (rc) return rc;
          // normalise (:236-238) and draw ai(N_P) = sample(paNt) (:241)
          double *ai_tap = w->AI ? w->AI + ((size_t)k * T + t) * N : nullptr;
          k_normalize<<<1, 1024, 0, ctx->stream>>>(N, 0, w->paNtLog, w->paNt, nullptr, nullptr, nullptr, nullptr,
                                                   nullptr, ai_tap);
          int *ai = ctx->d_Ahist + (size_t)(t % ctx->T_hist) * N;
          const size_t soff = ((size_t)k * T + t) * N;
          RngSrc rs;
          rs.U = ctx->have_U ? ctx->d_U + soff : nullptr;
          rs.seed = ctx->cfg.seed; rs.sweep = k; rs.t = t;
          const int *forced = ctx->have_forced ? ctx->d_forced + soff : nullptr;
          size_t smem = sizeof(double) * (size_t)N;
          smem = std::min(smem, std::min(ctx->smem_resample_max, (size_t)(96 << 10)));
          k_resample<<<1, 1024, smem, ctx->stream>>>(N, N - 1, 1, w->paNt, ctx->d_wc, rs, forced, ai, ctx->d_status);
          ctx->launches += 2;
          rb_phase_end(ctx);
        }
        rb_phase_begin(ctx, RB_PH_RESAMPLE);
        if ((rc = rb_plan_phase(ctx))) return rc;
        rb_phase_end(ctx);
        rb_phase_begin(ctx, RB_PH_PROPAGATE);
        if ((rc = rb_propagate_phase(ctx, k == 0 ? N : N - 1))) return rc;
        if (k > 0) {
          k_set_particle<<<1, 32, 0, ctx->stream>>>(ctx->d_Xhist + (size_t)(t % ctx->T_hist) * N * n, n, N - 1,
                                                    w->xnk + (size_t)t * n);
          ctx->launches += 1;
        }
        rb_phase_end(ctx);
      } else {
        k_plan_identity<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, ctx->d_slot[ctx->cs], ctx->d_src_slot,
                                                                 ctx->d_listB, ctx->d_counts);
        ctx->launches += 1;
      }
      rb_phase_begin(ctx, RB_PH_MEAS);
      if ((rc = rb_meas_phase(ctx, resampled))) return rc;
      rb_phase_end(ctx);
      rb_phase_begin(ctx, form == 1 ? RB_PH_INFO : RB_PH_KALMAN);
      if (form == 1) rc = rb_info_kalman_phase(ctx, ctx->d_y + (size_t)t * d, ctx->h_y.data() + (size_t)t * d, resampled);
      else rc = rb_kalman_phase(ctx, ctx->d_y + (size_t)t * d, resampled);
      rb_phase_end(ctx);
      if (rc) return rc;
      rb_phase_begin(ctx, RB_PH_NORMALIZE);
      if ((rc = rb_normalize_phase(ctx))) return rc;
      rb_phase_end(ctx);
      ctx->t = t + 1;
      if (ctx->step_fn) {
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->step_fn(ctx->step_user, k, t);
      }
    }
    // ak = sample(w) (:346); xnk, xlk, Pk (:347-354)
    if (!ctx->h_forced_ak.empty()) {
      h_ak[k] = ctx->h_forced_ak[k];
      CK(cudaMemcpyAsync(w->ak + k, &h_ak[k], sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    } else {
      RngSrc rs;
      rs.U = ctx->have_U ? w->Uend + k : nullptr;
      rs.seed = ctx->cfg.seed; rs.sweep = k; rs.t = T;
      size_t smem = sizeof(double) * (size_t)N;
      smem = std::min(smem, std::min(ctx->smem_resample_max, (size_t)(96 << 10)));
      k_resample<<<1, 1024, smem, ctx->stream>>>(N, 0, 1, ctx->d_w, ctx->d_wc, rs, nullptr, w->ak + k, ctx->d_status);
      ctx->launches += 1;
      if ((rc = rb_d2h(ctx, &h_ak[k], w->ak + k, sizeof(int)))) return rc;
    }
    if ((rc = rb_check_status(ctx))) return rc;
    k_trace<<<1, 32, 0, ctx->stream>>>(N, n, T, ctx->d_Xhist, ctx->d_Ahist, w->ak + k, w->xnk);
    ctx->launches += 1;
    if (out->XNK && (rc = rb_d2h(ctx, out->XNK + (size_t)k * n * T, w->xnk, sizeof(double) * n * T))) return rc;
    if (out->XLK && (rc = rb_d2h(ctx, out->XLK + (size_t)k * M, ctx->d_xl[ctx->cx] + (size_t)h_ak[k] * M, sizeof(double) * M))) return rc;
    if (out->PK) {
      if ((rc = rb_flush_pending(ctx))) return rc;
      double *tmp = ctx->d_scratch + 2 * (size_t)M + 64;
      k_pack_slabs<<<dim3(M, 1), 128, 0, ctx->stream>>>(M, ctx->ld, ctx->slab, ctx->d_P, ctx->d_slot[ctx->cs], h_ak[k], tmp, 0);
      ctx->launches += 1;
      if ((rc = rb_d2h(ctx, out->PK + (size_t)k * M * M, tmp, sizeof(double) * M * M))) return rc;
    }
    ctx->running = false;
    // sweep k is complete and its outputs are written: t == T tells the host's callback so
    // (src/particleSmoother.m:359-365: makePlots(xnk,xlk,k,XNK,XLK,PK), then the progress line)
    if (ctx->step_fn) ctx->step_fn(ctx->step_user, k, T);
  }
  if (out->ak) for (int k = 0; k < N_K; ++k) out->ak[k] = h_ak[k];
  if (out->AI && (rc = rb_d2h(ctx, out->AI, w->AI, sizeof(double) * (size_t)N_K * T * N))) return rc;
  return rb_check_status(ctx);
}

// ---------------------------------------------------------------------------
// kernel-level entry point: the measurement part of the ancestor weights for N particles,
// host buffers in and out.  Runs exactly the kernels of the sweeps (k_future_resid, k_dgemm x2,
// k_chol_solve for form 0; k_chol_solve on Imat_i + ImatAddt for form 1) at any size, so K6 / K7
// can be compared with the oracle's per-particle formulas at the C1 / C5 shapes where a whole
// smoother run is out of the oracle's reach.
// ---------------------------------------------------------------------------
__global__ void k_aw_combine(int N, int form, double ne, const double *__restrict__ sumlog,
                             const double *__restrict__ vtv, const double *__restrict__ q2,
                             const double *__restrict__ hld, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (form == 0) out[i] = -sumlog[i] - 0.5 * vtv[i] - ne / 2.0 * RB_LOG2PI;         // src/particleSmoother.m:229
  else out[i] = -0.5 * q2[i] - hld[i] - sumlog[i] + 0.5 * vtv[i];                    // ...InformationForm.m:234-236
}

extern "C" int rbslam_op_ancestor_weights(rbslam_ctx *ctx, int32_t form, int32_t N, int32_t ne, const double *A,
                                          const double *v, const double *S, const double *r, const double *R,
                                          const double *q2, const double *hld, double jitter, double *logwMeas) {
  if (!ctx || !A || !v || !S || !r || !logwMeas || N < 1 || (form != 0 && form != 1)) return RBSLAM_EARG;
  if (form == 0 && (ne < 1 || !R)) return ctx->fail(RBSLAM_EARG, "op_ancestor_weights form 0: need ne >= 1 and R");
  if (form == 1 && (!q2 || !hld)) return ctx->fail(RBSLAM_EARG, "op_ancestor_weights form 1: need q2 and halfLogDetP");
  if (form == 0 && ctx->mc.family == FAM_SPARSE_VISUAL2D)
    return ctx->fail(RBSLAM_EARG, "op_ancestor_weights: dense families only (the sparse family re-linearises per particle)");
  CK(cudaSetDevice(ctx->cfg.device));
  const int M = ctx->M, d = ctx->d;
  const int n = form == 0 ? ne : M;          // order of the matrix that is factored
  if (form == 0 && ne % d) return ctx->fail(RBSLAM_EARG, "op_ancestor_weights form 0: ne must be a multiple of d");
  int rc;
  struct Buf { void *p = nullptr; ~Buf() { if (p) cudaFree(p); } };
  Buf bA, bv, bS, br, bR, bW, bSS, bL, be, bsl, bvv, bq, bh, bo;
  auto alloc = [&](Buf &b, size_t bytes) -> int {
    if (cudaMalloc(&b.p, bytes ? bytes : 8) != cudaSuccess) { cudaGetLastError(); return ctx->fail(RBSLAM_ECUDA, "op_ancestor_weights: out of device memory"); }
    return RBSLAM_OK;
  };
  const size_t MM = (size_t)M * M, ldl = chol_ldl(n);
  // form 0: the covariances and D' get the slabs' padded leading dimension, as in the sweeps (the cp.async
  // staging of k_dgemm_pipe needs 16-byte aligned columns)
  const size_t ldp = form == 0 ? (size_t)ctx->ld : (size_t)M;
  if ((rc = alloc(bA, ldp * M * N * 8)) || (rc = alloc(bv, (size_t)M * N * 8)) || (rc = alloc(bsl, (size_t)N * 8)) ||
      (rc = alloc(bvv, (size_t)N * 8)) || (rc = alloc(bo, (size_t)N * 8)) || (rc = alloc(bL, ldl * n * N * 8)))
    return rc;
  if (form == 0) {
    CK(cudaMemsetAsync(bA.p, 0, ldp * M * N * 8, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy2D(bA.p, ldp * 8, A, (size_t)M * 8, (size_t)M * 8, (size_t)M * N, cudaMemcpyHostToDevice));
  } else if ((rc = rb_h2d(ctx, bA.p, A, MM * N * 8))) {
    return rc;
  }
  if ((rc = rb_h2d(ctx, bv.p, v, (size_t)M * N * 8))) return rc;
  CholArgs c{};
  c.n = n; c.L = (double *)bL.p; c.ldl = (int)ldl; c.strideL = ldl * n; c.jitter = jitter;
  c.sum_log_diag = (double *)bsl.p; c.vtv = (double *)bvv.p; c.status = ctx->d_status; c.t = 0;
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(DevStatus), ctx->stream));
  if (form == 0) {
    // S = D [ne x M] (MATLAB layout) -> Dt [M x ne]; r = stacked future measurements
    std::vector<double> Dt(ldp * ne, 0.0);
    for (int j = 0; j < ne; ++j)
      for (int cidx = 0; cidx < M; ++cidx) Dt[cidx + (size_t)j * ldp] = S[j + (size_t)cidx * ne];
    if ((rc = alloc(bS, Dt.size() * 8)) || (rc = alloc(br, (size_t)ne * 8)) || (rc = alloc(bR, (size_t)d * d * 8)) ||
        (rc = alloc(bW, ldp * ne * N * 8)) || (rc = alloc(bSS, (size_t)ne * ne * N * 8)) ||
        (rc = alloc(be, (size_t)ne * N * 8)))
      return rc;
    if ((rc = rb_h2d(ctx, bS.p, Dt.data(), Dt.size() * 8)) || (rc = rb_h2d(ctx, br.p, r, (size_t)ne * 8)) ||
        (rc = rb_h2d(ctx, bR.p, R, (size_t)d * d * 8)))
      return rc;
    k_future_resid<<<dim3((ne + 3) / 4, N), 128, 0, ctx->stream>>>(M, ne, (const double *)bS.p, (int)ldp, (const double *)br.p,
                                                                  (const double *)bv.p, 0, (double *)be.p, ne);
    ctx->launches += 1;
    GemmArgs g1{};   // W = P_i * D'
    g1.m = M; g1.n = ne; g1.k = M; g1.A = (const double *)bA.p; g1.lda = (int)ldp; g1.strideA = ldp * M; g1.slotA = nullptr;
    g1.B = (const double *)bS.p; g1.ldb = (int)ldp; g1.strideB = 0;
    g1.C = (double *)bW.p; g1.ldc = (int)ldp; g1.strideC = ldp * ne; g1.Rblk = nullptr; g1.d = 1;
    GemmArgs g2{};   // SS = D * W + kron(I, R)
    g2.m = ne; g2.n = ne; g2.k = M; g2.A = (const double *)bS.p; g2.lda = (int)ldp; g2.strideA = 0; g2.slotA = nullptr;
    g2.B = (const double *)bW.p; g2.ldb = (int)ldp; g2.strideB = ldp * ne;
    g2.C = (double *)bSS.p; g2.ldc = ne; g2.strideC = (size_t)ne * ne; g2.Rblk = (const double *)bR.p; g2.d = d;
    g2.lower = 1;
    if ((rc = launch_gemm(ctx, false, g1, N)) || (rc = launch_gemm(ctx, true, g2, N))) return rc;
    c.A1 = (const double *)bSS.p; c.lda1 = ne; c.strideA1 = (size_t)ne * ne; c.slot1 = nullptr; c.A2 = nullptr; c.lda2 = 0;
    c.rhs = (const double *)be.p; c.stride_rhs = ne; c.rhs2 = nullptr;
  } else {
    // A = Imat [M x M x N], v = ivec [M x N], S = ImatAddt [M x M], r = ivecAddt [M]
    if ((rc = alloc(bS, MM * 8)) || (rc = alloc(br, (size_t)M * 8)) || (rc = alloc(bq, (size_t)N * 8)) ||
        (rc = alloc(bh, (size_t)N * 8)))
      return rc;
    if ((rc = rb_h2d(ctx, bS.p, S, MM * 8)) || (rc = rb_h2d(ctx, br.p, r, (size_t)M * 8)) ||
        (rc = rb_h2d(ctx, bq.p, q2, (size_t)N * 8)) || (rc = rb_h2d(ctx, bh.p, hld, (size_t)N * 8)))
      return rc;
    c.A1 = (const double *)bA.p; c.lda1 = M; c.strideA1 = MM; c.slot1 = nullptr; c.A2 = (const double *)bS.p; c.lda2 = M;
    c.rhs = (const double *)bv.p; c.stride_rhs = M; c.rhs2 = (const double *)br.p;
  }
  if ((rc = launch_chol(ctx, c, N))) return rc;
  k_aw_combine<<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, form, (double)ne, (const double *)bsl.p, (const double *)bvv.p,
                                                         (const double *)bq.p, (const double *)bh.p, (double *)bo.p);
  ctx->launches += 1;
  CK(cudaGetLastError());
  if ((rc = rb_d2h(ctx, logwMeas, bo.p, (size_t)N * 8))) return rc;
  return rb_check_status(ctx);
}
