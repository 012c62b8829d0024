// Per-step O(N) kernels: multinomial resampling (tools/sample.m), in-place slot
// planning for the covariance slabs, pose propagation, measurement Jacobians and
// log-sum-exp normalisation.
#pragma once
#include "models.cuh"

namespace rb {

// ---------------------------------------------------------------------------
// K5a  resample: wc = cumsum(w) in strict left-to-right fp64 order (bit-exact
// with MATLAB's cumsum / tools/sample.m:30), then ind = sum(wc < u) per draw.
// Single CTA: the scan is a serial dependency chain by contract.
// ---------------------------------------------------------------------------
struct RngSrc {
  const double *U;   // injected uniforms for this (sweep, step): [N], or nullptr
  uint64_t seed;
  uint32_t sweep, t;
};

__global__ void __launch_bounds__(1024)
k_resample(int N, int i0, int n_draws, const double *__restrict__ w, double *__restrict__ wc,
           RngSrc rng, const int *__restrict__ forced, int *__restrict__ ai,
           DevStatus *status, int only_if_ambiguous = 0) {
  if (only_if_ambiguous && !status->scan_ambig) return;   // the fast path of this step stands (k_scan_approx)
  // draws for particles i0 .. i0+n_draws-1 (uniform U[i] / Philox counter i, result ai[i]).
  // The scan runs chunk by chunk through shared memory: all threads stage a chunk, ONE thread
  // adds it up left to right (the rounding order is the contract), all threads write it back.
  extern __shared__ __align__(16) double s_buf[];
  __shared__ double s_carry;
  unsigned dyn;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
  const int chunk = (int)(dyn / sizeof(double));
  if (threadIdx.x == 0) s_carry = 0.0;
  for (int c0 = 0; c0 < N; c0 += chunk) {
    const int cn = min(chunk, N - c0);
    for (int j = threadIdx.x; j < cn; j += blockDim.x) s_buf[j] = w[c0 + j];
    __syncthreads();
    if (threadIdx.x == 0) {
      // the DADD chain is the critical path: the 8 addends of the NEXT block are loaded (as double2)
      // while the chain of this block runs, results leave as double2 stores.  (64 registers per
      // thread at 1024 threads: blocks of 8, not more.)
      double acc = s_carry;
      int j = 0;
      double2 nx[4];
      if (cn >= 8) {
#pragma unroll
        for (int q = 0; q < 4; ++q) nx[q] = reinterpret_cast<const double2 *>(s_buf)[q];
      }
      for (; j + 8 <= cn; j += 8) {
        double2 cur[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = nx[q];
        if (j + 16 <= cn) {
#pragma unroll
          for (int q = 0; q < 4; ++q) nx[q] = reinterpret_cast<const double2 *>(s_buf + j + 8)[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc += cur[q].x; cur[q].x = acc;
          acc += cur[q].y; cur[q].y = acc;
          reinterpret_cast<double2 *>(s_buf + j)[q] = cur[q];
        }
      }
      for (; j < cn; ++j) { acc += s_buf[j]; s_buf[j] = acc; }
      s_carry = acc;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < cn; j += blockDim.x) wc[c0 + j] = s_buf[j];
    __syncthreads();
  }
  if (n_draws < 0) return;                         // scan only: the draws run in k_resample_search
  const double *buf = (N <= chunk) ? s_buf : wc;   // single chunk: search in shared memory
  for (int i = i0 + threadIdx.x; i < i0 + n_draws; i += blockDim.x) {
    int idx;
    if (forced != nullptr) {
      idx = forced[i];
    } else {
      const double u = rng.U ? rng.U[i] : philox_uniform(rng.seed, rng.sweep, rng.t, i);
      // count of wc < u == lower bound (wc is non-decreasing)
      int lo = 0, hi = N;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (buf[mid] < u) lo = mid + 1; else hi = mid;
      }
      idx = lo;
      if (idx >= N) {  // reference would raise an index error (u > wc(end))
        idx = N - 1;
        atomicAdd(&status->clamp_sample, 1);
      }
    }
    ai[i] = idx;
  }
}

// the draws alone, one thread per draw over the whole grid (large N: the single scanning CTA
// would otherwise also do N binary searches)
__global__ void k_resample_search(int N, int i0, int n_draws, const double *__restrict__ wc, RngSrc rng,
                                  const int *__restrict__ forced, int *__restrict__ ai, DevStatus *status,
                                  int only_if_ambiguous = 0) {
  if (only_if_ambiguous) {
    if (!status->scan_ambig) {          // the fast path stands: its clamp count becomes the step's
      if (blockIdx.x == 0 && threadIdx.x == 0 && status->clamp_fast) atomicAdd(&status->clamp_sample, status->clamp_fast);
      return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&status->scan_fallbacks, 1);
  }
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i0 + n_draws) return;
  int idx;
  if (forced != nullptr) {
    idx = forced[i];
  } else {
    const double u = rng.U ? rng.U[i] : philox_uniform(rng.seed, rng.sweep, rng.t, i);
    int lo = 0, hi = N;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (wc[mid] < u) lo = mid + 1; else hi = mid;
    }
    idx = lo;
    if (idx >= N) { idx = N - 1; atomicAdd(&status->clamp_sample, 1); }
  }
  ai[i] = idx;
}

// ---------------------------------------------------------------------------
// K5a, fast path for large populations.  The strict left-to-right scan is one dependent fp64 add per weight
// (0.41 ms at 80 000 weights: the contract of `cumsum`, not parallelisable).  But an ancestor index is
// count(wc < u), and it can only depend on the rounding order of the sums if u lies within the rounding error
// of some wc(j).  So: (1) k_scan_approx -- a PARALLEL prefix sum wc' (blocked: per-thread segments + a tree over the threads);
// (2) k_search_checked -- idx = count(wc' < u) and the proof that the sequential scan gives the same index:
// wc'(idx-1) + delta < u <= wc'(idx) - delta, delta = eps ((idx+2) wc'(idx) + depth sum(w)) bounding both
// summations (wc is non-decreasing, so the two neighbours decide for all j); a draw that cannot prove itself sets scan_ambig; (3) the exact pair
// k_resample / k_resample_search runs ONLY IF the flag is set (a few per cent of the steps at N = 80 000; always for
// adversarial draws placed on a boundary, NaN weights, ...).  The ancestors are bit-identical to the sequential
// path in every case.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_scan_approx(int N, const double *__restrict__ w, double *__restrict__ wc, DevStatus *status) {
  __shared__ double s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { status->scan_ambig = 0; status->clamp_fast = 0; }
  const int per = (N + 1023) / 1024;
  const int b = min(N, tid * per), e = min(N, b + per);
  double s = 0.0;
  for (int i = b; i < e; ++i) s += w[i];
  double x = s;                                   // inclusive scan over the threads
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[wid] = x;
  __syncthreads();
  if (wid == 0) {
    double t = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    s_warp[lane] = t;
  }
  __syncthreads();
  double excl = __shfl_up_sync(0xffffffffu, x, 1);          // exclusive prefix of this thread's segment
  if (lane == 0) excl = 0.0;
  double run = (wid ? s_warp[wid - 1] : 0.0) + excl;
  for (int i = b; i < e; ++i) { run += w[i]; wc[i] = run; }
}

__global__ void k_search_checked(int N, int i0, int n_draws, const double *__restrict__ wc, RngSrc rng,
                                 const int *__restrict__ forced, int *__restrict__ ai, DevStatus *status) {
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i0 + n_draws) return;
  if (forced != nullptr) { ai[i] = forced[i]; return; }
  const double u = rng.U ? rng.U[i] : philox_uniform(rng.seed, rng.sweep, rng.t, i);
  int lo = 0, hi = N;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (wc[mid] < u) lo = mid + 1; else hi = mid;
  }
  const int idx = lo;
  // |sequential wc(j) - exact| <= eps * sum_{i<=j} wc(i) <= eps (j+1) wc(j)   (non-negative terms, non-decreasing sums);
  // |wc'(j) - exact| <= depth * eps * sum(w), depth <= 2 ceil(N/1024) + 12 additions on any path of k_scan_approx.
  // Both neighbours use the bound at the larger index; 5 % for the second-order terms.
  const double per = (double)((N + 1023) / 1024);
  const double delta = 1.05 * 1.1102230246251565e-16 *
                       ((double)(idx + 2) * wc[min(idx, N - 1)] + (2.0 * per + 16.0) * wc[N - 1]);
  // written so that any NaN makes the draw ambiguous
  const bool below = idx == 0 || (wc[idx - 1] + delta < u);
  const bool above = idx == N || (wc[idx] - delta >= u);
  if (!(below && above && delta >= 0.0)) status->scan_ambig = 1;
  if (idx >= N) { ai[i] = N - 1; atomicAdd(&status->clamp_fast, 1); }
  else ai[i] = idx;
}

// ---------------------------------------------------------------------------
// K5b  slot planning.  Covariance slabs are resampled IN PLACE: the first
// offspring (smallest i) of an ancestor keeps the ancestor's physical slab
// (src == dst); every further offspring is assigned the slab of a particle that
// died (src != dst).  Lists:  listA = offspring that copy (processed first, they
// only read surviving slabs and write dead ones), listB = offspring in place.
// Single CTA, deterministic.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int block_excl_scan(int v, int *s_tmp, int &total) {
  // exclusive scan of one int per thread across the block (blockDim <= 1024)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_tmp[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int t = (lane < (int)((blockDim.x + 31) >> 5)) ? s_tmp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    s_tmp[32 + lane] = t;  // inclusive
  }
  __syncthreads();
  const int base = wid == 0 ? 0 : s_tmp[32 + wid - 1];
  total = s_tmp[32 + ((blockDim.x + 31) >> 5) - 1];
  __syncthreads();
  return base + x - v;
}

__global__ void __launch_bounds__(1024)
k_plan_slots(int N, const int *__restrict__ ai, const int *__restrict__ slot_old,
             int *__restrict__ slot_new, int *__restrict__ src_slot, int *__restrict__ first_child,
             int *__restrict__ free_list, int *__restrict__ listA, int *__restrict__ listB,
             int *__restrict__ counts) {
  __shared__ int s_tmp[64];
  for (int a = threadIdx.x; a < N; a += blockDim.x) first_child[a] = 0x7fffffff;
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) atomicMin(&first_child[ai[i]], i);
  __syncthreads();
  // contiguous chunk per thread so ranks are ordered by index
  const int per = (N + blockDim.x - 1) / blockDim.x;
  const int b = min(N, (int)threadIdx.x * per), e = min(N, b + per);
  int n_free = 0, n_copy = 0;
  for (int a = b; a < e; ++a) n_free += (first_child[a] == 0x7fffffff);
  for (int i = b; i < e; ++i) n_copy += (first_child[ai[i]] != i);
  int tot_free, tot_copy;
  int off_free = block_excl_scan(n_free, s_tmp, tot_free);
  int off_copy = block_excl_scan(n_copy, s_tmp, tot_copy);
  for (int a = b; a < e; ++a)
    if (first_child[a] == 0x7fffffff) free_list[off_free++] = slot_old[a];
  __syncthreads();
  int off_keep = b - off_copy;  // in-place offspring before this chunk
  for (int i = b; i < e; ++i) {
    const int a = ai[i];
    const int s = slot_old[a];
    src_slot[i] = s;
    if (first_child[a] == i) {
      slot_new[i] = s;
      listB[off_keep++] = i;
    } else {
      slot_new[i] = free_list[off_copy];
      listA[off_copy++] = i;
    }
  }
  if (threadIdx.x == 0) {
    counts[0] = tot_copy;       // nA
    counts[1] = N - tot_copy;   // nB
  }
}

// identity plan for the first time step (no resampling): everything in place
__global__ void k_plan_identity(int N, const int *__restrict__ slot, int *__restrict__ src_slot,
                                int *__restrict__ listB, int *__restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { src_slot[i] = slot[i]; listB[i] = i; }
  if (i == 0) { counts[0] = 0; counts[1] = N; }
}

// ---------------------------------------------------------------------------
// K1  propagate: xn_new(:,i) = dynModel(xn_old(:,ai(i)), odometry(t-1,:), dt, Q)
// (src/particleFilter.m:104-109).  One thread per particle.
// ---------------------------------------------------------------------------
struct NormalSrc {
  const double *Z;   // injected normals for this (sweep, step): [nz x N] or nullptr
  uint64_t seed;
  uint32_t sweep, t;
};

__global__ void k_propagate(ModelConsts mc, int N, int n_prop, const double *__restrict__ xn_old,
                            const int *__restrict__ ai, const double *__restrict__ dx, double dt,
                            const double *__restrict__ Q, NormalSrc nsrc,
                            double *__restrict__ xn_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_prop) return;
  double z[6] = {0, 0, 0, 0, 0, 0};
  if (nsrc.Z) {
    for (int j = 0; j < mc.nz; ++j) z[j] = nsrc.Z[j + (size_t)i * mc.nz];
  } else {
    for (int p = 0; 2 * p < mc.nz; ++p)
      philox_normal_pair(nsrc.seed, nsrc.sweep, nsrc.t, i, p, z[2 * p], z[2 * p + 1]);
  }
  double xin[7], xo[7], dxl[7];
  const int a = ai[i];
  for (int j = 0; j < mc.n; ++j) xin[j] = xn_old[j + (size_t)a * mc.n];
  for (int j = 0; j < mc.n_odo; ++j) dxl[j] = dx[j];
  dyn_model(mc, xin, dxl, dt, Q, z, xo);
  for (int j = 0; j < mc.n; ++j) xn_new[j + (size_t)i * mc.n] = xo[j];
}

// logwDyn[i] = -0.5*||dynResNorm(xnk_t, xn(:,i), ...)||^2
__global__ void k_dyn_logweight(ModelConsts mc, int N, const double *__restrict__ xnk_t,
                                const double *__restrict__ xn, const double *__restrict__ dx,
                                double dt, const double *__restrict__ Q, int use_default,
                                double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double xk[7], xi[7], dxl[7];
  for (int j = 0; j < mc.n; ++j) { xk[j] = xnk_t[j]; xi[j] = xn[j + (size_t)i * mc.n]; }
  for (int j = 0; j < mc.n_odo; ++j) dxl[j] = dx[j];
  out[i] = dyn_logweight(mc, xk, xi, dxl, dt, Q, use_default != 0);
}

// ---------------------------------------------------------------------------
// K2  measurement Jacobian H_i [d x M] (row-contiguous, leading dim ldh) at each
// particle's pose.  One CTA per particle.  Separable eigenbasis: sin/cos tables
// per dimension in shared memory, then one product per basis function.
//   dense-mag  : H = Rnb(q)' * [I3, dPhi/dx; dPhi/dy; dPhi/dz]  (run_dense3D_magfield.m:265-279)
//   dense-radio: H = Phi(pos)                                    (run_dense2D_withHeading.m:168)
//   sparse     : pinhole projection + Jacobian wrt landmarks     (measurement.m:32-84)
// ---------------------------------------------------------------------------
#define RB_MAXTAB 96
__global__ void __launch_bounds__(128)
k_meas(ModelConsts mc, int N, const double *__restrict__ xn, const double *__restrict__ xl,
       int ldxl, const int *__restrict__ xl_index, double *__restrict__ H, size_t hs_p, int hs_a,
       int hs_c, int ldpad, double *__restrict__ yhat, const int *__restrict__ pidx = nullptr) {
  const int i = blockIdx.x;
  if (i >= N) return;
  __shared__ double s_sin[3][RB_MAXTAB], s_cos[3][RB_MAXTAB];
  __shared__ double s_x[7];
  // pidx: pose index of item i (sharded engine: item = local slab, pose = global particle)
  if (threadIdx.x < mc.n) s_x[threadIdx.x] = xn[threadIdx.x + (size_t)(pidx ? pidx[i] : i) * mc.n];
  __syncthreads();
  // H_i(a, c) lives at Hi[a*hs_a + c*hs_c]; columns M..ldpad-1 are zero padding
  double *Hi = H + (size_t)i * hs_p;
  if (mc.family == FAM_SPARSE_VISUAL2D) {
    // measurement.m:36-50 (projection) and :59-79 (Jacobian wrt map)
    const double *xli = xl + (size_t)(xl_index ? xl_index[i] : i) * ldxl;
    double s, c;
    sincos(s_x[2], &s, &c);
    for (int idx = threadIdx.x; idx < mc.d * ldpad; idx += blockDim.x)
      Hi[(size_t)(idx / ldpad) * hs_a + (size_t)(idx % ldpad) * hs_c] = 0.0;
    __syncthreads();
    for (int l = threadIdx.x; l < mc.m; l += blockDim.x) {
      const double m1 = xli[2 * l], m2 = xli[2 * l + 1];
      const double p1 = s_x[0], p2 = s_x[1];
      // u = K*[R' -R'*p]*[map;1],  R = [c -s; s c]
      const double t1 = -(c * p1 + s * p2), t2 = -(-s * p1 + c * p2);
      const double lx = c * m1 + s * m2 + t1;
      const double ly = -s * m1 + c * m2 + t2;
      const double u1 = mc.cam_f * lx + mc.cam_fp * ly;
      yhat[l + (size_t)i * mc.d] = u1 / ly;
      const double dv = m2 * c - p2 * c - m1 * s + p1 * s;
      const double div = dv * dv;
      Hi[(size_t)l * hs_a + (size_t)(2 * l) * hs_c] = (mc.cam_f * (m2 - p2)) / div;
      Hi[(size_t)l * hs_a + (size_t)(2 * l + 1) * hs_c] = -(mc.cam_f * (m1 - p1)) / div;
    }
    return;
  }
  // sin/cos tables: arg = pi*n*(x+L)/(2L)  (tools/domain_cartesian_dx.m:88-91)
  for (int idx = threadIdx.x; idx < mc.dim * RB_MAXTAB; idx += blockDim.x) {
    const int j = idx / RB_MAXTAB, nn = idx % RB_MAXTAB;
    if (nn >= 1 && nn <= mc.maxn[j]) {
      const double arg = (RB_PI * (double)nn) * (s_x[j] + mc.L[j]) / (2.0 * mc.L[j]);
      double s, c;
      sincos(arg, &s, &c);
      s_sin[j][nn] = s;
      s_cos[j][nn] = c;
    }
  }
  __syncthreads();
  if (mc.family == FAM_DENSE_RADIO2D) {
    const double r0 = sqrt(mc.L[0]), r1 = sqrt(mc.L[1]);
    for (int cidx = threadIdx.x; cidx < mc.m; cidx += blockDim.x) {
      const int n0 = mc.NN[cidx], n1 = mc.NN[cidx + mc.m];
      double v = 1.0;
      v = v * 1.0 / r0 * s_sin[0][n0];
      v = v * 1.0 / r1 * s_sin[1][n1];
      Hi[(size_t)cidx * hs_c] = v;
    }
    for (int cidx = mc.m + threadIdx.x; cidx < ldpad; cidx += blockDim.x) Hi[(size_t)cidx * hs_c] = 0.0;
    return;
  }
  // dense-mag
  double Rnb[3][3];
  quat2rmat(s_x + 3, Rnb);
  const double rL[3] = {sqrt(mc.L[0]), sqrt(mc.L[1]), sqrt(mc.L[2])};
  for (int cidx = threadIdx.x; cidx < ldpad; cidx += blockDim.x) {
    double g[3];
    if (cidx < 3) {
      g[0] = cidx == 0; g[1] = cidx == 1; g[2] = cidx == 2;
    } else if (cidx < mc.M) {
      const int b = cidx - 3;
      const int nn[3] = {mc.NN[b], mc.NN[b + mc.m], mc.NN[b + 2 * mc.m]};
#pragma unroll
      for (int di = 0; di < 3; ++di) {
        double v = 1.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j == di)  // tools/domain_cartesian_dx.m:153-154
            v = v * RB_PI * (double)nn[j] / (2.0 * mc.L[j] * rL[j]) * s_cos[j][nn[j]];
          else          // :157-158
            v = v * 1.0 / rL[j] * s_sin[j][nn[j]];
        }
        g[di] = v;
      }
    } else {
      g[0] = g[1] = g[2] = 0.0;
    }
    // H(a, c) = sum_b Rnb(b, a) * g_b   (Rnb' * dPhi)
#pragma unroll
    for (int a = 0; a < 3; ++a)
      Hi[(size_t)a * hs_a + (size_t)cidx * hs_c] = Rnb[0][a] * g[0] + Rnb[1][a] * g[1] + Rnb[2][a] * g[2];
  }
}

// ---------------------------------------------------------------------------
// K4  normalise (src/particleFilter.m:153-161): c=max, lse, w=exp(logw-lse),
// iw_max = first argmax(w), traj_max(:,t), traj_mean(:,t) = sum(xn.*w,2).
// Single CTA with a fixed reduction tree: the result does not depend on how
// many GPUs produced logw.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_normalize(int N, int n, const double *__restrict__ logw, double *__restrict__ w,
            const double *__restrict__ xn, double *__restrict__ traj_max_t,
            double *__restrict__ traj_mean_t, int *__restrict__ iw_max_out,
            double *__restrict__ logw_hist_t, double *__restrict__ w_hist_t) {
  __shared__ double s_red[32];
  __shared__ int s_idx[32];
  __shared__ double s_bcast;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double m = -INFINITY;
  {
    double m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int i = threadIdx.x;
    for (; i + 3 * (int)blockDim.x < N; i += 4 * blockDim.x) {
#pragma unroll
      for (int u = 0; u < 4; ++u) m4[u] = fmax(m4[u], logw[i + u * blockDim.x]);
    }
    for (; i < N; i += blockDim.x) m4[0] = fmax(m4[0], logw[i]);
    m = fmax(fmax(m4[0], m4[1]), fmax(m4[2], m4[3]));
  }
  m = warp_max(m);
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  if (wid == 0) {
    double v = lane < nwarp ? s_red[lane] : -INFINITY;
    v = warp_max(v);
    if (lane == 0) s_bcast = v;
  }
  __syncthreads();
  const double c = s_bcast;
  double s = 0.0;
  {   // four independent exp chains per thread (the single CTA is latency-bound, not throughput-bound);
      // the summation order is fixed by (blockDim, N) only, so it is the same on every GPU
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
    int i = threadIdx.x;
    for (; i + 3 * (int)blockDim.x < N; i += 4 * blockDim.x) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s4[u] += exp(logw[i + u * blockDim.x] - c);
    }
    for (; i < N; i += blockDim.x) s4[0] += exp(logw[i] - c);
    s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) s_red[wid] = s;
  __syncthreads();
  if (wid == 0) {
    double v = lane < nwarp ? s_red[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) s_bcast = c + log(v);
  }
  __syncthreads();
  const double lse = s_bcast;
  double best = -1.0;
  int bidx = 0x7fffffff;
  {
    int i = threadIdx.x;
    for (; i + 3 * (int)blockDim.x < N; i += 4 * blockDim.x) {
      double wi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) wi[u] = exp(logw[i + u * blockDim.x] - lse);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ii = i + u * blockDim.x;
        w[ii] = wi[u];
        if (logw_hist_t) logw_hist_t[ii] = logw[ii];
        if (w_hist_t) w_hist_t[ii] = wi[u];
        if (wi[u] > best) { best = wi[u]; bidx = ii; }   // ascending ii: first index kept on ties
      }
    }
    for (; i < N; i += blockDim.x) {
      const double wi = exp(logw[i] - lse);
      w[i] = wi;
      if (logw_hist_t) logw_hist_t[i] = logw[i];
      if (w_hist_t) w_hist_t[i] = wi;
      if (wi > best) { best = wi; bidx = i; }
    }
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
    // every weight NaN (chol failed even with jitter: the status word reports it): no comparison
    // succeeded.  MATLAB's max returns index 1 for an all-NaN vector; never leave an out-of-range
    // index behind for the kernels that gather with iw_max
    if (bidx < 0 || bidx >= N) bidx = 0;
    if (lane == 0) { s_idx[0] = bidx; if (iw_max_out) *iw_max_out = bidx; }
  }
  __syncthreads();
  const int imax = s_idx[0];
  if (xn == nullptr) return;
  if (threadIdx.x < n && traj_max_t) traj_max_t[threadIdx.x] = xn[threadIdx.x + (size_t)imax * n];
  // weighted mean: one warp per state component, fixed order
  if (traj_mean_t) {
    for (int j = wid; j < n; j += nwarp) {
      double acc = 0.0;
      for (int i = lane; i < N; i += 32) acc += xn[j + (size_t)i * n] * w[i];
      acc = warp_sum(acc);
      if (lane == 0) traj_mean_t[j] = acc;
    }
  }
}

// ---------------------------------------------------------------------------
// K4 for large populations (sharded filter, N in the tens of thousands): the same normalisation in
// fixed chunks of RB_NCHUNK weights, one CTA per chunk, partial results combined in chunk order --
// deterministic for a given N on every GPU (all ranks of a sharded filter normalise the same replicated
// log-weights and must agree bit for bit), but no longer one CTA walking 80 000 weights three times.
//   pass 1: chunk maxima;  pass 2: c = max, chunk sums of exp(logw - c);
//   pass 3: lse = c + log(sum), w = exp(logw - lse), chunk arg-max (first index) and chunk sums xn.*w;
//   final : arg-max over the chunks (first index on ties), traj_max, traj_mean.
// ---------------------------------------------------------------------------
#define RB_NCHUNK 4096
#define RB_NCHUNK_MAX 256
struct NormWs {
  double *pmax, *psum, *pbest, *pmean;   // [nchunk], [nchunk], [nchunk], [nchunk][8]
  int *pidx;                             // [nchunk]
};
__device__ __forceinline__ double block_reduce_1024(double v, bool is_max, double *s_red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double x = lane < (int)(blockDim.x >> 5) ? s_red[lane] : (is_max ? -INFINITY : 0.0);
    x = is_max ? warp_max(x) : warp_sum(x);
    if (lane == 0) s_red[32] = x;
  }
  __syncthreads();
  return s_red[32];
}
__global__ void __launch_bounds__(1024) k_norm_max(int N, const double *__restrict__ logw, NormWs ws) {
  __shared__ double s_red[33];
  const int c0 = blockIdx.x * RB_NCHUNK, c1 = min(N, c0 + RB_NCHUNK);
  double m = -INFINITY;
  for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) m = fmax(m, logw[i]);
  m = block_reduce_1024(m, true, s_red);
  if (threadIdx.x == 0) ws.pmax[blockIdx.x] = m;
}
__global__ void __launch_bounds__(1024) k_norm_sum(int N, const double *__restrict__ logw, NormWs ws) {
  __shared__ double s_red[33];
  double c = -INFINITY;
  for (int q = 0; q < (int)gridDim.x; ++q) c = fmax(c, ws.pmax[q]);
  const int c0 = blockIdx.x * RB_NCHUNK, c1 = min(N, c0 + RB_NCHUNK);
  double s = 0.0;
  for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) s += exp(logw[i] - c);
  s = block_reduce_1024(s, false, s_red);
  if (threadIdx.x == 0) ws.psum[blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024)
k_norm_write(int N, int n, const double *__restrict__ logw, double *__restrict__ w, const double *__restrict__ xn,
             NormWs ws, double *__restrict__ logw_hist_t, double *__restrict__ w_hist_t) {
  __shared__ double s_red[33];
  __shared__ double s_b[32];
  __shared__ int s_i[32];
  double c = -INFINITY, tot = 0.0;
  for (int q = 0; q < (int)gridDim.x; ++q) c = fmax(c, ws.pmax[q]);
  for (int q = 0; q < (int)gridDim.x; ++q) tot += ws.psum[q];
  const double lse = c + log(tot);
  const int c0 = blockIdx.x * RB_NCHUNK, c1 = min(N, c0 + RB_NCHUNK);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double best = -1.0;
  int bidx = 0x7fffffff;
  for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
    const double wi = exp(logw[i] - lse);
    w[i] = wi;
    if (logw_hist_t) logw_hist_t[i] = logw[i];
    if (w_hist_t) w_hist_t[i] = wi;
    if (wi > best) { best = wi; bidx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  if (lane == 0) { s_b[wid] = best; s_i[wid] = bidx; }
  __syncthreads();
  if (wid == 0) {
    best = lane < nwarp ? s_b[lane] : -1.0;
    bidx = lane < nwarp ? s_i[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { ws.pbest[blockIdx.x] = best; ws.pidx[blockIdx.x] = bidx; }
  }
  if (xn != nullptr) {   // chunk contribution to traj_mean = sum(xn .* w, 2)
    for (int j = 0; j < n; ++j) {
      double acc = 0.0;
      for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) acc += xn[j + (size_t)i * n] * w[i];
      acc = block_reduce_1024(acc, false, s_red);
      if (threadIdx.x == 0) ws.pmean[(size_t)blockIdx.x * 8 + j] = acc;
    }
  }
}
__global__ void k_norm_final(int N, int n, int nchunk, const double *__restrict__ xn, NormWs ws,
                             double *__restrict__ traj_max_t, double *__restrict__ traj_mean_t,
                             int *__restrict__ iw_max_out) {
  __shared__ int s_imax;
  if (threadIdx.x == 0) {
    double best = -1.0;
    int bidx = 0x7fffffff;
    for (int q = 0; q < nchunk; ++q)
      if (ws.pbest[q] > best || (ws.pbest[q] == best && ws.pidx[q] < bidx)) { best = ws.pbest[q]; bidx = ws.pidx[q]; }
    if (bidx < 0 || bidx >= N) bidx = 0;     // all-NaN weights: MATLAB's max returns index 1
    s_imax = bidx;
    if (iw_max_out) *iw_max_out = bidx;
  }
  __syncthreads();
  if (xn == nullptr) return;
  if ((int)threadIdx.x < n) {
    if (traj_max_t) traj_max_t[threadIdx.x] = xn[threadIdx.x + (size_t)s_imax * n];
    if (traj_mean_t) {
      double acc = 0.0;
      for (int q = 0; q < nchunk; ++q) acc += ws.pmean[(size_t)q * 8 + threadIdx.x];
      traj_mean_t[threadIdx.x] = acc;
    }
  }
}

// ---------------------------------------------------------------------------
// K2b  JacobianPhi3D (tools/JacobianPhi3D.m:29-64): Hessian of every basis function at
// every point, J [3 x 3 x m x N].  Same separable structure as K2: per point and axis one
// sin and one cos per distinct index (<= 3*RB_MAXTAB sincos calls instead of 6*m trig).
//   f_d  = pi*j_d/(b_d-a_d)                          (:41-44)
//   s_d  = sin(pi*j_d*(x_d-a_d)/(b_d-a_d)) * mult_d  (:51-56), mult_d = 1/sqrt((b_d-a_d)/2)
//   J_dd = -f_d^2 s1 s2 s3;  J_de = f_d f_e * (cos on axes d,e; sin on the third)  (:58-66)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_jacobian_phi3d(ModelConsts mc, int N, const double *__restrict__ x, double a0, double a1, double a2,
                 double b0, double b1, double b2, double *__restrict__ J) {
  const int i = blockIdx.x;
  if (i >= N) return;
  __shared__ double s_sin[3][RB_MAXTAB], s_cos[3][RB_MAXTAB];
  const double a[3] = {a0, a1, a2}, b[3] = {b0, b1, b2};
  for (int idx = threadIdx.x; idx < 3 * RB_MAXTAB; idx += blockDim.x) {
    const int dd = idx / RB_MAXTAB, nn = idx % RB_MAXTAB;
    if (nn >= 1 && nn <= mc.maxn[dd]) {
      const double core = (RB_PI * (double)nn) * (x[dd + 3 * (size_t)i] - a[dd]) / (b[dd] - a[dd]);
      const double mult = 1.0 / sqrt(0.5 * (b[dd] - a[dd]));
      double sv, cv;
      sincos(core, &sv, &cv);
      s_sin[dd][nn] = sv * mult;
      s_cos[dd][nn] = cv * mult;
    }
  }
  __syncthreads();
  for (int bidx = threadIdx.x; bidx < mc.m; bidx += blockDim.x) {
    double f[3], sn[3], cs[3];
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
      const int nn = mc.NN[bidx + dd * mc.m];
      f[dd] = (RB_PI * (double)nn) / (b[dd] - a[dd]);
      sn[dd] = s_sin[dd][nn];
      cs[dd] = s_cos[dd][nn];
    }
    double *o = J + 9 * ((size_t)bidx + (size_t)mc.m * i);   // J(r,c,bidx,i) at o[r + 3c]
    const double j12 = f[0] * f[1] * cs[0] * cs[1] * sn[2];
    const double j13 = f[0] * f[2] * cs[0] * sn[1] * cs[2];
    const double j23 = f[1] * f[2] * sn[0] * cs[1] * cs[2];
    o[0] = -(f[0] * f[0]) * sn[0] * sn[1] * sn[2];
    o[1] = f[1] * f[0] * cs[0] * cs[1] * sn[2];
    o[2] = f[2] * f[0] * cs[0] * sn[1] * cs[2];
    o[3] = j12;
    o[4] = -(f[1] * f[1]) * sn[0] * sn[1] * sn[2];
    o[5] = f[2] * f[1] * sn[0] * cs[1] * cs[2];
    o[6] = j13;
    o[7] = j23;
    o[8] = -(f[2] * f[2]) * sn[0] * sn[1] * sn[2];
  }
}

}  // namespace rb
