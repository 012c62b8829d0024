// K3: fused resampling gather + log-weight + Kalman map update
//   logw_i, xl_i += K e, P_i -= K SS K'      (src/particleFilter.m:126-151,164-204)
// for every particle, reading the ancestor's slab (src) and writing the
// particle's own slab (dst).  Two families of kernels:
//   * k_kalman_small : M*M*8 B fits one CTA's shared memory (C2 radio M=128, C3
//     visual M=40): single pass, exactly 16*M^2 B of HBM traffic per particle;
//     supports d<=32 measurements with NaN (unobserved) rows.
//   * k_ph / k_innov / k_downdate : streaming path for large M (C1 M=515, C4
//     M=1027), d<=4.
#pragma once
#include "common.cuh"

namespace rb {

struct KalmanArgs {
  int N, M, d, ld, ldh;
  size_t slab;              // doubles per covariance slab (ld*M)
  double *P;                // slabs
  const int *src_slot;      // [N] physical slab of the ancestor
  const int *dst_slot;      // [N] physical slab of the particle
  const double *xl_old;     // [M x N] logical order before resampling
  const int *ai;            // [N] ancestors (xl_old column), nullptr = identity
  double *xl_new;           // [M x N]
  const double *H;          // [N][d][ldh]
  const double *yhat;       // [d x N] (sparse family) or nullptr
  const double *y_t;        // [d]
  const double *R;          // [d x d] column-major
  double jitter;
  double *logw;             // [N]
  DevStatus *status;
  int t;
};

// ---------------------------------------------------------------------------
// small-M single-pass kernel
// ---------------------------------------------------------------------------
#define RB_DMAX 32
__global__ void __launch_bounds__(256)
k_kalman_small(KalmanArgs a, const int *__restrict__ list, const int *__restrict__ count) {
  extern __shared__ double sm[];
  const int M = a.M, d = a.d;
  double *sP = sm;                      // [M*M] column-major
  double *sH = sP + (size_t)M * M;      // [d][M]
  double *sPH = sH + (size_t)d * M;     // [d][M]   PH(r,a) = sum_c P(r,c) H(a,c)
  double *sG = sPH + (size_t)d * M;     // [d][M]   gain, compacted observed rows
  double *sKS = sG + (size_t)d * M;     // [d][M]   K*SS
  double *sxl = sKS + (size_t)d * M;    // [M]
  double *sS = sxl + M;                 // [32*32] chol factor (compact)
  double *sSS = sS + RB_DMAX * RB_DMAX; // [32*32] SS (compact, un-jittered)
  double *se = sSS + RB_DMAX * RB_DMAX; // [32]
  __shared__ int s_obs[RB_DMAX];
  __shared__ int s_nobs;
  const int n_items = *count;
  for (int p = blockIdx.x; p < n_items; p += gridDim.x) {
    const int i = list[p];
    const double *Ps = a.P + (size_t)a.src_slot[i] * a.slab;
    double *Pd = a.P + (size_t)a.dst_slot[i] * a.slab;
    const double *xls = a.xl_old + (size_t)(a.ai ? a.ai[i] : i) * M;
    const double *Hi = a.H + (size_t)i * d * a.ldh;
    for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) {
      const int r = idx % M, c = idx / M;
      sP[idx] = Ps[r + (size_t)c * a.ld];
    }
    for (int idx = threadIdx.x; idx < d * M; idx += blockDim.x)
      sH[idx] = Hi[(size_t)(idx / M) * a.ldh + (idx % M)];
    for (int r = threadIdx.x; r < M; r += blockDim.x) sxl[r] = xls[r];
    if (threadIdx.x == 0) {
      int n = 0;
      for (int q = 0; q < d; ++q)
        if (!isnan(a.y_t[q])) s_obs[n++] = q;   // src/particleFilter.m:134
      s_nobs = n;
    }
    __syncthreads();
    const int nobs = s_nobs;
    // PH for observed rows (compact index q)
    for (int idx = threadIdx.x; idx < nobs * M; idx += blockDim.x) {
      const int q = idx / M, r = idx % M;
      const double *h = sH + (size_t)s_obs[q] * M;
      double acc = 0.0;
      for (int c = 0; c < M; ++c) acc += sP[r + (size_t)c * M] * h[c];
      sPH[idx] = acc;
    }
    __syncthreads();
    // SS = H P H' + R (observed block), e = y - yhat
    for (int idx = threadIdx.x; idx < nobs * nobs + nobs; idx += blockDim.x) {
      if (idx < nobs * nobs) {
        const int qa = idx % nobs, qb = idx / nobs;
        const double *h = sH + (size_t)s_obs[qa] * M;
        const double *ph = sPH + (size_t)qb * M;
        double acc = 0.0;
        for (int c = 0; c < M; ++c) acc += h[c] * ph[c];
        acc += a.R[s_obs[qa] + s_obs[qb] * d];
        sSS[qa + qb * RB_DMAX] = acc;
        sS[qa + qb * RB_DMAX] = acc;
      } else {
        const int q = idx - nobs * nobs;
        const int oq = s_obs[q];
        double yh;
        if (a.yhat) {
          yh = a.yhat[oq + (size_t)i * d];
        } else {
          const double *h = sH + (size_t)oq * M;
          yh = 0.0;
          for (int c = 0; c < M; ++c) yh += h[c] * sxl[c];
        }
        se[q] = a.y_t[oq] - yh;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int flag = chol_small(sS, nobs, RB_DMAX);
      if (flag) {  // src/particleFilter.m:146-148
        for (int c = 0; c < nobs; ++c)
          for (int r = 0; r < nobs; ++r)
            sS[r + c * RB_DMAX] = sSS[r + c * RB_DMAX] + (r == c ? a.jitter : 0.0);
        atomicAdd(&a.status->used_jitter, 1);
        flag = chol_small(sS, nobs, RB_DMAX);
        if (flag && atomicCAS(&a.status->not_pd, 0, 1) == 0) {
          a.status->not_pd_step = a.t;
          a.status->not_pd_particle = i;
        }
      }
      // v = cS\e ; logw (src/particleFilter.m:149-150)
      double lw = 0.0, vv = 0.0;
      double v[RB_DMAX];
      for (int r = 0; r < nobs; ++r) {
        double s = se[r];
        for (int k = 0; k < r; ++k) s -= sS[r + k * RB_DMAX] * v[k];
        v[r] = s / sS[r + r * RB_DMAX];
        vv += v[r] * v[r];
        lw -= log(sS[r + r * RB_DMAX]);
      }
      a.logw[i] = lw - 0.5 * vv - 0.5 * nobs * RB_LOG2PI;
    }
    __syncthreads();
    // gain rows: G(r,:) = PH(r,:) / cS' / cS ; KS = G*SS ; xl += G e
    for (int r = threadIdx.x; r < M; r += blockDim.x) {
      for (int q = 0; q < nobs; ++q) {           // forward: cS z = ph
        double s = sPH[(size_t)q * M + r];
        for (int k = 0; k < q; ++k) s -= sS[q + k * RB_DMAX] * sG[(size_t)k * M + r];
        sG[(size_t)q * M + r] = s / sS[q + q * RB_DMAX];
      }
      for (int q = nobs - 1; q >= 0; --q) {      // backward: cS' g = z
        double s = sG[(size_t)q * M + r];
        for (int k = q + 1; k < nobs; ++k) s -= sS[k + q * RB_DMAX] * sG[(size_t)k * M + r];
        sG[(size_t)q * M + r] = s / sS[q + q * RB_DMAX];
      }
      double ge = 0.0;
      for (int q = 0; q < nobs; ++q) {
        ge += sG[(size_t)q * M + r] * se[q];
        double ks = 0.0;
        for (int k = 0; k < nobs; ++k) ks += sG[(size_t)k * M + r] * sSS[k + q * RB_DMAX];
        sKS[(size_t)q * M + r] = ks;
      }
      a.xl_new[(size_t)i * M + r] = sxl[r] + ge;   // src/particleFilter.m:197
    }
    __syncthreads();
    // P = P - K*SS*K'  (src/particleFilter.m:198)
    for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) {
      const int r = idx % M, c = idx / M;
      double acc = 0.0;
      for (int q = 0; q < nobs; ++q) acc += sKS[(size_t)q * M + r] * sG[(size_t)q * M + c];
      Pd[r + (size_t)c * a.ld] = sP[idx] - acc;
    }
    __syncthreads();
  }
}

static inline size_t kalman_small_smem(int M, int d) {
  return sizeof(double) * ((size_t)M * M + 4 * (size_t)d * M + M + 2 * RB_DMAX * RB_DMAX + RB_DMAX);
}

// ---------------------------------------------------------------------------
// large-M streaming path, d = D <= 4 (dense families)
// ---------------------------------------------------------------------------
#define RB_CWMAX 256
// partial PH over a column chunk: PHpart[i][s][a][r] = sum_{c in chunk s} P(r,c) H(a,c)
template <int D, int R2>
__global__ void __launch_bounds__(256)
k_ph(const double *__restrict__ P, size_t slab, int ld, int M, const int *__restrict__ src_slot,
     const double *__restrict__ H, int ldh, int nsplit, int cw, double *__restrict__ PHpart) {
  __shared__ double sH[D][RB_CWMAX];
  const int i = blockIdx.x, s = blockIdx.y;
  const int c0 = s * cw, c1 = min(M, c0 + cw);
  const double *Ps = P + (size_t)src_slot[i] * slab;
  for (int idx = threadIdx.x; idx < D * cw; idx += blockDim.x) {
    const int aa = idx / cw, c = idx % cw;
    sH[aa][c] = (c0 + c < c1) ? H[((size_t)i * D + aa) * ldh + c0 + c] : 0.0;
  }
  __syncthreads();
  const int npairs = ld >> 1;
  double2 acc[R2][D];
#pragma unroll
  for (int k = 0; k < R2; ++k)
#pragma unroll
    for (int aa = 0; aa < D; ++aa) acc[k][aa] = make_double2(0.0, 0.0);
  const int nc = c1 - c0;
  int c = 0;
  for (; c + 4 <= nc; c += 4) {
    double2 v[4][R2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < R2; ++k) {
        const int rp = threadIdx.x + k * blockDim.x;
        v[u][k] = rp < npairs
                      ? __ldg(reinterpret_cast<const double2 *>(Ps + (size_t)(c0 + c + u) * ld) + rp)
                      : make_double2(0.0, 0.0);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int aa = 0; aa < D; ++aa) {
        const double h = sH[aa][c + u];
#pragma unroll
        for (int k = 0; k < R2; ++k) {
          acc[k][aa].x = fma(v[u][k].x, h, acc[k][aa].x);
          acc[k][aa].y = fma(v[u][k].y, h, acc[k][aa].y);
        }
      }
  }
  for (; c < nc; ++c) {
#pragma unroll
    for (int k = 0; k < R2; ++k) {
      const int rp = threadIdx.x + k * blockDim.x;
      if (rp < npairs) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(Ps + (size_t)(c0 + c) * ld) + rp);
#pragma unroll
        for (int aa = 0; aa < D; ++aa) {
          const double h = sH[aa][c];
          acc[k][aa].x = fma(v.x, h, acc[k][aa].x);
          acc[k][aa].y = fma(v.y, h, acc[k][aa].y);
        }
      }
    }
  }
  double *out = PHpart + ((size_t)i * nsplit + s) * D * ld;
#pragma unroll
  for (int k = 0; k < R2; ++k) {
    const int rp = threadIdx.x + k * blockDim.x;
    if (rp < npairs) {
#pragma unroll
      for (int aa = 0; aa < D; ++aa)
        reinterpret_cast<double2 *>(out + (size_t)aa * ld)[rp] = acc[k][aa];
    }
  }
}

// per particle: reduce partials, SS, chol, logw, gain G, KS = G*SS, xl update
template <int D>
__global__ void __launch_bounds__(128)
k_innov(KalmanArgs a, int nsplit, const double *__restrict__ PHpart, double *__restrict__ Gg,
        double *__restrict__ KSg) {
  extern __shared__ double sm[];
  const int M = a.M, ld = a.ld;
  double *sPH = sm;  // [D][ld]
  __shared__ double s_red[4][D * D + D];
  __shared__ double s_L[D * D], s_SS[D * D], s_e[D];
  const int i = blockIdx.x;
  const double *Hi = a.H + (size_t)i * D * a.ldh;
  const double *xls = a.xl_old + (size_t)(a.ai ? a.ai[i] : i) * M;
  double part[D * D + D];
#pragma unroll
  for (int q = 0; q < D * D + D; ++q) part[q] = 0.0;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) {
    double ph[D], h[D];
#pragma unroll
    for (int aa = 0; aa < D; ++aa) {
      double s = 0.0;
      for (int sp = 0; sp < nsplit; ++sp)   // fixed order: deterministic
        s += PHpart[(((size_t)i * nsplit + sp) * D + aa) * ld + r];
      ph[aa] = (r < M) ? s : 0.0;
      sPH[aa * ld + r] = ph[aa];
      h[aa] = (r < M) ? Hi[(size_t)aa * a.ldh + r] : 0.0;
    }
    const double x = (r < M) ? xls[r] : 0.0;
#pragma unroll
    for (int aa = 0; aa < D; ++aa) {
#pragma unroll
      for (int bb = 0; bb < D; ++bb) part[aa + bb * D] = fma(h[aa], ph[bb], part[aa + bb * D]);
      part[D * D + aa] = fma(h[aa], x, part[D * D + aa]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < D * D + D; ++q) {
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
    double lw = 0.0, vv = 0.0, v[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double s = e[r];
      for (int k = 0; k < r; ++k) s -= Lc[r + k * D] * v[k];
      v[r] = s / Lc[r + r * D];
      vv += v[r] * v[r];
      lw -= log(Lc[r + r * D]);
    }
    a.logw[i] = lw - 0.5 * vv - 0.5 * D * RB_LOG2PI;
#pragma unroll
    for (int q = 0; q < D * D; ++q) s_L[q] = Lc[q];
#pragma unroll
    for (int aa = 0; aa < D; ++aa) s_e[aa] = e[aa];
  }
  __syncthreads();
  double *Gi = Gg + (size_t)i * D * ld;
  double *KSi = KSg + (size_t)i * D * ld;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) {
    double g[D];
#pragma unroll
    for (int q = 0; q < D; ++q) {   // forward: cS z = ph
      double s = sPH[q * ld + r];
#pragma unroll
      for (int k = 0; k < q; ++k) s -= s_L[q + k * D] * g[k];
      g[q] = s / s_L[q + q * D];
    }
#pragma unroll
    for (int q = D - 1; q >= 0; --q) {  // backward: cS' g = z
      double s = g[q];
#pragma unroll
      for (int k = q + 1; k < D; ++k) s -= s_L[k + q * D] * g[k];
      g[q] = s / s_L[q + q * D];
    }
    double ge = 0.0;
#pragma unroll
    for (int q = 0; q < D; ++q) {
      ge = fma(g[q], s_e[q], ge);
      double ks = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) ks = fma(g[k], s_SS[k + q * D], ks);
      Gi[(size_t)q * ld + r] = g[q];
      KSi[(size_t)q * ld + r] = ks;
    }
    if (r < M) a.xl_new[(size_t)i * M + r] = xls[r] + ge;
  }
}

// P_dst(r,c) = P_src(r,c) - sum_b KS(r,b) G(c,b) over a column chunk
template <int D, int R2>
__global__ void __launch_bounds__(256)
k_downdate(double *__restrict__ P, size_t slab, int ld, int M, const int *__restrict__ src_slot,
           const int *__restrict__ dst_slot, const int *__restrict__ list,
           const int *__restrict__ count, const double *__restrict__ Gg,
           const double *__restrict__ KSg, int cw) {
  if ((int)blockIdx.x >= *count) return;
  __shared__ double sG[D][RB_CWMAX];
  const int i = list[blockIdx.x], s = blockIdx.y;
  const int c0 = s * cw, c1 = min(M, c0 + cw);
  const double *Ps = P + (size_t)src_slot[i] * slab;
  double *Pd = P + (size_t)dst_slot[i] * slab;
  for (int idx = threadIdx.x; idx < D * cw; idx += blockDim.x) {
    const int aa = idx / cw, c = idx % cw;
    sG[aa][c] = (c0 + c < c1) ? Gg[((size_t)i * D + aa) * ld + c0 + c] : 0.0;
  }
  const int npairs = ld >> 1;
  double2 ks[R2][D];
#pragma unroll
  for (int k = 0; k < R2; ++k) {
    const int rp = threadIdx.x + k * blockDim.x;
#pragma unroll
    for (int aa = 0; aa < D; ++aa)
      ks[k][aa] = rp < npairs
                      ? reinterpret_cast<const double2 *>(KSg + ((size_t)i * D + aa) * ld)[rp]
                      : make_double2(0.0, 0.0);
  }
  __syncthreads();
  const int nc = c1 - c0;
  int c = 0;
  for (; c + 4 <= nc; c += 4) {
    double2 v[4][R2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < R2; ++k) {
        const int rp = threadIdx.x + k * blockDim.x;
        if (rp < npairs)
          v[u][k] = *(reinterpret_cast<const double2 *>(Ps + (size_t)(c0 + c + u) * ld) + rp);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < R2; ++k) {
        const int rp = threadIdx.x + k * blockDim.x;
        if (rp < npairs) {
          double2 o = v[u][k];
#pragma unroll
          for (int aa = 0; aa < D; ++aa) {
            const double g = sG[aa][c + u];
            o.x = fma(-ks[k][aa].x, g, o.x);
            o.y = fma(-ks[k][aa].y, g, o.y);
          }
          *(reinterpret_cast<double2 *>(Pd + (size_t)(c0 + c + u) * ld) + rp) = o;
        }
      }
  }
  for (; c < nc; ++c) {
#pragma unroll
    for (int k = 0; k < R2; ++k) {
      const int rp = threadIdx.x + k * blockDim.x;
      if (rp < npairs) {
        double2 o = *(reinterpret_cast<const double2 *>(Ps + (size_t)(c0 + c) * ld) + rp);
#pragma unroll
        for (int aa = 0; aa < D; ++aa) {
          const double g = sG[aa][c];
          o.x = fma(-ks[k][aa].x, g, o.x);
          o.y = fma(-ks[k][aa].y, g, o.y);
        }
        *(reinterpret_cast<double2 *>(Pd + (size_t)(c0 + c) * ld) + rp) = o;
      }
    }
  }
}

}  // namespace rb
