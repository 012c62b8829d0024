// Batched fp64 dense kernels for the smoothers' ancestor weights:
//   * k_dgemm      C_b = op(A_b) * B_b (+ kron(I, R))      (src/particleSmoother.m:191,214)
//   * k_chol_solve L_b = chol(A1_b + A2 [+ jitter I]); logdet, v = L\rhs, v'v
//                  (src/particleSmoother.m:221-229; ...InformationForm.m:225-236)
// SIMT fp64 (B200's fp64 FMA pipes); one CTA per output tile / per matrix.
#pragma once
#include "common.cuh"

namespace rb {

struct GemmArgs {
  int m, n, k;
  const double *A; int lda; size_t strideA;   // op(A) is m x k; TA: A stored k x m
  const int *slotA;                           // optional: A_b = A + slotA[b]*strideA
  const double *B; int ldb; size_t strideB;   // k x n
  double *C; int ldc; size_t strideC;         // m x n
  const double *Rblk; int d;                  // optional: C += kron(I, R) (d x d blocks)
};

template <bool TA>
__global__ void __launch_bounds__(256) k_dgemm(GemmArgs g) {
  __shared__ double As[16][64 + 4];
  __shared__ double Bs[16][64 + 4];
  const int b = blockIdx.z;
  const double *A = g.A + (size_t)(g.slotA ? g.slotA[b] : b) * g.strideA;
  const double *B = g.B + (size_t)b * g.strideB;
  double *C = g.C + (size_t)b * g.strideC;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < g.k; k0 += 16) {
    if (TA) {
      // A stored [k x m]: element (kk, i) at kk + i*lda
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kk = tid & 15, i = (tid >> 4) + 16 * q;
        const int gk = k0 + kk, gi = m0 + i;
        As[kk][i] = (gk < g.k && gi < g.m) ? A[gk + (size_t)gi * g.lda] : 0.0;
      }
    } else {
      // A stored [m x k]: element (i, kk) at i + kk*lda
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = tid & 63, kk = (tid >> 6) + 4 * q;
        const int gk = k0 + kk, gi = m0 + i;
        As[kk][i] = (gk < g.k && gi < g.m) ? A[gi + (size_t)gk * g.lda] : 0.0;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kk = tid & 15, j = (tid >> 4) + 16 * q;
      const int gk = k0 + kk, gj = n0 + j;
      Bs[kk][j] = (gk < g.k && gj < g.n) ? B[gk + (size_t)gj * g.ldb] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gj = n0 + ty * 4 + j;
    if (gj >= g.n) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = m0 + tx * 4 + i;
      if (gi >= g.m) continue;
      double v = acc[i][j];
      if (g.Rblk && (gi / g.d) == (gj / g.d)) v += g.Rblk[(gi % g.d) + (gj % g.d) * g.d];
      C[gi + (size_t)gj * g.ldc] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// batched Cholesky + forward solve, one CTA (256 threads) per matrix
// ---------------------------------------------------------------------------
struct CholArgs {
  int n;
  const double *A1; int lda1; size_t strideA1; const int *slot1;  // A1_b at A1 + slot1[b]*strideA1 (slot1 null: b)
  const double *A2; int lda2;                                     // shared addend or nullptr
  double *L; int ldl; size_t strideL;                             // workspace, factor (lower)
  const double *rhs; size_t stride_rhs;                           // [n] per matrix
  const double *rhs2;                                             // shared addend to rhs or nullptr
  double jitter;        // retry with A + jitter I on failure; < 0: no retry (failure is an error)
  double *sum_log_diag; // [batch]  sum(log(diag(L)))
  double *vtv;          // [batch]  v'v with v = L\rhs
  DevStatus *status;
  int t;
};

#define RB_CH_NB 32
__global__ void __launch_bounds__(256) k_chol_solve(CholArgs a) {
  extern __shared__ double sm[];
  double *sD = sm;                      // [32][33] diagonal block
  double *sR = sD + 32 * 33;            // [64][33] L21 rows of the row tile
  double *sC = sR + 64 * 33;            // [64][33] L21 rows of the column tile
  double *sv = sC + 64 * 33;            // [n] right-hand side / solution
  __shared__ int s_fail;
  const int b = blockIdx.x, n = a.n, tid = threadIdx.x;
  const double *A1 = a.A1 + (size_t)(a.slot1 ? a.slot1[b] : b) * a.strideA1;
  double *L = a.L + (size_t)b * a.strideL;
  const int ldl = a.ldl;
  bool ok = false;
  for (int attempt = 0; attempt < 2 && !ok; ++attempt) {
    const double jit = attempt ? a.jitter : 0.0;
    for (int c = 0; c < n; ++c)
      for (int r = c + tid; r < n; r += blockDim.x) {   // lower triangle only
        double v = A1[r + (size_t)c * a.lda1];
        if (a.A2) v += a.A2[r + (size_t)c * a.lda2];
        if (r == c) v += jit;
        L[r + (size_t)c * ldl] = v;
      }
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (int jb = 0; jb < n; jb += RB_CH_NB) {
      const int nb = min(RB_CH_NB, n - jb);
      // (1) diagonal block -> smem, unblocked factorisation by warp 0
      for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        const int r = idx % nb, c = idx / nb;
        sD[r * 33 + c] = (r >= c) ? L[(jb + r) + (size_t)(jb + c) * ldl] : 0.0;
      }
      __syncthreads();
      if (tid < 32) {
        const int r = tid;
        for (int j = 0; j < nb; ++j) {
          const double dj = sD[j * 33 + j];
          if (!(dj > 0.0)) { if (r == 0) s_fail = 1; break; }
          const double ljj = sqrt(dj);
          __syncwarp();
          if (r == j) sD[j * 33 + j] = ljj;
          if (r > j && r < nb) sD[r * 33 + j] /= ljj;
          __syncwarp();
          if (r > j && r < nb) {
            const double lrj = sD[r * 33 + j];
            for (int c = j + 1; c <= r; ++c) sD[r * 33 + c] -= lrj * sD[c * 33 + j];
          }
          __syncwarp();
        }
      }
      __syncthreads();
      if (s_fail) break;
      for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        const int r = idx % nb, c = idx / nb;
        if (r >= c) L[(jb + r) + (size_t)(jb + c) * ldl] = sD[r * 33 + c];
      }
      const int r0 = jb + nb;
      if (r0 >= n) break;
      // (2) panel solve: L21 = A21 * L11^-T, one row per thread
      for (int r = r0 + tid; r < n; r += blockDim.x) {
        double x[RB_CH_NB];
#pragma unroll
        for (int k = 0; k < RB_CH_NB; ++k) x[k] = L[r + (size_t)(jb + k) * ldl];
#pragma unroll
        for (int k = 0; k < RB_CH_NB; ++k) {
          double s = x[k];
#pragma unroll
          for (int q = 0; q < RB_CH_NB; ++q)
            if (q < k) s -= x[q] * sD[k * 33 + q];
          x[k] = s / sD[k * 33 + k];
        }
#pragma unroll
        for (int k = 0; k < RB_CH_NB; ++k) L[r + (size_t)(jb + k) * ldl] = x[k];
      }
      __syncthreads();
      // (3) trailing update A22 -= L21 L21' on lower 64x64 tiles
      const int tx = tid & 15, ty = tid >> 4;
      for (int tj = r0; tj < n; tj += 64) {
        for (int idx = tid; idx < 64 * 32; idx += blockDim.x) {
          const int i = idx & 63, k = idx >> 6;
          sC[i * 33 + k] = (tj + i < n) ? L[(tj + i) + (size_t)(jb + k) * ldl] : 0.0;
        }
        for (int ti = tj; ti < n; ti += 64) {
          __syncthreads();
          if (ti == tj) {
            for (int idx = tid; idx < 64 * 32; idx += blockDim.x) sR[(idx & 63) * 33 + (idx >> 6)] = sC[(idx & 63) * 33 + (idx >> 6)];
          } else {
            for (int idx = tid; idx < 64 * 32; idx += blockDim.x) {
              const int i = idx & 63, k = idx >> 6;
              sR[i * 33 + k] = (ti + i < n) ? L[(ti + i) + (size_t)(jb + k) * ldl] : 0.0;
            }
          }
          __syncthreads();
          double acc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 8
          for (int k = 0; k < 32; ++k) {
            double ar[4], bc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ar[i] = sR[(tx + 16 * i) * 33 + k];
#pragma unroll
            for (int j = 0; j < 4; ++j) bc[j] = sC[(ty + 16 * j) * 33 + k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fma(ar[i], bc[j], acc[i][j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int gc = tj + ty + 16 * j;
            if (gc >= n) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int gr = ti + tx + 16 * i;
              if (gr < n && gr >= gc) L[gr + (size_t)gc * ldl] -= acc[i][j];
            }
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    ok = !s_fail;
    if (!ok && attempt == 0) {
      if (a.jitter < 0.0) break;
      if (tid == 0) atomicAdd(&a.status->used_jitter, 1);
    }
    __syncthreads();
  }
  if (!ok) {
    if (tid == 0 && atomicCAS(&a.status->not_pd, 0, 1) == 0) {
      a.status->not_pd_step = a.t;
      a.status->not_pd_particle = b;
    }
    if (tid == 0) { a.sum_log_diag[b] = nan(""); a.vtv[b] = nan(""); }
    return;
  }
  // ---- sum log diag, forward solve v = L \ rhs, v'v ---------------------------
  for (int r = tid; r < n; r += blockDim.x)
    sv[r] = a.rhs[(size_t)b * a.stride_rhs + r] + (a.rhs2 ? a.rhs2[r] : 0.0);
  __syncthreads();
  for (int jb = 0; jb < n; jb += RB_CH_NB) {
    const int nb = min(RB_CH_NB, n - jb);
    if (tid < 32) {   // diagonal block: serial over rows, lanes share the dot product
      for (int j = 0; j < nb; ++j) {
        double part = (tid < j) ? L[(jb + j) + (size_t)(jb + tid) * ldl] * sv[jb + tid] : 0.0;
        part = warp_sum(part);
        if (tid == 0) sv[jb + j] = (sv[jb + j] - part) / L[(jb + j) + (size_t)(jb + j) * ldl];
        __syncwarp();
      }
    }
    __syncthreads();
    for (int r = jb + nb + tid; r < n; r += blockDim.x) {
      double s = 0.0;
      for (int k = 0; k < nb; ++k) s = fma(L[r + (size_t)(jb + k) * ldl], sv[jb + k], s);
      sv[r] -= s;
    }
    __syncthreads();
  }
  double ld = 0.0, vv = 0.0;
  for (int r = tid; r < n; r += blockDim.x) {
    ld += log(L[r + (size_t)r * ldl]);
    vv = fma(sv[r], sv[r], vv);
  }
  __shared__ double s_red[2][8];
  ld = warp_sum(ld); vv = warp_sum(vv);
  if ((tid & 31) == 0) { s_red[0][tid >> 5] = ld; s_red[1][tid >> 5] = vv; }
  __syncthreads();
  if (tid == 0) {
    double l2 = 0.0, v2 = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { l2 += s_red[0][q]; v2 += s_red[1][q]; }
    a.sum_log_diag[b] = l2;
    a.vtv[b] = v2;
  }
}

static inline size_t chol_solve_smem(int n) {
  return sizeof(double) * (32 * 33 + 2 * 64 * 33 + (size_t)n);
}

}  // namespace rb
