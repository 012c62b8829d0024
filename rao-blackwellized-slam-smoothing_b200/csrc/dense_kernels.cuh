// Batched fp64 dense kernels for the smoothers' ancestor weights:
//   * k_dgemm      C_b = op(A_b) * B_b (+ kron(I, R))      (src/particleSmoother.m:191,214)
//   * k_chol_solve L_b = chol(A1_b + A2 [+ jitter I]); logdet, v = L\rhs, v'v
//                  (src/particleSmoother.m:221-229; ...InformationForm.m:225-236)
// SIMT fp64 (B200's fp64 FMA pipes); one CTA per output tile / per matrix.
#pragma once
#include "common.cuh"
#include "kalman_stream.cuh"   // mbarrier + cp.async.bulk helpers

namespace rb {

// ---- fp64 tensor-core tile: 128x64 outputs, K=32, 8 warps (4 x 2), warp tile 32 x 32 -------
// mma.sync.aligned.m8n8k4.f64 (DMMA; SASS DMMA).  Measured on B200 (profiles/fp64_peaks_r1.json):
// DMMA 37.1 TFLOP/s = DFMA 36.6 TFLOP/s, but a DMMA needs 2 operand doubles per 256 MACs
// where a 4x4 SIMT tile needs 8 per 16, so shared-memory bandwidth stops being the limiter.
// Operand tiles in shared memory, row index contiguous: As[32][RB_LDA] (A[i][k] at k*RB_LDA+i),
// Bs[32][RB_LDB] (C = A * B' with B given by rows: B[j][k] at k*RB_LDB+j).  The leading
// dimensions are = 8 (mod 16) doubles so that a fragment read (8 rows x 4 k) costs the
// minimal two shared-memory wavefronts.
#define RB_LDA 136
#define RB_LDB 72
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// acc[mi][nj][e]: C(row = wr + 8 mi + g, col = wc + 8 nj + 2 tg + e), g = lane>>2, tg = lane&3
template <int LDA_ = RB_LDA>
__device__ __forceinline__ void mma_tile_k32(const double *__restrict__ As, const double *__restrict__ Bs,
                                             int wr, int wc, int lane, double (&acc)[4][4][2]) {
  const int g = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int kk = 0; kk < 32; kk += 4) {
    double a[4], b[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a[mi] = As[(kk + tg) * LDA_ + wr + 8 * mi + g];
#pragma unroll
    for (int nj = 0; nj < 4; ++nj) b[nj] = Bs[(kk + tg) * RB_LDB + wc + 8 * nj + g];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], a[mi], b[nj]);
  }
}
// asynchronous 16-byte global->shared copy; bytes beyond src_bytes are zero-filled
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// stage rows [row0, row0+ROWS) x 32 columns of a column-major matrix (16-byte aligned columns,
// row0 even) into a [32][LD] tile; rows >= nrows are zero
template <int ROWS, int LD, int NT = 256>
__device__ __forceinline__ void stage_rows_k32(double *tile, const double *__restrict__ src, int ld_src,
                                               int row0, int nrows, int tid) {
#pragma unroll
  for (int q = 0; q < (ROWS / 2) * 32 / NT; ++q) {
    const int idx = tid + q * NT;
    const int i2 = idx % (ROWS / 2), k = idx / (ROWS / 2);
    const int r = row0 + 2 * i2;
    const int valid = max(0, min(2, nrows - r));
    const double *gp = src + (size_t)k * ld_src + (valid ? r : 0);
    cp_async16(tile + k * LD + 2 * i2, gp, 8 * valid);
  }
}

struct GemmArgs {
  int m, n, k;
  const double *A; int lda; size_t strideA;   // op(A) is m x k; TA: A stored k x m
  const int *slotA;                           // optional: A_b = A + slotA[b]*strideA
  const double *B; int ldb; size_t strideB;   // k x n
  double *C; int ldc; size_t strideC;         // m x n
  const double *Rblk; int d;                  // optional: C += kron(I, R) (d x d blocks)
  int lower = 0;                              // C is symmetric and only read below the diagonal: tiles above it are skipped
};

template <bool TA>
__global__ void __launch_bounds__(256, 2) k_dgemm(GemmArgs g) {
  extern __shared__ double sm_gemm[];
  double *As = sm_gemm;                  // [32][RB_LDA]
  double *Bs = sm_gemm + 32 * RB_LDA;    // [32][RB_LDB]
  const int b = blockIdx.z;
  const double *A = g.A + (size_t)(g.slotA ? g.slotA[b] : b) * g.strideA;
  const double *B = g.B + (size_t)b * g.strideB;
  double *C = g.C + (size_t)b * g.strideC;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 64;
  if (g.lower && n0 >= m0 + 128) return;   // tile entirely above the diagonal of a symmetric result
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int k0 = 0; k0 < g.k; k0 += 32) {
    if (TA) {   // A stored [k x m]: (kk, i) at kk + i*lda, contiguous along kk
      for (int idx = tid; idx < 128 * 32; idx += 256) {
        const int kk = idx & 31, i = idx >> 5;
        const int gk = k0 + kk, gi = m0 + i;
        As[kk * RB_LDA + i] = (gk < g.k && gi < g.m) ? A[gk + (size_t)gi * g.lda] : 0.0;
      }
    } else {    // A stored [m x k]: (i, kk) at i + kk*lda, contiguous along i
      for (int idx = tid; idx < 128 * 32; idx += 256) {
        const int i = idx & 127, kk = idx >> 7;
        const int gk = k0 + kk, gi = m0 + i;
        As[kk * RB_LDA + i] = (gk < g.k && gi < g.m) ? A[gi + (size_t)gk * g.lda] : 0.0;
      }
    }
    for (int idx = tid; idx < 64 * 32; idx += 256) {   // B [k x n]: contiguous along kk
      const int kk = idx & 31, j = idx >> 5;
      const int gk = k0 + kk, gj = n0 + j;
      Bs[kk * RB_LDB + j] = (gk < g.k && gj < g.n) ? B[gk + (size_t)gj * g.ldb] : 0.0;
    }
    __syncthreads();
    mma_tile_k32(As, Bs, wr, wc, lane, acc);
    __syncthreads();
  }
  const int gq = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int nj = 0; nj < 4; ++nj)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gi = m0 + wr + 8 * mi + gq, gj = n0 + wc + 8 * nj + 2 * tg + e;
        if (gi < g.m && gj < g.n) {
          double v = acc[mi][nj][e];
          if (g.Rblk && (gi / g.d) == (gj / g.d)) v += g.Rblk[(gi % g.d) + (gj % g.d) * g.d];
          C[gi + (size_t)gj * g.ldc] = v;
        }
      }
}

// ---------------------------------------------------------------------------
// k_dgemm_pipe<AK>: the same product with the operands staged by cp.async (16-byte copies) into a
// double-buffered ring, so the loads of K-step i+1 fly while step i is multiplied.  k_dgemm above stages
// with synchronous scalar loads and is kept for operands whose leading dimensions / addresses are not
// 16-byte friendly (launch_gemm checks).
//   A: AK = false  stored [m x k] (row index contiguous)  -> tile [32 k][RB_LDA], as k_dgemm
//      AK = true   stored [k x m] (k contiguous)          -> tile [128 m][RB_LDK]
//   B: stored [k x n] (k contiguous)                      -> tile [64 n][RB_LDK]
// RB_LDK = 36 = 4 (mod 16) doubles: a fragment read (8 rows x 4 k) touches every 8-byte bank pair exactly
// twice, the minimum for 256 bytes.
// ---------------------------------------------------------------------------
#define RB_LDK 36
static inline size_t dgemm_pipe_smem(bool ak) {
  return sizeof(double) * 2 * ((ak ? 128 * RB_LDK : 32 * RB_LDA) + 64 * RB_LDK);
}
// rows [r0, r0+ROWS) x k [k0, k0+32) of a k-contiguous operand X(k, r) = X[k + r*ld] into tile[r][RB_LDK]
template <int ROWS>
__device__ __forceinline__ void stage_kmajor(double *tile, const double *__restrict__ X, int ld, int r0, int nrows,
                                             int k0, int K, int tid) {
#pragma unroll
  for (int q = 0; q < ROWS * 16 / 256; ++q) {
    const int idx = tid + q * 256;
    const int k2 = idx & 15, r = idx >> 4;
    const int gk = k0 + 2 * k2, gr = r0 + r;
    const int valid = gr < nrows ? max(0, min(2, K - gk)) : 0;
    const double *gp = X + (valid ? gk + (size_t)gr * ld : 0);
    cp_async16(tile + r * RB_LDK + 2 * k2, gp, 8 * valid);
  }
}
// rows [r0, r0+128) x k [k0, k0+32) of a row-contiguous operand X(r, k) = X[r + k*ld] into tile[k][RB_LDA]
__device__ __forceinline__ void stage_mmajor(double *tile, const double *__restrict__ X, int ld, int r0, int nrows,
                                             int k0, int K, int tid) {
#pragma unroll
  for (int q = 0; q < 64 * 32 / 256; ++q) {
    const int idx = tid + q * 256;
    const int i2 = idx & 63, k = idx >> 6;
    const int gr = r0 + 2 * i2, gk = k0 + k;
    const int valid = gk < K ? max(0, min(2, nrows - gr)) : 0;
    const double *gp = X + (valid ? gr + (size_t)gk * ld : 0);
    cp_async16(tile + k * RB_LDA + 2 * i2, gp, 8 * valid);
  }
}

template <bool AK>
__global__ void __launch_bounds__(256, 2) k_dgemm_pipe(GemmArgs g) {
  extern __shared__ __align__(16) double sm_gemm[];
  constexpr int ASZ = AK ? 128 * RB_LDK : 32 * RB_LDA, BSZ = 64 * RB_LDK;
  double *As = sm_gemm;                  // [2][ASZ]
  double *Bs = sm_gemm + 2 * ASZ;        // [2][BSZ]
  const int b = blockIdx.z;
  const double *A = g.A + (size_t)(g.slotA ? g.slotA[b] : b) * g.strideA;
  const double *B = g.B + (size_t)b * g.strideB;
  double *C = g.C + (size_t)b * g.strideC;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 64;
  if (g.lower && n0 >= m0 + 128) return;   // tile entirely above the diagonal of a symmetric result
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
  const int gq = lane >> 2, tg = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto stage = [&](int buf, int k0) {
    if (AK) stage_kmajor<128>(As + buf * ASZ, A, g.lda, m0, g.m, k0, g.k, tid);
    else stage_mmajor(As + buf * ASZ, A, g.lda, m0, g.m, k0, g.k, tid);
    stage_kmajor<64>(Bs + buf * BSZ, B, g.ldb, n0, g.n, k0, g.k, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int nk = (g.k + 31) / 32;
  stage(0, 0);
  for (int i = 0; i < nk; ++i) {
    const int buf = i & 1;
    if (i + 1 < nk) {
      stage(buf ^ 1, (i + 1) * 32);      // its buffer was released by the barrier that closed step i-1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double *Ab = As + buf * ASZ, *Bb = Bs + buf * BSZ;
#pragma unroll
    for (int kk = 0; kk < 32; kk += 4) {
      double a[4], bv[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
        a[mi] = AK ? Ab[(wr + 8 * mi + gq) * RB_LDK + kk + tg] : Ab[(kk + tg) * RB_LDA + wr + 8 * mi + gq];
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) bv[nj] = Bb[(wc + 8 * nj + gq) * RB_LDK + kk + tg];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], a[mi], bv[nj]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int nj = 0; nj < 4; ++nj)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gi = m0 + wr + 8 * mi + gq, gj = n0 + wc + 8 * nj + 2 * tg + e;
        if (gi < g.m && gj < g.n) {
          double v = acc[mi][nj][e];
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
  const int *only_failed = nullptr;   // k_chol_solve: process only matrices with only_failed[b] != 0
};

#define RB_CH_NB 32
#define RB_CH_LDB 40    // leading dimension of the panel-row operand tile ( = 8 mod 16: conflict-free fragment reads)
// Left-looking blocked Cholesky, NB = 32.  The right-hand side rides along as row n of the
// workspace (ldl >= n+1): the panel update and the panel solve applied to that row ARE the
// forward substitution, so after the last panel L(n, 0:n-1) = (L \ rhs)'.
//   per panel jb (columns jb .. jb+31, rows jb .. n):
//     (0) C = (A1 + A2 [+ jitter I])(rows, panel) - L(rows, 0:jb) L(panel, 0:jb)'   on the fp64 tensor
//         cores: row tiles of TR rows, the tile of C preloaded into the accumulators (negated), the two
//         operands streamed through a double-buffered cp.async stage ring in 32-column steps;
//     (1) 32x32 diagonal block in shared memory, left-looking by one warp;
//     (2) L21 = C21 L11^-T, one row per thread, registers.
// Every element of the factor is written once and the finished panels are only ever READ again
// (K = jb columns per panel): n^3 / (6 NB) * 8 B = 5.7 MB of operand reads per 515 x 515 matrix, mostly
// L2 hits, against 22.6 MB of read-modify-write traffic for the right-looking form this replaces
// (which re-streamed the whole trailing matrix after every panel and was HBM-bound at 21 % of the fp64
// peak, profiles/tuning_r1.md section 4).
// NT threads per matrix: 256 (row tiles of 128, 2 CTAs/SM) or 128 (row tiles of 64, 4 CTAs/SM --
// more CTAs per SM overlap one matrix's serial panel phases with another's tensor-core phase)
template <int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_chol_solve(CholArgs a) {
  constexpr int TR = NT / 2, LDA_T = TR + 8;
  extern __shared__ double sm[];
  double *sD = sm;                           // [32][33] (+ pad to a 16-byte boundary)
  double *As = sD + 32 * 34;                 // [2][32][LDA_T] rows of the row tile, columns k0 .. k0+31
  double *Bs = As + 2 * 32 * LDA_T;          // [2][32][RB_CH_LDB] rows of the panel
  __shared__ int s_fail;
  __shared__ double s_red[2][8];
  const int b = blockIdx.x, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.only_failed && !a.only_failed[b]) return;   // retry pass after the batched panel path
  const double *A1 = a.A1 + (size_t)(a.slot1 ? a.slot1[b] : b) * a.strideA1;
  double *L = a.L + (size_t)b * a.strideL;
  const int ldl = a.ldl;
  const int nr = n + 1;                      // rows incl. the right-hand-side row
  const int gq = lane >> 2, tg = lane & 3;
  const int wr = warp * 16;                  // warp tile: 16 rows x 32 columns = 2 x 4 DMMA tiles
  bool ok = false;
  for (int attempt = 0; attempt < 2 && !ok; ++attempt) {
    const double jit = attempt ? a.jitter : 0.0;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (int jb = 0; jb < n; jb += RB_CH_NB) {
      const int nb = min(RB_CH_NB, n - jb);
      // (0) panel update: rows jb .. n (row n = right-hand side), columns jb .. jb+nb-1, K = jb
      for (int ti = jb; ti < nr; ti += TR) {
        // C tile: all 32 loads of a thread are issued back to back (clamped addresses, no branches:
        // the loads are latency-bound, one dependent load at a time costs a third of the kernel)
        double acc[2][4][2];
        {
          double c1[2][4][2], c2[2][4][2];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int gr = min(ti + wr + 8 * mi + gq, n - 1), gc = min(jb + 8 * nj + 2 * tg + e, n - 1);
                c1[mi][nj][e] = A1[gr + (size_t)gc * a.lda1];
                c2[mi][nj][e] = a.A2 ? a.A2[gr + (size_t)gc * a.lda2] : 0.0;
              }
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int gr = ti + wr + 8 * mi + gq, gc = jb + 8 * nj + 2 * tg + e;
                double v = c1[mi][nj][e] + c2[mi][nj][e] + (gr == gc ? jit : 0.0);
                if (gr == n && gc < n) v = a.rhs[(size_t)b * a.stride_rhs + gc] + (a.rhs2 ? a.rhs2[gc] : 0.0);
                acc[mi][nj][e] = (gc < n && gr >= gc && gr <= n) ? -v : 0.0;   // acc = -C, acc += A B', C_new = -acc
              }
        }
        if (jb > 0) {
          __syncthreads();                       // the stage ring is free (previous tile / diagonal block done)
          stage_rows_k32<TR, LDA_T, NT>(As, L, ldl, ti, nr, tid);
          stage_rows_k32<32, RB_CH_LDB, NT>(Bs, L, ldl, jb, n, tid);
          asm volatile("cp.async.commit_group;" ::: "memory");
          int buf = 0;
          for (int k0 = 0; k0 < jb; k0 += 32, buf ^= 1) {
            if (k0 + 32 < jb) {                  // next step's operands fly while this one is multiplied
              stage_rows_k32<TR, LDA_T, NT>(As + (buf ^ 1) * 32 * LDA_T, L + (size_t)(k0 + 32) * ldl, ldl, ti, nr, tid);
              stage_rows_k32<32, RB_CH_LDB, NT>(Bs + (buf ^ 1) * 32 * RB_CH_LDB, L + (size_t)(k0 + 32) * ldl, ldl, jb, n, tid);
              asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");
            } else {
              asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
            const double *Ab = As + buf * 32 * LDA_T, *Bb = Bs + buf * 32 * RB_CH_LDB;
#pragma unroll
            for (int kk = 0; kk < 32; kk += 4) {
              double av[2], bv[4];
#pragma unroll
              for (int mi = 0; mi < 2; ++mi) av[mi] = Ab[(kk + tg) * LDA_T + wr + 8 * mi + gq];
#pragma unroll
              for (int nj = 0; nj < 4; ++nj) bv[nj] = Bb[(kk + tg) * RB_CH_LDB + 8 * nj + gq];
#pragma unroll
              for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], av[mi], bv[nj]);
            }
            __syncthreads();                     // buffer `buf` may be refilled two steps from now
          }
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nj = 0; nj < 4; ++nj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int gr = ti + wr + 8 * mi + gq, gc = jb + 8 * nj + 2 * tg + e;
              if (gr < nr && gc < n && gr >= gc) L[gr + (size_t)gc * ldl] = -acc[mi][nj][e];
            }
      }
      __syncthreads();
      // (1) diagonal block
      for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        const int r = idx % nb, c = idx / nb;
        sD[r * 33 + c] = (r >= c) ? L[(jb + r) + (size_t)(jb + c) * ldl] : 0.0;
      }
      __syncthreads();
      if (warp == 0) {
        const int r = lane;
        for (int j = 0; j < nb; ++j) {        // left-looking: column j from columns 0..j-1
          double s = 0.0;
          if (r >= j && r < nb) {
            s = sD[r * 33 + j];
            for (int k = 0; k < j; ++k) s = fma(-sD[r * 33 + k], sD[j * 33 + k], s);
          }
          const double dj = __shfl_sync(0xffffffffu, s, j);
          if (!(dj > 0.0)) { if (r == 0) s_fail = 1; break; }
          const double ljj = sqrt(dj);
          if (r == j) sD[j * 33 + j] = ljj;
          else if (r > j && r < nb) sD[r * 33 + j] = s / ljj;
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
      // (2) panel solve for rows r0..n (row n = right-hand side)
      if (nb == RB_CH_NB) {
#pragma unroll 1
        for (int r = r0 + tid; r < nr; r += blockDim.x) {
          asm volatile("" ::: "memory");   // keep the 496 L11 loads inside the iteration (no LICM spills)
          double x[RB_CH_NB];
#pragma unroll
          for (int k = 0; k < RB_CH_NB; ++k) x[k] = L[r + (size_t)(jb + k) * ldl];
#pragma unroll
          for (int k = 0; k < RB_CH_NB; ++k) {
            double sx = x[k];
#pragma unroll
            for (int q = 0; q < RB_CH_NB; ++q)
              if (q < k) sx = fma(-x[q], sD[k * 33 + q], sx);
            x[k] = sx / sD[k * 33 + k];
          }
#pragma unroll
          for (int k = 0; k < RB_CH_NB; ++k) L[r + (size_t)(jb + k) * ldl] = x[k];
        }
      } else {   // last, narrower panel: only the right-hand-side row is left below it
        if (tid == 0) {
          for (int k = 0; k < nb; ++k) {
            double sx = L[n + (size_t)(jb + k) * ldl];
            for (int q = 0; q < k; ++q) sx = fma(-L[n + (size_t)(jb + q) * ldl], sD[k * 33 + q], sx);
            L[n + (size_t)(jb + k) * ldl] = sx / sD[k * 33 + k];
          }
        }
      }
      __syncthreads();
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
  double ld = 0.0, vv = 0.0;
  for (int r = tid; r < n; r += blockDim.x) {
    ld += log(L[r + (size_t)r * ldl]);
    const double v = L[n + (size_t)r * ldl];    // v = L \ rhs
    vv = fma(v, v, vv);
  }
  ld = warp_sum(ld); vv = warp_sum(vv);
  if (lane == 0) { s_red[0][warp] = ld; s_red[1][warp] = vv; }
  __syncthreads();
  if (tid == 0) {
    double l2 = 0.0, v2 = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { l2 += s_red[0][q]; v2 += s_red[1][q]; }
    a.sum_log_diag[b] = l2;
    a.vtv[b] = v2;
  }
}

// ---------------------------------------------------------------------------
// k_chol_inv: the same left-looking factorisation with the two serial phases of k_chol_solve taken
// off the per-matrix critical path (ncu source page of k_chol_solve, profiles/tuning_r2.md section 3:
// 31 % of a CTA's time in the tensor-core loop, 35 % in the per-row forward substitution of the
// panel, 20 % with three warps waiting for the one that factors the diagonal block in shared memory):
//   * the C tile of a row tile never goes to global memory: accumulators -> shared memory in the
//     A-operand layout of the next product;
//   * the 32 x 32 diagonal block is factored AND inverted by all warps together, right-looking, two
//     columns per block barrier (a single warp in registers took 53 k cycles per block: 28 % of the
//     kernel with three warps waiting, ncu source page of the first version);
//   * the panel solve L21 = C21 L11^-T is one more tensor-core product C21 * (L11^-1)' per row
//     tile (K <= 32: the zero half of the triangular inverse is skipped), instead of 496 dependent
//     FMAs per row;
//   * the addend tile of A1 is prefetched towards L2 before the operand loop and read after it;
//   * operand ring of NS stages of KC columns, ONE block barrier per stage, stage 0 of the next tile
//     issued before the epilogue of the current one;
//   * products with swapped operands (a thread owns two consecutive rows of a column: 16-byte
//     accesses), ragged tiles skip padded warps and the tiles above the diagonal.
// A narrower last panel is padded with an identity block.  Same arguments and results as
// k_chol_solve (which stays as the retry kernel of the across-the-batch path and as
// RBSLAM_CHOL_KERNEL=solve).
// ---------------------------------------------------------------------------
// KC = columns of L per operand stage, NS = stages in the cp.async ring
static inline size_t chol_inv_smem(int kc, int ns, int nt = 128) {
  return sizeof(double) * ((size_t)ns * kc * (nt / 2 + 8 + RB_CH_LDB) + 32 * RB_CH_LDB);
}
// rows [row0, row0+ROWS) x KC columns of a column-major matrix into a [KC][LD] tile; rows >= nrows are zero
template <int ROWS, int KC, int LD, int NT>
__device__ __forceinline__ void stage_rows(double *tile, const double *__restrict__ src, int ld_src, int row0,
                                           int nrows, int tid) {
  constexpr int TOTAL = (ROWS / 2) * KC;     // 16-byte copies
#pragma unroll
  for (int q = 0; q < (TOTAL + NT - 1) / NT; ++q) {
    const int idx = tid + q * NT;
    if (TOTAL % NT != 0 && idx >= TOTAL) break;
    const int i2 = idx % (ROWS / 2), k = idx / (ROWS / 2);
    const int r = row0 + 2 * i2;
    const int valid = max(0, min(2, nrows - r));
    const double *gp = src + (size_t)k * ld_src + (valid ? r : 0);
    cp_async16(tile + k * LD + 2 * i2, gp, 8 * valid);
  }
}

// (A cp.async.bulk form of the ring -- one bulk copy per 512-byte operand column, mbarrier per stage -- was
// measured at 18.2 ms against 12.5 ms for the cp.async form at the C5 shape: copies this small are bound by
// the per-copy cost of the TMA unit.  profiles/tuning_r2.md section 3.)
// one operand stage of the panel update for a 16 x 32 warp tile; MASK selects the 8 x 8 tiles (bit 4 mi + nj).
// Operands swapped: the thread ends up with C(rows 8 mi + 2 tg + {0, 1}, column 8 nj + gq).
template <int KC, int MASK, int LDA_T>
__device__ __forceinline__ void chol_kstep(const double *__restrict__ Ab, const double *__restrict__ Bb,
                                           double (&acc)[2][4][2]) {
#pragma unroll
  for (int kk = 0; kk < KC; kk += 4) {
    double av[2], bv[4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
      if ((MASK >> (4 * mi)) & 0xf) av[mi] = Ab[kk * LDA_T + 8 * mi];
#pragma unroll
    for (int nj = 0; nj < 4; ++nj)
      if ((MASK >> nj) & 0x11) bv[nj] = Bb[kk * RB_CH_LDB + 8 * nj];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nj = 0; nj < 4; ++nj)
        if ((MASK >> (4 * mi + nj)) & 1) dmma884(acc[mi][nj][0], acc[mi][nj][1], bv[nj], av[mi]);
  }
}

// NT threads per matrix: 128 (row tiles of 64, 4 CTAs/SM: large batches) or 512 (row tiles of 256, one
// CTA per SM: batches that leave SMs idle anyway -- the C1 example has 100 particles -- get a whole SM's
// warps per matrix, which shortens the serial chain of row tiles per panel 4x)
// warp 0 of a panel's first tile when the panel's few last rows ("tail": at most 8 rows, staged into rows
// TR .. TR+7 of the A stages) ride along: its own three tiles (mask 0x31) plus the four tiles of the tail rows,
// kept in the accumulators of the tiles it skips: tail column tile nj -> acc[0][1], acc[0][2], acc[0][3], acc[1][2].
template <int KC, int LDA_T, int TR>
__device__ __forceinline__ void chol_kstep_tail(const double *__restrict__ Ab, const double *__restrict__ Bb,
                                                double (&acc)[2][4][2]) {
#pragma unroll
  for (int kk = 0; kk < KC; kk += 4) {
    const double a0 = Ab[kk * LDA_T], a1 = Ab[kk * LDA_T + 8], at = Ab[kk * LDA_T + TR];
    double bv[4];
#pragma unroll
    for (int nj = 0; nj < 4; ++nj) bv[nj] = Bb[kk * RB_CH_LDB + 8 * nj];
    dmma884(acc[0][0][0], acc[0][0][1], bv[0], a0);
    dmma884(acc[1][0][0], acc[1][0][1], bv[0], a1);
    dmma884(acc[1][1][0], acc[1][1][1], bv[1], a1);
    dmma884(acc[0][1][0], acc[0][1][1], bv[0], at);
    dmma884(acc[0][2][0], acc[0][2][1], bv[1], at);
    dmma884(acc[0][3][0], acc[0][3][1], bv[2], at);
    dmma884(acc[1][2][0], acc[1][2][1], bv[3], at);
  }
}

template <int KC, int NS, int MINB, int NT>
__global__ void __launch_bounds__(NT, MINB) k_chol_inv(CholArgs a) {
  constexpr int NW = NT / 32, TR = 16 * NW, LDA_T = TR + 8, LDB = RB_CH_LDB;
  constexpr int NQ = 32 / NW;                // columns of the diagonal block per thread (pairs: NQ even)
  static_assert(NW >= 4 && NW <= 16 && (NQ % 2) == 0, "4, 8 or 16 warps");
  static_assert(NS >= 2 && NS <= 5, "wait_group cases below");
  static_assert(NS * KC >= 32, "the C tile lives in the first 32 columns of the A ring");
  extern __shared__ __align__(128) double sm[];
  double *As = sm;                           // [NS][KC][LDA_T] operand ring; its first 32 columns double as the C tile
  double *Bs = As + NS * KC * LDA_T;         // [NS][KC][LDB]; doubles as the exchange line of the diagonal block
  double *sX = Bs + NS * KC * LDB;           // [32][LDB]: L11^-1 as a B operand, X(j, c) at c*LDB + j
  double *colb = Bs;
  __shared__ int s_fail;
  __shared__ double s_red[2][NW];
  const int b = blockIdx.x, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double *A1 = a.A1 + (size_t)(a.slot1 ? a.slot1[b] : b) * a.strideA1;
  double *L = a.L + (size_t)b * a.strideL;
  const int ldl = a.ldl;
  const int nr = n + 1;                      // rows incl. the right-hand-side row
  const int gq = lane >> 2, tg = lane & 3;
  const int wr = warp * 16;                  // warp tile: 16 rows x 32 columns = 2 x 4 DMMA tiles
  const bool v1ok = ((a.lda1 & 1) == 0) && ((a.strideA1 & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.A1) & 15) == 0);
  const bool v2ok = ((a.lda2 & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.A2) & 15) == 0);
  bool ok = false;
  for (int attempt = 0; attempt < 2 && !ok; ++attempt) {
    const double jit = attempt ? a.jitter : 0.0;
    if (tid == 0) s_fail = 0;
    bool pre = false;
    __syncthreads();
    for (int jb = 0; jb < n; jb += RB_CH_NB) {
      const int nb = min(RB_CH_NB, n - jb);
      // A panel whose row count leaves a remainder of at most 8 rows (n = 515: the last three rows and the
      // right-hand side, on every second panel) would spend a whole operand loop on a 64-row tile with one
      // useful 8-row slice.  Those "tail" rows ride with the panel's FIRST tile instead: they are staged into
      // the 8 padding rows of the A stages, warp 0 -- which skips five of its eight tiles there -- multiplies
      // them, and their C values and panel product use the padding rows of the C tile.
      const int trem = (nr - jb) % TR;
      const bool tail = NT == 128 && nr - jb > TR && trem > 0 && trem <= 8;
      const int trow0 = nr - trem;             // first tail row (even: jb + a multiple of TR)
      const int nr_t = tail ? trow0 : nr;      // rows covered by regular tiles
      for (int ti = jb; ti < nr_t; ti += TR) {
        // the addend tile of A1 towards L2 (one 64-byte piece per thread and half tile): it is read after
        // the operand loop, registers stay free for a fourth CTA per SM
        {
#pragma unroll
          for (int h = 0; h < 2; ++h) {        // 32 columns x TR / 8 pieces of 64 bytes = 2 per thread
            const int pc = (tid + h * NT) & 31, pr = ((tid + h * NT) >> 5) * 8;
            const int gc = min(jb + pc, n - 1), gr = min(ti + pr, n - 1);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A1 + gr + (size_t)gc * a.lda1));
          }
        }
        double acc[2][4][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nj = 0; nj < 4; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
        // the last tile of a panel is ragged: a warp whose 16 rows lie past the right-hand-side row skips the
        // products (with 64-row granularity 27 % of all tensor-core work of a 515 x 515 matrix would be padding,
        // most of it in the late panels where K is largest)
        const bool wact = ti + wr < nr;
        // tiles (mi, nj) of this warp that are multiplied: bit 4 mi + nj
        const bool tile0 = ti == jb;
        const bool tailw = tail && tile0 && warp == 0;
        const int tmask = !wact ? 0 : (!tile0 || warp > 1) ? 0xff : (warp == 0 ? (tail ? 0x131 : 0x31) : 0xf7);
        // tail rows of this panel into rows TR .. TR+7 of an A stage (4 row pairs x KC columns)
        auto stage_tail = [&](double *Astage, int k0, int row0) {
          if (tid < 4 * KC) {
            const int i2 = tid & 3, kc = tid >> 2;
            const int r = row0 + 2 * i2;
            const int valid = max(0, min(2, nr - r));
            cp_async16(Astage + kc * LDA_T + TR + 2 * i2, L + (size_t)(k0 + kc) * ldl + (valid ? r : 0), 8 * valid);
          }
        };
        __syncthreads();                         // ring + C tile + sX readers of the previous tile are done
        if (jb > 0) {
          const int nk = jb / KC;
          // stage i of this tile lives in buffer (sb + i) % NS.  pre: stage 0 was issued into the last buffer
          // while the previous tile finished (its C tile occupies the first buffers of the ring)
          const int sb = pre ? NS - 1 : 0;
#pragma unroll
          for (int s0 = 0; s0 < NS - 1; ++s0)
            if (s0 < nk && !(pre && s0 == 0)) {
              const int bq = (sb + s0) % NS;
              stage_rows<TR, KC, LDA_T, NT>(As + bq * KC * LDA_T, L + (size_t)(s0 * KC) * ldl, ldl, ti, nr, tid);
              stage_rows<32, KC, LDB, NT>(Bs + bq * KC * LDB, L + (size_t)(s0 * KC) * ldl, ldl, jb, n, tid);
              if (tail && tile0) stage_tail(As + bq * KC * LDA_T, s0 * KC, trow0);
              asm volatile("cp.async.commit_group;" ::: "memory");
            }
          int buf = sb, nbuf = (sb + NS - 1) % NS;   // buffer of stage i, buffer stage i + NS - 1 goes to
          for (int i = 0; i < nk; ++i) {
            // stage i has landed when at most min(NS - 2, nk - 1 - i) younger groups are outstanding
            {
              const int rem = nk - 1 - i;
              if (NS >= 3 && rem >= NS - 2) asm volatile("cp.async.wait_group %0;" ::"n"(NS >= 3 ? NS - 2 : 0) : "memory");
              else if (NS >= 5 && rem == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
              else if (NS >= 4 && rem == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
              else asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();                     // everyone's copies of stage i landed; stage i-1's buffer is free
            if (i + NS - 1 < nk) {
              const int k0 = (i + NS - 1) * KC;
              stage_rows<TR, KC, LDA_T, NT>(As + nbuf * KC * LDA_T, L + (size_t)k0 * ldl, ldl, ti, nr, tid);
              stage_rows<32, KC, LDB, NT>(Bs + nbuf * KC * LDB, L + (size_t)k0 * ldl, ldl, jb, n, tid);
              if (tail && tile0) stage_tail(As + nbuf * KC * LDA_T, k0, trow0);
              asm volatile("cp.async.commit_group;" ::: "memory");
            }
            const double *Ab = As + buf * KC * LDA_T + wr + gq + tg * LDA_T, *Bb = Bs + buf * KC * LDB + gq + tg * LDB;
            // the 8 x 8 tiles strictly above the diagonal of the diagonal block are never read: warps 0 and 1
            // of a panel's first tile skip them (6 of their 16 tiles)
            if (tmask == 0xff) chol_kstep<KC, 0xff, LDA_T>(Ab, Bb, acc);
            else if (tmask == 0x31) chol_kstep<KC, 0x31, LDA_T>(Ab, Bb, acc);
            else if (tmask == 0xf7) chol_kstep<KC, 0xf7, LDA_T>(Ab, Bb, acc);
            else if (tmask == 0x131) chol_kstep_tail<KC, LDA_T, TR>(Ab, Bb, acc);
            buf = (buf + 1 == NS) ? 0 : buf + 1;
            nbuf = (nbuf + 1 == NS) ? 0 : nbuf + 1;
          }
        }
        // the addends of this tile (A1 was prefetched towards L2, A2 is shared by every matrix): all loads of a
        // thread are issued BEFORE the barrier that frees the ring, so their latency overlaps the barrier and the
        // prefetch below.  A thread owns C(rows wr + 8 mi + 2 tg + {0, 1}, column 8 nj + gq) (swapped operands).
        const bool interior = ti >= jb + RB_CH_NB && ti + TR <= n;   // all rows below the diagonal block, above row n
        double2 c1[2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int nj = 0; nj < 4; ++nj) {
            const int gr = ti + wr + 8 * mi + 2 * tg, gc = min(jb + 8 * nj + gq, n - 1);
            const double *p1 = A1 + (size_t)gc * a.lda1;
            if (interior && v1ok) c1[mi][nj] = *reinterpret_cast<const double2 *>(p1 + gr);
            else c1[mi][nj] = make_double2(p1[min(gr, n - 1)], p1[min(gr + 1, n - 1)]);
          }
        if (jb > 0) __syncthreads();             // the ring is free: its head becomes the C tile
        // stage 0 of the NEXT tile into the last ring buffer: it lands while this tile's epilogue, diagonal
        // block and panel product run.  Next tile: the next 64 rows of this panel, or the first tile of the
        // next panel -- whose stage 0 (columns 0 .. KC-1) is final unless this is panel 0.
        pre = false;
        if (NS >= 3) {
          const bool same = ti + TR < nr_t;
          const int njb = same ? jb : jb + RB_CH_NB, nti = same ? ti + TR : jb + RB_CH_NB;
          if (jb > 0 && njb < n) {
            stage_rows<TR, KC, LDA_T, NT>(As + (NS - 1) * KC * LDA_T, L, ldl, nti, nr, tid);
            stage_rows<32, KC, LDB, NT>(Bs + (NS - 1) * KC * LDB, L, ldl, njb, n, tid);
            if (!same) {                         // the next panel's first tile: its tail rows, if it has a tail
              const int nrem = (nr - njb) % TR;
              if (NT == 128 && nr - njb > TR && nrem > 0 && nrem <= 8) stage_tail(As + (NS - 1) * KC * LDA_T, 0, nr - nrem);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            pre = true;
          }
        }
        if (a.A2) {
          double2 c2[2][4];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
              const int gr = ti + wr + 8 * mi + 2 * tg, gc = min(jb + 8 * nj + gq, n - 1);
              const double *p2 = a.A2 + (size_t)gc * a.lda2;
              if (interior && v2ok) c2[mi][nj] = *reinterpret_cast<const double2 *>(p2 + gr);
              else c2[mi][nj] = make_double2(p2[min(gr, n - 1)], p2[min(gr + 1, n - 1)]);
            }
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) { c1[mi][nj].x += c2[mi][nj].x; c1[mi][nj].y += c2[mi][nj].y; }
        }
        // C = A1 + A2 [+ jitter I] - L L' (row n: the right-hand side), into shared memory as an A operand:
        // conflict-free 16-byte stores (two consecutive rows of a column per thread)
        if (interior) {
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj)
              *reinterpret_cast<double2 *>(As + (8 * nj + gq) * LDA_T + wr + 8 * mi + 2 * tg) =
                  make_double2(c1[mi][nj].x - acc[mi][nj][0], c1[mi][nj].y - acc[mi][nj][1]);
        } else {
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
              double v2[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int gr = ti + wr + 8 * mi + 2 * tg + e, gc = jb + 8 * nj + gq;
                double v = (e ? c1[mi][nj].y : c1[mi][nj].x) + (gr == gc ? jit : 0.0);
                if (gr == n && gc < n) v = a.rhs[(size_t)b * a.stride_rhs + gc] + (a.rhs2 ? a.rhs2[gc] : 0.0);
                v -= acc[mi][nj][e];
                v2[e] = (gc < n && gr >= gc && gr <= n) ? v : 0.0;
              }
              *reinterpret_cast<double2 *>(As + (8 * nj + gq) * LDA_T + wr + 8 * mi + 2 * tg) = make_double2(v2[0], v2[1]);
            }
        }
        if (tailw) {
          // C values of the tail rows into the padding rows TR .. TR+7 of the C tile
          const double (*ts[4])[2] = {&acc[0][1], &acc[0][2], &acc[0][3], &acc[1][2]};
#pragma unroll
          for (int nj = 0; nj < 4; ++nj) {
            double v2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int gr = trow0 + 2 * tg + e, gc = jb + 8 * nj + gq;   // gr > gc always (rows below the panel)
              const int cr = min(gr, n - 1), cc = min(gc, n - 1);
              double v = A1[cr + (size_t)cc * a.lda1] + (a.A2 ? a.A2[cr + (size_t)cc * a.lda2] : 0.0);
              if (gr == n && gc < n) v = a.rhs[(size_t)b * a.stride_rhs + gc] + (a.rhs2 ? a.rhs2[gc] : 0.0);
              v -= (*ts[nj])[e];
              v2[e] = (gc < n && gr <= n) ? v : 0.0;
            }
            *reinterpret_cast<double2 *>(As + (8 * nj + gq) * LDA_T + TR + 2 * tg) = make_double2(v2[0], v2[1]);
          }
        }
        __syncthreads();
        if (ti == jb) {
          // diagonal block (rows 0..31 of this tile): Cholesky AND inverse by all four warps, right-looking,
          // TWO columns per block barrier.  Columns 2P, 2P+1 belong to warp P % NW; thread (lane, warp) holds
          // A(lane, c) and Y(c, lane) for its 32 / NW columns c = 2 NW q' + 2 warp + e (Y = L11^-1, built column-oriented
          // alongside: as soon as column j of L is final, row j of Y is, and both update what is to their
          // right / below).  Owner: pivot -> rsqrt (one Newton step for the square root) -> column j, its
          // effect on column j+1 inside the warp (shuffles), pivot j+1, both columns and both rows of Y into
          // the double-buffered exchange line; barrier; rank-2 updates by everybody.
          {
            double x[NQ], y[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              const int c = 2 * NW * (q >> 1) + 2 * warp + (q & 1);
              const double v = As[c * LDA_T + lane];
              x[q] = (lane < nb && c < nb) ? v : (c == lane ? 1.0 : 0.0);   // identity padding of a narrow last panel
              y[q] = (c == lane) ? 1.0 : 0.0;
            }
#pragma unroll
            for (int P = 0; P < 16; ++P) {
              const int j = 2 * P, qj = 2 * (P / NW), buf = P & 1;
              double *cb = colb + buf * 128;     // [4][32]: L(:, j), L(:, j+1), Y(j, :), Y(j+1, :)
              if (warp == (P % NW)) {
                const double d0 = __shfl_sync(0xffffffffu, x[qj], j);
                double l0 = 0.0, l1 = 0.0, y0 = 0.0, y1 = 0.0;
                bool bad = !(d0 > 0.0);
                if (!bad) {
                  // critical chain: rsqrt -> scaled column -> L(j+1, j) -> column j+1 -> its pivot -> rsqrt.
                  // The square roots themselves (one Newton step each) and the diagonal selects hang off it.
                  const double r0 = rsqrt(d0);
                  const double lo0 = x[qj] * r0;                             // column j below the diagonal (lanes > j)
                  const double l10 = __shfl_sync(0xffffffffu, lo0, j + 1);   // L(j+1, j)
                  const double x1 = fma(-lo0, l10, x[qj + 1]);               // column j+1 (lanes > j)
                  const double d1 = __shfl_sync(0xffffffffu, x1, j + 1);
                  double s0 = d0 * r0;
                  s0 = fma(fma(-s0, s0, d0), 0.5 * r0, s0);
                  l0 = (lane == j) ? s0 : (lane > j ? lo0 : 0.0);
                  y0 = y[qj] * r0;
                  bad = !(d1 > 0.0);
                  if (!bad) {
                    const double r1 = rsqrt(d1);
                    double s1 = d1 * r1;
                    s1 = fma(fma(-s1, s1, d1), 0.5 * r1, s1);
                    l1 = (lane == j + 1) ? s1 : (lane > j + 1 ? x1 * r1 : 0.0);
                    y1 = fma(-l10, y0, y[qj + 1]) * r1;
                  }
                }
                if (bad) {
                  if (lane == 0) s_fail = 1;
                } else {
                  x[qj] = l0; x[qj + 1] = l1; y[qj] = y0; y[qj + 1] = y1;
                  cb[lane] = l0; cb[32 + lane] = l1; cb[64 + lane] = y0; cb[96 + lane] = y1;
                }
              }
              __syncthreads();
              if (s_fail) break;
              const double lr0 = cb[lane], lr1 = cb[32 + lane];          // L(lane, j), L(lane, j+1)
              const double yl0 = cb[64 + lane], yl1 = cb[96 + lane];     // Y(j, lane), Y(j+1, lane)
#pragma unroll
              for (int q = qj; q < NQ; ++q) {
                const int k = 2 * NW * (q >> 1) + 2 * warp + (q & 1);
                const bool upd = k > j + 1;
                const double lk0 = upd ? cb[k] : 0.0, lk1 = upd ? cb[32 + k] : 0.0;   // L(k, j), L(k, j+1)
                x[q] = fma(-lr1, lk1, fma(-lr0, lk0, x[q]));     // A(lane, k) -= L(lane, j) L(k, j) + L(lane, j+1) L(k, j+1)
                y[q] = fma(-lk1, yl1, fma(-lk0, yl0, y[q]));     // Y(k, lane) -= L(k, j) Y(j, lane) + L(k, j+1) Y(j+1, lane)
              }
            }
            if (!s_fail) {
#pragma unroll
              for (int q = 0; q < NQ; ++q) {
                const int c = 2 * NW * (q >> 1) + 2 * warp + (q & 1);
                if (c <= lane && lane < nb) L[(jb + lane) + (size_t)(jb + c) * ldl] = x[q];
                sX[lane * LDB + c] = y[q];                   // X(row c, column lane) as a B operand
              }
            }
          }
          __syncthreads();
          if (s_fail) break;
        }
        // L21 = C21 X' on the tensor cores; rows of the diagonal block itself are skipped at the store
        if (ti + wr + 16 > jb + nb && wact) {    // warp-uniform: this warp owns rows below the diagonal block
          double out[2][4][2];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) out[mi][nj][0] = out[mi][nj][1] = 0.0;
          // X = L11^-1 is lower triangular: X(j, c) = 0 for c > j, so column tile nj only needs c < 8 (nj + 1)
#pragma unroll
          for (int kk = 0; kk < 32; kk += 4) {
            double av[2], bv[4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) av[mi] = As[(kk + tg) * LDA_T + wr + 8 * mi + gq];
#pragma unroll
            for (int nj = 0; nj < 4; ++nj)
              if (kk < 8 * (nj + 1)) bv[nj] = sX[(kk + tg) * LDB + 8 * nj + gq];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
              for (int nj = 0; nj < 4; ++nj)
                if (kk < 8 * (nj + 1)) dmma884(out[mi][nj][0], out[mi][nj][1], bv[nj], av[mi]);
          }
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
              const int gr = ti + wr + 8 * mi + 2 * tg, gc = jb + 8 * nj + gq;
              if (gc < n) {
                double *dst = L + gr + (size_t)gc * ldl;
                if (gr >= jb + nb && gr + 1 < nr) {
                  *reinterpret_cast<double2 *>(dst) = make_double2(out[mi][nj][0], out[mi][nj][1]);
                } else {
                  if (gr >= jb + nb && gr < nr) dst[0] = out[mi][nj][0];
                  if (gr + 1 >= jb + nb && gr + 1 < nr) dst[1] = out[mi][nj][1];
                }
              }
            }
        } else if (tailw) {
          // the tail rows' part of the panel: warp 0 has no rows below the diagonal block in this tile
          double outt[4][2];
#pragma unroll
          for (int nj = 0; nj < 4; ++nj) outt[nj][0] = outt[nj][1] = 0.0;
#pragma unroll
          for (int kk = 0; kk < 32; kk += 4) {
            const double at = As[(kk + tg) * LDA_T + TR + gq];
#pragma unroll
            for (int nj = 0; nj < 4; ++nj)
              if (kk < 8 * (nj + 1)) dmma884(outt[nj][0], outt[nj][1], sX[(kk + tg) * LDB + 8 * nj + gq], at);
          }
#pragma unroll
          for (int nj = 0; nj < 4; ++nj) {
            const int gr = trow0 + 2 * tg, gc = jb + 8 * nj + gq;
            if (gc < n) {
              double *dst = L + gr + (size_t)gc * ldl;
              if (gr + 1 < nr) {
                *reinterpret_cast<double2 *>(dst) = make_double2(outt[nj][0], outt[nj][1]);
              } else if (gr < nr) {
                dst[0] = outt[nj][0];
              }
            }
          }
        }
      }
      __syncthreads();
      if (s_fail) break;
    }
    __syncthreads();
    ok = !s_fail;
    if (!ok && attempt == 0) {
      if (a.jitter < 0.0) break;
      if (tid == 0) atomicAdd(&a.status->used_jitter, 1);
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_all;" ::: "memory");   // a prefetch may be in flight after a failed attempt
  if (!ok) {
    if (tid == 0 && atomicCAS(&a.status->not_pd, 0, 1) == 0) {
      a.status->not_pd_step = a.t;
      a.status->not_pd_particle = b;
    }
    if (tid == 0) { a.sum_log_diag[b] = nan(""); a.vtv[b] = nan(""); }
    return;
  }
  double ld = 0.0, vv = 0.0;
  for (int r = tid; r < n; r += blockDim.x) {
    ld += log(L[r + (size_t)r * ldl]);
    const double v = L[n + (size_t)r * ldl];    // v = L \ rhs
    vv = fma(v, v, vv);
  }
  ld = warp_sum(ld); vv = warp_sum(vv);
  if (lane == 0) { s_red[0][warp] = ld; s_red[1][warp] = vv; }
  __syncthreads();
  if (tid == 0) {
    double l2 = 0.0, v2 = 0.0;
    for (int q = 0; q < NW; ++q) { l2 += s_red[0][q]; v2 += s_red[1][q]; }
    a.sum_log_diag[b] = l2;
    a.vtv[b] = v2;
  }
}

// ---------------------------------------------------------------------------
// Batched Cholesky, panel by panel ACROSS the batch (large batches: C5, N = 4096).
// One CTA per matrix (k_chol_solve) keeps 3-4 matrices per SM in flight and is bound by the latency of
// its own serial chain (operand loads, one warp on the diagonal block, barriers; ncu: issue slots 17 %
// busy).  Here every panel step is a pair of kernels over the whole batch:
//   k_chol_panel_gemm   C(rows, panel) = (A1 + A2)(rows, panel) - L(rows, 0:jb) L(panel, 0:jb)'
//                       128 x 64 x jb tiles on the fp64 tensor cores, grid (row tiles, 1, batch):
//                       thousands of independent tiles, double-buffered cp.async operand staging;
//   k_chol_panel_factor diagonal 64 x 64 block (two 32 x 32 warp-level factorisations + the update
//                       between them) and the triangular solve of the rows below, one CTA per matrix;
// then k_chol_finalize (log-determinant, v'v).  Left-looking: every factor element is written once.
// ---------------------------------------------------------------------------
#define RB_CP_NB 64
struct CholPanelArgs {
  CholArgs c;
  int jb;          // first column of the panel
  int *fail;       // [batch] != 0: a pivot was not positive (matrix not PD)
};

__global__ void __launch_bounds__(256, 2) k_chol_panel_gemm(CholPanelArgs p) {
  extern __shared__ double sm_cp[];
  double *As = sm_cp;                        // [2][32][RB_LDA]
  double *Bs = sm_cp + 2 * 32 * RB_LDA;      // [2][32][RB_LDB]
  const CholArgs &a = p.c;
  const int b = blockIdx.z, n = a.n, nr = n + 1, jb = p.jb;
  const int ti = jb + 128 * blockIdx.x;
  if (ti >= nr) return;
  const double *A1 = a.A1 + (size_t)(a.slot1 ? a.slot1[b] : b) * a.strideA1;
  double *L = a.L + (size_t)b * a.strideL;
  const int ldl = a.ldl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
  const int gq = lane >> 2, tg = lane & 3;
  double acc[4][4][2];
  {   // C tile, negated; all loads issued back to back (clamped addresses, no branches)
    double c1[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj < 4; ++nj)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gr = min(ti + wr + 8 * mi + gq, n - 1), gc = min(jb + wc + 8 * nj + 2 * tg + e, n - 1);
          c1[mi][nj][e] = A1[gr + (size_t)gc * a.lda1] + (a.A2 ? a.A2[gr + (size_t)gc * a.lda2] : 0.0);
        }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj < 4; ++nj)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gr = ti + wr + 8 * mi + gq, gc = jb + wc + 8 * nj + 2 * tg + e;
          double v = c1[mi][nj][e];
          if (gr == n && gc < n) v = a.rhs[(size_t)b * a.stride_rhs + gc] + (a.rhs2 ? a.rhs2[gc] : 0.0);
          acc[mi][nj][e] = (gc < n && gr >= gc && gr <= n) ? -v : 0.0;
        }
  }
  if (jb > 0) {
    stage_rows_k32<128, RB_LDA, 256>(As, L, ldl, ti, nr, tid);
    stage_rows_k32<64, RB_LDB, 256>(Bs, L, ldl, jb, n, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
    int buf = 0;
    for (int k0 = 0; k0 < jb; k0 += 32, buf ^= 1) {
      if (k0 + 32 < jb) {
        stage_rows_k32<128, RB_LDA, 256>(As + (buf ^ 1) * 32 * RB_LDA, L + (size_t)(k0 + 32) * ldl, ldl, ti, nr, tid);
        stage_rows_k32<64, RB_LDB, 256>(Bs + (buf ^ 1) * 32 * RB_LDB, L + (size_t)(k0 + 32) * ldl, ldl, jb, n, tid);
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      mma_tile_k32(As + buf * 32 * RB_LDA, Bs + buf * 32 * RB_LDB, wr, wc, lane, acc);
      __syncthreads();
    }
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int nj = 0; nj < 4; ++nj)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gr = ti + wr + 8 * mi + gq, gc = jb + wc + 8 * nj + 2 * tg + e;
        if (gr < nr && gc < n && gr >= gc) L[gr + (size_t)gc * ldl] = -acc[mi][nj][e];
      }
}

// 32 x 32 Cholesky in shared memory by one warp (left-looking); sD row-major with stride 33.
// Returns false when a pivot is not positive.
__device__ __forceinline__ bool warp_chol32(double *sD, int nb, int lane) {
  for (int j = 0; j < nb; ++j) {
    double s = 0.0;
    if (lane >= j && lane < nb) {
      s = sD[lane * 33 + j];
      for (int k = 0; k < j; ++k) s = fma(-sD[lane * 33 + k], sD[j * 33 + k], s);
    }
    const double dj = __shfl_sync(0xffffffffu, s, j);
    if (!(dj > 0.0)) return false;
    const double ljj = sqrt(dj);
    if (lane == j) sD[j * 33 + j] = ljj;
    else if (lane > j && lane < nb) sD[lane * 33 + j] = s / ljj;
    __syncwarp();
  }
  return true;
}
// x (one row of 32) <- x L^-T by forward substitution, L in shared memory (row-major, stride 33)
__device__ __forceinline__ void row_solve32(double (&x)[32], const double *sD) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    double sx = x[k];
#pragma unroll
    for (int q = 0; q < 32; ++q)
      if (q < k) sx = fma(-x[q], sD[k * 33 + q], sx);
    x[k] = sx / sD[k * 33 + k];
  }
}

__global__ void __launch_bounds__(128, 3) k_chol_panel_factor(CholPanelArgs p) {
  __shared__ double sD1[32 * 33], sD2[32 * 33], sL21[32 * 33];   // L11, L22, L21 (row-major, stride 33)
  __shared__ int s_ok;
  const CholArgs &a = p.c;
  const int b = blockIdx.x, n = a.n, nr = n + 1, jb = p.jb, ldl = a.ldl;
  double *L = a.L + (size_t)b * a.strideL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = min(RB_CP_NB, n - jb), nb1 = min(32, nb), nb2 = nb - nb1;
  if (p.fail[b]) return;                     // an earlier panel failed: the retry path redoes this matrix
  for (int idx = tid; idx < 32 * 33; idx += blockDim.x) { sD1[idx] = 0.0; sD2[idx] = 0.0; sL21[idx] = 0.0; }
  if (tid == 0) s_ok = 1;
  __syncthreads();
  for (int idx = tid; idx < nb1 * nb1; idx += blockDim.x) {
    const int r = idx % nb1, c = idx / nb1;
    if (r >= c) sD1[r * 33 + c] = L[(jb + r) + (size_t)(jb + c) * ldl];
  }
  for (int idx = tid; idx < nb2 * nb1; idx += blockDim.x) {   // A21 of the diagonal block
    const int r = idx % nb2, c = idx / nb2;
    sL21[r * 33 + c] = L[(jb + 32 + r) + (size_t)(jb + c) * ldl];
  }
  for (int idx = tid; idx < nb2 * nb2; idx += blockDim.x) {
    const int r = idx % nb2, c = idx / nb2;
    if (r >= c) sD2[r * 33 + c] = L[(jb + 32 + r) + (size_t)(jb + 32 + c) * ldl];
  }
  __syncthreads();
  if (warp == 0 && !warp_chol32(sD1, nb1, lane) && lane == 0) s_ok = 0;
  __syncthreads();
  if (s_ok && nb2 > 0) {
    if (warp == 0) {   // L21 = A21 L11^-T: lane = row
      if (lane < nb2) {
        double x[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) x[k] = sL21[lane * 33 + k];
        row_solve32(x, sD1);
#pragma unroll
        for (int k = 0; k < 32; ++k) sL21[lane * 33 + k] = x[k];
      }
      __syncwarp();
      // A22 -= L21 L21'  (lower part), then factor it
      for (int idx = lane; idx < nb2 * nb2; idx += 32) {
        const int r = idx % nb2, c = idx / nb2;
        if (r >= c) {
          double s = sD2[r * 33 + c];
          for (int k = 0; k < 32; ++k) s = fma(-sL21[r * 33 + k], sL21[c * 33 + k], s);
          sD2[r * 33 + c] = s;
        }
      }
      __syncwarp();
      if (!warp_chol32(sD2, nb2, lane) && lane == 0) s_ok = 0;
    }
  }
  __syncthreads();
  if (!s_ok) { if (tid == 0) p.fail[b] = 1; return; }
  // the factored diagonal block
  for (int idx = tid; idx < nb1 * nb1; idx += blockDim.x) {
    const int r = idx % nb1, c = idx / nb1;
    if (r >= c) L[(jb + r) + (size_t)(jb + c) * ldl] = sD1[r * 33 + c];
  }
  for (int idx = tid; idx < nb2 * nb1; idx += blockDim.x) {
    const int r = idx % nb2, c = idx / nb2;
    L[(jb + 32 + r) + (size_t)(jb + c) * ldl] = sL21[r * 33 + c];
  }
  for (int idx = tid; idx < nb2 * nb2; idx += blockDim.x) {
    const int r = idx % nb2, c = idx / nb2;
    if (r >= c) L[(jb + 32 + r) + (size_t)(jb + 32 + c) * ldl] = sD2[r * 33 + c];
  }
  // rows below the diagonal block (row n = right-hand side): X = C L_JJ^-T, 32 columns at a time
  const int r0 = jb + nb;
  if (nb == RB_CP_NB) {
#pragma unroll 1
    for (int r = r0 + tid; r < nr; r += blockDim.x) {
      asm volatile("" ::: "memory");
      double x[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = L[r + (size_t)(jb + k) * ldl];
      row_solve32(x, sD1);
#pragma unroll
      for (int k = 0; k < 32; ++k) L[r + (size_t)(jb + k) * ldl] = x[k];
      double y[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) y[k] = L[r + (size_t)(jb + 32 + k) * ldl];
#pragma unroll
      for (int k = 0; k < 32; ++k) {     // second half: c2 - x1 L21'
        double sy = y[k];
#pragma unroll
        for (int q = 0; q < 32; ++q) sy = fma(-x[q], sL21[k * 33 + q], sy);
        y[k] = sy;
      }
      row_solve32(y, sD2);
#pragma unroll
      for (int k = 0; k < 32; ++k) L[r + (size_t)(jb + 32 + k) * ldl] = y[k];
    }
  } else {   // last, narrower panel: only the right-hand-side row is left below it
    if (tid == 0) {
      double x[RB_CP_NB];
      for (int k = 0; k < nb; ++k) x[k] = L[n + (size_t)(jb + k) * ldl];
      for (int k = 0; k < nb; ++k) {
        double sx = x[k];
        for (int q = 0; q < k; ++q) {
          const double lkq = k < 32 ? sD1[k * 33 + q] : (q < 32 ? sL21[(k - 32) * 33 + q] : sD2[(k - 32) * 33 + (q - 32)]);
          sx = fma(-x[q], lkq, sx);
        }
        x[k] = sx / (k < 32 ? sD1[k * 33 + k] : sD2[(k - 32) * 33 + (k - 32)]);
      }
      for (int k = 0; k < nb; ++k) L[n + (size_t)(jb + k) * ldl] = x[k];
    }
  }
}

// sum(log(diag(L))) and v'v with v = L \ rhs (row n of the workspace); failed matrices get NaN
__global__ void __launch_bounds__(128) k_chol_finalize(CholArgs a, const int *__restrict__ fail) {
  __shared__ double s_red[2][4];
  const int b = blockIdx.x, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double *L = a.L + (size_t)b * a.strideL;
  if (fail[b]) {
    if (tid == 0) { a.sum_log_diag[b] = nan(""); a.vtv[b] = nan(""); }
    return;
  }
  double ld = 0.0, vv = 0.0;
  for (int r = tid; r < n; r += blockDim.x) {
    ld += log(L[r + (size_t)r * a.ldl]);
    const double v = L[n + (size_t)r * a.ldl];
    vv = fma(v, v, vv);
  }
  ld = warp_sum(ld); vv = warp_sum(vv);
  if (lane == 0) { s_red[0][warp] = ld; s_red[1][warp] = vv; }
  __syncthreads();
  if (tid == 0) {
    a.sum_log_diag[b] = (s_red[0][0] + s_red[0][1]) + (s_red[0][2] + s_red[0][3]);
    a.vtv[b] = (s_red[1][0] + s_red[1][1]) + (s_red[1][2] + s_red[1][3]);
  }
}

static inline size_t chol_solve_smem(int nt) {
  return sizeof(double) * (32 * 34 + 2 * 32 * (nt / 2 + 8) + 2 * 32 * RB_CH_LDB);
}
// workspace leading dimension: n rows + the right-hand-side row, 64-byte aligned columns
static inline int chol_ldl(int n) { return ((n + 1 + 7) / 8) * 8; }

}  // namespace rb
