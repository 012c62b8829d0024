// Sibling fusion for the streaming Kalman pass.
//
// After resampling, every surviving ancestor slab is the source of k >= 1 offspring
// (Poisson-like: k = 1: 37 %, 2: 18 %, 3: 6 % of the population for near-uniform weights).
// k_stream_pass streams the ancestor slab once PER OFFSPRING.  Here all offspring of one
// ancestor form a FAMILY that one CTA handles together: the slab tile is read once, the
// ancestor's pending downdate is applied once, and the tile is multiplied with every
// sibling's own H and written to every sibling's slab (CB siblings per pass over the slab).
// HBM traffic per particle-step drops from 16 M^2 B to (8/k_avg_pass + 8) M^2 B, and the
// copy-before-in-place hazard disappears inside a family: the in-place sibling (keeper) is
// always in the last batch of its family, which the same CTA processes last.
// Families larger than KF siblings spill their surplus into "surplus families" (copies only)
// that run in an earlier launch, so one heavy ancestor cannot serialise a whole step.
#pragma once
#include "kalman_stream.cuh"
#include "shard_plan.cuh"

namespace rb {

#define RB_CB 3                // siblings per pass over the slab
#define RB_KF (2 * RB_CB)      // siblings handled by the main family item (incl. the keeper)

struct FamLists {
  int *work_counter;   // [1] dynamic work distribution (zeroed before the launch)
  const int *n_fam;    // [1] number of families
  const int *src;      // [n_fam] source slab
  const int *anc;      // [n_fam] ancestor index for the pending (G,KS) arrays
  const int *first;    // [n_fam] first entry in child[]
  const int *cnt;      // [n_fam] number of siblings (keeper, if any, is the last one)
  const int *child;    // item indices (H4 / PHp / dst_slot index)
};

// ---------------------------------------------------------------------------
// family construction (single CTA, after the slot plan)
// ---------------------------------------------------------------------------
struct FamBuildArgs {
  int n_items, n_slabs;
  const int *n_items_dev;   // optional device-side item count (overrides n_items)
  const int *src_slot;      // [n_items] source slab of each item
  const int *dst_slot;      // [n_items] or nullptr (= item index)
  const int *anc;           // [n_items] or nullptr (= source slab)
  const int *item_group;    // [n_items] or nullptr: only items with item_group[j] == group take part
  int group;
  int *s_cnt, *s_keeper, *s_cursor, *s_first, *s_fid, *s_xoff, *s_xfam;   // [n_slabs] scratch
  // main families (launch 2) and surplus families (launch 1)
  int *fb_src, *fb_anc, *fb_first, *fb_cnt, *fb_child, *n_fb;
  int *fa_src, *fa_anc, *fa_first, *fa_cnt, *fa_child, *n_fa;
  int *work_ctr;            // [2] dynamic work counters of the two k_stream_fam launches, zeroed here
  int cb, kf;               // siblings per pass / per main family (RB_CB, RB_KF; 2, 4 for the symmetric kernel)
};

__global__ void __launch_bounds__(1024) k_build_families(FamBuildArgs p) {
  __shared__ int s_w[32 * 4];
  const int n = p.n_items_dev ? *p.n_items_dev : p.n_items, ns = p.n_slabs, tid = threadIdx.x;
  for (int s = tid; s < ns; s += blockDim.x) { p.s_cnt[s] = 0; p.s_keeper[s] = -1; p.s_cursor[s] = 0; }
  __syncthreads();
  for (int j = tid; j < n; j += blockDim.x) {
    if (p.item_group && p.item_group[j] != p.group) continue;
    const int s = p.src_slot[j];
    atomicAdd(&p.s_cnt[s], 1);
    if ((p.dst_slot ? p.dst_slot[j] : j) == s) p.s_keeper[s] = j;
  }
  __syncthreads();
  const int per = (ns + blockDim.x - 1) / blockDim.x;
  const int b = min(ns, tid * per), e = min(ns, b + per);
  int v[4] = {0, 0, 0, 0}, tot[4];   // families, main children, surplus families, surplus children
  for (int s = b; s < e; ++s) {
    const int c = p.s_cnt[s];
    if (!c) continue;
    const int x = max(0, c - p.kf);
    v[0] += 1; v[1] += min(c, p.kf); v[2] += (x + p.cb - 1) / p.cb; v[3] += x;
  }
  block_scan_vec<4>(v, tot, s_w);
  for (int s = b; s < e; ++s) {
    const int c = p.s_cnt[s];
    if (!c) continue;
    const int x = max(0, c - p.kf), nm = min(c, p.kf);
    p.s_fid[s] = v[0]; p.s_first[s] = v[1]; p.s_xfam[s] = v[2]; p.s_xoff[s] = v[3];
    p.fb_src[v[0]] = s; p.fb_first[v[0]] = v[1]; p.fb_cnt[v[0]] = nm;
    for (int q = 0; q * p.cb < x; ++q) {
      p.fa_src[v[2] + q] = s; p.fa_first[v[2] + q] = v[3] + q * p.cb; p.fa_cnt[v[2] + q] = min(p.cb, x - q * p.cb);
    }
    v[0] += 1; v[1] += nm; v[2] += (x + p.cb - 1) / p.cb; v[3] += x;
  }
  if (tid == 0) { *p.n_fb = tot[0]; *p.n_fa = tot[2]; p.work_ctr[0] = 0; p.work_ctr[1] = 0; }
  __syncthreads();
  for (int j = tid; j < n; j += blockDim.x) {
    if (p.item_group && p.item_group[j] != p.group) continue;
    const int s = p.src_slot[j];
    const int c = p.s_cnt[s], nm = min(c, p.kf), kp = p.s_keeper[s];
    const int an = p.anc ? p.anc[j] : s;
    const int fid = p.s_fid[s];
    if (j == kp) {
      p.fb_child[p.s_first[s] + nm - 1] = j;      // the in-place sibling is processed last
      p.fb_anc[fid] = an;
    } else {
      const int pos = atomicAdd(&p.s_cursor[s], 1);
      const int lim = nm - (kp >= 0 ? 1 : 0);
      if (pos < lim) {
        p.fb_child[p.s_first[s] + pos] = j;
        if (kp < 0 && pos == 0) p.fb_anc[fid] = an;
      } else {
        const int x = pos - lim;
        p.fa_child[p.s_xoff[s] + x] = j;
        if (x % p.cb == 0) p.fa_anc[p.s_xfam[s] + x / p.cb] = an;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// the fused pass: one family per item, CB siblings per pass over the slab
// ---------------------------------------------------------------------------
template <int D, int R2, int KC, int S, int CB>
__global__ void __launch_bounds__(RB_STREAM_THREADS, 1)
k_stream_fam(StreamArgs a, FamLists f) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) uint64_t full[S];
  const int ld = a.ld, M = a.M;
  const int npairs = ld >> 1;
  const size_t stage_doubles = (size_t)KC * ld + (size_t)4 * KC * (1 + CB);
  double *stages = reinterpret_cast<double *>(smraw);
  const int tid = threadIdx.x;
  const int n_items = (*f.n_fam) * a.nsplit;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // ---- producer (thread 0): walks items -> batches -> column chunks, S stages ahead.
  // Items are claimed dynamically (atomic counter): families cost 1..KF/CB passes, a static
  // round-robin leaves CTAs idle at the end of the launch.  Claimed ids go through a small
  // shared ring so that the consumers process exactly the sequence the producer issued.
  __shared__ int s_ring[8];
  int p_claim = 0;                 // number of items claimed so far (producer)
  int p_it = 0, p_b = 0, p_c = 0, p_q = 0;
  auto claim = [&]() {
    p_it = atomicAdd(f.work_counter, 1);
    s_ring[p_claim & 7] = p_it;
    ++p_claim;
  };
  if (tid == 0) claim();
  auto issue = [&]() {
    if (p_it >= n_items) return;
    const int fam = p_it / a.nsplit, sp = p_it % a.nsplit;
    const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);
    const int c = c0 + p_c, ncols = min(KC, c1 - c);
    const int cnt = f.cnt[fam], first = f.first[fam];
    const int nbat = (cnt + CB - 1) / CB;
    const int nv = min(CB, cnt - p_b * CB);
    double *st = stages + (size_t)(p_q % S) * stage_doubles;
    uint64_t *bar = &full[p_q % S];
    const uint32_t bytes_p = (uint32_t)ncols * ld * 8u, bytes_v = (uint32_t)ncols * 32u;
    mbar_expect_tx(bar, bytes_p + (1 + nv) * bytes_v);
    tma_load_1d(st, a.P + (size_t)f.src[fam] * a.slab + (size_t)c * ld, bytes_p, bar);
    tma_load_1d(st + (size_t)KC * ld, a.G4prev + ((size_t)f.anc[fam] * ld + c) * 4, bytes_v, bar);
    for (int q = 0; q < nv; ++q) {
      const int ch = f.child[first + p_b * CB + q];
      tma_load_1d(st + (size_t)KC * ld + 4 * KC * (1 + q), a.H4 + ((size_t)ch * ld + c) * 4, bytes_v, bar);
    }
    ++p_q;
    p_c += KC;
    if (c0 + p_c >= c1) {
      p_c = 0;
      if (++p_b >= nbat) { p_b = 0; claim(); }
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) issue();
  }

  // ---- consumers
  int q = 0;
  __syncthreads();                 // first claim and the first S issues are visible
  for (int c_claim = 0;; ++c_claim) {
    const int it = s_ring[c_claim & 7];   // written by the producer at least one barrier ago
    if (it >= n_items) break;
    const int fam = it / a.nsplit, sp = it % a.nsplit;
    const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);
    const int an = f.anc[fam], cnt = f.cnt[fam], first = f.first[fam];
    double2 ks[R2][D];
#pragma unroll
    for (int k = 0; k < R2; ++k) {
      const int rp = tid + k * RB_STREAM_THREADS;
      if (rp < npairs) {
        const double4 *kp = reinterpret_cast<const double4 *>(a.KS4prev + ((size_t)an * ld + 2 * rp) * 4);
        const double4 k0 = kp[0], k1 = kp[1];
        const double r0[4] = {k0.x, k0.y, k0.z, k0.w}, r1[4] = {k1.x, k1.y, k1.z, k1.w};
#pragma unroll
        for (int bq = 0; bq < D; ++bq) ks[k][bq] = make_double2(r0[bq], r1[bq]);
      } else {
#pragma unroll
        for (int bq = 0; bq < D; ++bq) ks[k][bq] = make_double2(0.0, 0.0);
      }
    }
    for (int b0 = 0; b0 < cnt; b0 += CB) {
      const int nv = min(CB, cnt - b0);
      int child[CB];
      double *Pd[CB];
#pragma unroll
      for (int s = 0; s < CB; ++s) {
        child[s] = s < nv ? f.child[first + b0 + s] : -1;
        Pd[s] = s < nv ? a.P + (size_t)a.dst_slot[child[s]] * a.slab : nullptr;
      }
      double2 acc[CB][R2][D];
#pragma unroll
      for (int s = 0; s < CB; ++s)
#pragma unroll
        for (int k = 0; k < R2; ++k)
#pragma unroll
          for (int bq = 0; bq < D; ++bq) acc[s][k][bq] = make_double2(0.0, 0.0);
      for (int c = c0; c < c1; c += KC, ++q) {
        const double *st = stages + (size_t)(q % S) * stage_doubles;
        mbar_wait(&full[q % S], (uint32_t)((q / S) & 1));
        const int ncols = min(KC, c1 - c);
#pragma unroll
        for (int u = 0; u < KC; ++u) {
          if (u < ncols) {
            const double4 g4 = *reinterpret_cast<const double4 *>(st + (size_t)KC * ld + 4 * u);
            const double g[4] = {g4.x, g4.y, g4.z, g4.w};
            double h[CB][4];
#pragma unroll
            for (int s = 0; s < CB; ++s) {
              const double4 h4 = *reinterpret_cast<const double4 *>(st + (size_t)KC * ld + 4 * KC * (1 + s) + 4 * u);
              h[s][0] = h4.x; h[s][1] = h4.y; h[s][2] = h4.z; h[s][3] = h4.w;
            }
            const double2 *col = reinterpret_cast<const double2 *>(st + (size_t)u * ld);
#pragma unroll
            for (int k = 0; k < R2; ++k) {
              const int rp = tid + k * RB_STREAM_THREADS;
              if (rp < npairs) {
                double2 v = col[rp];
#pragma unroll
                for (int bq = 0; bq < D; ++bq) {   // the ancestor's pending downdate, once per tile
                  v.x = fma(-ks[k][bq].x, g[bq], v.x);
                  v.y = fma(-ks[k][bq].y, g[bq], v.y);
                }
#pragma unroll
                for (int s = 0; s < CB; ++s) {
                  if (s < nv) {                     // every sibling: own H, own slab
#pragma unroll
                    for (int bq = 0; bq < D; ++bq) {
                      acc[s][k][bq].x = fma(v.x, h[s][bq], acc[s][k][bq].x);
                      acc[s][k][bq].y = fma(v.y, h[s][bq], acc[s][k][bq].y);
                    }
                    reinterpret_cast<double2 *>(Pd[s] + (size_t)(c + u) * ld)[rp] = v;
                  }
                }
              }
            }
          }
        }
        __syncthreads();
        if (tid == 0) issue();
      }
#pragma unroll
      for (int s = 0; s < CB; ++s) {
        if (s < nv) {
          double *out = a.PHp + ((size_t)child[s] * a.nsplit + sp) * ld * 4;
#pragma unroll
          for (int k = 0; k < R2; ++k) {
            const int rp = tid + k * RB_STREAM_THREADS;
            if (rp < npairs) {
              double o0[4] = {0, 0, 0, 0}, o1[4] = {0, 0, 0, 0};
#pragma unroll
              for (int bq = 0; bq < D; ++bq) { o0[bq] = acc[s][k][bq].x; o1[bq] = acc[s][k][bq].y; }
              double4 *op = reinterpret_cast<double4 *>(out + (size_t)2 * rp * 4);
              op[0] = make_double4(o0[0], o0[1], o0[2], o0[3]);
              op[1] = make_double4(o1[0], o1[1], o1[2], o1[3]);
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Symmetric variant (kalman_variant 4, opt-in): stream only the LOWER triangle of each slab.
//
// P is symmetric and the downdate K SS K' = G SS G' is symmetric, so the strictly upper
// triangle carries no information.  This kernel reads and writes, per column c, only the
// rows r >= c (one TMA bulk copy per column, starting at the even row 2*(c>>1)); the upper
// triangle of the slab is never touched again after initialisation and is not valid.
// P H' is assembled from the triangle: element (r,c), r > c, contributes P(r,c) H(b,c) to
// row r ("row side", per-thread register accumulators as in k_stream_fam) and
// P(r,c) H(b,r) to row c ("column side": per column a block-wide reduction - warp shuffles,
// then one partial per warp in shared memory, summed after the stage).  The column-side
// sums go to a second partial slot of PHp (k_innov4 adds all slots).  HBM traffic per
// particle-step halves; the arithmetic per stored element doubles (still far under the
// fp64 ridge).  Results differ from the full-storage kernels by rounding only (the
// reference does not symmetrise P, its two triangles differ by O(eps)).
//   threads: 288 (9 warps) so that R2 = 2 row pairs per thread cover ld <= 1152 and the
//   sibling's H rows fit the register file next to the accumulators; CB = 2 siblings per pass.
// ---------------------------------------------------------------------------
#define RB_SYM_THREADS 288
#define RB_SYM_CB 2
#define RB_SYM_WARPS (RB_SYM_THREADS / 32)

template <int D, int R2, int KC, int S, int CB>
__global__ void __launch_bounds__(RB_SYM_THREADS, 1)
k_stream_fam_sym(StreamArgs a, FamLists f) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) uint64_t full[S];
  __shared__ double s_col[2][KC][RB_SYM_WARPS][CB * D];   // column-side partials, one per warp
  __shared__ int s_ring[8];
  const int ld = a.ld, M = a.M;
  const int npairs = ld >> 1;
  const size_t stage_doubles = (size_t)KC * ld + (size_t)4 * KC * (1 + CB);
  double *stages = reinterpret_cast<double *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_items = (*f.n_fam) * a.nsplit;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // ---- producer (thread 0), same item -> batch -> chunk walk as k_stream_fam --------------
  int p_claim = 0;
  int p_it = 0, p_b = 0, p_c = 0, p_q = 0;
  auto claim = [&]() {
    p_it = atomicAdd(f.work_counter, 1);
    s_ring[p_claim & 7] = p_it;
    ++p_claim;
  };
  if (tid == 0) claim();
  auto issue = [&]() {
    if (p_it >= n_items) return;
    const int fam = p_it / a.nsplit, sp = p_it % a.nsplit;
    const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);
    const int c = c0 + p_c, ncols = min(KC, c1 - c);
    const int cnt = f.cnt[fam], first = f.first[fam];
    const int nbat = (cnt + CB - 1) / CB;
    const int nv = min(CB, cnt - p_b * CB);
    double *st = stages + (size_t)(p_q % S) * stage_doubles;
    uint64_t *bar = &full[p_q % S];
    const uint32_t bytes_v = (uint32_t)ncols * 32u;
    const double *src = a.P + (size_t)f.src[fam] * a.slab;
    if (a.hints & 2) {   // diagnostic: whole columns in one copy (full-storage traffic, triangle arithmetic)
      const uint32_t bytes_p = (uint32_t)ncols * ld * 8u;
      mbar_expect_tx(bar, bytes_p + (1 + nv) * bytes_v);
      tma_load_1d(st, src + (size_t)c * ld, bytes_p, bar);
    } else {
      uint32_t bytes_p = 0;
      for (int u = 0; u < ncols; ++u) bytes_p += (uint32_t)(ld - (((c + u) >> 1) << 1)) * 8u;
      mbar_expect_tx(bar, bytes_p + (1 + nv) * bytes_v);
      for (int u = 0; u < ncols; ++u) {   // rows r0.. of column c+u, r0 = the even row at or above the diagonal
        const int r0 = ((c + u) >> 1) << 1;
        tma_load_1d(st + (size_t)u * ld + r0, src + (size_t)(c + u) * ld + r0, (uint32_t)(ld - r0) * 8u, bar);
      }
    }
    tma_load_1d(st + (size_t)KC * ld, a.G4prev + ((size_t)f.anc[fam] * ld + c) * 4, bytes_v, bar);
    for (int q = 0; q < nv; ++q) {
      const int ch = f.child[first + p_b * CB + q];
      tma_load_1d(st + (size_t)KC * ld + 4 * KC * (1 + q), a.H4 + ((size_t)ch * ld + c) * 4, bytes_v, bar);
    }
    ++p_q;
    p_c += KC;
    if (c0 + p_c >= c1) {
      p_c = 0;
      if (++p_b >= nbat) { p_b = 0; claim(); }
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) issue();
  }

  // ---- consumers ---------------------------------------------------------------------
  int q = 0;
  __syncthreads();
  for (int c_claim = 0;; ++c_claim) {
    const int it = s_ring[c_claim & 7];
    if (it >= n_items) break;
    const int fam = it / a.nsplit, sp = it % a.nsplit;
    const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);
    const int an = f.anc[fam], cnt = f.cnt[fam], first = f.first[fam];
    double2 ks[R2][D];
#pragma unroll
    for (int k = 0; k < R2; ++k) {
      const int rp = tid + k * RB_SYM_THREADS;
      if (rp < npairs) {
        const double4 *kp = reinterpret_cast<const double4 *>(a.KS4prev + ((size_t)an * ld + 2 * rp) * 4);
        const double4 k0 = kp[0], k1 = kp[1];
        const double r0[4] = {k0.x, k0.y, k0.z, k0.w}, r1[4] = {k1.x, k1.y, k1.z, k1.w};
#pragma unroll
        for (int bq = 0; bq < D; ++bq) ks[k][bq] = make_double2(r0[bq], r1[bq]);
      } else {
#pragma unroll
        for (int bq = 0; bq < D; ++bq) ks[k][bq] = make_double2(0.0, 0.0);
      }
    }
    for (int b0 = 0; b0 < cnt; b0 += CB) {
      const int nv = min(CB, cnt - b0);
      int child[CB];
      double *Pd[CB];
#pragma unroll
      for (int s = 0; s < CB; ++s) {
        child[s] = s < nv ? f.child[first + b0 + s] : -1;
        Pd[s] = s < nv ? a.P + (size_t)a.dst_slot[child[s]] * a.slab : nullptr;
      }
      // H of every sibling at this thread's own rows (column side) and zeroed accumulators
      double2 acc[CB][R2][D], hrow[CB][R2][D];
#pragma unroll
      for (int s = 0; s < CB; ++s)
#pragma unroll
        for (int k = 0; k < R2; ++k) {
          const int rp = tid + k * RB_SYM_THREADS;
          double h0[4] = {0, 0, 0, 0}, h1[4] = {0, 0, 0, 0};
          if (s < nv && rp < npairs) {
            const double4 *hp = reinterpret_cast<const double4 *>(a.H4 + ((size_t)child[s] * ld + 2 * rp) * 4);
            const double4 x0 = hp[0], x1 = hp[1];
            h0[0] = x0.x; h0[1] = x0.y; h0[2] = x0.z; h0[3] = x0.w;
            h1[0] = x1.x; h1[1] = x1.y; h1[2] = x1.z; h1[3] = x1.w;
          }
#pragma unroll
          for (int bq = 0; bq < D; ++bq) {
            hrow[s][k][bq] = make_double2(h0[bq], h1[bq]);
            acc[s][k][bq] = make_double2(0.0, 0.0);
          }
        }
      // the column-side slot of PHp: this item writes columns [c0, c1), everything else is zero
#pragma unroll
      for (int s = 0; s < CB; ++s) {
        if (s < nv) {
          double4 *oc = reinterpret_cast<double4 *>(a.PHp + ((size_t)child[s] * (2 * a.nsplit) + a.nsplit + sp) * ld * 4);
          for (int r = tid; r < ld; r += RB_SYM_THREADS)
            if (r < c0 || r >= c1) oc[r] = make_double4(0.0, 0.0, 0.0, 0.0);
        }
      }
      for (int c = c0; c < c1; c += KC, ++q) {
        const double *st = stages + (size_t)(q % S) * stage_doubles;
        mbar_wait(&full[q % S], (uint32_t)((q / S) & 1));
        const int ncols = min(KC, c1 - c);
#pragma unroll
        for (int u = 0; u < KC; ++u) {
          if (u < ncols) {
            const int cc = c + u, p0 = cc >> 1;
            const bool odd = cc & 1;
            const double4 g4 = *reinterpret_cast<const double4 *>(st + (size_t)KC * ld + 4 * u);
            const double g[4] = {g4.x, g4.y, g4.z, g4.w};
            double h[CB][4];
#pragma unroll
            for (int s = 0; s < CB; ++s) {
              const double4 h4 = *reinterpret_cast<const double4 *>(st + (size_t)KC * ld + 4 * KC * (1 + s) + 4 * u);
              h[s][0] = h4.x; h[s][1] = h4.y; h[s][2] = h4.z; h[s][3] = h4.w;
            }
            double colacc[CB][D];
#pragma unroll
            for (int s = 0; s < CB; ++s)
#pragma unroll
              for (int bq = 0; bq < D; ++bq) colacc[s][bq] = 0.0;
            const double2 *col = reinterpret_cast<const double2 *>(st + (size_t)u * ld);
#pragma unroll
            for (int k = 0; k < R2; ++k) {
              const int rp = tid + k * RB_SYM_THREADS;
              if (rp < npairs && rp >= p0) {
                double2 v = col[rp];
#pragma unroll
                for (int bq = 0; bq < D; ++bq) {   // the ancestor's pending downdate, once per element
                  v.x = fma(-ks[k][bq].x, g[bq], v.x);
                  v.y = fma(-ks[k][bq].y, g[bq], v.y);
                }
                // pair holding the diagonal: odd column -> (upper, diagonal), even -> (diagonal, lower)
                const bool dg = rp == p0;
                if (dg && odd) v.x = 0.0;                       // strictly upper: not part of the triangle
                const double wx = dg ? 0.0 : v.x;               // strictly lower elements feed the column side
                const double wy = (dg && odd) ? 0.0 : v.y;
#pragma unroll
                for (int s = 0; s < CB; ++s) {
                  if (s < nv) {
#pragma unroll
                    for (int bq = 0; bq < D; ++bq) {
                      acc[s][k][bq].x = fma(v.x, h[s][bq], acc[s][k][bq].x);      // row side: P(r,c) H(b,c)
                      acc[s][k][bq].y = fma(v.y, h[s][bq], acc[s][k][bq].y);
                      colacc[s][bq] = fma(wx, hrow[s][k][bq].x, colacc[s][bq]);   // column side: P(r,c) H(b,r)
                      colacc[s][bq] = fma(wy, hrow[s][k][bq].y, colacc[s][bq]);
                    }
                    reinterpret_cast<double2 *>(Pd[s] + (size_t)cc * ld)[rp] = v;
                  }
                }
              }
            }
#pragma unroll
            for (int s = 0; s < CB; ++s) {
              if (s < nv) {   // block-uniform: no shuffles for absent siblings
#pragma unroll
                for (int bq = 0; bq < D; ++bq) {
                  const double r = warp_sum(colacc[s][bq]);
                  if (lane == 0) s_col[q & 1][u][wid][s * D + bq] = r;
                }
              }
            }
          }
        }
        __syncthreads();
        if (tid == 0) issue();
        if (tid < KC * CB) {   // one thread per (column, sibling): add the warps' partials in fixed order
          const int u = tid / CB, s = tid % CB;
          if (u < ncols && s < nv) {
            double o[4] = {0, 0, 0, 0};
#pragma unroll
            for (int bq = 0; bq < D; ++bq) {
              double t = 0.0;
#pragma unroll
              for (int w = 0; w < RB_SYM_WARPS; ++w) t += s_col[q & 1][u][w][s * D + bq];
              o[bq] = t;
            }
            int ch = child[0];
#pragma unroll
            for (int s2 = 1; s2 < CB; ++s2) if (s == s2) ch = child[s2];
            double4 *oc = reinterpret_cast<double4 *>(a.PHp + ((size_t)ch * (2 * a.nsplit) + a.nsplit + sp) * ld * 4);
            oc[c + u] = make_double4(o[0], o[1], o[2], o[3]);
          }
        }
      }
#pragma unroll
      for (int s = 0; s < CB; ++s) {
        if (s < nv) {
          double *out = a.PHp + ((size_t)child[s] * (2 * a.nsplit) + sp) * ld * 4;
#pragma unroll
          for (int k = 0; k < R2; ++k) {
            const int rp = tid + k * RB_SYM_THREADS;
            if (rp < npairs) {
              double o0[4] = {0, 0, 0, 0}, o1[4] = {0, 0, 0, 0};
#pragma unroll
              for (int bq = 0; bq < D; ++bq) { o0[bq] = acc[s][k][bq].x; o1[bq] = acc[s][k][bq].y; }
              double4 *op = reinterpret_cast<double4 *>(out + (size_t)2 * rp * 4);
              op[0] = make_double4(o0[0], o0[1], o0[2], o0[3]);
              op[1] = make_double4(o1[0], o1[1], o1[2], o1[3]);
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Symmetric variant on the fp64 tensor cores (kalman_variant 5, opt-in).
//
// Same data movement as k_stream_fam_sym (lower triangle only; here in 8-row blocks, so a
// stage of 8 columns c..c+7 holds the row blocks j >= c/8), but every product is an
// mma.sync.m8n8k4.f64 on an 8x8 tile T = P_old(8j.., c..) read from the shared-memory stage:
//   downdate  T_new = T - KS_blk G_blk'           1 MMA, accumulators = T, stored to every sibling
//   row side  PHrow(8j.., n) += T(:, u) H_n(c+u)   2 MMAs (k = 4 columns each)
//   col side  PHcol(c.., n)  += T(r, :)' H_n(r)    2 MMAs (k = 4 rows each); the reduction over
//             rows, which costs the SIMT version a shuffle tree per column, happens in the MMA
// with n = (sibling, measurement row) = 2 x 4 output columns.  The products use the tile BEFORE
// its pending downdate; k_innov4 completes them (Innov4Args::G4prev).  Only the diagonal block
// needs masking.  Warp w owns the row blocks j = w (mod 8): every warp gets the same share of a
// triangular stage, and the row-side accumulators of its <= MAXQ blocks stay in registers.
// Stage columns are padded to ld + 2 doubles: every fragment read is two shared-memory wavefronts.
// ---------------------------------------------------------------------------
#define RB_SYMT_THREADS 256
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MAXQ>
__global__ void __launch_bounds__(RB_SYMT_THREADS, 1)
k_stream_fam_symt(StreamArgs a, FamLists f) {
  constexpr int KC = 8, S = 2, CB = 2, NW = RB_SYMT_THREADS / 32;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) uint64_t full[S];
  __shared__ __align__(16) double s_colp[2][NW][64];   // column-side accumulator tiles, one per warp
  __shared__ int s_ring[8];
  const int ld = a.ld, M = a.M, lds = a.ld + 2, nblk = a.ld >> 3;
  const size_t stage_doubles = (size_t)KC * lds + (size_t)4 * KC * (1 + CB);
  double *stages = reinterpret_cast<double *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;   // MMA fragment coordinates
  const int n_items = (*f.n_fam) * a.nsplit;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // ---- producer (thread 0): item -> batch -> 8-column chunk, S stages ahead -------------
  int p_claim = 0;
  int p_it = 0, p_b = 0, p_c = 0, p_q = 0;
  auto claim = [&]() {
    p_it = atomicAdd(f.work_counter, 1);
    s_ring[p_claim & 7] = p_it;
    ++p_claim;
  };
  if (tid == 0) claim();
  auto issue = [&]() {
    if (p_it >= n_items) return;
    const int fam = p_it / a.nsplit, sp = p_it % a.nsplit;
    const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);        // cw is a multiple of 8
    const int c = c0 + p_c, ncols = min(KC, c1 - c);
    const int cnt = f.cnt[fam], first = f.first[fam];
    const int nbat = (cnt + CB - 1) / CB;
    const int nv = min(CB, cnt - p_b * CB);
    double *st = stages + (size_t)(p_q % S) * stage_doubles;
    uint64_t *bar = &full[p_q % S];
    const uint32_t bytes_v = (uint32_t)ncols * 32u, bytes_c = (uint32_t)(ld - c) * 8u;
    mbar_expect_tx(bar, (uint32_t)ncols * bytes_c + (1 + nv) * bytes_v);
    const double *src = a.P + (size_t)f.src[fam] * a.slab;
    for (int u = 0; u < ncols; ++u)   // rows c.. (the diagonal block and everything below) of column c+u
      tma_load_1d(st + (size_t)u * lds + c, src + (size_t)(c + u) * ld + c, bytes_c, bar);
    tma_load_1d(st + (size_t)KC * lds, a.G4prev + ((size_t)f.anc[fam] * ld + c) * 4, bytes_v, bar);
    for (int q = 0; q < nv; ++q) {
      const int ch = f.child[first + p_b * CB + q];
      tma_load_1d(st + (size_t)KC * lds + 4 * KC * (1 + q), a.H4 + ((size_t)ch * ld + c) * 4, bytes_v, bar);
    }
    ++p_q;
    p_c += KC;
    if (c0 + p_c >= c1) {
      p_c = 0;
      if (++p_b >= nbat) { p_b = 0; claim(); }
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) issue();
  }

  // ---- consumers ---------------------------------------------------------------------
  int q = 0;
  __syncthreads();
  for (int c_claim = 0;; ++c_claim) {
    const int it = s_ring[c_claim & 7];
    if (it >= n_items) break;
    const int fam = it / a.nsplit, sp = it % a.nsplit;
    const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);
    const int an = f.anc[fam], cnt = f.cnt[fam], first = f.first[fam];
    const double *KSa = a.KS4prev + (size_t)an * ld * 4;
    for (int b0 = 0; b0 < cnt; b0 += CB) {
      const int nv = min(CB, cnt - b0);
      int child[CB];
      double *Pd[CB];
#pragma unroll
      for (int s = 0; s < CB; ++s) {
        child[s] = s < nv ? f.child[first + b0 + s] : -1;
        Pd[s] = s < nv ? a.P + (size_t)a.dst_slot[child[s]] * a.slab : nullptr;
      }
      // this lane's sibling in the B operands / outputs: output column n = 4 s + b
      const int sB = g >> 2;                          // B operand: n = g
      const double *HrB = sB < nv ? a.H4 + (size_t)child[sB] * ld * 4 + (g & 3) : nullptr;
      // the column-side slot of PHp: this item writes columns [c0, c1), everything else is zero
#pragma unroll
      for (int s = 0; s < CB; ++s) {
        if (s < nv) {
          double4 *oc = reinterpret_cast<double4 *>(a.PHp + ((size_t)child[s] * (2 * a.nsplit) + a.nsplit + sp) * ld * 4);
          for (int r = tid; r < ld; r += RB_SYMT_THREADS)
            if (r < c0 || r >= c1) oc[r] = make_double4(0.0, 0.0, 0.0, 0.0);
        }
      }
      double acc[MAXQ][2];   // row side: PHrow(8 j + g, n = 2 tg + e), j = wid + 8 qq
#pragma unroll
      for (int qq = 0; qq < MAXQ; ++qq) acc[qq][0] = acc[qq][1] = 0.0;
      for (int c = c0; c < c1; c += KC, ++q) {
        const double *st = stages + (size_t)(q % S) * stage_doubles;
        mbar_wait(&full[q % S], (uint32_t)((q / S) & 1));
        const int ncols = min(KC, c1 - c), j0 = c >> 3;
        const double *thin = st + (size_t)KC * lds;
        // per-stage B operands: G(c+g, tg) for the downdate, H_s(b, c + 4h + tg) for the row side;
        // columns that were not loaded (u >= ncols) and absent siblings are exact zeros
        const double gB = g < ncols ? thin[4 * g + tg] : 0.0;
        double hB[2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
          hB[h] = (4 * h + tg < ncols && sB < nv) ? thin[4 * KC * (1 + sB) + 4 * (4 * h + tg) + (g & 3)] : 0.0;
        double col0 = 0.0, col1 = 0.0;   // column side: PHcol(c + g, n = 2 tg + e)
#pragma unroll
        for (int qq = 0; qq < MAXQ; ++qq) {
          const int j = wid + NW * qq;
          if (j >= j0 && j < nblk) {
            const int r = 8 * j + g;
            const bool diag = j == j0;
            // downdate: T_new(r, u) = T(r, u) - sum_k KS(r, k) G(c+u, k)
            const double ksA = -KSa[(size_t)r * 4 + tg];
            double t0 = st[(size_t)(2 * tg) * lds + r], t1 = st[(size_t)(2 * tg + 1) * lds + r];
            dmma_m8n8k4(t0, t1, ksA, gB);
#pragma unroll
            for (int s = 0; s < CB; ++s) {
              if (s < nv) {
                if (2 * tg < ncols) Pd[s][(size_t)(c + 2 * tg) * ld + r] = t0;
                if (2 * tg + 1 < ncols) Pd[s][(size_t)(c + 2 * tg + 1) * ld + r] = t1;
              }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              // row side, A(row g, k tg) = T(r, u), u = 4h + tg: valid on and below the diagonal
              const int u = 4 * h + tg;
              double aR = st[(size_t)u * lds + r];
              if (u >= ncols || (diag && g < u)) aR = 0.0;
              dmma_m8n8k4(acc[qq][0], acc[qq][1], aR, hB[h]);
              // column side, A(row u = g, k tg) = T(8j + rho, c + g), rho = 4h + tg: strictly below
              const int rho = 4 * h + tg;
              double aC = st[(size_t)g * lds + 8 * j + rho];
              if (g >= ncols || (diag && rho <= g)) aC = 0.0;
              const double bC = HrB != nullptr ? HrB[(size_t)(8 * j + rho) * 4] : 0.0;
              dmma_m8n8k4(col0, col1, aC, bC);
            }
          }
        }
        *reinterpret_cast<double2 *>(&s_colp[q & 1][wid][g * 8 + 2 * tg]) = make_double2(col0, col1);
        __syncthreads();
        if (tid == 0) issue();
        if (tid < 64) {   // (column u, output n): add the warps' tiles in fixed order
          const int u = tid >> 3, n = tid & 7, s = n >> 2;
          if (u < ncols && s < nv) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += s_colp[q & 1][w][tid];
            int ch = child[0];
#pragma unroll
            for (int s2 = 1; s2 < CB; ++s2) if (s == s2) ch = child[s2];
            a.PHp[(((size_t)ch * (2 * a.nsplit) + a.nsplit + sp) * ld + (c + u)) * 4 + (n & 3)] = t;
          }
        }
      }
      // row-side slot: lane holds PHrow(8j+g, n = 2tg, 2tg+1) -> sibling tg>>1, entries 2(tg&1), +1
      {
        const int s = tg >> 1;
        if (s < nv) {
          int ch = child[0];
#pragma unroll
          for (int s2 = 1; s2 < CB; ++s2) if (s == s2) ch = child[s2];
          double *out = a.PHp + ((size_t)ch * (2 * a.nsplit) + sp) * ld * 4;
#pragma unroll
          for (int qq = 0; qq < MAXQ; ++qq) {
            const int j = wid + NW * qq;
            if (j < nblk)
              *reinterpret_cast<double2 *>(out + (size_t)(8 * j + g) * 4 + 2 * (tg & 1)) = make_double2(acc[qq][0], acc[qq][1]);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// kalman_variant 6 (opt-in; NOT YET RUN ON A GPU - written after the round's GPU budget was
// spent, tests are gated behind RBSLAM_TEST_UNVERIFIED=1): the tensor-core symmetric pass of
// k_stream_fam_symt with the data movement rebuilt around what limited it (tuning_r1.md, 7):
//   * a dedicated producer warp (warp 7 of 8); the 8 column copies of a stage are issued by 8 lanes
//     at once; full/empty mbarriers per ring slot instead of a block-wide barrier per stage,
//   * uniform stages of 8 columns x <= 512 rows (a panel taller than 512 rows is two stages),
//     six slots of 33 KB in flight, so the small stages at the narrow end of the triangle are
//     pipelined six deep instead of one,
//   * the consumers are driven by per-stage descriptors the producer leaves in shared memory
//     (item, batch, panel, row segment); the column-side tiles are reduced across the 7
//     consumer warps once per panel behind a named barrier,
//   * stage columns padded to 516 doubles (= 8 mod 32 words): every fragment read of the
//     three access patterns is exactly two shared-memory wavefronts.
// The arithmetic is k_stream_fam_symt's, statement for statement.
// ---------------------------------------------------------------------------
#define RB_SYMP_THREADS 256        // 7 consumer warps + 1 producer warp (8 warps: 255 registers each)
#define RB_SYMP_NW 7
#define RB_SYMP_SLOTS 6
#define RB_SYMP_ROWS 512           // rows per stage
#define RB_SYMP_SLD (RB_SYMP_ROWS + 4)
#define RB_SYMP_SLOT_DOUBLES (8 * RB_SYMP_SLD + 4 * 8 * 3)

struct SympDesc {   // what a ring slot holds; item < 0: no more work
  int item, b0, c, seg;
};

template <int MAXQ>
__global__ void __launch_bounds__(RB_SYMP_THREADS, 1)
k_stream_fam_symp(StreamArgs a, FamLists f) {
  constexpr int KC = 8, CB = 2, NW = RB_SYMP_NW, NS = RB_SYMP_SLOTS, SLD = RB_SYMP_SLD;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) uint64_t full[NS], empty[NS];
  __shared__ SympDesc s_desc[NS];
  __shared__ __align__(16) double s_colp[2][NW][64];
  const int ld = a.ld, M = a.M, nblk = a.ld >> 3;
  double *ring = reinterpret_cast<double *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int n_items = (*f.n_fam) * a.nsplit;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    mbar_fence_init();
  }
  __syncthreads();

  if (wid == NW) {
    // ================= producer warp ====================================================
    int it = 0;
    if (lane == 0) it = atomicAdd(f.work_counter, 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    int b = 0, c_off = 0, seg = 0;   // batch, column offset in the item's range, row segment
    for (int q = 0;; ++q) {
      const int slot = q % NS, round = q / NS;
      if (round > 0) mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));   // consumers are done with it
      double *st = ring + (size_t)slot * RB_SYMP_SLOT_DOUBLES;
      if (it >= n_items) {   // terminal descriptor: every consumer warp sees it in order
        if (lane == 0) {
          s_desc[slot].item = -1;
          mbar_arrive(&full[slot]);
        }
        break;
      }
      const int fam = it / a.nsplit, sp = it % a.nsplit;
      const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);
      const int c = c0 + c_off, ncols = min(KC, c1 - c);
      const int cnt = f.cnt[fam], first = f.first[fam];
      const int nbat = (cnt + CB - 1) / CB;
      const int nv = min(CB, cnt - b * CB);
      const int r_lo = c + RB_SYMP_ROWS * seg, r_hi = min(ld, r_lo + RB_SYMP_ROWS);
      const bool last_seg = r_hi == ld;
      const uint32_t bytes_v = (uint32_t)ncols * 32u, bytes_c = (uint32_t)(r_hi - r_lo) * 8u;
      if (lane == 0) {
        s_desc[slot].item = it; s_desc[slot].b0 = b * CB; s_desc[slot].c = c; s_desc[slot].seg = seg;
        mbar_expect_tx(&full[slot], (uint32_t)ncols * bytes_c + (1 + nv) * bytes_v);
        double *thin = st + (size_t)KC * SLD;
        tma_load_1d(thin, a.G4prev + ((size_t)f.anc[fam] * ld + c) * 4, bytes_v, &full[slot]);
        for (int s = 0; s < nv; ++s) {
          const int ch = f.child[first + b * CB + s];
          tma_load_1d(thin + 4 * KC * (1 + s), a.H4 + ((size_t)ch * ld + c) * 4, bytes_v, &full[slot]);
        }
      }
      __syncwarp();
      if (lane < ncols)   // rows [r_lo, r_hi) of column c + lane
        tma_load_1d(st + (size_t)lane * SLD, a.P + (size_t)f.src[fam] * a.slab + (size_t)(c + lane) * ld + r_lo,
                    bytes_c, &full[slot]);
      // advance: segment -> panel -> batch -> item
      if (!last_seg) {
        ++seg;
      } else {
        seg = 0;
        c_off += KC;
        if (c0 + c_off >= c1) {
          c_off = 0;
          if (++b >= nbat) {
            b = 0;
            if (lane == 0) it = atomicAdd(f.work_counter, 1);
            it = __shfl_sync(0xffffffffu, it, 0);
          }
        }
      }
    }
    return;
  }

  // ================= consumer warps ======================================================
  int cur_item = -1, cur_b0 = -1;
  int fam = 0, sp = 0, c0 = 0, c1 = 0, nv = 0;
  int child[CB] = {-1, -1};
  double *Pd[CB] = {nullptr, nullptr};
  const double *KSa = nullptr, *HrB = nullptr;
  const int sB = g >> 2;
  double acc[MAXQ][2];
#pragma unroll
  for (int qq = 0; qq < MAXQ; ++qq) acc[qq][0] = acc[qq][1] = 0.0;
  double col0 = 0.0, col1 = 0.0;
  int n_panel = 0;   // panels finished so far (parity selects the s_colp buffer)

  auto flush_rows = [&]() {   // row-side slot of the batch that just ended
    const int s = tg >> 1;
    if (s < nv) {
      const int ch = s == 0 ? child[0] : child[1];
      double *out = a.PHp + ((size_t)ch * (2 * a.nsplit) + sp) * ld * 4;
#pragma unroll
      for (int qq = 0; qq < MAXQ; ++qq) {
        const int j = wid + NW * qq;
        if (j < nblk)
          *reinterpret_cast<double2 *>(out + (size_t)(8 * j + g) * 4 + 2 * (tg & 1)) = make_double2(acc[qq][0], acc[qq][1]);
      }
    }
#pragma unroll
    for (int qq = 0; qq < MAXQ; ++qq) acc[qq][0] = acc[qq][1] = 0.0;
  };

  for (int q = 0;; ++q) {
    const int slot = q % NS;
    mbar_wait(&full[slot], (uint32_t)((q / NS) & 1));
    const SympDesc d = s_desc[slot];
    if (d.item < 0) break;
    if (d.item != cur_item || d.b0 != cur_b0) {   // a new batch starts (uniform over the consumers)
      if (cur_item >= 0) flush_rows();
      cur_item = d.item; cur_b0 = d.b0;
      fam = d.item / a.nsplit; sp = d.item % a.nsplit;
      c0 = sp * a.cw; c1 = min(M, c0 + a.cw);
      const int cnt = f.cnt[fam], first = f.first[fam];
      nv = min(CB, cnt - d.b0);
#pragma unroll
      for (int s = 0; s < CB; ++s) {
        child[s] = s < nv ? f.child[first + d.b0 + s] : -1;
        Pd[s] = s < nv ? a.P + (size_t)a.dst_slot[child[s]] * a.slab : nullptr;
      }
      KSa = a.KS4prev + (size_t)f.anc[fam] * ld * 4;
      HrB = sB < nv ? a.H4 + (size_t)(sB == 0 ? child[0] : child[1]) * ld * 4 + (g & 3) : nullptr;
#pragma unroll
      for (int s = 0; s < CB; ++s) {   // column-side slot: zero outside this item's columns
        if (s < nv) {
          double4 *oc = reinterpret_cast<double4 *>(a.PHp + ((size_t)child[s] * (2 * a.nsplit) + a.nsplit + sp) * ld * 4);
          for (int r = tid; r < ld; r += NW * 32)
            if (r < c0 || r >= c1) oc[r] = make_double4(0.0, 0.0, 0.0, 0.0);
        }
      }
    }
    const double *st = ring + (size_t)slot * RB_SYMP_SLOT_DOUBLES;
    const double *thin = st + (size_t)KC * SLD;
    const int c = d.c, ncols = min(KC, c1 - c), j0 = c >> 3;
    const int r_lo = c + RB_SYMP_ROWS * d.seg, r_hi = min(ld, r_lo + RB_SYMP_ROWS);
    const int jlo = r_lo >> 3, jhi = r_hi >> 3;
    const bool last_seg = r_hi == ld;
    const double gB = g < ncols ? thin[4 * g + tg] : 0.0;
    double hB[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
      hB[h] = (4 * h + tg < ncols && sB < nv) ? thin[4 * KC * (1 + sB) + 4 * (4 * h + tg) + (g & 3)] : 0.0;
#pragma unroll
    for (int qq = 0; qq < MAXQ; ++qq) {
      const int j = wid + NW * qq;
      if (j >= jlo && j < jhi) {
        const int r = 8 * j + g, rs = r - r_lo;       // global row, row inside the stage
        const bool diag = j == j0;
        const double ksA = -KSa[(size_t)r * 4 + tg];
        double t0 = st[(size_t)(2 * tg) * SLD + rs], t1 = st[(size_t)(2 * tg + 1) * SLD + rs];
        dmma_m8n8k4(t0, t1, ksA, gB);
#pragma unroll
        for (int s = 0; s < CB; ++s) {
          if (s < nv) {
            if (2 * tg < ncols) Pd[s][(size_t)(c + 2 * tg) * ld + r] = t0;
            if (2 * tg + 1 < ncols) Pd[s][(size_t)(c + 2 * tg + 1) * ld + r] = t1;
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int u = 4 * h + tg;
          double aR = st[(size_t)u * SLD + rs];
          if (u >= ncols || (diag && g < u)) aR = 0.0;
          dmma_m8n8k4(acc[qq][0], acc[qq][1], aR, hB[h]);
          const int rho = 4 * h + tg;
          double aC = st[(size_t)g * SLD + (8 * j - r_lo) + rho];
          if (g >= ncols || (diag && rho <= g)) aC = 0.0;
          const double bC = HrB != nullptr ? HrB[(size_t)(8 * j + rho) * 4] : 0.0;
          dmma_m8n8k4(col0, col1, aC, bC);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);   // this warp has read everything it needs from the slot
    if (last_seg) {   // the panel is complete: add the consumer warps' column-side tiles in fixed order
      const int buf = n_panel & 1;
      *reinterpret_cast<double2 *>(&s_colp[buf][wid][g * 8 + 2 * tg]) = make_double2(col0, col1);
      col0 = col1 = 0.0;
      named_barrier_sync(1, NW * 32);
      if (tid < 64) {
        const int u = tid >> 3, n = tid & 7, s = n >> 2;
        if (u < ncols && s < nv) {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w) t += s_colp[buf][w][tid];
          const int ch = s == 0 ? child[0] : child[1];
          a.PHp[(((size_t)ch * (2 * a.nsplit) + a.nsplit + sp) * ld + (c + u)) * 4 + (n & 3)] = t;
        }
      }
      ++n_panel;
    }
  }
  if (cur_item >= 0) flush_rows();
}

}  // namespace rb
