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
    tma_load_1d(st, src_base(a.st, a.st.P, a.P, f.src[fam], a.slab) + (size_t)c * ld, bytes_p, bar);
    tma_load_1d(st + (size_t)KC * ld, src_base(a.st, a.st.G4, a.G4prev, f.anc[fam], (size_t)ld * 4) + (size_t)c * 4, bytes_v, bar);
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
    const double *KSan = src_base(a.st, a.st.KS4, a.KS4prev, an, (size_t)ld * 4);
    double2 ks[R2][D];
#pragma unroll
    for (int k = 0; k < R2; ++k) {
      const int rp = tid + k * RB_STREAM_THREADS;
      if (rp < npairs) {
        const double4 *kp = reinterpret_cast<const double4 *>(KSan + (size_t)2 * rp * 4);
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

// fp64 tensor-core tile product D(8x8) += A(8x4) B(4x8): lane (g = lane>>2, tg = lane&3) holds
// A(g, tg), B(tg, g) and D(g, 2tg), D(g, 2tg+1)   (SASS: DMMA)
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

}  // namespace rb
