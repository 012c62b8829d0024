// K3 streaming path for large M (C1 M=515, C4 M=1027): ONE pass over each
// covariance slab per particle-step.
//
// The reference updates P_i in three sweeps per step (SS = H P H', K = P (..),
// P -= K SS K'; src/particleFilter.m:139-150,184-198).  SS needs all of P H'
// before any element of P can be downdated, so a literal implementation streams
// every 8.4 MB slab twice.  Here the downdate of step t is DEFERRED and applied
// while the slab is streamed for step t+1:
//
//   pass(t+1):  P_t   = P_{t-1} - KS_t G_t'        (pending rank-d downdate of step t)
//               PH    = P_t H_{t+1}'               (accumulated from the updated tile)
//   innov(t+1): SS, chol, logw, G_{t+1} = PH SS^-1, KS_{t+1} = G_{t+1} SS, xl += G e
//
// Every element sees exactly the reference's arithmetic, one step later.  Per
// particle-step the slab is read once and written once: 16*M^2 B, the algorithmic
// minimum.  The same pass performs the resampling gather: it reads the ANCESTOR's
// slab (src) with the ancestor's pending (G, KS) and writes the particle's slab
// (dst); offspring that copy run before offspring in place (see k_plan_slots).
//
// Data movement: the slab is column-major, so KC consecutive columns are one
// contiguous block of KC*ld*8 bytes; a single elected thread streams such blocks
// into a ring of shared-memory stages with TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx), S-1 stages in flight; all threads consume a stage from
// shared memory (LDS.128), and store the updated tile straight to HBM (STG.128).
#pragma once
#include "common.cuh"

namespace rb {

// ---- PTX helpers (mbarrier + TMA bulk copy) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among a subset of the CTA's warps (id 1..15, n_threads a multiple of 32)
__device__ __forceinline__ void named_barrier_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA, non-tensor form); bytes % 16 == 0, 16-B aligned
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                                 uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// Where the SOURCE of a family lives.  Single GPU (nloc == 0) or key < nloc: slab `key` of this GPU.
// Sharded filter, key >= nloc: migrant number key - nloc of this step -- its ancestor's state (slab,
// pending pair, mean) stays where it is, on rank fetch[4f+1] in slab fetch[4f+2], and is read in place
// through the peer mapping (NVLink) by the pass that needs it: no staging copy, no extra trip through
// local HBM.  The tables hold one base pointer per rank ([0] = this GPU, [1 + r] = rank r).
struct SrcTab {
  int nloc = 0;
  const int *fetch = nullptr;
  const double *const *P = nullptr, *const *G4 = nullptr, *const *KS4 = nullptr, *const *xl = nullptr;
};
__device__ __forceinline__ const double *src_base(const SrcTab &t, const double *const *tab, const double *local,
                                                  int key, size_t stride) {
  if (t.nloc == 0 || key < t.nloc) return local + (size_t)key * stride;
  const int f = key - t.nloc;
  return tab[1 + t.fetch[4 * f + 1]] + (size_t)t.fetch[4 * f + 2] * stride;
}

struct StreamArgs {
  SrcTab st;
  int hints;               // bit0: L2 evict_first on the slab loads, bit1: streaming (.cs) stores
  int gk_by_particle;      // 1: G4prev/KS4prev are indexed by the particle, not its ancestor
  int M, ld, cw, nsplit;
  size_t slab;
  double *P;
  const int *src_slot;     // [N] slab of the ancestor
  const int *dst_slot;     // [N] slab of the particle
  const int *anc;          // [N] ancestor's logical index (nullptr: identity)
  const double *G4prev;    // [N][ld][4] pending gain of the ancestor  (G(c,b))
  const double *KS4prev;   // [N][ld][4] pending K*SS of the ancestor  (KS(r,b))
  const double *H4;        // [N][ld][4] measurement Jacobian of the particle (H(a,c))
  double *PHp;             // [N][nsplit][ld][4] partial P H'
};

#define RB_STREAM_THREADS 192

// D: rank of the pending update; DA >= D: columns of H4 to multiply with (DA = D+1 carries
// P*ivec for the information form)
template <int D, int DA, int R2, int KC, int S>
__global__ void __launch_bounds__(RB_STREAM_THREADS, 1)
k_stream_pass(StreamArgs a, const int *__restrict__ list, const int *__restrict__ count,
              const int *__restrict__ list_off = nullptr) {
  if (list_off) list += *list_off;   // lists of later work groups start where the previous group ends
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) uint64_t full[S];
  const int ld = a.ld, M = a.M;
  const int npairs = ld >> 1;
  const size_t stage_doubles = (size_t)KC * ld + 8 * KC;
  double *stages = reinterpret_cast<double *>(smraw);
  const int tid = threadIdx.x;
  const int n_items = (*count) * a.nsplit;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // ---- producer (thread 0): next chunk to issue --------------------------------
  int p_it = blockIdx.x, p_c = 0, p_q = 0;   // item, column offset inside the item's chunk, seq no
  const uint64_t pol = l2_policy_evict_first();
  const bool hint_ld = a.hints & 1, hint_st = a.hints & 2;
  auto issue = [&]() {
    if (p_it >= n_items) return;
    const int i = list[p_it / a.nsplit], sp = p_it % a.nsplit;
    const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);
    const int c = c0 + p_c;
    const int ncols = min(KC, c1 - c);
    const int an = a.gk_by_particle ? i : (a.anc ? a.anc[i] : i);
    double *st = stages + (size_t)(p_q % S) * stage_doubles;
    uint64_t *bar = &full[p_q % S];
    const uint32_t bytes_p = (uint32_t)ncols * ld * 8u, bytes_v = (uint32_t)ncols * 32u;
    mbar_expect_tx(bar, bytes_p + 2 * bytes_v);
    if (hint_ld) tma_load_1d_hint(st, a.P + (size_t)a.src_slot[i] * a.slab + (size_t)c * ld, bytes_p, bar, pol);
    else tma_load_1d(st, a.P + (size_t)a.src_slot[i] * a.slab + (size_t)c * ld, bytes_p, bar);
    tma_load_1d(st + (size_t)KC * ld, a.G4prev + ((size_t)an * ld + c) * 4, bytes_v, bar);
    tma_load_1d(st + (size_t)KC * ld + 4 * KC, a.H4 + ((size_t)i * ld + c) * 4, bytes_v, bar);
    ++p_q;
    p_c += KC;
    if (c0 + p_c >= c1) { p_c = 0; p_it += gridDim.x; }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) issue();
  }

  // ---- consumers (all threads) --------------------------------------------------
  int q = 0;
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int i = list[it / a.nsplit], sp = it % a.nsplit;
    const int c0 = sp * a.cw, c1 = min(M, c0 + a.cw);
    const int an = a.gk_by_particle ? i : (a.anc ? a.anc[i] : i);
    double *Pd = a.P + (size_t)a.dst_slot[i] * a.slab;
    double2 ks[R2][D], acc[R2][DA];
#pragma unroll
    for (int k = 0; k < R2; ++k) {
      const int rp = tid + k * RB_STREAM_THREADS;
      if (rp < npairs) {
        const double4 *kp = reinterpret_cast<const double4 *>(a.KS4prev + ((size_t)an * ld + 2 * rp) * 4);
        const double4 k0 = kp[0], k1 = kp[1];
        const double r0[4] = {k0.x, k0.y, k0.z, k0.w}, r1[4] = {k1.x, k1.y, k1.z, k1.w};
#pragma unroll
        for (int b = 0; b < D; ++b) ks[k][b] = make_double2(r0[b], r1[b]);
      } else {
#pragma unroll
        for (int b = 0; b < D; ++b) ks[k][b] = make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int b = 0; b < DA; ++b) acc[k][b] = make_double2(0.0, 0.0);
    }
    for (int c = c0; c < c1; c += KC, ++q) {
      const double *st = stages + (size_t)(q % S) * stage_doubles;
      mbar_wait(&full[q % S], (uint32_t)((q / S) & 1));
      const int ncols = min(KC, c1 - c);
#pragma unroll
      for (int u = 0; u < KC; ++u) {
        if (u < ncols) {
          const double4 g4 = *reinterpret_cast<const double4 *>(st + (size_t)KC * ld + 4 * u);
          const double4 h4 = *reinterpret_cast<const double4 *>(st + (size_t)KC * ld + 4 * KC + 4 * u);
          const double g[4] = {g4.x, g4.y, g4.z, g4.w}, h[4] = {h4.x, h4.y, h4.z, h4.w};
          const double2 *col = reinterpret_cast<const double2 *>(st + (size_t)u * ld);
          double2 *dcol = reinterpret_cast<double2 *>(Pd + (size_t)(c + u) * ld);
#pragma unroll
          for (int k = 0; k < R2; ++k) {
            const int rp = tid + k * RB_STREAM_THREADS;
            if (rp < npairs) {
              double2 v = col[rp];
#pragma unroll
              for (int b = 0; b < D; ++b) {   // pending downdate: P(r,c) -= KS(r,b) G(c,b)
                v.x = fma(-ks[k][b].x, g[b], v.x);
                v.y = fma(-ks[k][b].y, g[b], v.y);
              }
#pragma unroll
              for (int b = 0; b < DA; ++b) {  // PH(r,b) += P(r,c) H(b,c)
                acc[k][b].x = fma(v.x, h[b], acc[k][b].x);
                acc[k][b].y = fma(v.y, h[b], acc[k][b].y);
              }
              if (hint_st) __stcs(dcol + rp, v); else dcol[rp] = v;
            }
          }
        }
      }
      __syncthreads();            // every thread is done reading this stage
      if (tid == 0) issue();      // refill it with the chunk S positions ahead
    }
    double *out = a.PHp + ((size_t)i * a.nsplit + sp) * ld * 4;
#pragma unroll
    for (int k = 0; k < R2; ++k) {
      const int rp = tid + k * RB_STREAM_THREADS;
      if (rp < npairs) {
        double o0[4] = {0, 0, 0, 0}, o1[4] = {0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < DA; ++b) { o0[b] = acc[k][b].x; o1[b] = acc[k][b].y; }
        double4 *op = reinterpret_cast<double4 *>(out + (size_t)2 * rp * 4);
        op[0] = make_double4(o0[0], o0[1], o0[2], o0[3]);
        op[1] = make_double4(o1[0], o1[1], o1[2], o1[3]);
      }
    }
  }
}

// innovation / gain for the streaming path (layouts [row][4]); one CTA per particle
struct Innov4Args {
  SrcTab st;
  int N, M, ld, nsplit;
  const double *PHp;       // [N][nsplit][ld][4]
  const double *H4;        // [N][ld][4]
  const double *xl_old;    // [M x N]
  const int *anc;          // ancestors or nullptr
  double *xl_new;
  double *G4new, *KS4new;  // [N][ld][4]
  const double *y_t, *R;
  double jitter;
  double *logw;
  DevStatus *status;
  int t;
  // tensor-core symmetric pass (k_stream_fam_symt): the partial sums are P_old H' of the slab
  // BEFORE its pending downdate; the innovation kernel completes them,
  //   P_new H' = P_old H' - KS_anc (G_anc' H'),   W = G_anc' H' is d x d per particle
  const double *G4prev = nullptr, *KS4prev = nullptr;   // [N][ld][4] pending pair, indexed by ancestor
};

template <int D>
__global__ void __launch_bounds__(128)
k_innov4(Innov4Args a) {
  extern __shared__ double sm[];
  const int M = a.M, ld = a.ld;
  double *sPH = sm;  // [ld][4]
  __shared__ double s_red[4][D * D + D];
  __shared__ double s_L[D * D], s_SS[D * D], s_e[D];
  const int i = blockIdx.x;
  const double *Hi = a.H4 + (size_t)i * ld * 4;
  const double *xls = src_base(a.st, a.st.xl, a.xl_old, a.anc ? a.anc[i] : i, M);
  __shared__ double s_W[D * D];   // W(k,b) = sum_c G_anc(c,k) H(b,c)
  const double *KSa = nullptr;
  if (a.G4prev != nullptr) {
    const int an = a.anc ? a.anc[i] : i;
    const double *Ga = src_base(a.st, a.st.G4, a.G4prev, an, (size_t)ld * 4);
    KSa = src_base(a.st, a.st.KS4, a.KS4prev, an, (size_t)ld * 4);
    double wp[D * D];
#pragma unroll
    for (int q = 0; q < D * D; ++q) wp[q] = 0.0;
    for (int c = threadIdx.x; c < M; c += blockDim.x) {
      const double4 gv = *reinterpret_cast<const double4 *>(Ga + (size_t)c * 4);
      const double4 hv = *reinterpret_cast<const double4 *>(Hi + (size_t)c * 4);
      const double g[4] = {gv.x, gv.y, gv.z, gv.w}, h[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int k = 0; k < D; ++k)
#pragma unroll
        for (int bb = 0; bb < D; ++bb) wp[k + bb * D] = fma(g[k], h[bb], wp[k + bb * D]);
    }
    __shared__ double s_wred[4][D * D];
#pragma unroll
    for (int q = 0; q < D * D; ++q) {
      const double v = warp_sum(wp[q]);
      if ((threadIdx.x & 31) == 0) s_wred[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < D * D)
      s_W[threadIdx.x] = (s_wred[0][threadIdx.x] + s_wred[1][threadIdx.x]) +
                         (s_wred[2][threadIdx.x] + s_wred[3][threadIdx.x]);
    __syncthreads();
  }
  double part[D * D + D];
#pragma unroll
  for (int q = 0; q < D * D + D; ++q) part[q] = 0.0;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) {
    double ph[4] = {0, 0, 0, 0};
    for (int sp = 0; sp < a.nsplit; ++sp) {   // fixed order: deterministic
      const double4 v = *reinterpret_cast<const double4 *>(a.PHp + (((size_t)i * a.nsplit + sp) * ld + r) * 4);
      ph[0] += v.x; ph[1] += v.y; ph[2] += v.z; ph[3] += v.w;
    }
    if (KSa != nullptr) {   // complete P_new H' (see Innov4Args)
      const double4 kv = *reinterpret_cast<const double4 *>(KSa + (size_t)r * 4);
      const double ks[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int bb = 0; bb < D; ++bb)
#pragma unroll
        for (int k = 0; k < D; ++k) ph[bb] = fma(-ks[k], s_W[k + bb * D], ph[bb]);
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
    if (flag) {   // src/particleFilter.m:146-148
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
#pragma unroll
      for (int k = 0; k < D; ++k) if (k < r) s -= Lc[r + k * D] * v[k];
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
  double *Gi = a.G4new + (size_t)i * ld * 4;
  double *KSi = a.KS4new + (size_t)i * ld * 4;
  for (int r = threadIdx.x; r < ld; r += blockDim.x) {
    double g[4] = {0, 0, 0, 0}, ksv[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < D; ++q) {   // forward: cS z = ph
      double s = sPH[(size_t)r * 4 + q];
#pragma unroll
      for (int k = 0; k < D; ++k) if (k < q) s -= s_L[q + k * D] * g[k];
      g[q] = s / s_L[q + q * D];
    }
#pragma unroll
    for (int q = D - 1; q >= 0; --q) {  // backward: cS' g = z
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
    if (r < M) a.xl_new[(size_t)i * M + r] = xls[r] + ge;
  }
}

// apply (and thereby clear) the pending downdate of every particle in place:
// P(:,:,i) -= KS_i G_i'.  Used before the state is read out as a whole.
template <int D>
__global__ void __launch_bounds__(256)
k_apply_pending(double *__restrict__ P, size_t slab, int ld, int M, const int *__restrict__ slot,
                const double *__restrict__ G4, const double *__restrict__ KS4) {
  const int i = blockIdx.y;
  double *Pi = P + (size_t)slot[i] * slab;
  const double *Gi = G4 + (size_t)i * ld * 4, *KSi = KS4 + (size_t)i * ld * 4;
  for (int c = blockIdx.x; c < M; c += gridDim.x) {
    double g[D];
#pragma unroll
    for (int b = 0; b < D; ++b) g[b] = Gi[(size_t)c * 4 + b];
    for (int r = threadIdx.x; r < M; r += blockDim.x) {
      double v = Pi[r + (size_t)c * ld];
#pragma unroll
      for (int b = 0; b < D; ++b) v = fma(-KSi[(size_t)r * 4 + b], g[b], v);
      Pi[r + (size_t)c * ld] = v;
    }
  }
}

}  // namespace rb
