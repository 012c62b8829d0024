// Packed symmetric covariance slabs and their streaming Kalman pass (kalman_variant 7).
//
// P_i and the downdate K SS K' (src/particleFilter.m:198) are symmetric, so half of a full
// [ld x M] slab is redundant.  Layout "PT" (packed tiles) stores the lower block triangle only,
// as 8x8 tiles:
//
//   nb = ld / 8 row blocks; panel p = the tiles (j, p), j = p .. nb-1, of block column p;
//   panels follow each other, tiles inside a panel are ordered by j:
//       tile index  tix(j, p) = p nb - p (p - 1) / 2 + (j - p),       64 doubles per tile
//   -> a slab is ONE linear stream of nb (nb + 1) / 2 tiles (4.29 MB at M = 1027 instead of
//   8.52 MB).  The diagonal tile (p, p) keeps all 64 entries (both triangles evolve as in the
//   reference, which never symmetrises P); for j > p the tile holds P(8j.., 8p..) and stands
//   for its mirror image as well.
//   Inside a tile element (r, c) sits at  8 sigma(r) + (c ^ (r & 4)),  sigma = 0 1 3 2 4 5 7 6:
//   the two fragment reads the tensor-core pass needs -- (row g, columns 2tg, 2tg+1) as one
//   16-byte access and (rows 2tg, 2tg+1, column g) as two 8-byte accesses -- are both free of
//   shared-memory bank conflicts, and the first one is also the layout of the store: a warp
//   writes an updated tile as one contiguous 512-byte block.
//
// The pass (k_stream_fam_pt) is the sibling-fused, deferred-downdate pass of family_kernels.cuh
// re-done on this layout:
//   * data movement: a dedicated producer warp streams the slab as uniform stages of TS tiles
//     (one exactly-sized cp.async.bulk each, whatever panel boundaries it crosses) into an
//     NS-deep ring; full/empty mbarriers per slot, no block-wide barrier in the steady state;
//   * arithmetic: mma.sync.m8n8k4.f64 per tile T (n = (sibling, measurement row) = 2 x 4):
//       row side   PHrow(8j.., n) += T H_n(8p..)      2 DMMA, A operand = the loaded fragment
//       col side   PHcol(8p.., n) += T' H_n(8j..)     2 DMMA, A operand = the transposed read
//       downdate   T -= KS(8j.., :) G(8p.., :)'       1 DMMA, accumulator = the fragment
//     and one 16-byte store per sibling.  The products use the tile BEFORE its pending
//     downdate; k_innov4 completes them (Innov4Args::G4prev), exactly as kalman_variant 5 did;
//   * warp w owns the row blocks j = w (mod NW): its row-side accumulators and its KS fragments
//     stay in registers for the whole pass; the siblings' H live in shared memory in fragment
//     order, G of the ancestor is prefetched one panel ahead; the column-side tiles of the
//     consumer warps go through a small ring of buffers to a reducer warp that adds them in
//     fixed order (deterministic) - the consumers never wait for each other inside a batch.
// HBM traffic per particle-step: half of k_stream_fam's (4.3 MB read per pass over a family's slab,
// 4.3 MB written per particle at M = 1027).
#pragma once
#include "family_kernels.cuh"

namespace rb {

enum { RB_LAYOUT_FULL = 0, RB_LAYOUT_SYM = 1, RB_LAYOUT_PT = 2 };

__host__ __device__ __forceinline__ size_t pt_panel_off(int nb, int p) {   // first tile of panel p
  return (size_t)p * nb - (size_t)p * (p - 1) / 2;
}
__host__ __device__ __forceinline__ int pt_pos(int r, int c) {             // inside a tile
  return 8 * (r ^ ((r >> 1) & 1)) + (c ^ (r & 4));
}
__host__ __device__ __forceinline__ size_t pt_slab_doubles(int ld) {
  const int nb = ld >> 3;
  return (size_t)nb * (nb + 1) / 2 * 64;
}
// offset of the stored copy of element (r, c) in a slab of the given layout
__host__ __device__ __forceinline__ size_t slab_elem(int layout, int ld, int r, int c) {
  if (layout == RB_LAYOUT_FULL) return (size_t)r + (size_t)c * ld;
  if (layout == RB_LAYOUT_SYM) return (size_t)(r > c ? r : c) + (size_t)(r > c ? c : r) * ld;
  int bj = r >> 3, bp = c >> 3, rr = r & 7, cc = c & 7;
  if (bj < bp) { int t = bj; bj = bp; bp = t; t = rr; rr = cc; cc = t; }
  return (pt_panel_off(ld >> 3, bp) + (size_t)(bj - bp)) * 64 + pt_pos(rr, cc);
}
// (row, column) of the pending-downdate factors that belong to the stored copy of (r, c)
__host__ __device__ __forceinline__ void slab_elem_rc(int layout, int r, int c, int &rs, int &cs) {
  rs = r; cs = c;
  if (layout == RB_LAYOUT_SYM && r < c) { rs = c; cs = r; }
  if (layout == RB_LAYOUT_PT && (r >> 3) < (c >> 3)) { rs = c; cs = r; }
}

// ---------------------------------------------------------------------------
// layout-aware helpers (initialisation, read-out)
// ---------------------------------------------------------------------------
// P[slot] = P0 for every slab; grid (tile chunk, particle chunk)
__global__ void k_init_slabs_pt(double *__restrict__ P, size_t slab, int ld, int M, int N,
                                const double *__restrict__ P0) {
  const int nb = ld >> 3;
  const int ntile = nb * (nb + 1) / 2;
  for (int i = blockIdx.y; i < N; i += gridDim.y) {
    double *dst = P + (size_t)i * slab;
    for (int tl = blockIdx.x; tl < ntile; tl += gridDim.x) {
      // invert tl -> (j, p): panels are short, a linear walk from an estimate is cheap
      int p = (int)(((2.0 * nb + 1.0) - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * tl)) * 0.5);
      while (p > 0 && pt_panel_off(nb, p) > (size_t)tl) --p;
      while (pt_panel_off(nb, p + 1) <= (size_t)tl) ++p;
      const int j = p + (tl - (int)pt_panel_off(nb, p));
      for (int e = threadIdx.x; e < 64; e += blockDim.x) {
        const int r = 8 * j + (e >> 3), c = 8 * p + (e & 7);
        dst[(size_t)tl * 64 + pt_pos(e >> 3, e & 7)] = (r < M && c < M) ? P0[r + (size_t)c * M] : 0.0;
      }
    }
  }
}

// dense [M x M x cnt] (logical order) -> packed slabs (rbslam_op_kalman_update)
__global__ void k_unpack_slabs_pt(int M, int ld, size_t slab, double *__restrict__ P, const double *__restrict__ in) {
  const int nb = ld >> 3, p = blockIdx.x, i = blockIdx.y;
  double *dst = P + (size_t)i * slab + pt_panel_off(nb, p) * 64;
  const double *src = in + (size_t)i * M * M;
  for (int idx = threadIdx.x; idx < (nb - p) * 64; idx += blockDim.x) {
    const int tl = idx >> 6, e = idx & 63;
    const int r = 8 * (p + tl) + (e >> 3), c = 8 * p + (e & 7);
    dst[(size_t)tl * 64 + pt_pos(e >> 3, e & 7)] = (r < M && c < M) ? src[r + (size_t)c * M] : 0.0;
  }
}

// apply (and thereby clear) the pending downdate of every stored element: P -= KS G'
__global__ void __launch_bounds__(256)
k_apply_pending_pt(double *__restrict__ P, size_t slab, int ld, const int *__restrict__ slot,
                   const double *__restrict__ G4, const double *__restrict__ KS4) {
  const int nb = ld >> 3, i = blockIdx.y;
  double *Pi = P + (size_t)slot[i] * slab;
  const double *Gi = G4 + (size_t)i * ld * 4, *KSi = KS4 + (size_t)i * ld * 4;
  for (int p = blockIdx.x; p < nb; p += gridDim.x) {
    double *pan = Pi + pt_panel_off(nb, p) * 64;
    for (int idx = threadIdx.x; idx < (nb - p) * 64; idx += blockDim.x) {
      const int tl = idx >> 6, e = idx & 63;
      const int r = 8 * (p + tl) + (e >> 3), c = 8 * p + (e & 7);
      double v = pan[(size_t)tl * 64 + pt_pos(e >> 3, e & 7)];
#pragma unroll
      for (int b = 0; b < 4; ++b) v = fma(-KSi[(size_t)r * 4 + b], Gi[(size_t)c * 4 + b], v);
      pan[(size_t)tl * 64 + pt_pos(e >> 3, e & 7)] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// the streaming pass
// ---------------------------------------------------------------------------
#define RB_PT_CB 2                          // siblings per pass: n = 2 x 4 output columns
#define RB_PT_MAXSPLIT 8
#define RB_PT_NBUF 4                        // column-side partial tiles in flight (panels)

struct PtArgs {
  SrcTab st;
  int M, ld, nb, nsplit;
  int rev;                        // families are walked from the last to the first (sharded filter: migrants, whose
                                  // slabs come over NVLink, have the highest source keys and start first)
  int ts, ns;                     // tiles per stage, ring slots
  int psplit[RB_PT_MAXSPLIT + 1]; // item sp streams the panels [psplit[sp], psplit[sp+1])
  size_t slab;                    // doubles per slab
  double *P;
  const int *dst_slot;            // [N] slab of the particle
  const double *G4prev;           // [N][ld][4] pending gain of the ancestor  (G(c,b))
  const double *KS4prev;          // [N][ld][4] pending K*SS of the ancestor  (KS(r,b))
  const double *H4;               // [N][ld][4] measurement Jacobian of the particle (H(a,c))
  double *PHp;                    // [N][nsplit + 1][ld][4]: row-side partials per item, then the column side
};

struct PtDesc {   // what a ring slot holds; item < 0: no more work
  int item, b0, t0, nt, p, j;
};
struct PtBatch {     // state of the batch the consumers work on (uniform over the CTA)
  int item, b0, sp, ch0, ch1;
  double *Pd0, *Pd1;
  const double *Ga;    // pending gain of the family's ancestor, [ld][4]: G(8p + g, tg) at Ga[32 p + lane]
};
struct PtColDesc {   // what a column-side buffer holds: panel p of the batch (ch0, ch1); p < 0: no more work
  int p, ch0, ch1;
};

static inline size_t pt_smem_bytes(int ld, int ts, int ns, int nw) {
  const int nb = ld >> 3;
  return (size_t)ns * ts * 512 + (size_t)nb * 32 * 16 + (size_t)RB_PT_NBUF * nw * 64 * 8 + (size_t)nw * 2 * 32 * 8;
}

// Warp roles: NW consumer warps and one service warp (lane 0 = producer, lanes 1..31 = reducer).
// NW + 1 is a multiple of 4 (registers are allocated to warps in groups of four: 16 warps x 128).
template <int NW, int MAXQ>
__global__ void __launch_bounds__(32 * (NW + 1), 1)
k_stream_fam_pt(PtArgs a, FamLists f) {
  constexpr int CB = RB_PT_CB, NBUF = RB_PT_NBUF;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) uint64_t full[8], empty[8], colfull[NBUF], colfree[NBUF];
  __shared__ PtDesc s_desc[8];
  __shared__ PtColDesc s_cdesc[NBUF];
  __shared__ PtBatch s_bat;
  __shared__ int s_psplit[RB_PT_MAXSPLIT + 1];   // (a parameter array indexed at run time would live in local memory)
  const int ld = a.ld, nb = a.nb, TS = a.ts, NS = a.ns;
  double *ring = reinterpret_cast<double *>(smraw);                             // [NS][TS][64]
  double2 *s_bC = reinterpret_cast<double2 *>(ring + (size_t)NS * TS * 64);     // [nb][32] fragment order
  double *s_colp = reinterpret_cast<double *>(s_bC + (size_t)nb * 32);          // [NBUF][NW][64]
  double *s_gpre = s_colp + (size_t)NBUF * NW * 64;                             // [NW][2][32] G fragments, prefetched
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int n_fam = *f.n_fam;
  const int n_items = n_fam * a.nsplit;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    for (int s = 0; s < NBUF; ++s) { mbar_init(&colfull[s], NW); mbar_init(&colfree[s], 1); }
    mbar_fence_init();
#pragma unroll
    for (int s = 0; s <= RB_PT_MAXSPLIT; ++s) s_psplit[s] = a.psplit[s];
    s_bat.item = -1; s_bat.b0 = -1; s_bat.ch0 = s_bat.ch1 = -1;
  }
  __syncthreads();

  if (wid == NW && lane == 0) {
    // ================= producer: one lane walks item -> batch -> stage ===================
    int it = atomicAdd(f.work_counter, 1);
    int b = 0;
    bool fresh = true;         // (it, b) just changed: (re)start the tile walk
    int t_cur = 0, t_end = 0, p = 0, j = 0;
    int slot = 0, round = 0;
    for (;;) {
      if (round > 0) mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));   // consumers are done with it
      if (it >= n_items) {     // terminal descriptor: every consumer warp sees it in order
        s_desc[slot].item = -1;
        mbar_arrive(&full[slot]);
        break;
      }
      const int fam = a.rev ? n_fam - 1 - it / a.nsplit : it / a.nsplit, sp = it % a.nsplit;
      if (fresh) {
        p = s_psplit[sp]; j = p;
        t_cur = (int)pt_panel_off(nb, p); t_end = (int)pt_panel_off(nb, s_psplit[sp + 1]);
        fresh = false;
      }
      const int nt = min(TS, t_end - t_cur);
      PtDesc d;
      d.item = it; d.b0 = b * CB; d.t0 = t_cur; d.nt = nt; d.p = p; d.j = j;
      s_desc[slot] = d;
      mbar_expect_tx(&full[slot], (uint32_t)nt * 512u);
      tma_load_1d(ring + (size_t)slot * TS * 64, src_base(a.st, a.st.P, a.P, f.src[fam], a.slab) + (size_t)t_cur * 64,
                  (uint32_t)nt * 512u, &full[slot]);
      // advance (p, j) by nt tiles
      t_cur += nt;
      int left = nt;
      while (left > 0) {
        const int seg = min(left, nb - j);
        left -= seg; j += seg;
        if (j == nb) { ++p; j = p; }
      }
      if (t_cur >= t_end) {
        fresh = true;
        const int nbat = (f.cnt[fam] + CB - 1) / CB;
        if (++b >= nbat) { b = 0; it = atomicAdd(f.work_counter, 1); }
      }
      if (++slot == NS) { slot = 0; ++round; }
    }
    return;
  }

  if (wid == NW) {
    // ================= reducer (lanes 1..31 of the producer's warp; the two roles only ever wait
    // on mbarriers, independent thread scheduling interleaves them): adds the consumers'
    // column-side tiles, panel by panel, in fixed warp order (deterministic), and stores
    // PHcol(8p.., n) of both siblings ========================================================
    constexpr unsigned RMASK = 0xfffffffeu;
    const int col_slot = a.nsplit;
    for (int n = 0;; ++n) {
      const int buf = n % NBUF;
      mbar_wait(&colfull[buf], (uint32_t)((n / NBUF) & 1));
      const PtColDesc cd = s_cdesc[buf];
      if (cd.p < 0) break;
      for (int e = lane - 1; e < 32; e += 31) {   // fragment slot e = 4 g' + tg' holds PHcol(8p + g', n = 2tg', 2tg' + 1)
        const double2 *part = reinterpret_cast<const double2 *>(s_colp + (size_t)buf * NW * 64) + e;
        double2 sum = make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < NW; ++w) { const double2 v = part[w * 32]; sum.x += v.x; sum.y += v.y; }
        const int ge = e >> 2, te = e & 3;
        const int ch = (te >> 1) == 0 ? cd.ch0 : cd.ch1;   // sibling te >> 1, entries 2 (te & 1), +1
        if (ch >= 0)
          *reinterpret_cast<double2 *>(a.PHp + (((size_t)ch * (a.nsplit + 1) + col_slot) * ld + (8 * cd.p + ge)) * 4 + 2 * (te & 1)) = sum;
      }
      __syncwarp(RMASK);
      if (lane == 1) mbar_arrive(&colfree[buf]);
    }
    return;
  }

  // ================= consumer warps ======================================================
  // Everything that is uniform over the CTA for the length of a batch lives in shared memory
  // (s_bat), not in registers: the register file is needed for the MAXQ row-block accumulators
  // and KS fragments, and a spilled scalar costs an L2 round trip in this kernel (the L1 is a
  // few KB next to a 220 KB shared-memory carve-out).
  // fragment offsets inside a tile: (row g, cols 2tg..2tg+1) and (rows 2tg, 2tg+1, col g)
  const int offC = pt_pos(g, 2 * tg), offT0 = pt_pos(2 * tg, g), offT1 = pt_pos(2 * tg + 1, g);
  double acc[MAXQ][2], ksA[MAXQ];
#pragma unroll
  for (int qq = 0; qq < MAXQ; ++qq) { acc[qq][0] = acc[qq][1] = 0.0; ksA[qq] = 0.0; }
  double col0 = 0.0, col1 = 0.0, gB = 0.0, hB0 = 0.0, hB1 = 0.0;
  // G(8p + g, tg) of the next panel travels global -> shared memory with cp.async while the current
  // panel is processed (a register prefetch would be waited for at the loop's back edge)
  double *const gpre = s_gpre + (size_t)wid * 64 + lane;   // two slots of 32 doubles per warp
  auto prefetch_g = [&](const double *Ga, int p) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\ncp.async.commit_group;" ::"r"(smem_u32(gpre + (p & 1) * 32)),
                 "l"(Ga + (size_t)p * 32 + lane)
                 : "memory");
  };
  int n_panel = 0;   // panels finished so far by this warp (selects the column-side buffer)
  int cur_p = -1;    // panel whose gB / hB are loaded

  auto flush_rows = [&]() {   // row-side slot of the batch that just ended: lane holds PHrow(8j+g, n = 2tg, 2tg+1)
    const int ch = (tg >> 1) == 0 ? s_bat.ch0 : s_bat.ch1;
    if (ch >= 0) {
      double *out = a.PHp + ((size_t)ch * (a.nsplit + 1) + s_bat.sp) * ld * 4;
#pragma unroll
      for (int qq = 0; qq < MAXQ; ++qq) {
        const int jj = wid + NW * qq;
        if (jj < nb)
          *reinterpret_cast<double2 *>(out + (size_t)(8 * jj + g) * 4 + 2 * (tg & 1)) = make_double2(acc[qq][0], acc[qq][1]);
      }
    }
#pragma unroll
    for (int qq = 0; qq < MAXQ; ++qq) acc[qq][0] = acc[qq][1] = 0.0;
  };
  // hand this warp's column-side tile of panel p to the reducer (p < 0: end of work)
  auto post_col = [&](int p) {
    const int buf = n_panel % NBUF;
    if (n_panel >= NBUF) mbar_wait(&colfree[buf], (uint32_t)((n_panel / NBUF - 1) & 1));
    *reinterpret_cast<double2 *>(s_colp + ((size_t)(buf * NW + wid) * 64 + 2 * lane)) = make_double2(col0, col1);
    if (wid == 0 && lane == 0) { PtColDesc cd; cd.p = p; cd.ch0 = s_bat.ch0; cd.ch1 = s_bat.ch1; s_cdesc[buf] = cd; }
    col0 = col1 = 0.0;
    __syncwarp();
    if (lane == 0) mbar_arrive(&colfull[buf]);
    ++n_panel;
  };

  int slot = 0, round = 0;
  for (;;) {
    mbar_wait(&full[slot], (uint32_t)(round & 1));
    const int d_item = s_desc[slot].item;
    if (d_item < 0) break;
    if (d_item != s_bat.item || s_desc[slot].b0 != s_bat.b0) {   // a new batch starts (uniform over the consumers)
      if (s_bat.item >= 0) flush_rows();
      const int fam = a.rev ? n_fam - 1 - d_item / a.nsplit : d_item / a.nsplit, b0 = s_desc[slot].b0;
      const int first = f.first[fam], nv = min(CB, f.cnt[fam] - b0);
      const int ch0 = f.child[first + b0], ch1 = nv > 1 ? f.child[first + b0 + 1] : -1;
      const double *Ga = src_base(a.st, a.st.G4, a.G4prev, f.anc[fam], (size_t)ld * 4);
      if (d_item != s_bat.item) {   // KS fragments of the family's ancestor: registers for the whole item
        const double *KSa = src_base(a.st, a.st.KS4, a.KS4prev, f.anc[fam], (size_t)ld * 4);
#pragma unroll
        for (int qq = 0; qq < MAXQ; ++qq) {
          const int jj = wid + NW * qq;
          ksA[qq] = jj < nb ? -KSa[(size_t)jj * 32 + lane] : 0.0;   // -KS(8jj + g, tg)
        }
      }
      named_barrier_sync(1, NW * 32);   // every consumer is done with the previous batch (operands, s_bat)
      if (tid == 0) {
        PtBatch nbt;
        nbt.item = d_item; nbt.b0 = b0; nbt.sp = d_item % a.nsplit; nbt.ch0 = ch0; nbt.ch1 = ch1;
        nbt.Pd0 = a.P + (size_t)a.dst_slot[ch0] * a.slab;
        nbt.Pd1 = ch1 >= 0 ? a.P + (size_t)a.dst_slot[ch1] * a.slab : nullptr;
        nbt.Ga = Ga;
        s_bat = nbt;
      }
      {   // siblings' H in fragment order: s_bC[j][lane] = H_s(b, 8j + 2tg + {0, 1}), n = g = 4s + b
        const double *H0 = a.H4 + (size_t)ch0 * ld * 4;
        const double *H1 = ch1 >= 0 ? a.H4 + (size_t)ch1 * ld * 4 : nullptr;
        for (int idx = tid; idx < nb * 32; idx += NW * 32) {
          const int l2 = idx & 31, jb = idx >> 5, g2 = l2 >> 2, t2 = l2 & 3;
          const double *Hs = (g2 >> 2) == 0 ? H0 : H1;
          double2 v = make_double2(0.0, 0.0);
          if (Hs != nullptr) {
            const size_t o = (size_t)(8 * jb + 2 * t2) * 4 + (g2 & 3);
            v = make_double2(Hs[o], Hs[o + 4]);
          }
          s_bC[idx] = v;
        }
      }
      cur_p = -1;
      asm volatile("cp.async.wait_group 0;" ::: "memory");   // a prefetch of the previous batch may still be landing
      prefetch_g(Ga, s_desc[slot].p);
      named_barrier_sync(1, NW * 32);
    }
    const double *st = ring + (size_t)slot * TS * 64;
    const int d_nt = s_desc[slot].nt;
    int t = 0, p = s_desc[slot].p, j = s_desc[slot].j;
    while (t < d_nt) {
      const int seg = min(d_nt - t, nb - j), jend = j + seg;
      if (p != cur_p) {   // operands of the panel: G(8p + g, tg) (prefetched one panel ahead) and the siblings' H at columns 8p..
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        gB = gpre[(p & 1) * 32];
        if (p + 1 < nb) prefetch_g(s_bat.Ga, p + 1);
        const double2 hv = s_bC[(size_t)p * 32 + lane];
        hB0 = hv.x; hB1 = hv.y;
        cur_p = p;
      }
      const double *tb = st + (long long)(t - j) * 64 + offC;                    // fragment of tile (jj, p) at tb + jj * 64
      const size_t gt = (size_t)(s_desc[slot].t0 + t - j) * 64 + offC;           // same, in the slab
      double *const P0 = s_bat.Pd0 + gt, *const P1 = s_bat.Pd1 ? s_bat.Pd1 + gt : nullptr;
      // this warp's row blocks in [j, jend): jj = wid + NW qq, qq in [qlo, qhi)
      const int qlo = j > wid ? (j - wid + NW - 1) / NW : 0;
      const int qhi = jend > wid ? (jend - wid + NW - 1) / NW : 0;
#pragma unroll
      for (int qq = 0; qq < MAXQ; ++qq) {
        if (qq >= qlo && qq < qhi) {
          const int jj = wid + NW * qq;
          const double *tp = tb + (size_t)jj * 64;
          double2 tv = *reinterpret_cast<const double2 *>(tp);
          dmma_m8n8k4(acc[qq][0], acc[qq][1], tv.x, hB0);      // row side, tile before its downdate
          dmma_m8n8k4(acc[qq][0], acc[qq][1], tv.y, hB1);
          if (jj != p) {                                       // column side: strictly lower tiles
            const double a0 = tp[offT0 - offC], a1 = tp[offT1 - offC];
            const double2 xv = s_bC[(size_t)jj * 32 + lane];
            dmma_m8n8k4(col0, col1, a0, xv.x);
            dmma_m8n8k4(col0, col1, a1, xv.y);
          }
          dmma_m8n8k4(tv.x, tv.y, ksA[qq], gB);                // the ancestor's pending downdate
          *reinterpret_cast<double2 *>(P0 + (size_t)jj * 64) = tv;
          if (P1 != nullptr) *reinterpret_cast<double2 *>(P1 + (size_t)jj * 64) = tv;
        }
      }
      t += seg;
      if (jend == nb) {   // the panel is complete
        post_col(p);
        ++p; j = p;
      } else {
        j = jend;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);   // this warp has read everything it needs from the slot
    if (++slot == NS) { slot = 0; ++round; }
  }
  if (s_bat.item >= 0) flush_rows();
  post_col(-1);   // terminal marker for the reducer
}

}  // namespace rb
