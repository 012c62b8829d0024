// Device-side plan of one sharded resampling step (one thread-block cluster, deterministic:
// every ordering decision comes from exclusive scans in particle order).  Same policy
// and same results as the host reference rbslam_plan_shard (sharded.cu): offspring stay on the
// ancestor's rank up to the capacity N/world (in particle order), the surplus fills the
// deficits of the other ranks in rank order; the first staying offspring keeps the ancestor's
// slab; migrants take dead slabs, remaining local copies take dead then exported-only slabs.
// On top it emits this rank's work lists in two groups (safe / deferred) and the fetch list.
// Running it on the device removes the only host round trip of the sharded step.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace rb {

#define RB_PW 8        // max ranks
#define RB_PLAN_CL 8   // CTAs of the planning cluster (portable maximum)

struct PlanArgs {
  int N, world, rank;
  const int *ai;                          // [N] ancestors of the new particles
  const int *owner_old, *lslot_old;       // [N]
  int *owner_new, *lslot_new;             // [N]
  int *n_child, *keeper, *unsafe, *inv;   // [N] scratch ([Nloc] for inv)
  int *dead_list, *expo_list;             // [world][Nloc] scratch
  // this rank's work
  int *src_slot, *glob;                   // [Nloc]
  int *listA, *listB;                     // [Nloc] each: group 0 first, then group 1
  int *fetch;                             // [Nloc][4]: dst slot, source rank, source slot, 0
  int *item_group;                        // [Nloc] or nullptr: 0 = safe, 1 = deferred behind the barrier
  int *counts;                            // [8]: nA0, nB0, nA1, nB1, nFetch, nMigTotal
  int fused;                              // 1: migrants are not fetched ahead of the pass: their source key is
                                          // Nloc + (index in fetch[]) and they belong to the safe group, which
                                          // runs before any exporter overwrites a slab a peer reads
};

// exclusive scan across the block of W ints per thread; totals (all threads get them)
template <int W>
__device__ __forceinline__ void block_scan_vec(int (&v)[W], int (&tot)[W], int *s_w /*[32][W]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int incl[W];
#pragma unroll
  for (int c = 0; c < W; ++c) {
    int x = v[c];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    incl[c] = x;
  }
  __syncthreads();
  if (lane == 31) {
#pragma unroll
    for (int c = 0; c < W; ++c) s_w[wid * W + c] = incl[c];
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < W; ++c) {
    int base = 0, t = 0;
    for (int q = 0; q < nw; ++q) {
      const int x = s_w[q * W + c];
      if (q < wid) base += x;
      t += x;
    }
    tot[c] = t;
    v[c] = base + incl[c] - v[c];
  }
  __syncthreads();
}

// the same scan across all CTAs of the cluster: block scan, block totals exchanged through
// distributed shared memory, two cluster barriers
template <int W>
__device__ __forceinline__ void cluster_scan_vec(int (&v)[W], int (&tot)[W], int *s_w /*[32][W]*/,
                                                 int *s_blk /*[W]*/, int *s_all /*[RB_PLAN_CL][W]*/) {
  namespace cg = cooperative_groups;
  block_scan_vec<W>(v, tot, s_w);
  cg::cluster_group cl = cg::this_cluster();
  const int nb = (int)cl.num_blocks(), rk = (int)cl.block_rank();
  if (nb == 1) return;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int c = 0; c < W; ++c) s_blk[c] = tot[c];
  }
  cl.sync();
  for (int q = threadIdx.x; q < nb * W; q += blockDim.x) s_all[q] = *cl.map_shared_rank(&s_blk[q % W], q / W);
  __syncthreads();
#pragma unroll
  for (int c = 0; c < W; ++c) {
    int base = 0, t = 0;
    for (int r = 0; r < nb; ++r) {
      const int x = s_all[r * W + c];
      if (r < rk) base += x;
      t += x;
    }
    v[c] += base;
    tot[c] = t;
  }
  cl.sync();   // s_blk / s_all are reused by the next scan
}

// Work group of a local offspring i of the local ancestor a (0 = safe: runs while peers may still read this
// rank's slabs; 1 = deferred behind the peer barrier).  Fetch-then-pass mode defers the whole family of an
// ancestor that is exported or has a copy landing in a slab a peer reads (one family = one group, copies
// before the in-place update).  Fused mode defers per item: only what WRITES a slab a peer reads -- the
// in-place offspring of such an ancestor, and a copy whose destination is such a slab; the other copies
// stay in the safe group (the family splits, its safe part simply runs first).
__device__ __forceinline__ int plan_group(const PlanArgs &p, int i, int a, int inplace) {
  const int u = p.unsafe[a];
  if (!p.fused) return u ? 1 : 0;
  if (inplace) return u ? 1 : 0;
  const int old = p.inv[p.lslot_new[i]];
  return (old != a && old >= 0 && p.n_child[old] > 0) ? 1 : 0;
}

__global__ void __launch_bounds__(1024) k_plan_shard(PlanArgs p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  __shared__ int s_w[32 * 16], s_blk[16], s_all[RB_PLAN_CL * 16];
  __shared__ int s_defp[RB_PW + 1], s_nmig[RB_PW], s_ndead[RB_PW];
  const int N = p.N, W = p.world, me = p.rank, cap = N / W, Nloc = cap;
  // threads of the whole cluster own consecutive chunks of particles: scans run in particle order
  const int GT = (int)cl.num_blocks() * blockDim.x;
  const int tid = (int)cl.block_rank() * blockDim.x + threadIdx.x;
  const int per = (N + GT - 1) / GT;
  const int b = min(N, tid * per), e = min(N, b + per);
  const int NONE = 0x7fffffff;
  for (int a = tid; a < N; a += GT) { p.n_child[a] = 0; p.keeper[a] = NONE; p.unsafe[a] = 0; }
  for (int j = tid; j < Nloc; j += GT) p.inv[j] = -1;
  cl.sync();
  // ---- 1. who stays: position of each child among the children of its ancestor's rank
  int cnt[RB_PW], tot[RB_PW];
#pragma unroll
  for (int r = 0; r < RB_PW; ++r) cnt[r] = 0;
  for (int i = b; i < e; ++i) {
    const int a = p.ai[i];
    atomicAdd(&p.n_child[a], 1);
    const int r = p.owner_old[a];
#pragma unroll
    for (int q = 0; q < RB_PW; ++q) cnt[q] += (q == r);
  }
  cluster_scan_vec<RB_PW>(cnt, tot, s_w, s_blk, s_all);
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < W; ++r) { s_defp[r] = acc; acc += cap - min(cap, tot[r]); }
    s_defp[W] = acc;
  }
  int n_sur = 0;
  for (int i = b; i < e; ++i) {
    const int a = p.ai[i], r = p.owner_old[a];
    int pos = 0;
#pragma unroll
    for (int q = 0; q < RB_PW; ++q) if (q == r) pos = cnt[q]++;
    if (pos < cap) { p.owner_new[i] = r; atomicMin(&p.keeper[a], i); }
    else { p.owner_new[i] = -1; ++n_sur; }
  }
  {
    int v1[1] = {n_sur}, t1[1];
    cluster_scan_vec<1>(v1, t1, s_w, s_blk, s_all);
    int k = v1[0];
    for (int i = b; i < e; ++i) {
      if (p.owner_new[i] >= 0) continue;
      int r2 = 0;
      while (r2 + 1 < W && k >= s_defp[r2 + 1]) ++r2;
      p.owner_new[i] = r2;
      atomicOr(&p.unsafe[p.ai[i]], 1);   // bit 0: the ancestor's slab is exported (a peer reads it)
      ++k;
    }
    if (tid == 0) p.counts[5] = t1[0];
  }
  cl.sync();
  // ---- 2. free slabs per rank: dead (no offspring anywhere) and exported-only, ordered by a
  int cf[16], tf[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) cf[q] = 0;
  for (int a = b; a < e; ++a) {
    if (p.keeper[a] != NONE) continue;
    const int r = p.owner_old[a], kind = p.n_child[a] == 0 ? 0 : 8;
#pragma unroll
    for (int q = 0; q < 16; ++q) cf[q] += (q == r + kind);
  }
  cluster_scan_vec<16>(cf, tf, s_w, s_blk, s_all);
  if (threadIdx.x == 0)
    for (int r = 0; r < W; ++r) s_ndead[r] = tf[r];
  for (int a = b; a < e; ++a) {
    if (p.keeper[a] != NONE) continue;
    const int r = p.owner_old[a];
    const bool dead = p.n_child[a] == 0;
    int pos = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) if (q == r + (dead ? 0 : 8)) pos = cf[q]++;
    (dead ? p.dead_list : p.expo_list)[(size_t)r * Nloc + pos] = p.lslot_old[a];
    if (r == me) { /* slot becomes free on this rank */ }
  }
  for (int a = b; a < e; ++a)
    if (p.owner_old[a] == me) p.inv[p.lslot_old[a]] = a;
  cl.sync();
  // ---- 3. slab of every new particle
  int cm[16], tm[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) cm[q] = 0;
  for (int i = b; i < e; ++i) {
    const int a = p.ai[i], r = p.owner_new[i];
    const bool mig = p.owner_old[a] != r;
    const bool copy = !mig && p.keeper[a] != i;
    if (!mig && !copy) continue;
#pragma unroll
    for (int q = 0; q < 16; ++q) cm[q] += (q == r + (mig ? 0 : 8));
  }
  cluster_scan_vec<16>(cm, tm, s_w, s_blk, s_all);
  if (threadIdx.x == 0)
    for (int r = 0; r < W; ++r) s_nmig[r] = tm[r];
  cl.sync();
  for (int i = b; i < e; ++i) {
    const int a = p.ai[i], r = p.owner_new[i];
    const bool mig = p.owner_old[a] != r;
    if (!mig && p.keeper[a] == i) { p.lslot_new[i] = p.lslot_old[a]; continue; }
    int pos = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) if (q == r + (mig ? 0 : 8)) pos = cm[q]++;
    if (mig) {
      p.lslot_new[i] = p.dead_list[(size_t)r * Nloc + pos];
    } else {
      const int qd = s_nmig[r] + pos;
      p.lslot_new[i] = qd < s_ndead[r] ? p.dead_list[(size_t)r * Nloc + qd]
                                       : p.expo_list[(size_t)r * Nloc + qd - s_ndead[r]];
    }
  }
  cl.sync();
  // ---- 4. deferred group: ancestors with an exported slab, or with a copy that lands in a
  //         slab a peer may still be reading
  for (int i = b; i < e; ++i) {
    if (p.owner_new[i] != me) continue;
    const int a = p.ai[i];
    if (p.owner_old[a] != me || p.keeper[a] == i) continue;
    const int old = p.inv[p.lslot_new[i]];
    if (old != a && old >= 0 && p.n_child[old] > 0) atomicOr(&p.unsafe[a], 2);   // bit 1: a copy lands in a slab a peer reads
  }
  cl.sync();
  // ---- 5. this rank's lists, in particle order
  int cl_[5], tl[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) cl_[q] = 0;
  for (int i = b; i < e; ++i) {
    if (p.owner_new[i] != me) continue;
    const int a = p.ai[i];
    if (p.owner_old[a] != me) { ++cl_[3]; ++cl_[4]; continue; }
    const int inplace = p.keeper[a] == i ? 1 : 0;
    const int grp = plan_group(p, i, a, inplace);
#pragma unroll
    for (int q = 0; q < 4; ++q) cl_[q] += (q == 2 * grp + inplace);
  }
  cluster_scan_vec<5>(cl_, tl, s_w, s_blk, s_all);
  for (int i = b; i < e; ++i) {
    if (p.owner_new[i] != me) continue;
    const int a = p.ai[i], j = p.lslot_new[i];
    p.glob[j] = i;
    if (p.owner_old[a] != me) {
      const int f = cl_[4]++;
      p.fetch[4 * f] = j; p.fetch[4 * f + 1] = p.owner_old[a]; p.fetch[4 * f + 2] = p.lslot_old[a];
      p.fetch[4 * f + 3] = 0;
      p.src_slot[j] = p.fused ? Nloc + f : j;
      if (p.item_group) p.item_group[j] = p.fused ? 0 : 1;
      p.listB[tl[1] + cl_[3]++] = j;
    } else {
      const int inplace = p.keeper[a] == i ? 1 : 0;
      const int grp = plan_group(p, i, a, inplace);
      p.src_slot[j] = p.lslot_old[a];
      if (p.item_group) p.item_group[j] = grp;
      const int q = 2 * grp + inplace;
      int pos = 0;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) if (qq == q) pos = cl_[qq]++;
      (inplace ? p.listB : p.listA)[(grp ? tl[inplace] : 0) + pos] = j;
    }
  }
  if (tid == 0) {
    p.counts[0] = tl[0]; p.counts[1] = tl[1]; p.counts[2] = tl[2]; p.counts[3] = tl[3]; p.counts[4] = tl[4];
  }
}

// one cluster of RB_PLAN_CL CTAs x 1024 threads
static inline cudaError_t launch_plan_shard(const PlanArgs &pa, cudaStream_t stream) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(RB_PLAN_CL); lc.blockDim = dim3(1024); lc.dynamicSmemBytes = 0; lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = RB_PLAN_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  return cudaLaunchKernelEx(&lc, k_plan_shard, pa);
}

}  // namespace rb
