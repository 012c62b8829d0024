// librbslam: one particle filter sharded over the GPUs of a node (one process per GPU).
//
// Particles are exchangeable within a step; ranks are coupled only through the weight
// normalisation and the resampling.  Every rank owns N/G covariance slabs (plus the thin
// per-slab arrays) and REPLICATES everything that is O(N): poses, weights, ancestors,
// the owner/slot maps.  Per step:
//   1. every rank draws the same ancestors from the same weights (k_resample, bit-exact),
//   2. every rank derives the same migration plan on the host (rbslam_plan_shard):
//      offspring stay on their ancestor's rank up to the rank's capacity, only the surplus
//      migrates; a rank either exports or imports, so migrants always land in DEAD slabs,
//   3. importers pull the migrants' ancestor state (slab, pending (G,KS), xl) straight out
//      of the exporter's HBM over NVLink with plain peer loads (k_peer_fetch; buffers are
//      mapped with CUDA IPC), then a peer barrier (flags in peer memory, system-scope
//      release/acquire) lets exporters reuse the slabs that were read,
//   4. the local streaming Kalman pass runs exactly as on one GPU (slot-indexed),
//   5. every rank stores its log-weights into every peer's replicated array (peer stores),
//      a second peer barrier, and every rank normalises the identical N log-weights with
//      the same fixed reduction tree -> identical weights on all ranks, independent of G.
// No NCCL call sits in the step loop; NCCL/gloo are only used by the host to exchange the
// 64-byte IPC handles once.
#include <vector>
#include <algorithm>
#include <cstring>
#include "engine_internal.h"
#include "step_kernels.cuh"
#include "shard_plan.cuh"
#include "packed_kernels.cuh"

using namespace rb;


struct PeerTable {
  void *p[SH_COUNT][RB_MAXW];
};

struct ShardWs {
  int world = 1, rank = 0, gN = 0, Nloc = 0;
  double *g_Xhist = nullptr, *g_w = nullptr, *g_wc = nullptr, *g_logw = nullptr;
  double *traj_max = nullptr, *traj_mean = nullptr;
  int *g_Ahist = nullptr, *iwmax = nullptr;
  unsigned long long *flags = nullptr;
  int *d_glob = nullptr, *d_fetch = nullptr;
  int *d_own[2] = {nullptr, nullptr}, *d_lsl[2] = {nullptr, nullptr};   // replicated owner / slot maps (device)
  int *d_group = nullptr;   // [Nloc] work group of each local item
  int *d_nchild = nullptr, *d_keeper = nullptr, *d_unsafe = nullptr, *d_inv = nullptr, *d_dead = nullptr, *d_expo = nullptr;
  bool host_plan = false;
  bool overlap = false;    // RBSLAM_OVERLAP=1: migrants travel while the safe work group runs (measured slower)
  bool fused = true;       // migrants' ancestor state is read in place from the exporter by the Kalman pass
                           // itself (no k_peer_fetch, no staging copy); RBSLAM_FUSED=0 restores fetch-then-pass
  const double **d_tab = nullptr;   // [SH_COUNT][1 + RB_MAXW] base pointers: [0] this rank, [1 + r] rank r
  PeerTable peers{};
  bool imported[SH_COUNT][RB_MAXW] = {};
  std::vector<int> owner[2], lslot[2];
  int cur = 0;
  unsigned long long epoch = 0;
  std::vector<int> h_ai, h_src, h_listA, h_listB, h_glob, h_fetch, h_nchild, h_inv, h_lists[4];
  std::vector<unsigned char> h_unsafe;
  std::vector<int> h_group;
  int64_t migrated = 0;
};

static ShardWs *sh_of(rbslam_ctx *ctx) { return static_cast<ShardWs *>(ctx->shard_ws); }

void rb_shard_free(rbslam_ctx *ctx) {
  ShardWs *s = sh_of(ctx);
  if (!s) return;
  for (int w = 0; w < SH_COUNT; ++w)
    for (int r = 0; r < RB_MAXW; ++r)
      if (s->imported[w][r]) cudaIpcCloseMemHandle(s->peers.p[w][r]);
  void *ptrs[] = {s->g_Xhist, s->g_w, s->g_wc, s->g_logw, s->traj_max, s->traj_mean, s->g_Ahist, s->iwmax,
                  s->flags, s->d_glob, s->d_fetch, s->d_own[0], s->d_own[1], s->d_lsl[0], s->d_lsl[1], s->d_nchild,
                  s->d_keeper, s->d_unsafe, s->d_inv, s->d_dead, s->d_expo, s->d_group, (void *)s->d_tab};
  for (void *p : ptrs) if (p) cudaFree(p);
  delete s;
  ctx->shard_ws = nullptr;
}

// ---------------------------------------------------------------------------
// host planner (pure integer logic; exported for the CPU tests)
// ---------------------------------------------------------------------------
extern "C" int rbslam_plan_shard(int32_t N, int32_t world, const int32_t *ai, const int32_t *owner_old,
                                 const int32_t *lslot_old, int32_t *owner_new, int32_t *lslot_new,
                                 int32_t *n_migrate) {
  if (N < 1 || world < 1 || N % world || !ai || !owner_old || !lslot_old || !owner_new || !lslot_new)
    return RBSLAM_EARG;
  int nm = 0;
  int rc = rbslam_plan_migration(N, world, ai, owner_old, owner_new, &nm);
  if (rc) return rc;
  const int Nloc = N / world;
  // children anywhere / children staying on the ancestor's rank
  std::vector<int> n_child(N, 0), keeper(N, -1);
  for (int i = 0; i < N; ++i) {
    ++n_child[ai[i]];
    if (owner_new[i] == owner_old[ai[i]] && keeper[ai[i]] < 0) keeper[ai[i]] = i;   // first local child
  }
  // free slots per rank as flat, rank-major lists: dead (no children anywhere) and
  // exported-only (children only on other ranks)
  std::vector<int> dcnt(world + 1, 0), ecnt(world + 1, 0);
  for (int a = 0; a < N; ++a) {
    if (keeper[a] >= 0) continue;
    ++(n_child[a] == 0 ? dcnt : ecnt)[owner_old[a] + 1];
  }
  for (int r = 0; r < world; ++r) { dcnt[r + 1] += dcnt[r]; ecnt[r + 1] += ecnt[r]; }
  std::vector<int> dead(dcnt[world]), expo(ecnt[world]), dfill(dcnt.begin(), dcnt.end() - 1),
      efill(ecnt.begin(), ecnt.end() - 1);
  for (int a = 0; a < N; ++a) {
    if (keeper[a] >= 0) continue;
    if (n_child[a] == 0) dead[dfill[owner_old[a]]++] = lslot_old[a];
    else expo[efill[owner_old[a]]++] = lslot_old[a];
  }
  std::vector<int> nd(dcnt.begin(), dcnt.end() - 1), ne(ecnt.begin(), ecnt.end() - 1);   // cursors
  // migrants first: they are fetched before the barrier and must land in dead slots
  if (nm > 0) {
    for (int i = 0; i < N; ++i) {
      const int r = owner_new[i];
      if (owner_old[ai[i]] == r) continue;
      if (nd[r] >= dcnt[r + 1]) return RBSLAM_EARG;   // cannot happen with the locality planner
      lslot_new[i] = dead[nd[r]++];
    }
  }
  for (int i = 0; i < N; ++i) {
    const int r = owner_new[i], a = ai[i];
    if (owner_old[a] != r) continue;
    if (keeper[a] == i) { lslot_new[i] = lslot_old[a]; continue; }
    if (nd[r] < dcnt[r + 1]) lslot_new[i] = dead[nd[r]++];
    else if (ne[r] < ecnt[r + 1]) lslot_new[i] = expo[ne[r]++];
    else return RBSLAM_EARG;
  }
  for (int r = 0; r < world; ++r)
    if (nd[r] != dcnt[r + 1] || ne[r] != ecnt[r + 1]) return RBSLAM_EARG;
  (void)Nloc;
  if (n_migrate) *n_migrate = nm;
  return RBSLAM_OK;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// pull the ancestor state of every migrant out of the exporter's HBM (NVLink peer loads).
// Work is flattened to (migrant, 64 KiB chunk) units and grid-strided so that a bounded grid
// (2 CTAs per SM when overlapping: 2 x 256 threads x 32 registers + the 192 x 232 registers of
// the streaming Kalman CTA still fit the 64 K register file, so both kernels are co-resident on
// every SM) keeps ~5 MB of loads in flight while the safe work group runs concurrently.
#define RB_FETCH_CHUNK2 4096   // double2 elements per unit = 64 KiB
__global__ void __launch_bounds__(256)
k_peer_fetch(const int *__restrict__ n_fetch_p, const int *__restrict__ fetch, PeerTable pt, int cg, int cx,
             size_t slab, int ld, int M, double *__restrict__ P, double *__restrict__ G4,
             double *__restrict__ KS4, double *__restrict__ xl) {
  const int n_fetch = *n_fetch_p;
  const size_t n2 = slab / 2;
  const int nchunk = (int)((n2 + RB_FETCH_CHUNK2 - 1) / RB_FETCH_CHUNK2);
  const long long units = (long long)n_fetch * (nchunk + 1);
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const int f = (int)(u / (nchunk + 1)), c = (int)(u % (nchunk + 1));
    const int j = fetch[4 * f], sr = fetch[4 * f + 1], ss = fetch[4 * f + 2];
    if (c < nchunk) {
      const double2 *srcP =
          reinterpret_cast<const double2 *>(static_cast<const double *>(pt.p[SH_P][sr]) + (size_t)ss * slab);
      double2 *dstP = reinterpret_cast<double2 *>(P + (size_t)j * slab);
      const size_t lo = (size_t)c * RB_FETCH_CHUNK2, hi = lo + RB_FETCH_CHUNK2 < n2 ? lo + RB_FETCH_CHUNK2 : n2;
      size_t idx = lo + threadIdx.x;
      for (; idx + 3 * 256 < hi; idx += 4 * 256) {   // four independent 16 B loads in flight per thread
        const double2 v0 = srcP[idx], v1 = srcP[idx + 256], v2 = srcP[idx + 512], v3 = srcP[idx + 768];
        dstP[idx] = v0; dstP[idx + 256] = v1; dstP[idx + 512] = v2; dstP[idx + 768] = v3;
      }
      for (; idx < hi; idx += 256) dstP[idx] = srcP[idx];
    } else {   // the thin per-particle arrays: pending downdate pair and the mean
      const double *sg = static_cast<const double *>(pt.p[cg ? SH_G4B : SH_G4A][sr]) + (size_t)ss * ld * 4;
      const double *sk = static_cast<const double *>(pt.p[cg ? SH_KS4B : SH_KS4A][sr]) + (size_t)ss * ld * 4;
      const double *sx = static_cast<const double *>(pt.p[cx ? SH_XLB : SH_XLA][sr]) + (size_t)ss * M;
      for (int idx = threadIdx.x; idx < ld * 4; idx += blockDim.x) {
        G4[(size_t)j * ld * 4 + idx] = sg[idx];
        KS4[(size_t)j * ld * 4 + idx] = sk[idx];
      }
      for (int idx = threadIdx.x; idx < M; idx += blockDim.x) xl[(size_t)j * M + idx] = sx[idx];
    }
  }
}

// every rank stores its local log-weights into every peer's replicated array
__global__ void k_logw_scatter(int Nloc, int world, const int *__restrict__ glob,
                               const double *__restrict__ logw_loc, PeerTable pt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Nloc) return;
  const double v = logw_loc[j];
  const int g = glob[j];
  for (int r = 0; r < world; ++r) static_cast<double *>(pt.p[SH_LOGW][r])[g] = v;
}

// barrier across the GPUs of the node: flags live in every rank's memory, mapped by all
// A rank that fails (not-PD error, CUDA error, host exception) never arrives: the spin is bounded
// (~20 s of SM clock) and a time-out is reported through the status word instead of hanging the
// stream.  Slots 16.. of the flags page carry every rank's not-PD flag so that all ranks, rank 0
// in particular, see a failure that happened on any shard.
#define RB_BARRIER_SPIN_CYCLES 40000000000ll
__global__ void k_peer_barrier(int world, int rank, unsigned long long epoch, PeerTable pt, DevStatus *status) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const unsigned long long bad = status->not_pd ? 1ull : 0ull;
  unsigned long long *rflags = static_cast<unsigned long long *>(pt.p[SH_FLAGS][r]);
  if (bad) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(rflags + 16 + rank), "l"(bad) : "memory");
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(rflags + rank), "l"(epoch) : "memory");
  const unsigned long long *mine = static_cast<const unsigned long long *>(pt.p[SH_FLAGS][rank]);
  unsigned long long v;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + r) : "memory");
    if (v < epoch && clock64() - t0 > RB_BARRIER_SPIN_CYCLES) { status->peer_timeout = 1; break; }
  } while (v < epoch);
  __threadfence_system();
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + 16 + r) : "memory");
  if (v && !status->not_pd) { status->not_pd = 1; status->not_pd_step = -1; status->not_pd_particle = -1 - r; }
}

__global__ void k_final_means_sharded(int gN, int M, const int *__restrict__ owner,
                                      const int *__restrict__ lslot, PeerTable pt, int cx,
                                      const double *__restrict__ w, const int *__restrict__ iw_max,
                                      double *__restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int which = cx ? SH_XLB : SH_XLA;
  const int im = *iw_max;
  out[r] = static_cast<const double *>(pt.p[which][owner[im]])[(size_t)lslot[im] * M + r];
  double acc = 0.0;
  for (int i = 0; i < gN; ++i)
    acc = fma(static_cast<const double *>(pt.p[which][owner[i]])[(size_t)lslot[i] * M + r], w[i], acc);
  out[M + r] = acc;
}

__global__ void k_final_cov_sharded(int M, int ld, const double *__restrict__ Pm, const double *__restrict__ Pl,
                                    const double *__restrict__ G4m, const double *__restrict__ KS4m,
                                    const double *__restrict__ G4l, const double *__restrict__ KS4l,
                                    const double *__restrict__ xll, double wl, const double *__restrict__ means,
                                    double *__restrict__ Pmax, double *__restrict__ Pmean, int layout) {
  const int c = blockIdx.x;
  const double *xmean = means + M;
  const double dc = xmean[c] - xll[c];
  for (int r = threadIdx.x; r < M; r += blockDim.x) {
    int rr, cc;   // packed symmetric slabs store (r,c) of an upper block as its mirror image
    slab_elem_rc(layout, r, c, rr, cc);
    const size_t off = slab_elem(layout, ld, r, c);
    double pm = Pm[off], pl = Pl[off];
    for (int b = 0; b < 4; ++b) {
      pm = fma(-KS4m[(size_t)rr * 4 + b], G4m[(size_t)cc * 4 + b], pm);
      pl = fma(-KS4l[(size_t)rr * 4 + b], G4l[(size_t)cc * 4 + b], pl);
    }
    Pmax[r + (size_t)c * M] = pm;
    Pmean[r + (size_t)c * M] = wl * (pl + (xmean[r] - xll[r]) * dc);
  }
}

// ---------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------
int rb_shard_create(rbslam_ctx *ctx, int world, int rank, int gN) {
  if (world > RB_MAXW) return ctx->fail(RBSLAM_EARG, "at most 8 ranks (one node)");
  if (ctx->kpath != 1) return ctx->fail(RBSLAM_EARG, "sharding needs the streaming Kalman path (dense model, M > ~160 or kalman_variant=2)");
  if (!ctx->cfg.keep_history) return ctx->fail(RBSLAM_EARG, "sharding needs keep_history=1");
  ShardWs *s = new ShardWs();
  ctx->shard_ws = s;
  s->world = world; s->rank = rank; s->gN = gN; s->Nloc = gN / world;
  const int T = ctx->T, n = ctx->n;
  RB_ALLOC(s->g_Xhist, (size_t)T * gN * n);
  RB_ALLOC(s->g_Ahist, (size_t)T * gN);
  RB_ALLOC(s->g_w, gN); RB_ALLOC(s->g_wc, gN); RB_ALLOC(s->g_logw, gN);
  RB_ALLOC(s->traj_max, (size_t)T * n); RB_ALLOC(s->traj_mean, (size_t)T * n); RB_ALLOC(s->iwmax, T);
  RB_ALLOC(s->flags, 64);
  RB_ALLOC(s->d_glob, s->Nloc); RB_ALLOC(s->d_fetch, (size_t)4 * s->Nloc);
  for (int b = 0; b < 2; ++b) { RB_ALLOC(s->d_own[b], gN); RB_ALLOC(s->d_lsl[b], gN); }
  RB_ALLOC(s->d_nchild, gN); RB_ALLOC(s->d_keeper, gN); RB_ALLOC(s->d_unsafe, gN); RB_ALLOC(s->d_inv, s->Nloc);
  RB_ALLOC(s->d_dead, gN); RB_ALLOC(s->d_expo, gN); RB_ALLOC(s->d_group, s->Nloc);
  s->host_plan = getenv("RBSLAM_HOST_PLAN") != nullptr;
  s->overlap = getenv("RBSLAM_OVERLAP") != nullptr;
  if (const char *e = getenv("RBSLAM_FUSED")) s->fused = atoi(e) != 0;
  if (s->overlap || s->host_plan || !ctx->use_fam) s->fused = false;   // the fused path is the family path with the device planner
  RB_ALLOC(s->d_tab, (size_t)SH_COUNT * (RB_MAXW + 1));
  CK(cudaMemset(s->flags, 0, 64 * sizeof(unsigned long long)));
  CK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->ev_fetch, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_plan, cudaEventDisableTiming));
  void *own[SH_COUNT] = {ctx->d_P, ctx->d_G4[0], ctx->d_G4[1], ctx->d_KS4[0], ctx->d_KS4[1], ctx->d_xl[0],
                         ctx->d_xl[1], s->g_logw, s->flags};
  for (int w = 0; w < SH_COUNT; ++w) s->peers.p[w][rank] = own[w];
  return RBSLAM_OK;
}

extern "C" int rbslam_ipc_export(rbslam_ctx *ctx, int32_t which, void *handle64) {
  if (!ctx || !handle64 || which < 0 || which >= SH_COUNT) return RBSLAM_EARG;
  ShardWs *s = sh_of(ctx);
  if (!s) return ctx->fail(RBSLAM_EARG, "context is not sharded (world == 1)");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, s->peers.p[which][s->rank]));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return RBSLAM_OK;
}

extern "C" int rbslam_ipc_import(rbslam_ctx *ctx, int32_t peer, int32_t which, const void *handle64) {
  if (!ctx || !handle64 || which < 0 || which >= SH_COUNT) return RBSLAM_EARG;
  ShardWs *s = sh_of(ctx);
  if (!s || peer < 0 || peer >= s->world || peer == s->rank) return ctx->fail(RBSLAM_EARG, "bad peer");
  CK(cudaSetDevice(ctx->cfg.device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  s->peers.p[which][peer] = p;
  s->imported[which][peer] = true;
  return RBSLAM_OK;
}

extern "C" int rbslam_ipc_count(void) { return SH_COUNT; }

// single-process group (rbslam_create_group): peers are wired with plain device pointers
// (cudaDeviceEnablePeerAccess between different devices; nothing to do on one device)
int rb_shard_wire(rbslam_ctx *a, rbslam_ctx *b) {
  ShardWs *sa = sh_of(a), *sb = sh_of(b);
  if (!sa || !sb) return RBSLAM_EARG;
  for (int w = 0; w < SH_COUNT; ++w) sa->peers.p[w][sb->rank] = sb->peers.p[w][sb->rank];
  return RBSLAM_OK;
}

// ---------------------------------------------------------------------------
// sharded filter
// ---------------------------------------------------------------------------
static int peer_barrier(rbslam_ctx *ctx) {
  ShardWs *s = sh_of(ctx);
  ++s->epoch;
  k_peer_barrier<<<1, 32, 0, ctx->stream>>>(s->world, s->rank, s->epoch, s->peers, ctx->d_status);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return RBSLAM_OK;
}

// between the safe and the deferred group: migrants have landed, every peer is done reading
static int shard_group_hook(rbslam_ctx *ctx, int) {
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_fetch, 0));
  return peer_barrier(ctx);
}

// phase 0: everything (one process per GPU); 1: all but the closing barrier; 2: the barrier only.
// A single process that drives several shards (rbslam_create_group) must finish phase 1 on every
// shard -- it allocates, frees and synchronises -- before any shard's barrier kernel starts to spin.
// fused migration: every peer is done reading the slabs that group 1 is about to overwrite
static int shard_barrier_hook(rbslam_ctx *ctx, int) { return peer_barrier(ctx); }

int rb_shard_begin(rbslam_ctx *ctx, const rbslam_inputs *in, int phase) {
  ShardWs *s = sh_of(ctx);
  if (phase == 2) return peer_barrier(ctx);
  for (int w = 0; w < SH_COUNT; ++w)
    for (int r = 0; r < s->world; ++r)
      if (!s->peers.p[w][r]) return ctx->fail(RBSLAM_EARG, "peer buffers not imported (rbslam_ipc_import)");
  if (in->x0_lin_cols != 1) return ctx->fail(RBSLAM_EARG, "sharded filter takes a single x0_lin column");
  if (in->U || in->Z || in->forced_ancestors)
    return ctx->fail(RBSLAM_EARG, "sharded filter uses the device Philox stream (rng_mode PHILOX)");
  int rc = rb_upload_inputs(ctx, in, 1);
  if (rc) return rc;
  ctx->jitter = 1e-3;
  ctx->sweep = 0;
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(DevStatus), ctx->stream));
  if ((rc = rb_init_state(ctx, false))) return rc;   // local slabs, xl, pending = 0, slot = identity
  const int gN = s->gN, n = ctx->n, Nloc = s->Nloc;
  double *x0d = ctx->d_scratch;
  CK(cudaMemcpyAsync(x0d, ctx->h_x0n.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  k_init_xn<<<(n * gN + 255) / 256, 256, 0, ctx->stream>>>(s->g_Xhist, n, gN, x0d);
  k_fill<<<(gN + 255) / 256, 256, 0, ctx->stream>>>(s->g_w, (size_t)gN, 1.0 / gN);
  ctx->launches += 2;
  // block distribution: particle i lives on rank i / Nloc at slot i % Nloc
  for (int b = 0; b < 2; ++b) { s->owner[b].assign(gN, 0); s->lslot[b].assign(gN, 0); }
  for (int i = 0; i < gN; ++i) { s->owner[0][i] = i / Nloc; s->lslot[0][i] = i % Nloc; }
  s->cur = 0;
  CK(cudaMemcpyAsync(s->d_own[0], s->owner[0].data(), sizeof(int) * gN, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(s->d_lsl[0], s->lslot[0].data(), sizeof(int) * gN, cudaMemcpyHostToDevice, ctx->stream));
  s->h_glob.assign(Nloc, 0);
  for (int j = 0; j < Nloc; ++j) s->h_glob[j] = s->rank * Nloc + j;
  CK(cudaMemcpyAsync(s->d_glob, s->h_glob.data(), sizeof(int) * Nloc, cudaMemcpyHostToDevice, ctx->stream));
  {   // base pointers of every shared buffer: [0] this rank, [1 + r] rank r (fused migration reads through them)
    std::vector<const double *> tab((size_t)SH_COUNT * (RB_MAXW + 1), nullptr);
    for (int w = 0; w < SH_COUNT; ++w) {
      tab[(size_t)w * (RB_MAXW + 1)] = static_cast<const double *>(s->peers.p[w][s->rank]);
      for (int r = 0; r < s->world; ++r) tab[(size_t)w * (RB_MAXW + 1) + 1 + r] = static_cast<const double *>(s->peers.p[w][r]);
    }
    if ((rc = rb_h2d(ctx, s->d_tab, tab.data(), sizeof(double *) * tab.size()))) return rc;
    ctx->d_peer_tab = s->d_tab;
  }
  s->migrated = 0;
  ctx->running = true;
  ctx->t = 0;
  CK(cudaMemsetAsync(s->flags + 16, 0, 16 * sizeof(unsigned long long), ctx->stream));   // peers' failure flags of an earlier run
  if (phase == 1) { CK(cudaStreamSynchronize(ctx->stream)); return RBSLAM_OK; }
  if ((rc = peer_barrier(ctx))) return rc;   // nobody starts before everybody is initialised
  return RBSLAM_OK;
}

int rb_shard_step(rbslam_ctx *ctx) {
  ShardWs *s = sh_of(ctx);
  if (!ctx->running) return ctx->fail(RBSLAM_EARG, "filter_step without filter_begin");
  if (ctx->t >= ctx->run_T) return ctx->fail(RBSLAM_EARG, "all time steps already processed");
  CK(cudaSetDevice(ctx->cfg.device));
  const int gN = s->gN, Nloc = s->Nloc, t = ctx->t, n = ctx->n, d = ctx->d, M = ctx->M;
  int rc;
  double *xn_t = s->g_Xhist + (size_t)t * gN * n;
  const bool resampled = t > 0;
  if (resampled) {
    rb_phase_begin(ctx, RB_PH_RESAMPLE);
    int *ai = s->g_Ahist + (size_t)t * gN;
    RngSrc rs;
    rs.U = nullptr; rs.seed = ctx->cfg.seed; rs.sweep = 0; rs.t = t;
    size_t smem = sizeof(double) * (size_t)gN;
    smem = std::min(smem, std::min(ctx->smem_resample_max, (size_t)(96 << 10)));
    const int fast = (rb_fast_scan() && gN >= 4096) ? 1 : 0;   // see rb_resample_phase
    if (fast) {
      k_scan_approx<<<1, 1024, 0, ctx->stream>>>(gN, s->g_w, s->g_wc, ctx->d_status);
      k_search_checked<<<(gN + 255) / 256, 256, 0, ctx->stream>>>(gN, 0, gN, s->g_wc, rs, nullptr, ai, ctx->d_status);
      ctx->launches += 2;
    }
    k_resample<<<1, 1024, smem, ctx->stream>>>(gN, 0, -1, s->g_w, s->g_wc, rs, nullptr, ai, ctx->d_status, fast);
    k_resample_search<<<(gN + 255) / 256, 256, 0, ctx->stream>>>(gN, 0, gN, s->g_wc, rs, nullptr, ai, ctx->d_status, fast);
    ctx->launches += 2;
    rb_phase_end(ctx);
    rb_phase_begin(ctx, RB_PH_INFO);   // reported as the "plan" phase of the sharded filter
    const int o = s->cur, nw = 1 - s->cur;
    if (!s->host_plan) {
      // plan on the device: no host round trip in the step (shard_plan.cuh)
      PlanArgs pa = {};
      pa.N = gN; pa.world = s->world; pa.rank = s->rank; pa.ai = ai;
      pa.owner_old = s->d_own[o]; pa.lslot_old = s->d_lsl[o]; pa.owner_new = s->d_own[nw]; pa.lslot_new = s->d_lsl[nw];
      pa.n_child = s->d_nchild; pa.keeper = s->d_keeper; pa.unsafe = s->d_unsafe; pa.inv = s->d_inv;
      pa.dead_list = s->d_dead; pa.expo_list = s->d_expo;
      pa.src_slot = ctx->d_src_slot; pa.glob = s->d_glob; pa.listA = ctx->d_listA; pa.listB = ctx->d_listB;
      pa.fetch = s->d_fetch; pa.counts = ctx->d_counts; pa.item_group = s->d_group;
      pa.fused = s->fused ? 1 : 0;
      CK(launch_plan_shard(pa, ctx->stream));
      ctx->launches += 1;
      s->cur = nw;
      ctx->stream_groups = 2;
      ctx->group_off[0][0] = ctx->group_off[0][1] = ctx->group_off[1][0] = ctx->group_off[1][1] = 0;
      ctx->group_off_dev[0][0] = ctx->group_off_dev[0][1] = nullptr;
      ctx->group_off_dev[1][0] = ctx->d_counts + 0;   // group 1 lists start after group 0's
      ctx->group_off_dev[1][1] = ctx->d_counts + 1;
    } else {
      // identical plan on every rank (host, pure integer logic)
      s->h_ai.resize(gN);
      if ((rc = rb_d2h(ctx, s->h_ai.data(), ai, sizeof(int) * gN))) return rc;
      int nmig = 0;
      if (rbslam_plan_shard(gN, s->world, s->h_ai.data(), s->owner[o].data(), s->lslot[o].data(),
                            s->owner[nw].data(), s->lslot[nw].data(), &nmig) != RBSLAM_OK)
        return ctx->fail(RBSLAM_EARG, "internal: shard plan failed");
      s->migrated += nmig;
      // Work lists in two groups.  Group 0 ("safe") touches no slab that a peer may still be
      // reading and runs WHILE migrants are in flight; group 1 (everything that involves an
      // exported slab, plus the migrants themselves) runs after the peer barrier.  All local
      // offspring of one ancestor are kept in the same group so that copies still precede the
      // in-place update of their source slab.
      s->h_src.assign(Nloc, 0); s->h_fetch.clear(); s->h_group.assign(Nloc, 0);
      for (int q = 0; q < 4; ++q) s->h_lists[q].clear();   // A0, B0, A1, B1
      s->h_nchild.assign(gN, 0); s->h_unsafe.assign(gN, 0); s->h_inv.assign(Nloc, -1);
      for (int i = 0; i < gN; ++i) {
        const int a = s->h_ai[i];
        ++s->h_nchild[a];
        if (s->owner[nw][i] != s->owner[o][a]) s->h_unsafe[a] = 1;        // ancestor's slab is exported
        if (s->owner[o][i] == s->rank) s->h_inv[s->lslot[o][i]] = i;      // old particle living in each slot
      }
      for (int i = 0; i < gN; ++i) {   // a copy landing in an exported-only slot must wait as well
        if (s->owner[nw][i] != s->rank) continue;
        const int a = s->h_ai[i];
        if (s->owner[o][a] != s->rank) continue;
        const int old = s->h_inv[s->lslot[nw][i]];
        if (old != a && old >= 0 && s->h_nchild[old] > 0) s->h_unsafe[a] = 1;
      }
      for (int i = 0; i < gN; ++i) {
        if (s->owner[nw][i] != s->rank) continue;
        const int j = s->lslot[nw][i], a = s->h_ai[i];
        s->h_glob[j] = i;
        if (s->owner[o][a] != s->rank) {          // migrant: fetched into slot j, then updated in place
          s->h_fetch.push_back(j); s->h_fetch.push_back(s->owner[o][a]); s->h_fetch.push_back(s->lslot[o][a]);
          s->h_fetch.push_back(0);
          s->h_src[j] = j; s->h_lists[3].push_back(j); s->h_group[j] = 1;
        } else {
          s->h_src[j] = s->lslot[o][a];
          const int grp = s->h_unsafe[a] ? 1 : 0;
          s->h_group[j] = grp;
          s->h_lists[2 * grp + (s->h_src[j] == j ? 1 : 0)].push_back(j);
        }
      }
      s->cur = nw;
      const int nF = (int)s->h_fetch.size() / 4;
      const int counts[4] = {(int)s->h_lists[0].size(), (int)s->h_lists[1].size(), (int)s->h_lists[2].size(),
                             (int)s->h_lists[3].size()};
      ctx->stream_groups = 2;
      ctx->group_off[0][0] = 0; ctx->group_off[0][1] = 0;
      ctx->group_off[1][0] = counts[0]; ctx->group_off[1][1] = counts[1];
      CK(cudaMemcpyAsync(ctx->d_src_slot, s->h_src.data(), sizeof(int) * Nloc, cudaMemcpyHostToDevice, ctx->stream));
      for (int q = 0; q < 4; ++q) {
        if (!counts[q]) continue;
        int *dst = (q & 1 ? ctx->d_listB : ctx->d_listA) + (q >= 2 ? counts[q - 2] : 0);
        CK(cudaMemcpyAsync(dst, s->h_lists[q].data(), sizeof(int) * counts[q], cudaMemcpyHostToDevice, ctx->stream));
      }
      CK(cudaMemcpyAsync(ctx->d_counts, counts, sizeof counts, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(s->d_glob, s->h_glob.data(), sizeof(int) * Nloc, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(s->d_group, s->h_group.data(), sizeof(int) * Nloc, cudaMemcpyHostToDevice, ctx->stream));
      if (nF) CK(cudaMemcpyAsync(s->d_fetch, s->h_fetch.data(), sizeof(int) * 4 * nF, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));   // the host vectors are reused next step
      ctx->h2d += (int64_t)sizeof(int) * (3 * Nloc + 4 * nF);
      CK(cudaMemcpyAsync(s->d_own[nw], s->owner[nw].data(), sizeof(int) * gN, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(s->d_lsl[nw], s->lslot[nw].data(), sizeof(int) * gN, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(ctx->d_counts + 4, &nF, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->group_off_dev[1][0] = ctx->group_off_dev[1][1] = nullptr;
    }
    rb_phase_end(ctx);
    rb_phase_begin(ctx, RB_PH_PROPAGATE);
    NormalSrc ns;
    ns.Z = nullptr; ns.seed = ctx->cfg.seed; ns.sweep = 0; ns.t = t;
    const double *Qp = ctx->d_Q + (ctx->Q_pages > 1 ? (size_t)(t - 1) * ctx->nw * ctx->nw : 0);
    k_propagate<<<(gN + 127) / 128, 128, 0, ctx->stream>>>(ctx->mc, gN, gN, s->g_Xhist + (size_t)(t - 1) * gN * n, ai,
                                                           ctx->d_odo + (size_t)(t - 1) * ctx->n_odo, ctx->h_dt[t - 1],
                                                           Qp, ns, xn_t);
    ctx->launches += 1;
    rb_phase_end(ctx);
    // migrants travel on a second stream, concurrently with the "safe" group of the Kalman pass
    // migrants travel on a second stream while the families that touch no exported slab
    // (work group 0) are processed; group 1 waits for the landing + the peer barrier
    cudaStream_t fs = s->overlap ? ctx->stream2 : ctx->stream;
    if (s->fused) {
      // nothing travels ahead of the pass: the Kalman pass reads each migrant's ancestor slab, pending pair
      // and mean in place from the exporter (SrcTab).  Group 0 = families no peer reads + the migrants;
      // peer barrier; group 1 = families that overwrite a slab a peer had to read first.
    } else {
    if (s->overlap) {
      CK(cudaEventRecord(ctx->ev_plan, ctx->stream));
      CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_plan, 0));
    }
    k_peer_fetch<<<(s->overlap ? 2 : 8) * ctx->num_sms, 256, 0, fs>>>(
        ctx->d_counts + 4, s->d_fetch, s->peers, ctx->cg, ctx->cx, ctx->slab, ctx->ld, M, ctx->d_P,
        ctx->d_G4[ctx->cg], ctx->d_KS4[ctx->cg], ctx->d_xl[ctx->cx]);
    ctx->launches += 1;
    if (s->overlap) CK(cudaEventRecord(ctx->ev_fetch, ctx->stream2));
    else if ((rc = peer_barrier(ctx))) return rc;   // land everything, then one pass over all families
    }
  } else {
    k_plan_identity<<<(Nloc + 255) / 256, 256, 0, ctx->stream>>>(Nloc, ctx->d_slot[ctx->cs], ctx->d_src_slot,
                                                                ctx->d_listB, ctx->d_counts);
    ctx->launches += 1;
    ctx->stream_groups = 1;
    ctx->group_off[0][0] = ctx->group_off[0][1] = 0;
  }
  rb_phase_begin(ctx, RB_PH_MEAS);
  k_meas<<<Nloc, 128, 0, ctx->stream>>>(ctx->mc, Nloc, xn_t, nullptr, M, nullptr, ctx->d_H, ctx->hs_p, ctx->hs_a,
                                        ctx->hs_c, ctx->ld, nullptr, s->d_glob);
  ctx->launches += 1;
  rb_phase_end(ctx);
  rb_phase_begin(ctx, RB_PH_KALMAN);
  ctx->anc_override = ctx->d_src_slot;     // thin arrays are slot-indexed: ancestor index = source slot
  ctx->stream_groups = resampled ? 2 : 1;
  const bool fused_step = s->fused && resampled;
  ctx->group_hook = fused_step ? shard_barrier_hook : (s->overlap ? shard_group_hook : nullptr);
  ctx->item_group = (fused_step || s->overlap) ? s->d_group : nullptr;
  ctx->st_nloc = fused_step ? Nloc : 0;    // source keys >= Nloc: migrants, read in place from their exporter
  ctx->st_fetch = s->d_fetch;
  ctx->fam_slabs = fused_step ? 2 * Nloc : 0;
  rc = rb_kalman_phase(ctx, ctx->d_y + (size_t)t * d, resampled);
  ctx->anc_override = nullptr;
  ctx->group_hook = nullptr;
  ctx->item_group = nullptr;
  ctx->st_nloc = 0; ctx->fam_slabs = 0;
  ctx->stream_groups = 1;
  ctx->group_off[0][0] = ctx->group_off[0][1] = 0;
  rb_phase_end(ctx);
  if (rc) return rc;
  rb_phase_begin(ctx, RB_PH_NORMALIZE);
  k_logw_scatter<<<(Nloc + 127) / 128, 128, 0, ctx->stream>>>(Nloc, s->world, s->d_glob, ctx->d_logw, s->peers);
  ctx->launches += 1;
  if ((rc = peer_barrier(ctx))) return rc;
  if ((rc = rb_normalize(ctx, gN, n, s->g_logw, s->g_w, xn_t, s->traj_max + (size_t)t * n, s->traj_mean + (size_t)t * n,
                         s->iwmax + t, nullptr, nullptr)))
    return rc;
  rb_phase_end(ctx);
  CK(cudaGetLastError());
  ctx->t += 1;
  return RBSLAM_OK;
}

// phase 0: everything; 1: status + outputs; 2: enqueue the closing barrier; 3: wait for it
int rb_shard_end(rbslam_ctx *ctx, rbslam_filter_outputs *out, int phase) {
  ShardWs *s = sh_of(ctx);
  const int gN = s->gN, M = ctx->M, n = ctx->n, T = ctx->t, ld = ctx->ld;
  if (T < 1) return ctx->fail(RBSLAM_EARG, "no step was run");
  int rc;
  if (phase == 2) return peer_barrier(ctx);
  if (phase == 3) {
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->running = false;
    return rb_check_status(ctx);
  }
  rc = rb_check_status(ctx);
  if (rc) return rc;
  if (s->rank == 0) {
    const int cur = s->cur;
    const int *iw = s->iwmax + (T - 1);
    double *means = ctx->d_scratch;
    double *Pmax = ctx->d_scratch + 2 * (size_t)M + 64, *Pmean = Pmax + (size_t)M * M;
    k_final_means_sharded<<<(M + 127) / 128, 128, 0, ctx->stream>>>(gN, M, s->d_own[cur], s->d_lsl[cur], s->peers, ctx->cx,
                                                                   s->g_w, iw, means);
    ctx->launches += 1;
    int im = 0;
    double wl = 0.0;
    if ((rc = rb_d2h(ctx, &im, iw, sizeof(int)))) return rc;
    if ((rc = rb_d2h(ctx, &wl, s->g_w + (gN - 1), sizeof(double)))) return rc;
    int own2[2], lsl2[2];   // owner / slot of the two particles whose slabs are read out
    const int sel[2] = {im, gN - 1};
    for (int q = 0; q < 2; ++q) {
      if ((rc = rb_d2h(ctx, &own2[q], s->d_own[cur] + sel[q], sizeof(int)))) return rc;
      if ((rc = rb_d2h(ctx, &lsl2[q], s->d_lsl[cur] + sel[q], sizeof(int)))) return rc;
    }
    auto at = [&](int which, int i, size_t stride) {
      const int q = (i == im) ? 0 : 1;
      return static_cast<const double *>(s->peers.p[which][own2[q]]) + (size_t)lsl2[q] * stride;
    };
    const int gsel = ctx->cg ? SH_G4B : SH_G4A, ksel = ctx->cg ? SH_KS4B : SH_KS4A, xsel = ctx->cx ? SH_XLB : SH_XLA;
    const size_t t4 = (size_t)ld * 4;
    k_final_cov_sharded<<<M, 128, 0, ctx->stream>>>(M, ld, at(SH_P, im, ctx->slab), at(SH_P, gN - 1, ctx->slab),
                                                    at(gsel, im, t4), at(ksel, im, t4), at(gsel, gN - 1, t4),
                                                    at(ksel, gN - 1, t4), at(xsel, gN - 1, M), wl, means, Pmax, Pmean,
                                                    ctx->layout);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if (out->traj_max && (rc = rb_d2h(ctx, out->traj_max, s->traj_max, sizeof(double) * n * T))) return rc;
    if (out->traj_mean && (rc = rb_d2h(ctx, out->traj_mean, s->traj_mean, sizeof(double) * n * T))) return rc;
    if (out->xl_max && (rc = rb_d2h(ctx, out->xl_max, means, sizeof(double) * M))) return rc;
    if (out->xl_mean && (rc = rb_d2h(ctx, out->xl_mean, means + M, sizeof(double) * M))) return rc;
    if (out->P_max && (rc = rb_d2h(ctx, out->P_max, Pmax, sizeof(double) * M * M))) return rc;
    if (out->P_mean && (rc = rb_d2h(ctx, out->P_mean, Pmean, sizeof(double) * M * M))) return rc;
    if (out->traj_sample_iwmax) {
      double *tmp = nullptr;
      RB_ALLOC(tmp, (size_t)n * T);
      k_trace<<<1, 32, 0, ctx->stream>>>(gN, n, T, s->g_Xhist, s->g_Ahist, iw, tmp);
      ctx->launches += 1;
      rc = rb_d2h(ctx, out->traj_sample_iwmax, tmp, sizeof(double) * n * T);
      cudaFree(tmp);
      if (rc) return rc;
    }
    if (out->xn_traj) {
      double *tmp = nullptr;
      RB_ALLOC(tmp, (size_t)n * gN * T);
      k_trace<<<(gN + 127) / 128, 128, 0, ctx->stream>>>(gN, n, T, s->g_Xhist, s->g_Ahist, nullptr, tmp);
      ctx->launches += 1;
      rc = rb_d2h(ctx, out->xn_traj, tmp, sizeof(double) * n * gN * T);
      cudaFree(tmp);
      if (rc) return rc;
    }
    if (out->ancestors && (rc = rb_d2h(ctx, out->ancestors, s->g_Ahist, sizeof(int) * (size_t)gN * T))) return rc;
  }
  if (phase == 1) return RBSLAM_OK;
  // peers keep their slabs alive until rank 0 has read what it needs
  if ((rc = peer_barrier(ctx))) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->running = false;
  return rb_check_status(ctx);
}

int64_t rb_shard_migrated(rbslam_ctx *ctx) { return sh_of(ctx) ? sh_of(ctx)->migrated : 0; }

// kernel-level entry point (parity test): run the DEVICE planner on host arrays
extern "C" int rbslam_op_plan_shard(int32_t device, int32_t N, int32_t world, int32_t rank, const int32_t *ai,
                                    const int32_t *owner_old, const int32_t *lslot_old, int32_t *owner_new,
                                    int32_t *lslot_new, int32_t *src_slot, int32_t *glob, int32_t *listA,
                                    int32_t *listB, int32_t *fetch, int32_t *counts8) {
  if (N < 1 || world < 1 || world > RB_PW || N % world || rank < 0 || rank >= world || !ai || !owner_old ||
      !lslot_old || !owner_new || !lslot_new || !src_slot || !glob || !listA || !listB || !fetch || !counts8)
    return RBSLAM_EARG;
  if (cudaSetDevice(device) != cudaSuccess) return RBSLAM_ECUDA;
  const int Nloc = N / world;
  int *buf = nullptr;
  const size_t total = (size_t)10 * N + (size_t)8 * Nloc + 8;
  if (cudaMalloc(&buf, total * sizeof(int)) != cudaSuccess) return RBSLAM_ECUDA;
  int *p = buf;
  auto take = [&](size_t n) { int *q = p; p += n; return q; };
  PlanArgs pa = {};
  pa.N = N; pa.world = world; pa.rank = rank;
  int *d_ai = take(N), *d_oo = take(N), *d_lo = take(N);
  pa.ai = d_ai; pa.owner_old = d_oo; pa.lslot_old = d_lo;
  pa.owner_new = take(N); pa.lslot_new = take(N); pa.n_child = take(N); pa.keeper = take(N); pa.unsafe = take(N);
  pa.dead_list = take(N); pa.expo_list = take(N);
  pa.inv = take(Nloc); pa.src_slot = take(Nloc); pa.glob = take(Nloc); pa.listA = take(Nloc); pa.listB = take(Nloc);
  pa.fetch = take((size_t)3 * Nloc); pa.counts = take(8);
  // fetch needs 4*Nloc ints: it shares the tail with counts only if 3*Nloc+8 >= 4*Nloc; allocate separately
  int *d_fetch = nullptr;
  if (cudaMalloc(&d_fetch, (size_t)4 * Nloc * sizeof(int)) != cudaSuccess) { cudaFree(buf); return RBSLAM_ECUDA; }
  pa.fetch = d_fetch;
  cudaMemcpy(d_ai, ai, sizeof(int) * N, cudaMemcpyHostToDevice);
  cudaMemcpy(d_oo, owner_old, sizeof(int) * N, cudaMemcpyHostToDevice);
  cudaMemcpy(d_lo, lslot_old, sizeof(int) * N, cudaMemcpyHostToDevice);
  cudaMemset(pa.counts, 0, 8 * sizeof(int));
  cudaError_t e = launch_plan_shard(pa, 0);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaMemcpy(owner_new, pa.owner_new, sizeof(int) * N, cudaMemcpyDeviceToHost);
  cudaMemcpy(lslot_new, pa.lslot_new, sizeof(int) * N, cudaMemcpyDeviceToHost);
  cudaMemcpy(src_slot, pa.src_slot, sizeof(int) * Nloc, cudaMemcpyDeviceToHost);
  cudaMemcpy(glob, pa.glob, sizeof(int) * Nloc, cudaMemcpyDeviceToHost);
  cudaMemcpy(listA, pa.listA, sizeof(int) * Nloc, cudaMemcpyDeviceToHost);
  cudaMemcpy(listB, pa.listB, sizeof(int) * Nloc, cudaMemcpyDeviceToHost);
  cudaMemcpy(fetch, pa.fetch, sizeof(int) * 4 * Nloc, cudaMemcpyDeviceToHost);
  cudaMemcpy(counts8, pa.counts, sizeof(int) * 8, cudaMemcpyDeviceToHost);
  cudaFree(d_fetch);
  cudaFree(buf);
  return e == cudaSuccess ? RBSLAM_OK : RBSLAM_ECUDA;
}
