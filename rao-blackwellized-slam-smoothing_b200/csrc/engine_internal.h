// Internal functions shared between engine.cu (filter) and smoother.cu.
#pragma once
#include "engine.h"
#include <cstdio>

template <typename T>
static inline int dev_alloc(rbslam_ctx *c, T **p, size_t count) {
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
  if (e != cudaSuccess) {
    c->fail_cuda(e, "cudaMalloc", __FILE__, __LINE__);
    char buf[128];
    snprintf(buf, sizeof buf, " (requested %.3f GB)", count * sizeof(T) / 1e9);
    c->err += buf;
    return RBSLAM_ECUDA;
  }
  return RBSLAM_OK;
}
#define RB_ALLOC(ptr, count)                                   \
  do {                                                         \
    int rc__ = dev_alloc(ctx, &(ptr), (size_t)(count));        \
    if (rc__) return rc__;                                     \
  } while (0)
#define CK(call)                                               \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) {                                  \
      ctx->fail_cuda(e__, #call, __FILE__, __LINE__);          \
      return RBSLAM_ECUDA;                                     \
    }                                                          \
  } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per DEVICE,
// so the "already done" flag is kept per (call site = kernel instantiation, device ordinal).
#define RB_OPTIN_SMEM(kern, bytes)                                                              \
  do {                                                                                          \
    static bool done__[64] = {false};                                                           \
    const int dv__ = ctx->cfg.device & 63;                                                      \
    if (!done__[dv__]) {                                                                        \
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      done__[dv__] = true;                                                                      \
    }                                                                                           \
  } while (0)

bool rb_fast_scan();   // resampling: provably exact parallel path first (RBSLAM_EXACT_SCAN=1 turns it off)
int rb_h2d(rbslam_ctx *ctx, void *dst, const void *src, size_t bytes);
int rb_d2h(rbslam_ctx *ctx, void *dst, const void *src, size_t bytes);
void rb_phase_begin(rbslam_ctx *ctx, int id);
void rb_phase_end(rbslam_ctx *ctx);
int rb_check_status(rbslam_ctx *ctx);
int rb_upload_inputs(rbslam_ctx *ctx, const rbslam_inputs *in, int K);
int rb_init_state(rbslam_ctx *ctx, bool info_form);
int rb_resample_phase(rbslam_ctx *ctx, int n_draws);
int rb_plan_phase(rbslam_ctx *ctx);
int rb_propagate_phase(rbslam_ctx *ctx, int n_prop);
int rb_meas_phase(rbslam_ctx *ctx, bool resampled);
int rb_kalman_phase(rbslam_ctx *ctx, const double *y_t_dev, bool resampled);
int rb_normalize_phase(rbslam_ctx *ctx);
int rb_normalize(rbslam_ctx *ctx, int N, int n, const double *logw, double *w, const double *xn, double *traj_max_t,
                 double *traj_mean_t, int *iw_max, double *logw_hist_t, double *w_hist_t);
int rb_read_slabs(rbslam_ctx *ctx, const double *slabs, double *host);
int rb_flush_pending(rbslam_ctx *ctx);
// buffers every rank of a sharded filter shares with its peers (sharded.cu)
#define RB_MAXW 8
enum { SH_P = 0, SH_G4A, SH_G4B, SH_KS4A, SH_KS4B, SH_XLA, SH_XLB, SH_LOGW, SH_FLAGS, SH_COUNT };
// sharded.cu
int rb_shard_create(rbslam_ctx *ctx, int world, int rank, int gN);
void rb_shard_free(rbslam_ctx *ctx);
int rb_shard_begin(rbslam_ctx *ctx, const rbslam_inputs *in, int phase = 0);
int rb_shard_step(rbslam_ctx *ctx);
int rb_shard_end(rbslam_ctx *ctx, rbslam_filter_outputs *out, int phase = 0);
int rb_shard_wire(rbslam_ctx *a, rbslam_ctx *b);
int64_t rb_shard_migrated(rbslam_ctx *ctx);
// smoother.cu
int rb_info_init(rbslam_ctx *ctx);
void rb_smoother_free(rbslam_ctx *ctx);
void rb_replicas_free(rbslam_ctx *leader);
