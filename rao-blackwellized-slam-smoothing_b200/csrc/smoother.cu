// placeholder until the smoother engine lands
#include "engine_internal.h"
int rb_info_init(rbslam_ctx *ctx) { return ctx->fail(RBSLAM_EARG, "information form not built yet"); }
void rb_smoother_free(rbslam_ctx *) {}
extern "C" int rbslam_smoother_run(rbslam_ctx *ctx, const rbslam_inputs *, int32_t, int32_t, rbslam_smoother_outputs *) {
  return ctx ? ctx->fail(RBSLAM_EARG, "smoother not built yet") : RBSLAM_EARG;
}
