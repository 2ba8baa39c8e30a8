#include "bpt_context.h"
namespace bpt {
int build_accel(Context* ctx) { return ctx->fail(BPT_ERROR_NOT_READY, "build_accel: not implemented yet"); }
int intersect_batch(Context* ctx, int64_t, const float*, const float*, const float*, const float*, int32_t*, float*, float*, uint8_t*) { return ctx->fail(BPT_ERROR_NOT_READY, "intersect: not implemented yet"); }
}
