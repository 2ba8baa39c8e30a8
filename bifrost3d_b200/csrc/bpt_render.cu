#include "bpt_context.h"
namespace bpt {
int render(Context* ctx, const bpt_camera*, const bpt_settings*, int, int, uint32_t, uint32_t, int) { return ctx->fail(BPT_ERROR_NOT_READY, "render: not implemented yet"); }
int resolve_half4(Context* ctx, uint16_t*, int) { return ctx->fail(BPT_ERROR_NOT_READY, "resolve: not implemented yet"); }
int resolve_float4(Context* ctx, float*) { return ctx->fail(BPT_ERROR_NOT_READY, "resolve: not implemented yet"); }
void release_wavefront(Context*) {}
}
