// render.cu -- tile rendering entry points (placeholder until the wavefront kernels land).
#include "dscene.cuh"

using namespace izpi;

struct RenderState {};
void render_state_free(izpi_ctx* ctx) { delete ctx->render; ctx->render = nullptr; }

extern "C" {
int izpi_render_setup(izpi_ctx*, const izpi_render_config*) { set_error("render: not implemented"); return IZPI_ESTATE; }
int izpi_render_tiles(izpi_ctx*, int32_t, const uint32_t*, double*) { set_error("render: not implemented"); return IZPI_ESTATE; }
int izpi_render_canvas_device(izpi_ctx*, double**) { set_error("render: not implemented"); return IZPI_ESTATE; }
int izpi_render_finish(izpi_ctx*, double*, uint64_t*) { set_error("render: not implemented"); return IZPI_ESTATE; }
}
