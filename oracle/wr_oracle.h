/* TEST INFRASTRUCTURE — CPU oracle for the WebRender frame-draw hot path.
 *
 * Plain-C restatement of the reference's software implementation of the path
 * (SWGL: swgl/src/rasterize.h, blend.h, swgl_ext.h, texture.h + the `#ifdef
 * SWGL` branches of webrender/res/*.glsl), exposing the same call surface as
 * include/wrcu.h with a `wro_` prefix so a test can run one frame description
 * through the oracle and through the CUDA backend and compare bytes.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  It is never linked into or
 * called from the product (webrender_b200/).
 *
 * Pinning: tests/test_oracle_vs_swgl.py checks this restatement bit-for-bit
 * against oracle/_ref/libswgl_ref.so = the unmodified reference rasteriser
 * built from /root/reference/swgl/src/gl.cc (see oracle/Makefile), and against
 * committed golden buffers generated from that build (tests/golden/).
 */
#ifndef WR_ORACLE_H
#define WR_ORACLE_H
#include "../include/wrcu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wro_ctx wro_ctx;

int wro_ctx_create(wro_ctx** out);
void wro_ctx_destroy(wro_ctx* ctx);
const char* wro_last_error_string(wro_ctx* ctx);
int wro_texture_create(wro_ctx* ctx, int format, int width, int height, wrcu_tex* out);
int wro_texture_set_filter(wro_ctx* ctx, wrcu_tex tex, int filter);
int wro_texture_upload(wro_ctx* ctx, wrcu_tex tex, int x, int y, int w, int h,
                       const void* data, size_t src_stride);
int wro_texture_upload_batch(wro_ctx* ctx, wrcu_tex tex, const wrcu_upload_rect* rects, size_t n_rects,
                             const void* staging, size_t staging_bytes);
int wro_texture_copy(wro_ctx* ctx, wrcu_tex src, wrcu_tex dst, const int32_t src_rect[4], int dst_x, int dst_y);
int wro_gpu_cache_update(wro_ctx* ctx, int height, int clear, const wrcu_gpu_cache_copy* updates, size_t n_updates,
                         const float* blocks, size_t n_blocks);
int wro_texture_destroy(wro_ctx* ctx, wrcu_tex tex);
int wro_read_pixels(wro_ctx* ctx, wrcu_tex tex, int x, int y, int w, int h,
                    void* out, size_t dst_stride);
int wro_frame_begin(wro_ctx* ctx, const wrcu_frame_tables* tables);
int wro_frame_end(wro_ctx* ctx);
int wro_target_bind(wro_ctx* ctx, wrcu_tex color, wrcu_tex depth,
                    const float projection[16], const int32_t viewport[4]);
int wro_clear(wro_ctx* ctx, const int32_t rect[4], const float color[4],
              const float* depth);
int wro_draw_batch(wro_ctx* ctx, int kind, uint32_t features,
                   const wrcu_draw_state* state, const void* instances,
                   size_t instance_stride, int n_instances);
/* pixel-layers (destination pixels written per instance) since creation */
uint64_t wro_shaded_pixels(wro_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
