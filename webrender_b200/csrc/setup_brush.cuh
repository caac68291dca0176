// setup_brush.cuh — vertex stage of the brush_* programs
// (webrender/res/brush.glsl:95-222 brush_shader_main_vs/main,
//  prim_shared.glsl:44-200 decode/fetch_prim_header/write_vertex/
//  clip_and_init_antialiasing/write_clip), SWGL branches: clip masks and edge
// AA are handed to the rasteriser (swgl_clipMask / swgl_antiAlias).
#pragma once
#include "setup_common.cuh"

#define WR_BRUSH_FLAG_FORCE_AA 1024

struct DevPrimHeader {
  float lr[4], lcr[4];  // local_rect, local_clip_rect (x0,y0,x1,y1)
  float z;
  int specific_prim_address, transform_id, picture_task_address;
  int user_data[4];
};

WRD DevPrimHeader wr_fetch_prim_header(const FrameTablesDev& T, int index) {
  DevPrimHeader ph;
  float4 f0 = wr_fetch(T.prim_headers_f, T.n_prim_headers_f, index * 2);
  float4 f1 = wr_fetch(T.prim_headers_f, T.n_prim_headers_f, index * 2 + 1);
  int4 i0 = wr_fetchi(T.prim_headers_i, T.n_prim_headers_i, index * 2);
  int4 i1 = wr_fetchi(T.prim_headers_i, T.n_prim_headers_i, index * 2 + 1);
  ph.lr[0] = f0.x; ph.lr[1] = f0.y; ph.lr[2] = f0.z; ph.lr[3] = f0.w;
  ph.lcr[0] = f1.x; ph.lcr[1] = f1.y; ph.lcr[2] = f1.z; ph.lcr[3] = f1.w;
  ph.z = (float)i0.x;
  ph.specific_prim_address = i0.y;
  ph.transform_id = i0.z;
  ph.picture_task_address = i0.w;
  ph.user_data[0] = i1.x; ph.user_data[1] = i1.y; ph.user_data[2] = i1.z; ph.user_data[3] = i1.w;
  return ph;
}

struct BrushVS {
  float2 local_pos[4];
  float4 world_pos[4];
  DevPrimHeader ph;
  float segment_rect[4];
  float4 segment_data;
  int brush_flags, edge_flags, resource_address;
  DevTransform transform;
  DevPictureTask task;
};

// write_clip → swgl_clipMask (prim_shared.glsl:183-190, swgl_ext.h:1867-1877)
WRD void wr_write_clip(const FrameTablesDev& T, int clip_address, const DevPictureTask& task,
                                     QuadOut& q) {
  if (clip_address >= 0x7FFFFFFF) return;  // CLIP_TASK_EMPTY → zero rect → ignored
  float4 a = wr_fetch(T.render_tasks, T.n_render_tasks, clip_address * 2);
  float4 b = wr_fetch(T.render_tasks, T.n_render_tasks, clip_address * 2 + 1);
  float sx = a.z - a.x, sy = a.w - a.y;
  if (sx != 0.0f || sy != 0.0f) {
    q.flags |= CMD_MASK;
    float ox = (task.tx0 - task.ox) - (a.x - b.y);
    float oy = (task.ty0 - task.oy) - (a.y - b.z);
    q.cm_off[0] = (int)ox;
    q.cm_off[1] = (int)oy;
    q.cm_bb[0] = (int)a.x;
    q.cm_bb[1] = (int)a.y;
    q.cm_bb[2] = (int)sx;
    q.cm_bb[3] = (int)sy;
  }
}

WRD void wr_brush_vertex(const SetupArgs& a, int4 aData, int vecs_per_specific_brush,
                                       QuadOut& q, BrushVS& o) {
  const FrameTablesDev& T = a.tabs;
  int prim_header_address = aData.x, clip_address = aData.y;
  int segment_index = aData.z & 0xffff, flags = aData.z >> 16;
  o.resource_address = aData.w & 0xffffff;
  o.ph = wr_fetch_prim_header(T, prim_header_address);
  o.transform = wr_fetch_transform(T, o.ph.transform_id);
  o.task = wr_fetch_picture_task(T, o.ph.picture_task_address);
  int edge_flags = (flags >> 12) & 0xf;
  int brush_flags = flags & 0xfff;
  o.edge_flags = edge_flags;
  o.brush_flags = brush_flags;
  float seg[4];
  if (segment_index == 0xffff) {
    seg[0] = o.ph.lr[0]; seg[1] = o.ph.lr[1]; seg[2] = o.ph.lr[2]; seg[3] = o.ph.lr[3];
    o.segment_data = make_float4(0, 0, 0, 0);
  } else {
    int sa = o.ph.specific_prim_address + vecs_per_specific_brush + segment_index * 2;
    float4 s0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, sa);
    o.segment_data = wr_fetch(T.gpu_cache, T.n_gpu_cache, sa + 1);
    seg[0] = s0.x + o.ph.lr[0]; seg[1] = s0.y + o.ph.lr[1];
    seg[2] = s0.z + o.ph.lr[0]; seg[3] = s0.w + o.ph.lr[1];
  }
  for (int i = 0; i < 4; i++) o.segment_rect[i] = seg[i];
  float adj[4] = {seg[0], seg[1], seg[2], seg[3]};
  bool antialiased = !o.transform.is_axis_aligned || (brush_flags & WR_BRUSH_FLAG_FORCE_AA) != 0;
  q.flags = 0;
  q.aa_edge_mask = 0;
  if (antialiased) {
    const float* cr = o.ph.lcr;
    int m = edge_flags | (cr[0] > adj[0] ? 1 : 0) | (cr[1] > adj[1] ? 2 : 0) | (cr[2] < adj[2] ? 4 : 0) |
            (cr[3] < adj[3] ? 8 : 0);
    q.aa_edge_mask = m;
    if (m) q.flags |= CMD_AA;
    adj[0] = wr_clamp(adj[0], cr[0], cr[2]); adj[1] = wr_clamp(adj[1], cr[1], cr[3]);
    adj[2] = wr_clamp(adj[2], cr[0], cr[2]); adj[3] = wr_clamp(adj[3], cr[1], cr[3]);
    o.ph.lcr[0] = o.ph.lcr[1] = -1.0e16f;
    o.ph.lcr[2] = o.ph.lcr[3] = 1.0e16f;
  }
  float fox = -o.task.ox + o.task.tx0, foy = -o.task.oy + o.task.ty0;
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int i = 0; i < 4; i++) {
    float lpx = (adj[2] - adj[0]) * ax[i] + adj[0], lpy = (adj[3] - adj[1]) * ay[i] + adj[1];
    lpx = wr_clamp(lpx, o.ph.lcr[0], o.ph.lcr[2]);
    lpy = wr_clamp(lpy, o.ph.lcr[1], o.ph.lcr[3]);
    float4 world = wr_mat_mul(o.transform.m, make_float4(lpx, lpy, 0.0f, 1.0f));
    float dpx = world.x * o.task.device_pixel_scale, dpy = world.y * o.task.device_pixel_scale;
    q.pos[i] = wr_mat_mul(a.tgt.proj, make_float4(dpx + fox * world.w, dpy + foy * world.w,
                                                  o.ph.z * world.w, world.w));
    o.local_pos[i] = make_float2(lpx, lpy);
    o.world_pos[i] = world;
  }
  wr_write_clip(T, clip_address, o.task, q);
}

WRD void wr_pack_color(QuadOut& q, const float* col) {
  q.col[0] = (uint16_t)wr_round_pixel(col[2], 255.0f);
  q.col[1] = (uint16_t)wr_round_pixel(col[1], 255.0f);
  q.col[2] = (uint16_t)wr_round_pixel(col[0], 255.0f);
  q.col[3] = (uint16_t)wr_round_pixel(col[3], 255.0f);
}

WRD void wr_finish_setup(const SetupArgs& a, int unsupported) {
  if (unsupported) {
    atomicAdd(&a.info->unsupported, 1);
    atomicAdd(a.err_counter, 1);
  }
}

// brush_solid (brush_solid.glsl:22-38): colour = gpu_cache[prim] * opacity
WRD void wr_setup_brush_solid_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  QuadOut q;
  BrushVS vs;
  memset(&q, 0, sizeof q);
  wr_brush_vertex(a, aData, 1, q, vs);
  float4 c = wr_fetch(a.tabs.gpu_cache, a.tabs.n_gpu_cache, vs.ph.specific_prim_address);
  float opacity = (float)vs.ph.user_data[0] / 65535.0f;
  float col[4] = {c.x * opacity, c.y * opacity, c.z * opacity, c.w * opacity};
  wr_pack_color(q, col);
  q.flags |= CMD_SPAN_SOLID | CMD_CONST_COLOR;  // swgl_drawSpanRGBA8/R8 both commit solid spans
  int unsupported = 0;
  wr_emit_quad(a, idx, q, &unsupported);
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_brush_solid)

// brush_image (brush_image.glsl:57-283, without WR_FEATURE_REPETITION)
WRD void wr_setup_brush_image_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  bool alpha_pass = (a.features & WRCU_FEAT_ALPHA_PASS) != 0;
  const bool repetition = (a.features & WRCU_FEAT_REPETITION) != 0;
  QuadOut q;
  BrushVS vs;
  memset(&q, 0, sizeof q);
  wr_brush_vertex(a, aData, 3, q, vs);
  const FrameTablesDev& T = a.tabs;
  float4 d0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.ph.specific_prim_address);
  float4 d2 = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.ph.specific_prim_address + 2);
  float image_color[4] = {d0.x, d0.y, d0.z, d0.w};
  float stretch[2] = {d2.x, d2.y};
  float tw = (float)a.color0.w, th = (float)a.color0.h;
  float4 r0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.resource_address);
  float uv0[2] = {r0.x, r0.y}, uv1[2] = {r0.z, r0.w};
  float lr[4] = {vs.ph.lr[0], vs.ph.lr[1], vs.ph.lr[2], vs.ph.lr[3]};
  if (stretch[0] < 0.0f) { stretch[0] = lr[2] - lr[0]; stretch[1] = lr[3] - lr[1]; }
  if (vs.brush_flags & 2) {  // SEGMENT_RELATIVE
    for (int i = 0; i < 4; i++) lr[i] = vs.segment_rect[i];
    stretch[0] = lr[2] - lr[0]; stretch[1] = lr[3] - lr[1];
    if (vs.brush_flags & 512) {  // TEXEL_RECT
      float usx = r0.z - r0.x, usy = r0.w - r0.y;
      uv0[0] = r0.x + vs.segment_data.x * usx; uv0[1] = r0.y + vs.segment_data.y * usy;
      uv1[0] = r0.x + vs.segment_data.z * usx; uv1[1] = r0.y + vs.segment_data.w * usy;
    }
    if (repetition) {  // brush_image.glsl:99-160
      const float* pr = vs.ph.lr;
      const float* sr = vs.segment_rect;
      if (vs.brush_flags & 512) {
        float rss[2] = {stretch[0], stretch[1]};
        float hus[2] = {uv1[0] - uv0[0], uv1[1] - uv0[1]}, vus[2] = {uv1[0] - uv0[0], uv1[1] - uv0[1]};
        if (vs.brush_flags & 256) {  // NINEPATCH_MIDDLE
          rss[0] = sr[0] - pr[0];
          rss[1] = sr[1] - pr[1];
          const float epsilon = 0.001f;
          vus[0] = uv0[0] - r0.x;
          if (vus[0] < epsilon || rss[0] < epsilon) {
            vus[0] = r0.z - uv1[0];
            rss[0] = pr[2] - sr[2];
          }
          hus[1] = uv0[1] - r0.y;
          if (hus[1] < epsilon || rss[1] < epsilon) {
            hus[1] = r0.w - uv1[1];
            rss[1] = pr[3] - sr[3];
          }
        }
        if (vs.brush_flags & 4) stretch[0] = rss[1] * (hus[0] / hus[1]);
        if (vs.brush_flags & 8) stretch[1] = rss[0] * (vus[1] / vus[0]);
      } else {
        if (vs.brush_flags & 4) stretch[0] = vs.segment_data.z - vs.segment_data.x;
        if (vs.brush_flags & 8) stretch[1] = vs.segment_data.w - vs.segment_data.y;
      }
      if (vs.brush_flags & 16) {
        float wdt = sr[2] - sr[0];
        float nx = wr_max(1.0f, roundf(wdt / stretch[0]));
        stretch[0] = wdt / nx;
      }
      if (vs.brush_flags & 32) {
        float hgt = sr[3] - sr[1];
        float ny = wr_max(1.0f, roundf(hgt / stretch[1]));
        stretch[1] = hgt / ny;
      }
    }
  }
  float perspective = (vs.brush_flags & 1) ? 1.0f : 0.0f;
  if (vs.brush_flags & 2048) { uv0[0] *= tw; uv0[1] *= th; uv1[0] *= tw; uv1[1] *= th; }
  float minu[2] = {wr_min(uv0[0], uv1[0]), wr_min(uv0[1], uv1[1])};
  float maxu[2] = {wr_max(uv0[0], uv1[0]), wr_max(uv0[1], uv1[1])};
  float fcold[8], gcold[8];
  fcold[0] = (minu[0] + 0.5f) / tw; fcold[1] = (minu[1] + 0.5f) / th;
  fcold[2] = (maxu[0] - 0.5f) / tw; fcold[3] = (maxu[1] - 0.5f) / th;
  int color_mode = vs.ph.user_data[0] & 0xffff, blend_mode = vs.ph.user_data[0] >> 16;
  int raster_space = vs.ph.user_data[1];
  float repeat[2] = {(lr[2] - lr[0]) / stretch[0], (lr[3] - lr[1]) / stretch[1]};
  float noff[2] = {0.0f, 0.0f};  // normalized_offset (brush_image.glsl:213-249)
  if (repetition) {
    if (vs.brush_flags & 64) { float h = repeat[0] * 0.5f + 0.5f; noff[0] = 1.0f - (h - floorf(h)); }
    if (vs.brush_flags & 128) { float h = repeat[1] * 0.5f + 0.5f; noff[1] = 1.0f - (h - floorf(h)); }
  }
  for (int k = 0; k < 4; k++) {
    float fx = (vs.local_pos[k].x - lr[0]) / (lr[2] - lr[0]);
    float fy = (vs.local_pos[k].y - lr[1]) / (lr[3] - lr[1]);
    if (raster_space == 1) {  // get_image_quad_uv
      float4 tl = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.resource_address + 2);
      float4 tr = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.resource_address + 3);
      float4 bl = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.resource_address + 4);
      float4 br = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.resource_address + 5);
      float Xx = (tr.x - tl.x) * fx + tl.x, Xy = (tr.y - tl.y) * fx + tl.y, Xw = (tr.w - tl.w) * fx + tl.w;
      float Yx = (br.x - bl.x) * fx + bl.x, Yy = (br.y - bl.y) * fx + bl.y, Yw = (br.w - bl.w) * fx + bl.w;
      float Zx = (Yx - Xx) * fy + Xx, Zy = (Yy - Xy) * fy + Xy, Zw = (Yw - Xw) * fy + Xw;
      fx = Zx / Zw;
      fy = Zy / Zw;
    }
    float ux = ((uv1[0] - uv0[0]) * fx + uv0[0]) - minu[0];
    float uy = ((uv1[1] - uv0[1]) * fy + uv0[1]) - minu[1];
    ux *= repeat[0]; uy *= repeat[1];
    if (repetition) {
      ux += noff[0] * (maxu[0] - minu[0]);
      uy += noff[1] * (maxu[1] - minu[1]);
    }
    ux /= tw; uy /= th;
    if (perspective == 0.0f) { ux *= vs.world_pos[k].w; uy *= vs.world_pos[k].w; }
    if (repetition) {  // brush_image.glsl:260-265
      ux /= (maxu[0] / tw - minu[0] / tw);
      uy /= (maxu[1] / th - minu[1] / th);
    }
    q.interp[k][0] = ux;
    q.interp[k][1] = uy;
  }
  q.n_interp = 2;
  fcold[4] = minu[0] / tw; fcold[5] = minu[1] / th;
  fcold[6] = perspective;
  float fw = 1.0f / q.pos[0].w;
  if (!isfinite(fw)) fw = 0.0f;
  fcold[7] = fw;
  float vcolor[4] = {1.0f, 1.0f, 1.0f, 1.0f}, swz[2] = {1.0f, 0.0f};
  bool drop_shadow = false;
  uint16_t shadow[4] = {0, 0, 0, 0};
  if (alpha_pass) {
    float opacity = (float)vs.ph.user_data[2] / 65535.0f;
    if (blend_mode == 0) image_color[3] *= opacity;
    else for (int i = 0; i < 4; i++) image_color[i] *= opacity;
    switch (color_mode) {
      case 0: case 2:
        drop_shadow = a.blend_enabled != 0;  // swgl_blendDropShadow needs blending on
        shadow[0] = (uint16_t)wr_round_pixel(image_color[2], 255.0f); shadow[1] = (uint16_t)wr_round_pixel(image_color[1], 255.0f);
        shadow[2] = (uint16_t)wr_round_pixel(image_color[0], 255.0f); shadow[3] = (uint16_t)wr_round_pixel(image_color[3], 255.0f);
        break;
      case 4: for (int i = 0; i < 4; i++) vcolor[i] = image_color[i]; break;
      case 3: for (int i = 0; i < 4; i++) vcolor[i] = image_color[3]; break;
      case 1: swz[0] = image_color[3]; swz[1] = 0.0f; for (int i = 0; i < 4; i++) vcolor[i] = image_color[i]; break;
      case 5: swz[0] = -image_color[3]; swz[1] = image_color[3]; for (int i = 0; i < 4; i++) vcolor[i] = image_color[i]; break;
      default: swz[0] = swz[1] = 0.0f; break;
    }
  }
  for (int i = 0; i < 4; i++) gcold[i] = vcolor[i];
  gcold[4] = swz[0]; gcold[5] = swz[1];
  gcold[6] = alpha_pass ? 1.0f : 0.0f;
  // swgl_drawSpanRGBA8 bails for non-(1,0) swizzles in the alpha pass
  gcold[7] = (!alpha_pass || (swz[0] == 1.0f && swz[1] == 0.0f)) ? 1.0f : 0.0f;
  wr_pack_color(q, vcolor);
  q.flags |= CMD_TEXTURED;
  if (drop_shadow) q.flags |= CMD_DROP_SHADOW;
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 8; i++) { k->f[i] = fcold[i]; k->g[i] = gcold[i]; }
    k->i[0] = (int)shadow[0] | ((int)shadow[1] << 16);
    k->i[1] = (int)shadow[2] | ((int)shadow[3] << 16);
    k->g[8] = maxu[0] / tw;  // v_uv_bounds.zw
    k->g[9] = maxu[1] / th;
    k->g[10] = (repetition && alpha_pass) ? repeat[0] + noff[0] : 0.0f;  // v_tile_repeat_bounds
    k->g[11] = (repetition && alpha_pass) ? repeat[1] + noff[1] : 0.0f;
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_brush_image)

// ps_split_composite (ps_split_composite.glsl:64-118; SplitCompositeInstance, gpu_types.rs:531-552): one
// polygon of a plane-split preserve-3d picture — four local points from the GPU cache, transformed
// (usually with perspective: wr_emit_quad's draw_perspective branch) and textured from the picture's
// surface.  The fragment / span stage is brush_image's opaque variant with uv bounds offset 0
// (ps_split_composite.glsl:121-136), so the command is shaded by ImageShader.
WRD void wr_setup_split_composite_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  const FrameTablesDev& T = a.tabs;
  const int prim_header_index = aData.x, polygons_address = aData.y, render_task_index = aData.w;
  const float ci_z = (float)aData.z;
  QuadOut q;
  memset(&q, 0, sizeof q);
  const float4 data0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, polygons_address);
  const float4 data1 = wr_fetch(T.gpu_cache, T.n_gpu_cache, polygons_address + 1);
  const float lx[4] = {data0.x, data0.z, data1.x, data1.z}, ly[4] = {data0.y, data0.w, data1.y, data1.w};
  DevPrimHeader ph = wr_fetch_prim_header(T, prim_header_index);
  DevPictureTask task = wr_fetch_picture_task(T, render_task_index);
  DevTransform transform = wr_fetch_transform(T, ph.transform_id);
  const float4 r0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, ph.user_data[0]);
  const float uv0[2] = {r0.x, r0.y}, uv1[2] = {r0.z, r0.w};
  const float dox = task.tx0 - task.ox, doy = task.ty0 - task.oy;  // dest_origin
  const float tw = (float)a.color0.w, th = (float)a.color0.h;
  const float4 tl = wr_fetch(T.gpu_cache, T.n_gpu_cache, ph.user_data[0] + 2);
  const float4 tr = wr_fetch(T.gpu_cache, T.n_gpu_cache, ph.user_data[0] + 3);
  const float4 bl = wr_fetch(T.gpu_cache, T.n_gpu_cache, ph.user_data[0] + 4);
  const float4 br = wr_fetch(T.gpu_cache, T.n_gpu_cache, ph.user_data[0] + 5);
  const float perspective = (float)ph.user_data[1];
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int i = 0; i < 4; i++) {
    // bilerp(local[0], local[1], local[3], local[2], aPosition.y, aPosition.x)
    const float t = ax[i], sgm = ay[i];
    const float xx = (lx[1] - lx[0]) * t + lx[0], xy = (ly[1] - ly[0]) * t + ly[0];
    const float yx = (lx[2] - lx[3]) * t + lx[3], yy = (ly[2] - ly[3]) * t + ly[3];
    const float lpx = (yx - xx) * sgm + xx, lpy = (yy - xy) * sgm + xy;
    const float4 world = wr_mat_mul(transform.m, make_float4(lpx, lpy, 0.0f, 1.0f));
    q.pos[i] = wr_mat_mul(a.tgt.proj, make_float4(dox * world.w + world.x * task.device_pixel_scale,
                                                  doy * world.w + world.y * task.device_pixel_scale, world.w * ci_z, world.w));
    float fx = (lpx - ph.lr[0]) / (ph.lr[2] - ph.lr[0]);
    float fy = (lpy - ph.lr[1]) / (ph.lr[3] - ph.lr[1]);
    {  // get_image_quad_uv (prim_shared.glsl:202-210)
      const float Xx = (tr.x - tl.x) * fx + tl.x, Xy = (tr.y - tl.y) * fx + tl.y, Xw = (tr.w - tl.w) * fx + tl.w;
      const float Yx = (br.x - bl.x) * fx + bl.x, Yy = (br.y - bl.y) * fx + bl.y, Yw = (br.w - bl.w) * fx + bl.w;
      const float Zx = (Yx - Xx) * fy + Xx, Zy = (Yy - Xy) * fy + Xy, Zw = (Yw - Xw) * fy + Xw;
      fx = Zx / Zw;
      fy = Zy / Zw;
    }
    const float ux = (uv1[0] - uv0[0]) * fx + uv0[0], uy = (uv1[1] - uv0[1]) * fy + uv0[1];
    const float mixw = (1.0f - q.pos[i].w) * perspective + q.pos[i].w;  // mix(gl_Position.w, 1.0, perspective_interpolate)
    q.interp[i][0] = ux / tw * mixw;
    q.interp[i][1] = uy / th * mixw;
  }
  q.n_interp = 2;
  q.flags = CMD_TEXTURED;
  q.aa_edge_mask = 0;
  wr_write_clip(T, ph.user_data[3], task, q);
  const float minu[2] = {wr_min(uv0[0], uv1[0]), wr_min(uv0[1], uv1[1])};
  const float maxu[2] = {wr_max(uv0[0], uv1[0]), wr_max(uv0[1], uv1[1])};
  float fcold[8], gcold[8];
  fcold[0] = (minu[0] + 0.5f) / tw; fcold[1] = (minu[1] + 0.5f) / th;
  fcold[2] = (maxu[0] - 0.5f) / tw; fcold[3] = (maxu[1] - 0.5f) / th;
  fcold[4] = 0.0f; fcold[5] = 0.0f;
  fcold[6] = perspective;
  float fw = 1.0f / q.pos[0].w;
  if (!isfinite(fw)) fw = 0.0f;
  fcold[7] = fw;
  const float one[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  for (int i = 0; i < 4; i++) gcold[i] = 1.0f;
  gcold[4] = 1.0f; gcold[5] = 0.0f;
  gcold[6] = 0.0f;  // colour = alpha (1) * texel
  gcold[7] = 1.0f;  // swgl_drawSpanRGBA8 always commits (RGBA8 surface)
  wr_pack_color(q, one);
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 8; i++) { k->f[i] = fcold[i]; k->g[i] = gcold[i]; }
    k->i[0] = k->i[1] = 0;
    k->g[8] = maxu[0] / tw;
    k->g[9] = maxu[1] / th;
    k->g[10] = k->g[11] = 0.0f;
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_split_composite)

// ps_text_run main (ps_text_run.glsl:98-264), no GLYPH_TRANSFORM
WRD void wr_setup_text_run_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  const FrameTablesDev& T = a.tabs;
  int prim_header_address = aData.x, clip_address = aData.y;
  int glyph_index = aData.z & 0xffff, flags = aData.z >> 16;
  int resource_address = aData.w & 0xffffff;
  int subpx_dir = (flags >> 8) & 0xff, color_mode = flags & 0xff;
  QuadOut q;
  memset(&q, 0, sizeof q);
  DevPrimHeader ph = wr_fetch_prim_header(T, prim_header_address);
  DevTransform transform = wr_fetch_transform(T, ph.transform_id);
  DevPictureTask task = wr_fetch_picture_task(T, ph.picture_task_address);
  float4 text_color = wr_fetch(T.gpu_cache, T.n_gpu_cache, ph.specific_prim_address);
  float tox = ph.lr[2], toy = ph.lr[3];  // text_offset = local_rect.p1
  float4 gd = wr_fetch(T.gpu_cache, T.n_gpu_cache, ph.specific_prim_address + 1 + (int)((unsigned)glyph_index / 2U));
  float gox = ((unsigned)glyph_index % 2U == 1U) ? gd.z : gd.x;
  float goy = ((unsigned)glyph_index % 2U == 1U) ? gd.w : gd.y;
  gox += ph.lr[0];
  goy += ph.lr[1];
  float4 r0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, resource_address);
  float4 r1 = wr_fetch(T.gpu_cache, T.n_gpu_cache, resource_address + 1);
  float res_scale = r1.z;
  float snx = 0.5f, sny = 0.5f;
  if (subpx_dir == 1) snx = 0.125f;
  else if (subpx_dir == 2) sny = 0.125f;
  else if (subpx_dir == 3) snx = sny = 0.125f;
  float raster_scale = (float)ph.user_data[0] / 65535.0f;
  float grs = raster_scale * task.device_pixel_scale;
  float gsi = res_scale / grs;
  float rgx = floorf(gox * grs + snx) / res_scale, rgy = floorf(goy * grs + sny) / res_scale;
  float ox = gsi * (r1.x + rgx) + tox, oy = gsi * (r1.y + rgy) + toy;
  float gr[4] = {ox, oy, ox + gsi * (r0.z - r0.x), oy + gsi * (r0.w - r0.y)};
  // WR_FEATURE_GLYPH_TRANSFORM (ps_text_run.glsl:129-167, 222-231): the glyph rect lives in the
  // transformed (glyph raster) space and the quad is its pre-image
  const bool glyph_transform = (a.features & WRCU_FEAT_GLYPH_TRANSFORM) != 0;
  float gt[4] = {1, 0, 0, 1}, gti[4] = {1, 0, 0, 1};  // column major: c0.x, c0.y, c1.x, c1.y
  float lrect[4] = {gr[0], gr[1], gr[2], gr[3]};
  bool inside = false;
  if (glyph_transform) {
    const float dps = task.device_pixel_scale;
    const float* m = transform.m;
    gt[0] = m[0] * dps; gt[1] = m[1] * dps; gt[2] = m[4] * dps; gt[3] = m[5] * dps;
    const float gtx = m[12] * dps, gty = m[13] * dps;
    const float det = gt[0] * gt[3] - gt[1] * gt[2];
    const float idet = (float)(1.0 / (double)det);  // glsl.h:2905-2908: `* (1. / det)`
    gti[0] = gt[3] * idet; gti[1] = -gt[1] * idet; gti[2] = -gt[2] * idet; gti[3] = gt[0] * idet;
    const float rgx2 = floorf((gt[0] * gox + gt[2] * goy) + snx), rgy2 = floorf((gt[1] * gox + gt[3] * goy) + sny);
    const float rtx = floorf(((gt[0] * tox + gt[2] * toy) + gtx) + 0.5f) - gtx;
    const float rty = floorf(((gt[1] * tox + gt[3] * toy) + gty) + 0.5f) - gty;
    const float orx = (r1.x + rgx2) + rtx, ory = (r1.y + rgy2) + rty;
    gr[0] = orx; gr[1] = ory;
    gr[2] = (orx + r0.z) - r0.x;
    gr[3] = (ory + r0.w) - r0.y;
    const float szx = gr[2] - gr[0], szy = gr[3] - gr[1];
    const float hcx = gr[0] + szx * 0.5f, hcy = gr[1] + szy * 0.5f;
    const float cx = gti[0] * hcx + gti[2] * hcy, cy = gti[1] * hcx + gti[3] * hcy;
    const float hsx = szx * 0.5f, hsy = szy * 0.5f;
    const float rdx = fabsf(gti[0]) * hsx + fabsf(gti[2]) * hsy, rdy = fabsf(gti[1]) * hsx + fabsf(gti[3]) * hsy;
    lrect[0] = cx - rdx; lrect[1] = cy - rdy; lrect[2] = cx + rdx; lrect[3] = cy + rdy;
    inside = ph.lcr[0] <= lrect[0] && ph.lcr[1] <= lrect[1] && lrect[2] <= ph.lcr[2] && lrect[3] <= ph.lcr[3];
  }
  float fox = -task.ox + task.tx0, foy = -task.oy + task.ty0;
  float tw = (float)a.color0.w, th = (float)a.color0.h;
  float st0x = r0.x / tw, st0y = r0.y / th, st1x = r0.z / tw, st1y = r0.w / th;
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int k = 0; k < 4; k++) {
    float lpx = (gr[2] - gr[0]) * ax[k] + gr[0], lpy = (gr[3] - gr[1]) * ay[k] + gr[1];
    if (glyph_transform) {
      if (inside) {
        const float gx = lpx, gy = lpy;
        lpx = gti[0] * gx + gti[2] * gy;
        lpy = gti[1] * gx + gti[3] * gy;
      } else {
        lpx = (lrect[2] - lrect[0]) * ax[k] + lrect[0];
        lpy = (lrect[3] - lrect[1]) * ay[k] + lrect[1];
      }
    }
    lpx = wr_clamp(lpx, ph.lcr[0], ph.lcr[2]);
    lpy = wr_clamp(lpy, ph.lcr[1], ph.lcr[3]);
    float4 world = wr_mat_mul(transform.m, make_float4(lpx, lpy, 0.0f, 1.0f));
    float dpx = world.x * task.device_pixel_scale, dpy = world.y * task.device_pixel_scale;
    q.pos[k] = wr_mat_mul(a.tgt.proj, make_float4(dpx + fox * world.w, dpy + foy * world.w, ph.z * world.w, world.w));
    float fx = (lpx - gr[0]) / (gr[2] - gr[0]), fy = (lpy - gr[1]) / (gr[3] - gr[1]);
    if (glyph_transform) {
      fx = ((gt[0] * lpx + gt[2] * lpy) - gr[0]) / (gr[2] - gr[0]);
      fy = ((gt[1] * lpx + gt[3] * lpy) - gr[1]) / (gr[3] - gr[1]);
      q.interp[k][2] = fx;  // gl_ClipDistance[0..3]
      q.interp[k][3] = fy;
      q.interp[k][4] = 1.0f - fx;
      q.interp[k][5] = 1.0f - fy;
    }
    q.interp[k][0] = (st1x - st0x) * fx + st0x;
    q.interp[k][1] = (st1y - st0y) * fy + st0y;
  }
  q.n_interp = glyph_transform ? 6 : 2;
  q.flags = CMD_TEXTURED | (glyph_transform ? CMD_CLIP_DIST : 0u);
  wr_write_clip(T, clip_address, task, q);
  bool dual = (a.features & WRCU_FEAT_DUAL_SOURCE_BLENDING) != 0;
  float vcolor[4] = {1.0f, 1.0f, 1.0f, 1.0f}, swz[3] = {0.0f, 0.0f, 0.0f};
  uint16_t bc[4] = {0, 0, 0, 0};
  float tc[4] = {text_color.x, text_color.y, text_color.z, text_color.w};
  switch (color_mode) {
    case 0: swz[1] = 1.0f; swz[2] = 1.0f; for (int i = 0; i < 4; i++) vcolor[i] = tc[i]; break;
    case 2:
    case 1:
      if (a.blend_enabled) q.flags |= (color_mode == 2 ? CMD_DROP_SHADOW : CMD_SUBPIXEL_TEXT);
      bc[0] = (uint16_t)wr_round_pixel(tc[2], 255.0f); bc[1] = (uint16_t)wr_round_pixel(tc[1], 255.0f);
      bc[2] = (uint16_t)wr_round_pixel(tc[0], 255.0f); bc[3] = (uint16_t)wr_round_pixel(tc[3], 255.0f);
      swz[0] = 1.0f;
      break;
    case 3: swz[0] = 1.0f; for (int i = 0; i < 4; i++) vcolor[i] = tc[3]; break;
    default: break;
  }
  float packc[4] = {vcolor[0], vcolor[1], vcolor[2], vcolor[3]};
  if (dual) packc[0] = packc[1] = packc[2] = packc[3] = 1.0f;  // span commit without colour
  wr_pack_color(q, packc);
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    k->f[0] = (r0.x + 0.5f) / tw; k->f[1] = (r0.y + 0.5f) / th;
    k->f[2] = (r0.z + -0.5f) / tw; k->f[3] = (r0.w + -0.5f) / th;
    for (int i = 0; i < 4; i++) k->g[i] = vcolor[i];
    k->g[4] = swz[0]; k->g[5] = swz[1]; k->g[6] = swz[2];
    k->g[7] = dual ? 1.0f : 0.0f;
    k->i[0] = (int)bc[0] | ((int)bc[1] << 16);
    k->i[1] = (int)bc[2] | ((int)bc[3] << 16);
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_text_run)

// brush_linear_gradient (brush_linear_gradient.glsl:18-70, gradient_shared.glsl:20-60)
WRD void wr_setup_brush_linear_gradient_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  QuadOut q;
  BrushVS vs;
  memset(&q, 0, sizeof q);
  wr_brush_vertex(a, aData, 2, q, vs);
  const FrameTablesDev& T = a.tabs;
  float4 g0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.ph.specific_prim_address);
  float4 g1 = wr_fetch(T.gpu_cache, T.n_gpu_cache, vs.ph.specific_prim_address + 1);
  int extend_mode = (int)g1.x;
  float stretch[2] = {g1.y, g1.z};
  const float* lr = vs.ph.lr;
  const float* sr = vs.segment_rect;
  for (int k = 0; k < 4; k++) {
    float vx, vy;
    if (vs.brush_flags & 2) {  // SEGMENT_RELATIVE
      vx = (vs.local_pos[k].x - sr[0]) / (sr[2] - sr[0]);
      vy = (vs.local_pos[k].y - sr[1]) / (sr[3] - sr[1]);
      vx = vx * (vs.segment_data.z - vs.segment_data.x) + vs.segment_data.x;
      vy = vy * (vs.segment_data.w - vs.segment_data.y) + vs.segment_data.y;
      vx = vx * (lr[2] - lr[0]);
      vy = vy * (lr[3] - lr[1]);
    } else {
      vx = vs.local_pos[k].x - lr[0];
      vy = vs.local_pos[k].y - lr[1];
    }
    q.interp[k][0] = vx / stretch[0];
    q.interp[k][1] = vy / stretch[1];
  }
  q.n_interp = 2;
  float dirx = g0.z - g0.x, diry = g0.w - g0.y;
  float dd = dirx * dirx + diry * diry;
  float sdx = dirx / dd, sdy = diry / dd;
  int address = vs.ph.user_data[0];
  uint32_t merge[5];
  bool valid = wr_grad_validate_merge(T, address, merge);
  float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  wr_pack_color(q, white);
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    k->f[2] = g0.x * sdx + g0.y * sdy;
    k->f[0] = sdx * stretch[0];
    k->f[1] = sdy * stretch[1];
    k->f[3] = (float)(extend_mode == 1);
    k->i[0] = address;
    k->i[1] = valid ? 1 : 0;
    k->i[2] = 0;  // tileRepeat on
    for (int i = 0; i < 5; i++) k->g[i] = __uint_as_float(merge[i]);
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_brush_linear_gradient)
