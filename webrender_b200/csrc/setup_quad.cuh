// setup_quad.cuh — vertex stage of the ps_quad_* programs
// (webrender/res/ps_quad.glsl:239-389 quad_primive_info/write_vertex/main,
//  ps_quad_textured.glsl:15-36 pattern_vertex, sample_color0.glsl:12-21),
// SWGL branches (SWGL_ANTIALIAS: edge AA is requested from the rasteriser via
// swgl_antiAlias instead of AA varyings).  One thread per instance.
#pragma once
#include "setup_common.cuh"
#include "setup_clip.cuh"

#define WR_QF_IS_OPAQUE 1
#define WR_QF_APPLY_DEVICE_CLIP 2
#define WR_QF_IGNORE_DEVICE_SCALE 4
#define WR_QF_USE_AA_SEGMENTS 8
#define WR_QF_IS_MASK 16
#define WR_AA_PIXEL_RADIUS 2.0f

struct QuadPrimInfo {
  float2 local_pos[4];  // PrimitiveInfo.local_pos (after pattern scale/offset), per lane
  float seg_rect[4], seg_uv_rect[4];
  float prim_bounds[4], prim_clip[4];
  float color[4];
  int edge_flags, quad_flags;
  int pattern_input[2];
};

WRD void wr_quad_primitive_info(const SetupArgs& a, int4 aData, QuadOut& q,
                                              QuadPrimInfo& pi) {
  const FrameTablesDev& T = a.tabs;
  int prim_address_i = aData.x, prim_address_f = aData.y;
  int quad_flags = (aData.z >> 24) & 0xff, edge_flags = (aData.z >> 16) & 0xff;
  int part_index = (aData.z >> 8) & 0xff, segment_index = aData.z & 0xff;
  int picture_task_address = aData.w;
  int4 header = wr_fetchi(T.gpu_buffer_i, T.n_gpu_buffer_i, prim_address_i);
  int transform_id = header.x, z_id = header.y;
  pi.pattern_input[0] = header.z;
  pi.pattern_input[1] = header.w;
  DevTransform transform = wr_fetch_transform(T, transform_id);
  DevPictureTask task = wr_fetch_picture_task(T, picture_task_address);
  float4 t0 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, prim_address_f);
  float4 t1 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, prim_address_f + 1);
  float4 t2 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, prim_address_f + 2);
  float4 so = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, prim_address_f + 3);
  float4 t4 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, prim_address_f + 4);
  float z = (float)z_id;
  float4 seg_rect, seg_uv;
  if (segment_index == 0xff) {
    seg_rect = t0;
    seg_uv = t2;
  } else {
    seg_rect = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, prim_address_f + 5 + segment_index * 2);
    seg_uv = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, prim_address_f + 5 + segment_index * 2 + 1);
  }
  float c0x = wr_max(seg_rect.x, t1.x), c0y = wr_max(seg_rect.y, t1.y);
  float c1x = wr_min(seg_rect.z, t1.z), c1y = wr_min(seg_rect.w, t1.w);
  c1x = wr_max(c0x, c1x);
  c1y = wr_max(c0y, c1y);
  int aa_mask = 0;
  switch (part_index) {
    case 1: c1x = c0x + WR_AA_PIXEL_RADIUS; aa_mask = 1; break;
    case 2:
      c0x = c0x + WR_AA_PIXEL_RADIUS; c1x = c1x - WR_AA_PIXEL_RADIUS;
      c1y = c0y + WR_AA_PIXEL_RADIUS; aa_mask = 2; break;
    case 3: c0x = c1x - WR_AA_PIXEL_RADIUS; aa_mask = 4; break;
    case 4:
      c0x = c0x + WR_AA_PIXEL_RADIUS; c1x = c1x - WR_AA_PIXEL_RADIUS;
      c0y = c1y - WR_AA_PIXEL_RADIUS; aa_mask = 8; break;
    case 0:
      c0x += (edge_flags & 1) ? WR_AA_PIXEL_RADIUS : 0.0f;
      c1x -= (edge_flags & 4) ? WR_AA_PIXEL_RADIUS : 0.0f;
      c0y += (edge_flags & 2) ? WR_AA_PIXEL_RADIUS : 0.0f;
      c1y -= (edge_flags & 8) ? WR_AA_PIXEL_RADIUS : 0.0f;
      break;
    default: aa_mask = edge_flags; break;
  }
  q.aa_edge_mask = aa_mask;
  q.flags = aa_mask ? CMD_AA : 0;
  float dps = task.device_pixel_scale;
  if (quad_flags & WR_QF_IGNORE_DEVICE_SCALE) dps = 1.0f;
  float fox = -task.ox + task.tx0, foy = -task.oy + task.ty0;
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int i = 0; i < 4; i++) {
    float lpx = (c1x - c0x) * ax[i] + c0x, lpy = (c1y - c0y) * ay[i] + c0y;
    float4 world = wr_mat_mul(transform.m, make_float4(lpx, lpy, 0.0f, 1.0f));
    float dpx = world.x * dps, dpy = world.y * dps;
    float vix = lpx, viy = lpy;
    if (quad_flags & WR_QF_APPLY_DEVICE_CLIP) {
      float d1x = task.ox + task.tx1 - task.tx0, d1y = task.oy + task.ty1 - task.ty0;
      dpx = wr_clamp(dpx, task.ox, d1x);
      dpy = wr_clamp(dpy, task.oy, d1y);
      float4 r = wr_mat_mul(transform.inv_m, make_float4(dpx / dps, dpy / dps, 0.0f, 1.0f));
      vix = r.x;
      viy = r.y;
    }
    q.pos[i] = wr_mat_mul(a.tgt.proj, make_float4(dpx + fox * world.w, dpy + foy * world.w,
                                                  z * world.w, world.w));
    pi.local_pos[i] = make_float2(vix * so.x + so.z, viy * so.y + so.w);
  }
  pi.color[0] = t4.x; pi.color[1] = t4.y; pi.color[2] = t4.z; pi.color[3] = t4.w;
  pi.seg_rect[0] = seg_rect.x * so.x + so.z; pi.seg_rect[1] = seg_rect.y * so.y + so.w;
  pi.seg_rect[2] = seg_rect.z * so.x + so.z; pi.seg_rect[3] = seg_rect.w * so.y + so.w;
  pi.seg_uv_rect[0] = seg_uv.x; pi.seg_uv_rect[1] = seg_uv.y;
  pi.seg_uv_rect[2] = seg_uv.z; pi.seg_uv_rect[3] = seg_uv.w;
  pi.prim_bounds[0] = t0.x * so.x + so.z; pi.prim_bounds[1] = t0.y * so.y + so.w;
  pi.prim_bounds[2] = t0.z * so.x + so.z; pi.prim_bounds[3] = t0.w * so.y + so.w;
  pi.prim_clip[0] = t1.x * so.x + so.z; pi.prim_clip[1] = t1.y * so.y + so.w;
  pi.prim_clip[2] = t1.z * so.x + so.z; pi.prim_clip[3] = t1.w * so.y + so.w;
  pi.edge_flags = edge_flags;
  pi.quad_flags = quad_flags;
}

WRD void wr_setup_quad_textured_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  QuadOut q;
  QuadPrimInfo pi;
  memset(&q, 0, sizeof q);
  wr_quad_primitive_info(a, aData, q, pi);
  float col[4] = {pi.color[0], pi.color[1], pi.color[2], pi.color[3]};
  float bounds[4] = {0, 0, 0, 0};
  q.n_interp = 0;
  if (pi.quad_flags & WR_QF_IS_MASK) q.flags |= CMD_OUT_RRRR;
  if (pi.seg_uv_rect[0] != pi.seg_uv_rect[2] || pi.seg_uv_rect[1] != pi.seg_uv_rect[3]) {
    q.flags |= CMD_TEXTURED;
    col[0] = col[1] = col[2] = col[3] = 1.0f;
    float tw = (float)a.color0.w, th = (float)a.color0.h;
    for (int i = 0; i < 4; i++) {
      float fx = (pi.local_pos[i].x - pi.seg_rect[0]) / (pi.seg_rect[2] - pi.seg_rect[0]);
      float fy = (pi.local_pos[i].y - pi.seg_rect[1]) / (pi.seg_rect[3] - pi.seg_rect[1]);
      float ux = (pi.seg_uv_rect[2] - pi.seg_uv_rect[0]) * fx + pi.seg_uv_rect[0];
      float uy = (pi.seg_uv_rect[3] - pi.seg_uv_rect[1]) * fy + pi.seg_uv_rect[1];
      q.interp[i][0] = ux / tw;
      q.interp[i][1] = uy / th;
    }
    q.n_interp = 2;
    bounds[0] = (pi.seg_uv_rect[0] + 0.5f) / tw;
    bounds[1] = (pi.seg_uv_rect[1] + 0.5f) / th;
    bounds[2] = (pi.seg_uv_rect[2] - 0.5f) / tw;
    bounds[3] = (pi.seg_uv_rect[3] - 0.5f) / th;
  }
  // pack_span / packColor: round_pixel then 16-bit lanes (B,G,R,A order)
  q.col[0] = (uint16_t)wr_round_pixel(col[2], 255.0f);
  q.col[1] = (uint16_t)wr_round_pixel(col[1], 255.0f);
  q.col[2] = (uint16_t)wr_round_pixel(col[0], 255.0f);
  q.col[3] = (uint16_t)wr_round_pixel(col[3], 255.0f);
  // swgl_drawSpanRGBA8 → swgl_commitSolidRGBA8 for untextured quads on RGBA8 targets
  if (!(q.flags & CMD_TEXTURED) && a.tgt.fmt == WRCU_FMT_RGBA8) q.flags |= CMD_SPAN_SOLID;
  if (!(q.flags & CMD_TEXTURED)) q.flags |= CMD_CONST_COLOR;
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    k->f[0] = bounds[0]; k->f[1] = bounds[1]; k->f[2] = bounds[2]; k->f[3] = bounds[3];
  }
  if (unsupported) {
    atomicAdd(&a.info->unsupported, 1);
    atomicAdd(a.err_counter, 1);
  }
}
WR_SETUP_KERNEL(wr_setup_quad_textured)

// ps_quad_mask (ps_quad_mask.glsl:66-137): MaskInstance = quad instance +
// aClipData [clip_transform_id, clip_address, clip_space, _].
WRD void wr_setup_quad_mask_one(const SetupArgs& a, int idx) {
  const int* inst = (const int*)(a.instances + (size_t)idx * a.stride);
  int4 aData = make_int4(inst[0], inst[1], inst[2], inst[3]);
  int clip_transform_id = inst[4], index = inst[5], space = inst[6];
  bool fast = (a.features & WRCU_FEAT_FAST_PATH) != 0;
  QuadOut q;
  QuadPrimInfo pi;
  memset(&q, 0, sizeof q);
  wr_quad_primitive_info(a, aData, q, pi);
  const FrameTablesDev& T = a.tabs;
  float4 t0 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, index);
  float4 t1 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, index + 1);
  float4 t2 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, index + 2);
  float4 t3 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, index + 3);
  float lr[4] = {t0.x, t0.y, t0.z, t0.w};
  DevTransform ct = wr_fetch_transform(T, clip_transform_id);
  for (int k = 0; k < 4; k++) {
    float4 p = wr_mat_mul(ct.m, make_float4(pi.local_pos[k].x, pi.local_pos[k].y, 0.0f, 1.0f));
    q.interp[k][0] = p.x; q.interp[k][1] = p.y; q.interp[k][2] = p.z; q.interp[k][3] = p.w;
  }
  q.n_interp = 4;
  float g[40];
  for (int i = 0; i < 40; i++) g[i] = 0.0f;
  g[CR_FAST] = fast ? 1.0f : 0.0f;
  if (fast) {
    g[CR_MODE] = t2.x;
    float hx = 0.5f * (lr[2] - lr[0]), hy = 0.5f * (lr[3] - lr[1]);
    float radius = t1.x;
    for (int k = 0; k < 4; k++) {
      q.interp[k][0] -= (hx + lr[0]) * q.interp[k][3];
      q.interp[k][1] -= (hy + lr[1]) * q.interp[k][3];
    }
    g[CR_PARAMS] = hx - radius; g[CR_PARAMS + 1] = hy - radius; g[CR_PARAMS + 2] = radius;
  } else {
    g[CR_MODE] = t3.x;
    if (space == 0) {
      g[CR_BOUNDS] = lr[0]; g[CR_BOUNDS + 1] = lr[1]; g[CR_BOUNDS + 2] = lr[2]; g[CR_BOUNDS + 3] = lr[3];
    } else {
      g[CR_BOUNDS] = wr_max(lr[0], pi.prim_clip[0]); g[CR_BOUNDS + 1] = wr_max(lr[1], pi.prim_clip[1]);
      g[CR_BOUNDS + 2] = wr_min(lr[2], pi.prim_clip[2]); g[CR_BOUNDS + 3] = wr_min(lr[3], pi.prim_clip[3]);
    }
    float r_tl[2] = {t1.x, t1.y}, r_tr[2] = {t1.z, t1.w}, r_bl[2] = {t2.x, t2.y}, r_br[2] = {t2.z, t2.w};
    float* cr = g + CR_CORNER;
    cr[0] = lr[0] + r_tl[0]; cr[1] = lr[1] + r_tl[1]; wr_inverse_radii_squared(r_tl, cr + 2);
    cr[4] = lr[2] - r_tr[0]; cr[5] = lr[1] + r_tr[1]; wr_inverse_radii_squared(r_tr, cr + 6);
    cr[8] = lr[2] - r_br[0]; cr[9] = lr[3] - r_br[1]; wr_inverse_radii_squared(r_br, cr + 10);
    cr[12] = lr[0] + r_bl[0]; cr[13] = lr[3] - r_bl[1]; wr_inverse_radii_squared(r_bl, cr + 14);
    float* pl = g + CR_PLANE;
    float n_tl[2] = {-r_tl[1], -r_tl[0]}, n_tr[2] = {r_tr[1], -r_tr[0]};
    float n_br[2] = {r_br[1], r_br[0]}, n_bl[2] = {-r_bl[1], r_bl[0]};
    pl[0] = n_tl[0]; pl[1] = n_tl[1]; pl[2] = n_tl[0] * lr[0] + n_tl[1] * (lr[1] + r_tl[1]);
    pl[3] = n_tr[0]; pl[4] = n_tr[1]; pl[5] = n_tr[0] * (lr[2] - r_tr[0]) + n_tr[1] * lr[1];
    pl[6] = n_br[0]; pl[7] = n_br[1]; pl[8] = n_br[0] * lr[2] + n_br[1] * (lr[3] - r_br[1]);
    pl[9] = n_bl[0]; pl[10] = n_bl[1]; pl[11] = n_bl[0] * (lr[0] + r_bl[0]) + n_bl[1] * lr[3];
  }
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 40; i++) k->g[i] = g[i];
  }
  if (unsupported) {
    atomicAdd(&a.info->unsupported, 1);
    atomicAdd(a.err_counter, 1);
  }
}
WR_SETUP_KERNEL(wr_setup_quad_mask)

// ps_clear (ps_clear.glsl:13-18): quad-based clear, depth forced to the far plane
WRD void wr_setup_clear_one(const SetupArgs& a, int idx) {
  const float* f = (const float*)(a.instances + (size_t)idx * a.stride);
  QuadOut q;
  memset(&q, 0, sizeof q);
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int k = 0; k < 4; k++) {
    float px = (f[2] - f[0]) * ax[k] + f[0], py = (f[3] - f[1]) * ay[k] + f[1];
    q.pos[k] = wr_mat_mul(a.tgt.proj, make_float4(px, py, 0.0f, 1.0f));
    q.pos[k].z = q.pos[k].w;
  }
  q.n_interp = 0;
  q.col[0] = (uint16_t)wr_round_pixel(f[6], 255.0f); q.col[1] = (uint16_t)wr_round_pixel(f[5], 255.0f);
  q.col[2] = (uint16_t)wr_round_pixel(f[4], 255.0f); q.col[3] = (uint16_t)wr_round_pixel(f[7], 255.0f);
  q.flags = CMD_CONST_COLOR;
  int unsupported = 0;
  wr_emit_quad(a, idx, q, &unsupported);
  if (unsupported) {
    atomicAdd(&a.info->unsupported, 1);
    atomicAdd(a.err_counter, 1);
  }
}
WR_SETUP_KERNEL(wr_setup_clear)

// ps_quad_radial_gradient / ps_quad_conic_gradient pattern_vertex
// (ps_quad_radial_gradient.glsl:37-58, ps_quad_conic_gradient.glsl:46-60) on top of
// the quad vertex stage.  Cold layout: shader_cs_gradient.cuh (+ g[8..11] v_color, i[3] = 1).
WRD void wr_setup_quad_gradient_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  QuadOut q;
  QuadPrimInfo pi;
  memset(&q, 0, sizeof q);
  wr_quad_primitive_info(a, aData, q, pi);
  const FrameTablesDev& T = a.tabs;
  const bool radial = a.kind == WRCU_KIND_QUAD_RADIAL_GRADIENT;
  float4 d0 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, pi.pattern_input[0]);
  float4 d1 = wr_fetch(T.gpu_buffer_f, T.n_gpu_buffer_f, pi.pattern_input[0] + 1);
  const int address = pi.pattern_input[1];
  float dd = d1.y - d1.x;
  float rscale = dd != 0.0f ? 1.0f / dd : 0.0f;
  float fc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  fc[3] = d1.w;
  for (int v = 0; v < 4; v++) {
    float px = (pi.local_pos[v].x - pi.prim_bounds[0]) * d0.z - d0.x;
    float py = (pi.local_pos[v].y - pi.prim_bounds[1]) * d0.w - d0.y;
    if (radial) {
      px = px * rscale;
      py = py * rscale;
      py *= d1.z;
    }
    q.interp[v][0] = px;
    q.interp[v][1] = py;
  }
  q.n_interp = 2;
  if (radial) {
    fc[0] = d1.x * rscale;
  } else {
    fc[5] = rscale;
    fc[4] = 3.141592653589793f / 2.0f - d1.z;
    fc[2] = d1.x * rscale;
  }
  if (pi.quad_flags & WR_QF_IS_MASK) q.flags |= CMD_OUT_RRRR;
  uint32_t merge[5];
  bool valid = wr_grad_validate_merge(T, address, merge);
  float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  wr_pack_color(q, white);
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 8; i++) k->f[i] = fc[i];
    for (int i = 0; i < 5; i++) k->g[i] = __uint_as_float(merge[i]);
    for (int i = 0; i < 4; i++) k->g[8 + i] = pi.color[i];
    k->i[0] = address;
    k->i[1] = valid ? 1 : 0;
    k->i[2] = 1;
    k->i[3] = 1;
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_quad_gradient)
