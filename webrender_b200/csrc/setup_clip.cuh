// setup_clip.cuh — vertex stage of the clip-mask programs
// (webrender/res/clip_shared.glsl:43-78 write_clip_tile_vertex,
//  transform.glsl:48-86 get_node_pos, cs_clip_rectangle.glsl:81-153 main).
#pragma once
#include "setup_common.cuh"
#include "shader_clip_rect.cuh"

WRD float4 wr_get_node_pos(float px, float py, const DevTransform& t) {
  float4 ah = wr_mat_mul(t.m, make_float4(0.0f, 0.0f, 0.0f, 1.0f));
  float ax = ah.x / ah.w, ay = ah.y / ah.w, az = ah.z / ah.w;
  float nx = t.inv_m[0] * 0.0f + t.inv_m[1] * 0.0f + t.inv_m[2] * 1.0f;
  float ny = t.inv_m[4] * 0.0f + t.inv_m[5] * 0.0f + t.inv_m[6] * 1.0f;
  float nz = t.inv_m[8] * 0.0f + t.inv_m[9] * 0.0f + t.inv_m[10] * 1.0f;
  float pz = -10000.0f;
  float tt = 0.0f;
  float denom = nx * 0.0f + ny * 0.0f + nz * 1.0f;
  if (fabsf(denom) > 1e-6f) {
    float dx = ax - px, dy = ay - py, dz = az - pz;
    tt = (dx * nx + dy * ny + dz * nz) / denom;
  }
  float z = pz + 1.0f * tt;
  return wr_mat_mul(t.inv_m, make_float4(px, py, z, 1.0f));
}

// common: ClipMaskInstanceCommon at the start of every clip instance
WRD void wr_clip_tile_vertex(const SetupArgs& a, const float* f, QuadOut& q, float4 local_out[4]) {
  const int* ids = (const int*)(f + 9);
  DevTransform clip_transform = wr_fetch_transform(a.tabs, ids[0]);
  DevTransform prim_transform = wr_fetch_transform(a.tabs, ids[1]);
  float dps = f[8];
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int i = 0; i < 4; i++) {
    float mxp = (f[2] - f[0]) * ax[i] + f[0], myp = (f[3] - f[1]) * ay[i] + f[1];
    float dpx = f[6] + mxp, dpy = f[7] + myp;
    float wx = dpx / dps, wy = dpy / dps;
    float4 pos = wr_mat_mul(prim_transform.m, make_float4(wx, wy, 0.0f, 1.0f));
    pos.x /= pos.w;
    pos.y /= pos.w;
    pos.z /= pos.w;
    float4 p = wr_get_node_pos(pos.x, pos.y, clip_transform);
    local_out[i] = make_float4(p.x * pos.w, p.y * pos.w, p.z * pos.w, p.w * pos.w);
    q.pos[i] = wr_mat_mul(a.tgt.proj, make_float4(f[4] + mxp, f[5] + myp, 0.0f, 1.0f));
  }
}

WRD void wr_inverse_radii_squared(const float* r, float* out) {
  out[0] = 1.0f / wr_max(r[0] * r[0], 1.0e-6f);
  out[1] = 1.0f / wr_max(r[1] * r[1], 1.0e-6f);
}

WRD void wr_setup_clip_rectangle_one(const SetupArgs& a, int idx) {
  const float* f = (const float*)(a.instances + (size_t)idx * a.stride);
  bool fast = (a.features & WRCU_FEAT_FAST_PATH) != 0;
  QuadOut q;
  memset(&q, 0, sizeof q);
  float4 lp[4];
  wr_clip_tile_vertex(a, f, q, lp);
  float clx = f[11], cly = f[12];
  float lr[4] = {f[13], f[14], f[15], f[16]};
  float mode = f[17];
  const float* corner[4] = {f + 18, f + 26, f + 34, f + 42};  // TL, TR, BL, BR
  float diffx = clx - lr[0], diffy = cly - lr[1];
  lr[0] = clx; lr[1] = cly; lr[2] += diffx; lr[3] += diffy;
  float g[40];
  for (int i = 0; i < 40; i++) g[i] = 0.0f;
  g[CR_MODE] = mode;
  g[CR_FAST] = fast ? 1.0f : 0.0f;
  g[CR_BOUNDS] = lr[0]; g[CR_BOUNDS + 1] = lr[1]; g[CR_BOUNDS + 2] = lr[2]; g[CR_BOUNDS + 3] = lr[3];
  for (int k = 0; k < 4; k++) {
    q.interp[k][0] = lp[k].x; q.interp[k][1] = lp[k].y; q.interp[k][2] = lp[k].z; q.interp[k][3] = lp[k].w;
  }
  q.n_interp = 4;
  if (fast) {
    float hx = 0.5f * (lr[2] - lr[0]), hy = 0.5f * (lr[3] - lr[1]);
    float radius = corner[0][4];
    for (int k = 0; k < 4; k++) {
      q.interp[k][0] -= (hx + clx) * lp[k].w;
      q.interp[k][1] -= (hy + cly) * lp[k].w;
    }
    g[CR_PARAMS] = hx - radius; g[CR_PARAMS + 1] = hy - radius; g[CR_PARAMS + 2] = radius;
  } else {
    const float* r_tl = corner[0] + 4; const float* r_tr = corner[1] + 4;
    const float* r_bl = corner[2] + 4; const float* r_br = corner[3] + 4;
    float* cr = g + CR_CORNER;  // TL, TR, BR, BL
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
WR_SETUP_KERNEL(wr_setup_clip_rectangle)

// cs_clip_box_shadow vertex stage (cs_clip_box_shadow.glsl:59-124)
WRD void wr_setup_clip_box_shadow_one(const SetupArgs& a, int idx) {
  const float* f = (const float*)(a.instances + (size_t)idx * a.stride);
  const int* iv = (const int*)f;
  const uint16_t* ra = (const uint16_t*)(f + 11);
  QuadOut q;
  memset(&q, 0, sizeof q);
  float4 lp[4];
  wr_clip_tile_vertex(a, f, q, lp);
  float src_w = f[12], src_h = f[13];
  int clip_mode = iv[14], smx = iv[15], smy = iv[16];
  float dest[4] = {f[17], f[18], f[19], f[20]};
  // fetch_image_source_direct: the address is a texel (x, y) of the gpu cache
  float4 res0 = wr_fetch(a.tabs.gpu_cache, a.tabs.n_gpu_cache, (int)ra[1] * 1024 + (int)ra[0]);
  float tw = (float)a.color0.w, th = (float)a.color0.h;
  float dsx = dest[2] - dest[0], dsy = dest[3] - dest[1];
  for (int k = 0; k < 4; k++) {
    float lpx = lp[k].x / lp[k].w, lpy = lp[k].y / lp[k].w;
    float ux = (smx == 0) ? (lpx - dest[0]) / src_w : (lpx - dest[0]) / dsx;
    float uy = (smy == 0) ? (lpy - dest[1]) / src_h : (lpy - dest[1]) / dsy;
    q.interp[k][0] = lp[k].x; q.interp[k][1] = lp[k].y; q.interp[k][2] = lp[k].z; q.interp[k][3] = lp[k].w;
    q.interp[k][4] = ux * lp[k].w;
    q.interp[k][5] = uy * lp[k].w;
  }
  q.n_interp = 6;
  float edge[4];
  if (smx == 0) { edge[0] = 0.5f; edge[2] = (dsx / src_w) - 0.5f; } else { edge[0] = 1.0f; edge[2] = 1.0f; }
  if (smy == 0) { edge[1] = 0.5f; edge[3] = (dsy / src_h) - 0.5f; } else { edge[1] = 1.0f; edge[3] = 1.0f; }
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 4; i++) k->f[i] = edge[i];
    k->f[4] = (res0.x + 0.5f) / tw; k->f[5] = (res0.y + 0.5f) / th;
    k->f[6] = (res0.z - 0.5f) / tw; k->f[7] = (res0.w - 0.5f) / th;
    k->g[0] = res0.x / tw; k->g[1] = res0.y / th; k->g[2] = res0.z / tw; k->g[3] = res0.w / th;
    for (int i = 0; i < 4; i++) k->g[4 + i] = dest[i];
    k->g[8] = (float)clip_mode;
  }
  if (unsupported) {
    atomicAdd(&a.info->unsupported, 1);
    atomicAdd(a.err_counter, 1);
  }
}
WR_SETUP_KERNEL(wr_setup_clip_box_shadow)
