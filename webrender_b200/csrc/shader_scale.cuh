// shader_scale.cuh — cs_scale (webrender/res/cs_scale.glsl): scaled copy of a
// source rect.  RGBA8 span body: swgl_commitTextureLinearRGBA8; otherwise main().
#pragma once
#include "raster.cuh"
#include "setup_common.cuh"

// CmdCold: f[0..3] vUvRect
struct ScaleShader {
  struct Row {
    float o[2], step[2];
    TexRow tr;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    int len = c.x1 - c.x0;
    int body_len = (rgba && len >= 4) ? (len & ~3) : 0;
    float u[4], v[4];
    for (int j = 0; j < 4; j++) {
      float uv[2];
      wr_interp_at<2>(a, r.o, r.step, j, uv);
      u[j] = uv[0];
      v[j] = uv[1];
    }
    wr_tex_row_setup(a.color0, k.f, false, body_len, u, v, max(tx0, (int)c.x0) - (int)c.x0, r.tr);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    int rel = x - c.x0;
    if (rel < r.tr.body_len) return wr_tex_body(t, r.tr, rel);
    float uv[2];
    wr_interp_at<2>(a, r.o, r.step, rel, uv);
    float col[4];
    wr_tex_fragment(t, wr_clamp(uv[0], k.f[0], k.f[2]), wr_clamp(uv[1], k.f[1], k.f[3]), col);
    Px o;
    o.r = wr_round_pixel(col[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(col[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(col[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(col[3], 255.0f) & 0xFFFF;
    return o;
  }
};

// cs_scale vertex stage (cs_scale.glsl:24-53)
WRD void wr_setup_scale_one(const SetupArgs& a, int idx) {
  const float* f = (const float*)(a.instances + (size_t)idx * a.stride);
  const float* trg = f;
  const float* src = f + 4;
  bool unnorm = (int)f[8] == 1;
  float r[4] = {wr_min(src[0], src[2]), wr_min(src[1], src[3]), wr_max(src[0], src[2]), wr_max(src[1], src[3])};
  float tw = (float)a.color0.w, th = (float)a.color0.h;
  if (unnorm) {
    r[0] += 0.5f; r[1] += 0.5f; r[2] -= 0.5f; r[3] -= 0.5f;
    r[0] /= tw; r[1] /= th; r[2] /= tw; r[3] /= th;
  }
  QuadOut q;
  memset(&q, 0, sizeof q);
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int k = 0; k < 4; k++) {
    float px = (trg[2] - trg[0]) * ax[k] + trg[0], py = (trg[3] - trg[1]) * ay[k] + trg[1];
    q.pos[k] = wr_mat_mul(a.tgt.proj, make_float4(px, py, 0.0f, 1.0f));
    float ux = src[0] + (src[2] - src[0]) * ax[k], uy = src[1] + (src[3] - src[1]) * ay[k];
    if (unnorm) { ux /= tw; uy /= th; }
    q.interp[k][0] = ux;
    q.interp[k][1] = uy;
  }
  q.n_interp = 2;
  q.flags = CMD_TEXTURED;
  q.col[0] = q.col[1] = q.col[2] = q.col[3] = 255;
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 4; i++) k->f[i] = r[i];
  }
  if (unsupported) {
    atomicAdd(&a.info->unsupported, 1);
    atomicAdd(a.err_counter, 1);
  }
}
WR_SETUP_KERNEL(wr_setup_scale)
