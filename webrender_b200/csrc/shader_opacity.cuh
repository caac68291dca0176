// shader_opacity.cuh — brush_opacity [ALPHA_PASS] (webrender/res/
// brush_opacity.glsl).  Span body: swgl_commitTextureLinearColorRGBA8 with the
// opacity in all four colour lanes; tail: brush_fs in float.
#pragma once
#include "raster.cuh"
#include "setup_brush.cuh"

// CmdCold: f[0..3] v_uv_sample_bounds, f[4] v_opacity, f[5] v_perspective, f[6] gl_FragCoord.w
struct OpacityShader {
  struct Row {
    float o[2], step[2];
    float pd;
    TexRow tr;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    r.pd = (1.0f - k.f[6]) * k.f[5] + k.f[6];
    int len = c.x1 - c.x0;
    int body_len = (rgba && len >= 4 && !a.persp) ? (len & ~3) : 0;  // (perspective rows have no span body)
    float u[4], v[4];
    for (int j = 0; j < 4; j++) {
      float uv[2];
      wr_interp_at<2>(a, r.o, r.step, j, uv);
      u[j] = uv[0] * r.pd;
      v[j] = uv[1] * r.pd;
    }
    wr_tex_row_setup(a.color0, k.f, false, body_len, u, v, max(tx0, (int)c.x0) - (int)c.x0, r.tr);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    int rel = x - c.x0;
    if (rel < r.tr.body_len) {
      Px col{c.col[0], c.col[1], c.col[2], c.col[3]};
      return px_apply_color(wr_tex_body(t, r.tr, rel), col);
    }
    float uv[2];
    wr_interp_at<2>(a, r.o, r.step, rel, uv);
    float texel[4];
    float pd = r.pd;
    if (a.persp) {  // gl_FragCoord.w varies per sample
      const float fw = wr_persp_zw(*a.persp, 1, rel);
      pd = (1.0f - fw) * k.f[5] + fw;
    }
    wr_tex_fragment(t, wr_clamp(uv[0] * pd, k.f[0], k.f[2]), wr_clamp(uv[1] * pd, k.f[1], k.f[3]), texel);
    Px o;
    o.r = wr_round_pixel(k.f[4] * texel[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(k.f[4] * texel[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(k.f[4] * texel[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(k.f[4] * texel[3], 255.0f) & 0xFFFF;
    return o;
  }
};

// get_image_quad_uv (prim_shared.glsl:204-210)
WRD void wr_image_quad_uv(const FrameTablesDev& T, int address, float& fx, float& fy) {
  float4 tl = wr_fetch(T.gpu_cache, T.n_gpu_cache, address + 2);
  float4 tr = wr_fetch(T.gpu_cache, T.n_gpu_cache, address + 3);
  float4 bl = wr_fetch(T.gpu_cache, T.n_gpu_cache, address + 4);
  float4 br = wr_fetch(T.gpu_cache, T.n_gpu_cache, address + 5);
  float Xx = (tr.x - tl.x) * fx + tl.x, Xy = (tr.y - tl.y) * fx + tl.y, Xw = (tr.w - tl.w) * fx + tl.w;
  float Yx = (br.x - bl.x) * fx + bl.x, Yy = (br.y - bl.y) * fx + bl.y, Yw = (br.w - bl.w) * fx + bl.w;
  float Zx = (Yx - Xx) * fy + Xx, Zy = (Yy - Xy) * fy + Xy, Zw = (Yw - Xw) * fy + Xw;
  fx = Zx / Zw;
  fy = Zy / Zw;
}

// brush_opacity vertex stage (brush_opacity.glsl:23-52)
WRD void wr_setup_brush_opacity_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  QuadOut q;
  BrushVS vs;
  memset(&q, 0, sizeof q);
  wr_brush_vertex(a, aData, 3, q, vs);
  const FrameTablesDev& T = a.tabs;
  int src = vs.ph.user_data[0];
  float4 r0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, src);
  float tw = (float)a.color0.w, th = (float)a.color0.h;
  const float* lr = vs.ph.lr;
  float persp = (vs.brush_flags & 1) ? 1.0f : 0.0f;
  for (int k = 0; k < 4; k++) {
    float fx = (vs.local_pos[k].x - lr[0]) / (lr[2] - lr[0]);
    float fy = (vs.local_pos[k].y - lr[1]) / (lr[3] - lr[1]);
    wr_image_quad_uv(T, src, fx, fy);
    float ux = (r0.z - r0.x) * fx + r0.x, uy = (r0.w - r0.y) * fy + r0.y;
    float m = (1.0f - vs.world_pos[k].w) * persp + vs.world_pos[k].w;
    q.interp[k][0] = ux / tw * m;
    q.interp[k][1] = uy / th * m;
  }
  q.n_interp = 2;
  float opacity = wr_clamp((float)vs.ph.user_data[1] / 65536.0f, 0.0f, 1.0f);
  float oc[4] = {opacity, opacity, opacity, opacity};
  wr_pack_color(q, oc);
  q.flags |= CMD_TEXTURED;
  float fw = 1.0f / q.pos[0].w;
  if (!isfinite(fw)) fw = 0.0f;
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    k->f[0] = (r0.x + 0.5f) / tw; k->f[1] = (r0.y + 0.5f) / th;
    k->f[2] = (r0.z - 0.5f) / tw; k->f[3] = (r0.w - 0.5f) / th;
    k->f[4] = opacity;
    k->f[5] = persp;
    k->f[6] = fw;
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_brush_opacity)

template <> struct WrRun<OpacityShader> {
  enum { n = 2 };
  WRD_MEMBER int drawn(const OpacityShader::Row& r) { return r.tr.body_len; }
};
