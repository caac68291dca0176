// shader_composite.cuh — composite [FAST_PATH] (webrender/res/composite.glsl
// :73-234, RGB path): picture-cache tiles and external RGB surfaces into the
// framebuffer.  Span body: swgl_commitTexture[Color]RGBA8; tail: main().
#pragma once
#include "raster.cuh"
#include "setup_common.cuh"

// CmdCold: f[0..3] uv bounds used by the span shader, g[0..3] vColor,
//          g[4] FAST_PATH, g[5..8] vUVBounds (fragment clamp),
//          g[12..19] the instance's sColor0 view (tile lists carry one texture per instance)
WRD const TexView& wr_composite_tex(const CmdCold& k) { return *(const TexView*)&k.g[12]; }
static_assert(sizeof(TexView) <= 32, "TexView must fit CmdCold::g[12..19]");
struct CompositeShader {
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
      wr_interp_at<2>(r.o, r.step, j, uv);
      u[j] = uv[0];
      v[j] = uv[1];
    }
    wr_tex_row_setup(wr_composite_tex(k), k.f, true, body_len, u, v, max(tx0, (int)c.x0) - (int)c.x0, r.tr);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = wr_composite_tex(k);
    int rel = x - c.x0;
    if (rel < r.tr.body_len) {
      Px col{c.col[0], c.col[1], c.col[2], c.col[3]};
      return px_apply_color(wr_tex_body(t, r.tr, rel), col);
    }
    float uv[2];
    wr_interp_at<2>(r.o, r.step, rel, uv);
    bool fast = k.g[4] != 0.0f;
    float cu = uv[0], cv = uv[1];
    if (!fast) {
      cu = wr_clamp(cu, k.g[5], k.g[7]);
      cv = wr_clamp(cv, k.g[6], k.g[8]);
    }
    float texel[4], col[4];
    wr_tex_fragment(t, cu, cv, texel);
    for (int i = 0; i < 4; i++) col[i] = fast ? texel[i] : k.g[i] * texel[i];
    Px o;
    o.r = wr_round_pixel(col[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(col[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(col[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(col[3], 255.0f) & 0xFFFF;
    return o;
  }
};

// composite vertex stage (composite.glsl:73-159)
WRD void wr_setup_composite_one(const SetupArgs& a, int idx) {
  const float* f = (const float*)(a.instances + (size_t)idx * a.stride);
  const float* dr = f;
  const float* cr = f + 4;
  const float* uvr = f + 16;
  float flipx = f[28], flipy = f[29];
  QuadOut q;
  memset(&q, 0, sizeof q);
  float rect[4] = {(dr[2] - dr[0]) * flipx + dr[0], (dr[3] - dr[1]) * flipy + dr[1],
                   (dr[0] - dr[2]) * flipx + dr[2], (dr[1] - dr[3]) * flipy + dr[3]};
  bool fast = (a.features & WRCU_FEAT_FAST_PATH) != 0;
  float ub[4] = {wr_min(uvr[0], uvr[2]), wr_min(uvr[1], uvr[3]), wr_max(uvr[0], uvr[2]), wr_max(uvr[1], uvr[3])};
  bool unnorm = (int)f[13] == 1;
  const TexView tex0 = a.tex_list ? a.tex_list[idx] : a.color0;
  float tw = (float)tex0.w, th = (float)tex0.h;
  if (unnorm) {
    ub[0] += 0.5f; ub[1] += 0.5f; ub[2] += -0.5f; ub[3] += -0.5f;
    ub[0] /= tw; ub[1] /= th; ub[2] /= tw; ub[3] /= th;
  }
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int k = 0; k < 4; k++) {
    float wx = (rect[2] - rect[0]) * ax[k] + rect[0], wy = (rect[3] - rect[1]) * ay[k] + rect[1];
    float cx = wr_clamp(wx, cr[0], cr[2]), cy = wr_clamp(wy, cr[1], cr[3]);
    float ux = (cx - rect[0]) / (rect[2] - rect[0]), uy = (cy - rect[1]) / (rect[3] - rect[1]);
    ux = (uvr[2] - uvr[0]) * ux + uvr[0];
    uy = (uvr[3] - uvr[1]) * uy + uvr[1];
    if (unnorm) { ux /= tw; uy /= th; }
    q.interp[k][0] = ux;
    q.interp[k][1] = uy;
    q.pos[k] = wr_mat_mul(a.tgt.proj, make_float4(cx, cy, 0.0f, 1.0f));
  }
  q.n_interp = 2;
  q.flags = CMD_TEXTURED;
  float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  const float* color = fast ? white : f + 8;
  // swgl_drawSpanRGBA8: colour modulation only when color != vec4(1.0)
  bool is_white = color[0] == 1.0f && color[1] == 1.0f && color[2] == 1.0f && color[3] == 1.0f;
  if (is_white) { q.col[0] = q.col[1] = q.col[2] = q.col[3] = 255; }
  else {
    q.col[0] = (uint16_t)wr_round_pixel(color[2], 255.0f); q.col[1] = (uint16_t)wr_round_pixel(color[1], 255.0f);
    q.col[2] = (uint16_t)wr_round_pixel(color[0], 255.0f); q.col[3] = (uint16_t)wr_round_pixel(color[3], 255.0f);
  }
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    if (fast) { k->f[0] = 0.0f; k->f[1] = 0.0f; k->f[2] = 1.0f; k->f[3] = 1.0f; }
    else for (int i = 0; i < 4; i++) k->f[i] = ub[i];
    for (int i = 0; i < 4; i++) { k->g[i] = f[8 + i]; k->g[5 + i] = ub[i]; }
    k->g[4] = fast ? 1.0f : 0.0f;
    *(TexView*)&k->g[12] = tex0;
  }
  if (unsupported) {
    atomicAdd(&a.info->unsupported, 1);
    atomicAdd(a.err_counter, 1);
  }
}
WR_SETUP_KERNEL(wr_setup_composite)

#ifndef WRCU_HOSTEMU
template <> struct WrMinCtas<CompositeShader> { enum { v = 3 }; };
#endif
