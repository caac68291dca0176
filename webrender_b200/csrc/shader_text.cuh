// shader_text.cuh — ps_text_run (webrender/res/ps_text_run.glsl:278-337): glyph
// blit from the glyph atlas (R8 coverage expanded to RGBA8, or RGBA8 colour /
// subpixel glyphs) modulated by the text colour.
// CmdCold: f[0..3] v_uv_bounds; g[0..3] v_color, g[4..6] v_mask_swizzle, g[7] DUAL variant
#pragma once
#include "raster.cuh"

struct TextShader {
  struct Row {
    float o[2], step[2];
    TexRow tr;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    int len = c.x1 - c.x0;
    bool span_ok = k.g[4] == 0.0f || k.g[4] == 1.0f;  // swgl_drawSpanRGBA8 guard on v_mask_swizzle.x
    int body_len = (rgba && len >= 4 && span_ok) ? (len & ~3) : 0;
    float u[4], v[4];
    for (int j = 0; j < 4; j++) {
      float uv[2];
      wr_interp_at<2>(a, r.o, r.step, j, uv);
      u[j] = uv[0];
      v[j] = uv[1];
    }
    int tile_rel = max(tx0, (int)c.x0) - (int)c.x0;
    if (k.g[7] == 0.0f && a.color0.fmt == WRCU_FMT_R8) wr_tex_row_setup_r8(a.color0, k.f, body_len, u, v, tile_rel, r.tr);
    else wr_tex_row_setup(a.color0, k.f, false, body_len, u, v, tile_rel, r.tr);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    int rel = x - c.x0;
    if (rel < r.tr.body_len) {
      Px col{c.col[0], c.col[1], c.col[2], c.col[3]};  // 255 lanes for the DUAL variant (NoColor)
      return px_apply_color(wr_tex_body(t, r.tr, rel), col);
    }
    float uv[2];
    wr_interp_at<2>(a, r.o, r.step, rel, uv);
    float mask[4], col[4];
    wr_tex_fragment(t, wr_clamp(uv[0], k.f[0], k.f[2]), wr_clamp(uv[1], k.f[1], k.f[3]), mask);
    if (k.g[6] != 0.0f) mask[1] = mask[2] = mask[3] = mask[0];
    if (k.g[7] == 0.0f) {
      float al = mask[3];
      for (int i = 0; i < 3; i++) mask[i] = mask[i] * k.g[4] + al * k.g[5];
    }
    for (int i = 0; i < 4; i++) col[i] = k.g[i] * mask[i] * 1.0f;
    Px o;
    o.r = wr_round_pixel(col[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(col[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(col[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(col[3], 255.0f) & 0xFFFF;
    return o;
  }
};

#ifndef WRCU_HOSTEMU
template <> struct WrMinCtas<TextShader> { enum { v = 3 }; };
template <> struct WrNarrowSpans<TextShader> { enum { v = 1 }; };  // glyph rows: ~12 pixels of a 128-pixel warp row
#endif

template <> struct WrRun<TextShader> {
  enum { n = 2 };
  WRD_MEMBER int drawn(const TextShader::Row& r) { return r.tr.body_len; }
};
