// shader_image.cuh — brush_image [ALPHA_PASS] (non-REPETITION variants;
// webrender/res/brush_image.glsl:319-429).  Span body: swgl_commitTexture
// [Color]RGBA8 (filter chosen by the sampler); tail: brush_fs in float.
#pragma once
#include "raster.cuh"

// CmdCold for image brushes:
//  f[0..3] v_uv_sample_bounds, f[4..5] v_uv_bounds.xy, f[6] v_perspective.x, f[7] gl_FragCoord.w
//  g[0..3] v_color (float), g[4..5] v_mask_swizzle, g[6] alpha_pass, g[7] span shader allowed
struct ImageShader {
  struct Row {
    float o[2], step[2];
    float pd;
    TexRow tr;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    // mix(gl_FragCoord.w, 1.0, v_perspective.x) = (1 - w) * p + w
    r.pd = (1.0f - k.f[7]) * k.f[6] + k.f[7];
    int len = c.x1 - c.x0;
    int body_len = (rgba && len >= 4 && k.g[7] != 0.0f) ? (len & ~3) : 0;
    float u[4], v[4];
    for (int j = 0; j < 4; j++) {
      float uv[2];
      wr_interp_at<2>(r.o, r.step, j, uv);
      u[j] = uv[0] * r.pd + k.f[4];
      v[j] = uv[1] * r.pd + k.f[5];
    }
    wr_tex_row_setup(a.color0, k.f, true, body_len, u, v, max(tx0, (int)c.x0) - (int)c.x0, r.tr);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    int rel = x - c.x0;
    if (rel < r.tr.body_len) {
      // packed v_color (255 lanes == NoColor: muldiv255(255, s) == s)
      Px col{c.col[0], c.col[1], c.col[2], c.col[3]};
      return px_apply_color(wr_tex_body(t, r.tr, rel), col);
    }
    float uv[2];
    wr_interp_at<2>(r.o, r.step, rel, uv);
    float ru = uv[0] * r.pd + k.f[4], rv = uv[1] * r.pd + k.f[5];
    float texel[4], col[4];
    wr_tex_fragment(t, wr_clamp(ru, k.f[0], k.f[2]), wr_clamp(rv, k.f[1], k.f[3]), texel);
    if (k.g[6] != 0.0f) {
      float al = texel[3];
      for (int i = 0; i < 3; i++) texel[i] = texel[i] * k.g[4] + al * k.g[5];
      for (int i = 0; i < 4; i++) col[i] = k.g[i] * (texel[i] * 1.0f) * 1.0f;
    } else {
      for (int i = 0; i < 4; i++) col[i] = texel[i];
    }
    Px o;
    o.r = wr_round_pixel(col[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(col[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(col[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(col[3], 255.0f) & 0xFFFF;
    return o;
  }
};
