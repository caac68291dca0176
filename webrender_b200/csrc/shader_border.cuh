// shader_border.cuh — border and line-decoration render tasks drawn by
// draw_texture_cache_target (renderer/mod.rs:4015-4083):
//   cs_border_solid    (webrender/res/cs_border_solid.glsl)
//   cs_border_segment  (cs_border_segment.glsl)
//   cs_line_decoration (cs_line_decoration.glsl)
// Fragment shaders only: every pixel evaluates the float shader on the varying
// position, which the reference advances once per 4-pixel chunk (a running sum,
// replayed by wr_chunk_base / the per-pixel walk below).  compute_aa_range under
// SWGL is 1 / (|dFdx(pos.x)| + |dFdx(pos.y)|) taken from lanes 0 and 1 of the
// pixel's own chunk (glsl.h:765-768, shared.glsl:145-148).
//
// CmdCold layout (borders):
//   g[0..15] four colours, g[16..19] vColorLine, g[20..21] normalize(vColorLine.zw),
//   g[22..25] vClipCenter_Sign, g[26..29] vClipRadii
//   solid:   g[30..33] vHorizontalClipCenter_Sign, g[34..35] vHorizontalClipRadii,
//            f[0..3] vVerticalClipCenter_Sign, f[4..5] vVerticalClipRadii, i[0] vMixColors.x
//   segment: g[30..33] vEdgeReference, g[34..37] vPartialWidths, f[0..2] vClipParams1.xyz,
//            f[4..5] normalize(vClipParams1.zw), f[6..7] vClipParams2.xy,
//            g[38..39] normalize(vClipParams2.zw),
//            i[0] segment | clip_mode << 8 | style0 << 16 | style1 << 24, i[1] edge axes
// line decoration: f[0..3] vParams, i[0] vStyle.x
#pragma once
#include "raster.cuh"
#include "setup_common.cuh"

WRD float wr_hypotf(float x, float y) {
#ifdef WRCU_HOSTEMU
  return hypotf(x, y);
#else
  // libm hypotf: the exact double sum of squares, one sqrt, one rounding
  return (float)sqrt((double)x * (double)x + (double)y * (double)y);
#endif
}
WRD float wr_distance_aa(float aa_range, float sd) { return wr_clamp(0.5f - sd * aa_range, 0.0f, 1.0f); }
WRD float wr_mixf(float x, float y, float a) { return (y - x) * a + x; }
// ellipse.glsl:7-46
WRD float wr_distance_to_ellipse(float px, float py, float rx, float ry) {
  float ix = 1.0f / wr_max(rx * rx, 1.0e-6f), iy = 1.0f / wr_max(ry * ry, 1.0e-6f);
  float scale = (rx > 0.0f && ry > 0.0f) ? 1.0f : 0.0f;
  float prx = px * ix, pry = py * iy;
  float g = (px * prx + py * pry) - scale;
  float dgx = (1.0f + scale) * prx, dgy = (1.0f + scale) * pry;
  return g * (1.0f / sqrtf(dgx * dgx + dgy * dgy));
}
// distance_to_line (shared.glsl:110-113) with the flat direction normalised in setup
WRD float wr_distance_to_line_n(float p0x, float p0y, float nx, float ny, float px, float py) {
  float dx = p0x - px, dy = p0y - py;
  return nx * dx + ny * dy;
}
WRD float wr_glsl_mod(float a, float b) { return a - b * floorf(a / b); }
WRD float wr_glsl_step(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }

// Position of pixel `rel` and the AA range of its chunk.
struct PosRow {
  float o[2], step[2];
  float base[4][2];
  int kb;
};
WRD void wr_pos_row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, PosRow& r) {
  const CmdCold& k = a.cold[c.cold];
  wr_row_interp<2>(a, k, c, y, r.o, r.step);
  r.kb = wr_chunk_base<2>(a, r.o, r.step, c, tx0, r.base);
}
WRD void wr_pos_at(const PosRow& r, int rel, float* p, float* aa_range) {
  const int kc = rel >> 2, j = rel & 3;
  float l0[2], l1[2], lj[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float is = __fmul_rn(r.step[i], 4.0f);
    float v0 = r.base[0][i], v1 = r.base[1][i], vj = r.base[j][i];
    for (int s = r.kb; s < kc; s++) {
      v0 = __fadd_rn(v0, is);
      v1 = __fadd_rn(v1, is);
      vj = __fadd_rn(vj, is);
    }
    l0[i] = v0; l1[i] = v1; lj[i] = vj;
  }
  *aa_range = 1.0f / (fabsf(l1[0] - l0[0]) + fabsf(l1[1] - l0[1]));
  p[0] = lj[0];
  p[1] = lj[1];
}
WRD Px wr_pack_rgba(const float* col) {
  Px o;
  o.r = wr_round_pixel(col[0], 255.0f) & 0xFFFF;
  o.g = wr_round_pixel(col[1], 255.0f) & 0xFFFF;
  o.b = wr_round_pixel(col[2], 255.0f) & 0xFFFF;
  o.a = wr_round_pixel(col[3], 255.0f) & 0xFFFF;
  return o;
}

// ---- cs_line_decoration (cs_line_decoration.glsl:100-162) ----------------------------
struct LineDecorationShader {
  typedef PosRow Row;
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    wr_pos_row_setup(a, c, y, tx0, r);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    float p[2], aa_range;
    wr_pos_at(r, x - c.x0, p, &aa_range);
    float px = p[0], py = p[1];
    const float* P = k.f;
    float alpha = 1.0f;
    switch (k.i[0]) {
      case 2:
        alpha = wr_glsl_step(floorf(px + 0.5f), P[1]);
        break;
      case 1: {
        float rx = px - P[1], ry = py - P[2];
        alpha = wr_distance_aa(aa_range, sqrtf(rx * rx + ry * ry) - P[1]);
        break;
      }
      case 3: {
        float half_line_thickness = P[0], slope_length = P[1], flat_length = P[2], vertical_bounds = P[3];
        float half_period = slope_length + flat_length;
        float mid_height = vertical_bounds / 2.0f;
        float peak_offset = mid_height - half_line_thickness;
        float flip = -2.0f * (wr_glsl_step(wr_glsl_mod(px, 2.0f * half_period), half_period) - 0.5f);
        peak_offset *= flip;
        float peak_height = mid_height + peak_offset;
        px = wr_glsl_mod(px, half_period);
        float dist[3];
        const float p0x[3] = {0.0f, 0.0f, flat_length}, dirx[3] = {1.0f, 0.0f, -1.0f};
#pragma unroll
        for (int i = 0; i < 3; i++) {
          float nx = dirx[i], ny = -flip;
          float l = sqrtf(nx * nx + ny * ny);
          nx = nx / l;
          ny = ny / l;
          dist[i] = nx * (p0x[i] - px) + ny * (peak_height - py);
        }
        float dd = fabsf(wr_max(wr_max(dist[0], dist[1]), dist[2]));
        alpha = wr_distance_aa(aa_range, dd - half_line_thickness);
        if (half_line_thickness <= 1.0f) alpha = 1.0f - wr_glsl_step(alpha, 0.5f);
        break;
      }
      default:
        break;
    }
    float col[4] = {alpha, alpha, alpha, alpha};
    return wr_pack_rgba(col);
  }
};

// ---- cs_border_solid (cs_border_solid.glsl:137-177) -----------------------------------
struct BorderSolidShader {
  typedef PosRow Row;
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    wr_pos_row_setup(a, c, y, tx0, r);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const float* g = k.g;
    float p[2], aa_range;
    wr_pos_at(r, x - c.x0, p, &aa_range);
    const float px = p[0], py = p[1];
    const int mix_colors = k.i[0];
    const bool do_aa = mix_colors != 2;
    float mix_factor = 0.0f;
    if (mix_colors != 0) {
      float d_line = wr_distance_to_line_n(g[16], g[17], g[20], g[21], px, py);
      if (do_aa) mix_factor = wr_distance_aa(aa_range, -d_line);
      else mix_factor = d_line + 0.0001f >= 0.0f ? 1.0f : 0.0f;
    }
    float d = -1.0f;
    float cx = px - g[22], cy = py - g[23];
    if (g[24] * cx < 0.0f && g[25] * cy < 0.0f)
      d = wr_max(wr_distance_to_ellipse(cx, cy, g[26], g[27]), -wr_distance_to_ellipse(cx, cy, g[28], g[29]));
    cx = px - g[30];
    cy = py - g[31];
    if (g[32] * cx < 0.0f && g[33] * cy < 0.0f) d = wr_max(wr_distance_to_ellipse(cx, cy, g[34], g[35]), d);
    cx = px - k.f[0];
    cy = py - k.f[1];
    if (k.f[2] * cx < 0.0f && k.f[3] * cy < 0.0f) d = wr_max(wr_distance_to_ellipse(cx, cy, k.f[4], k.f[5]), d);
    float alpha = do_aa ? wr_distance_aa(aa_range, d) : 1.0f;
    float col[4];
#pragma unroll
    for (int ch = 0; ch < 4; ch++) col[ch] = wr_mixf(g[ch], g[4 + ch], mix_factor) * alpha;
    return wr_pack_rgba(col);
  }
};

// ---- cs_border_segment (cs_border_segment.glsl:258-449) ---------------------------------
// evaluate_color_for_style_in_corner
WRD void wr_border_color_corner(const float* g, float cx, float cy, int style, const float* color0,
                                const float* color1, float mix_factor, int segment, float aa_range, float* out) {
  switch (style) {
    case 2: {
      float da = wr_distance_to_ellipse(cx, cy, g[26] - g[34], g[27] - g[35]);
      float db = wr_distance_to_ellipse(cx, cy, g[26] - 2.0f * g[34], g[27] - 2.0f * g[35]);
      float al = wr_distance_aa(aa_range, wr_min(-da, db));
      for (int ch = 0; ch < 4; ch++) out[ch] = color0[ch] * al;
      break;
    }
    case 6:
    case 7: {
      float alpha = wr_distance_aa(aa_range, wr_distance_to_ellipse(cx, cy, g[26] - g[36], g[27] - g[37]));
      float sf;
      switch (segment) {
        case 0: sf = 0.0f; break;
        case 1: sf = mix_factor; break;
        case 2: sf = 1.0f; break;
        case 3: sf = 1.0f - mix_factor; break;
        default: sf = 0.0f; break;
      }
      for (int ch = 0; ch < 4; ch++) {
        float c0 = wr_mixf(color1[ch], color0[ch], sf), c1 = wr_mixf(color0[ch], color1[ch], sf);
        out[ch] = wr_mixf(c0, c1, alpha);
      }
      break;
    }
    default:
      for (int ch = 0; ch < 4; ch++) out[ch] = color0[ch];
      break;
  }
}
// evaluate_color_for_style_in_edge
WRD void wr_border_color_edge(const float* g, float px, float py, int style, const float* color0, const float* color1,
                              float aa_range, int edge_axis_id, float* out) {
  const float ex = edge_axis_id != 0 ? 0.0f : 1.0f, ey = edge_axis_id != 0 ? 1.0f : 0.0f;
  const float pos = px * ex + py * ey;
  switch (style) {
    case 2: {
      float d = -1.0f;
      float partial_width = g[34] * ex + g[35] * ey;
      if (partial_width >= 1.0f) {
        float r0 = (g[30] * ex + g[31] * ey) + partial_width;
        float r1 = (g[32] * ex + g[33] * ey) - partial_width;
        d = wr_min(pos - r0, r1 - pos);
      }
      float al = wr_distance_aa(aa_range, d);
      for (int ch = 0; ch < 4; ch++) out[ch] = color0[ch] * al;
      break;
    }
    case 6:
    case 7: {
      float ref = (g[30] + g[36]) * ex + (g[31] + g[37]) * ey;
      float alpha = wr_distance_aa(aa_range, pos - ref);
      for (int ch = 0; ch < 4; ch++) out[ch] = wr_mixf(color0[ch], color1[ch], alpha);
      break;
    }
    default:
      for (int ch = 0; ch < 4; ch++) out[ch] = color0[ch];
      break;
  }
}

struct BorderSegmentShader {
  typedef PosRow Row;
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    wr_pos_row_setup(a, c, y, tx0, r);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const float* g = k.g;
    float p[2], aa_range;
    wr_pos_at(r, x - c.x0, p, &aa_range);
    const float px = p[0], py = p[1];
    const int segment = k.i[0] & 0xff, clip_mode = (k.i[0] >> 8) & 0xff;
    const int style0 = (k.i[0] >> 16) & 0xff, style1 = (k.i[0] >> 24) & 0xff;
    const int ea0 = k.i[1] & 1, ea1 = (k.i[1] >> 1) & 1;
    float mix_factor = 0.0f;
    if (ea0 != ea1) mix_factor = wr_distance_aa(aa_range, -wr_distance_to_line_n(g[16], g[17], g[20], g[21], px, py));
    const float cx = px - g[22], cy = py - g[23];
    const bool in_clip_region = g[24] * cx < 0.0f && g[25] * cy < 0.0f;
    float d = -1.0f;
    switch (clip_mode) {
      case 3: {
        float dx = k.f[0] - px, dy = k.f[1] - py;
        d = sqrtf(dx * dx + dy * dy) - k.f[2];
        break;
      }
      case 2: {
        bool is_vertical = k.f[0] == 0.0f;
        float half_dash = is_vertical ? k.f[1] : k.f[0];
        float pos = is_vertical ? py : px;
        bool in_dash = pos < half_dash || pos > 3.0f * half_dash;
        if (!in_dash) d = 1.0f;
        break;
      }
      case 1: {
        float d0 = wr_distance_to_line_n(k.f[0], k.f[1], k.f[4], k.f[5], px, py);
        float d1 = wr_distance_to_line_n(k.f[6], k.f[7], g[38], g[39], px, py);
        d = wr_max(d0, -d1);
        break;
      }
      default:
        break;
    }
    float color0[4], color1[4];
    if (in_clip_region) {
      float da = wr_distance_to_ellipse(cx, cy, g[26], g[27]), db = wr_distance_to_ellipse(cx, cy, g[28], g[29]);
      d = wr_max(d, wr_max(da, -db));
      wr_border_color_corner(g, cx, cy, style0, g, g + 4, mix_factor, segment, aa_range, color0);
      wr_border_color_corner(g, cx, cy, style1, g + 8, g + 12, mix_factor, segment, aa_range, color1);
    } else {
      wr_border_color_edge(g, px, py, style0, g, g + 4, aa_range, ea0, color0);
      wr_border_color_edge(g, px, py, style1, g + 8, g + 12, aa_range, ea1, color1);
    }
    float alpha = wr_distance_aa(aa_range, d);
    float col[4];
#pragma unroll
    for (int ch = 0; ch < 4; ch++) col[ch] = wr_mixf(color0[ch], color1[ch], mix_factor) * alpha;
    return wr_pack_rgba(col);
  }
};

// ---- vertex stages -------------------------------------------------------------------------
// cs_line_decoration.glsl:46-93
WRD void wr_setup_line_decoration_one(const SetupArgs& a, int idx) {
  const float* f = (const float*)(a.instances + (size_t)idx * a.stride);
  const int style = ((const int*)f)[7];
  const float axis = f[8];
  float size[2] = {wr_mixf(f[4], f[5], axis), wr_mixf(f[5], f[4], axis)};
  float params[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  switch (style) {
    case 2:
      params[0] = size[0]; params[1] = 0.5f * size[0];
      break;
    case 1:
      params[0] = size[1] * 2.0f; params[1] = size[1] / 2.0f; params[2] = 0.5f * size[1];
      break;
    case 3: {
      float line_thickness = wr_max(f[6], 1.0f);
      params[0] = line_thickness / 2.0f;
      params[1] = size[1] - line_thickness;
      params[2] = wr_max((line_thickness - 1.0f) * 2.0f, 1.0f);
      params[3] = size[1];
      break;
    }
    default:
      break;
  }
  QuadOut q;
  memset(&q, 0, sizeof q);
  const float axs[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ays[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int v = 0; v < 4; v++) {
    float ax = axs[v], ay = ays[v];
    q.pos[v] = wr_mat_mul(a.tgt.proj, make_float4((f[2] - f[0]) * ax + f[0], (f[3] - f[1]) * ay + f[1], 0.0f, 1.0f));
    q.interp[v][0] = wr_mixf(ax, ay, axis) * size[0];
    q.interp[v][1] = wr_mixf(ay, ax, axis) * size[1];
  }
  q.n_interp = 2;
  float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  wr_pack_color(q, white);
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 4; i++) k->f[i] = params[i];
    k->i[0] = style;
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_line_decoration)

// mod_color / get_colors_for_side (cs_border_segment.glsl:113-157)
WRD void wr_colors_for_side(const float* color, int style, float* r0, float* r1) {
  bool is_black = color[0] == 0.0f && color[1] == 0.0f && color[2] == 0.0f;
  float lighter[4], darker[4];
  for (int ch = 0; ch < 3; ch++) {
    lighter[ch] = is_black ? 0.7f : color[ch] * 1.0f;
    darker[ch] = is_black ? 0.3f : color[ch] * 0.66666666f;
  }
  lighter[3] = darker[3] = color[3];
  for (int ch = 0; ch < 4; ch++) {
    r0[ch] = style == 6 ? lighter[ch] : (style == 7 ? darker[ch] : color[ch]);
    r1[ch] = style == 6 ? darker[ch] : (style == 7 ? lighter[ch] : color[ch]);
  }
}

// cs_border_solid.glsl:85-133, cs_border_segment.glsl:159-254
WRD void wr_setup_border_one(const SetupArgs& a, int idx) {
  const float* f = (const float*)(a.instances + (size_t)idx * a.stride);
  const float* origin = f;
  const float* rect = f + 2;
  const float* color0 = f + 6;
  const float* color1 = f + 10;
  const int flags = ((const int*)f)[14];
  const float* widths = f + 15;
  const float* radii = f + 17;
  const float* cp1 = f + 19;
  const float* cp2 = f + 23;
  const bool solid = a.kind == WRCU_KIND_BORDER_SOLID;
  const int segment = flags & 0xff;
  float fc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, g[40];
  for (int i = 0; i < 40; i++) g[i] = 0.0f;
  int ic[2] = {0, 0};
  const float ox[4] = {0.0f, 1.0f, 1.0f, 0.0f}, oy[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  const float os0 = segment < 4 ? ox[segment] : 0.0f, os1 = segment < 4 ? oy[segment] : 0.0f;
  const float size[2] = {rect[2] - rect[0], rect[3] - rect[1]};
  const float outer[2] = {os0 * size[0], os1 * size[1]};
  const float cs[2] = {1.0f - 2.0f * os0, 1.0f - 2.0f * os1};
  g[22] = outer[0] + cs[0] * radii[0];
  g[23] = outer[1] + cs[1] * radii[1];
  g[24] = cs[0];
  g[25] = cs[1];
  g[26] = radii[0];
  g[27] = radii[1];
  g[28] = wr_max(radii[0] - widths[0], 0.0f);
  g[29] = wr_max(radii[1] - widths[1], 0.0f);
  g[16] = outer[0];
  g[17] = outer[1];
  g[18] = widths[1] * -cs[1];
  g[19] = widths[0] * cs[0];
  {
    float l = wr_hypotf(g[18], g[19]);
    g[20] = g[18] / l;
    g[21] = g[19] / l;
  }
  const float axs[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ays[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  float vpos[4][2];
  for (int v = 0; v < 4; v++) {
    vpos[v][0] = size[0] * axs[v];
    vpos[v][1] = size[1] * ays[v];
  }
  if (solid) {
    const bool do_aa = ((flags >> 24) & 0xf0) != 0;
    ic[0] = segment < 4 ? (do_aa ? 1 : 2) : 0;
    for (int ch = 0; ch < 4; ch++) { g[ch] = color0[ch]; g[4 + ch] = color1[ch]; }
    const float hs[2] = {-cs[0], cs[1]}, vs[2] = {cs[0], -cs[1]};
    g[30] = cp1[0] + hs[0] * cp1[2];
    g[31] = cp1[1] + hs[1] * cp1[3];
    g[32] = hs[0];
    g[33] = hs[1];
    g[34] = cp1[2];
    g[35] = cp1[3];
    fc[0] = cp2[0] + vs[0] * cp2[2];
    fc[1] = cp2[1] + vs[1] * cp2[3];
    fc[2] = vs[0];
    fc[3] = vs[1];
    fc[4] = cp2[2];
    fc[5] = cp2[3];
  } else {
    const int style0 = (flags >> 8) & 0xff, style1 = (flags >> 16) & 0xff, clip_mode = (flags >> 24) & 0x0f;
    int ea0 = 0, ea1 = 0;
    float er[2] = {0.0f, 0.0f};
    switch (segment) {
      case 0: ea0 = 0; ea1 = 1; er[0] = outer[0]; er[1] = outer[1]; break;
      case 1: ea0 = 1; ea1 = 0; er[0] = outer[0] - widths[0]; er[1] = outer[1]; break;
      case 2: ea0 = 0; ea1 = 1; er[0] = outer[0] - widths[0]; er[1] = outer[1] - widths[1]; break;
      case 3: ea0 = 1; ea1 = 0; er[0] = outer[0]; er[1] = outer[1] - widths[1]; break;
      case 5: case 7: ea0 = 1; ea1 = 1; break;
      default: break;
    }
    ic[0] = segment | (clip_mode << 8) | (style0 << 16) | (style1 << 24);
    ic[1] = ea0 | (ea1 << 1);
    wr_colors_for_side(color0, style0, g, g + 4);
    wr_colors_for_side(color1, style1, g + 8, g + 12);
    g[30] = er[0];
    g[31] = er[1];
    g[32] = er[0] + widths[0];
    g[33] = er[1] + widths[1];
    g[34] = widths[0] / 3.0f;
    g[35] = widths[1] / 3.0f;
    g[36] = widths[0] / 2.0f;
    g[37] = widths[1] / 2.0f;
    fc[0] = cp1[0];
    fc[1] = cp1[1];
    fc[2] = cp1[2];
    {
      float l = wr_hypotf(cp1[2], cp1[3]);
      fc[4] = cp1[2] / l;
      fc[5] = cp1[3] / l;
      l = wr_hypotf(cp2[2], cp2[3]);
      g[38] = cp2[2] / l;
      g[39] = cp2[3] / l;
    }
    fc[6] = cp2[0];
    fc[7] = cp2[1];
    if (clip_mode == 3) {
      float radius = cp1[2];
      if (radius > 0.5f) radius += 2.0f;
      for (int v = 0; v < 4; v++) {
        vpos[v][0] = wr_clamp(cp1[0] + radius * (2.0f * axs[v] - 1.0f), 0.0f, size[0]);
        vpos[v][1] = wr_clamp(cp1[1] + radius * (2.0f * ays[v] - 1.0f), 0.0f, size[1]);
      }
    } else if (clip_mode == 1) {
      const float center[2] = {(cp1[0] + cp2[0]) * 0.5f, (cp1[1] + cp2[1]) * 0.5f};
      const float dash_length = wr_hypotf(cp1[0] - cp2[0], cp1[1] - cp2[1]);
      const float r = wr_max(dash_length, wr_max(widths[0], widths[1])) + 2.0f;
      for (int v = 0; v < 4; v++) {
        vpos[v][0] = wr_clamp(vpos[v][0], center[0] - r, center[0] + r);
        vpos[v][1] = wr_clamp(vpos[v][1], center[1] - r, center[1] + r);
      }
    }
  }
  QuadOut q;
  memset(&q, 0, sizeof q);
  for (int v = 0; v < 4; v++) {
    q.pos[v] = wr_mat_mul(a.tgt.proj, make_float4((origin[0] + rect[0]) + vpos[v][0], (origin[1] + rect[1]) + vpos[v][1],
                                                  0.0f, 1.0f));
    q.interp[v][0] = vpos[v][0];
    q.interp[v][1] = vpos[v][1];
  }
  q.n_interp = 2;
  float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  wr_pack_color(q, white);
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 8; i++) k->f[i] = fc[i];
    for (int i = 0; i < 40; i++) k->g[i] = g[i];
    k->i[0] = ic[0];
    k->i[1] = ic[1];
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_border)
