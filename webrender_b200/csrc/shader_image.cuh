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
    // (perspective rows have no span body: every chunk is the fragment shader, rasterize.h:1262-1270)
    int body_len = (rgba && len >= 4 && k.g[7] != 0.0f && !a.persp) ? (len & ~3) : 0;
    float u[4], v[4];
    for (int j = 0; j < 4; j++) {
      float uv[2];
      wr_interp_at<2>(a, r.o, r.step, j, uv);
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
    wr_interp_at<2>(a, r.o, r.step, rel, uv);
    float pd = r.pd;
    if (a.persp) {  // gl_FragCoord.w varies per sample
      const float fw = wr_persp_zw(*a.persp, 1, rel);
      pd = (1.0f - fw) * k.f[6] + fw;
    }
    float ru = uv[0] * pd + k.f[4], rv = uv[1] * pd + k.f[5];
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

// ---- brush_image with WR_FEATURE_REPETITION (tiled images, border-image segments) ------
// (brush_image.glsl:99-160, 213-265 vertex; 319-352 fragment; span: swgl_commitTextureRepeat
// [Color]RGBA8 = blendTextureLinearRepeat / blendTextureNearestRepeat<true>, swgl_ext.h:671-853).
// The reference walks the span as alternating pieces: a run of chunks that stays inside one tile
// (drawn with the ordinary filtered span code on fract(uv) mapped into the tile's uv rect) and one
// chunk that may straddle a tile boundary (each lane wrapped on its own).  Where a piece starts
// depends on the float state of the walk, so the walk is replayed: warp-uniformly up to the tile,
// then per pixel (a lane keeps the partition of its current run between its four pixels).
// Extra CmdCold fields: g[8..9] v_uv_bounds.zw, g[10..11] v_tile_repeat_bounds.
struct RepWalk {
  float u[4], v[4];  // unwrapped, unscaled uv lanes at `pos`
  int pos;           // span-relative pixel where the next piece starts
};
WRD int wr_no_repeat_steps(const float* lanes, float step, float tile_repeat, int steps) {  // swgl_ext.h:683-701
  float lo = lanes[0], hi = lanes[3];
  if (hi < lo) { float t = lo; lo = hi; hi = t; }
  float limit = floorf(lo) + 1.0f;
  if (tile_repeat > 0.0f) limit = wr_min(limit, tile_repeat);
  if (!(lo >= 0.0f && hi < limit)) return 0;
  return step != 0.0f ? (int)wr_clamp((limit - lo) / step, 0.0f, (float)steps) : steps;
}
WRD void wr_tile_repeat_uv(float u, float v, float trx, float try_, float* ou, float* ov) {  // swgl_ext.h:671-680
  if (trx > 0.0f) {
    u = wr_clamp(u, 0.0f, trx - 1.0e-6f);
    v = wr_clamp(v, 0.0f, try_ - 1.0e-6f);
  }
  *ou = u - floorf(u);
  *ov = v - floorf(v);
}
struct RepConst {
  bool linear;
  int filter, span;
  float su, sv;                       // chunk step of the unscaled uv
  float scx, scy, ofx, ofy;           // uv_scale / uv_offset (quantised for linear, texel units for nearest)
  float minu, minv, maxu, maxv;       // clamp bounds (linear)
  float trx, try_;
  bool solid;                         // nearest: single-texel span
};
// one piece of the walk: `steps` chunks inside a tile from w.pos (may be 0), then one wrapped chunk
WRD int wr_rep_steps(const RepConst& rc, const RepWalk& w) {
  int steps = (rc.span - w.pos) / 4;
  steps = wr_no_repeat_steps(w.u, rc.su, rc.trx, steps);
  if (steps > 0) steps = wr_no_repeat_steps(w.v, rc.sv, rc.try_, steps);
  return steps;
}
WRD void wr_rep_advance(const RepConst& rc, RepWalk& w, int steps) {
  if (steps > 0) {
    const float fs = (float)steps;
#pragma unroll
    for (int j = 0; j < 4; j++) { w.u[j] = w.u[j] + fs * rc.su; w.v[j] = w.v[j] + fs * rc.sv; }
    w.pos += steps * 4;
    if (w.pos >= rc.span) return;
  }
#pragma unroll
  for (int j = 0; j < 4; j++) { w.u[j] = w.u[j] + rc.su; w.v[j] = w.v[j] + rc.sv; }
  w.pos += 4;
}

struct ImageRepeatShader {
  struct Row {
    float o[2], step[2];
    float pd;
    int body_len;
    RepConst rc;
    RepWalk w;            // walk state at the piece that reaches this tile
    // per-lane cache of the current run's partition
    mutable int run_pos, run_steps;
    mutable TexRow tr;
    mutable float nu[4], nv[4];  // nearest: lanes at the run start (texel units)
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    r.pd = (1.0f - k.f[7]) * k.f[6] + k.f[7];
    int len = c.x1 - c.x0;
    r.body_len = (rgba && len >= 4 && k.g[7] != 0.0f && t.fmt == WRCU_FMT_RGBA8) ? (len & ~3) : 0;
    r.run_pos = -1;
    if (!r.body_len) return;
    RepConst& rc = r.rc;
    RepWalk& w = r.w;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float uv[2];
      wr_interp_at<2>(a, r.o, r.step, j, uv);
      w.u[j] = uv[0] * r.pd;
      w.v[j] = uv[1] * r.pd;
    }
    w.pos = 0;
    rc.span = r.body_len;
    rc.su = 4.0f * (w.u[1] - w.u[0]);
    rc.sv = 4.0f * (w.v[1] - w.v[0]);
    rc.trx = k.g[10];
    rc.try_ = k.g[11];
    const float tw = (float)t.w, th = (float)t.h;
    const float rx0 = k.f[4], ry0 = k.f[5], rx1 = k.g[8], ry1 = k.g[9];  // uv_repeat = v_uv_bounds
    rc.linear = t.filter == WRCU_LINEAR && t.w >= 2;
    rc.solid = false;
    if (rc.linear) {
      float scx = rx1 - rx0, scy = ry1 - ry0;
      rc.filter = wr_needs_texture_linear(t, w.u[0] * scx + rx0, w.u[1] * scx + rx0, w.v[0] * scy + ry0,
                                          w.v[1] * scy + ry0, rc.span);
      rc.scx = scx * tw * 128.0f;
      rc.scy = scy * th * 128.0f;
      rc.ofx = wr_linear_quantize(rx0, t.w);
      rc.ofy = wr_linear_quantize(ry0, t.h);
      rc.minu = wr_max(wr_linear_quantize(k.f[0], t.w), 0.0f);
      rc.minv = wr_max(wr_linear_quantize(k.f[1], t.h), 0.0f);
      rc.maxu = wr_max(wr_linear_quantize(k.f[2], t.w), rc.minu);
      rc.maxv = wr_max(wr_linear_quantize(k.f[3], t.h), rc.minv);
    } else {
      rc.filter = 0;
      rc.ofx = rx0 * tw;  // min_uv
      rc.ofy = ry0 * th;
      const float mxu = rx1 * tw, mxv = ry1 * th;
      rc.scx = mxu - rc.ofx;
      rc.scy = mxv - rc.ofy;
      rc.minu = rc.minv = rc.maxu = rc.maxv = 0.0f;
      rc.solid = ((int)rc.ofx + 1 >= (int)mxu || fabsf(rc.su) * (float)rc.span * rc.scx < 0.5f) &&
                 ((int)rc.ofy + 1 >= (int)mxv || fabsf(rc.sv) * (float)rc.span * rc.scy < 0.5f);
    }
    if (rc.solid) return;
    // walk to the piece covering the first pixel this tile draws
    const int first = max(tx0, (int)c.x0) - (int)c.x0;
    if (first >= rc.span) return;
    for (;;) {
      const int steps = wr_rep_steps(rc, w);
      if (first < w.pos + steps * 4 + 4) break;
      wr_rep_advance(rc, w, steps);
      if (w.pos >= rc.span) break;
    }
  }
  WRD_MEMBER Px nearest_at(const TexView& t, float fu, float fv) {
    int cx = wr_clamp_coord((int)fu, t.w), cy = wr_clamp_coord((int)fv, t.h);
    return px_unpack(__ldg((const uint32_t*)(t.ptr + (size_t)cy * t.pitch) + cx));
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    const int rel = x - c.x0;
    const RepConst& rc = r.rc;
    if (rel < r.body_len) {
      const Px col{c.col[0], c.col[1], c.col[2], c.col[3]};
      const int j = rel & 3;
      if (rc.solid) {
        float ru, rv;
        wr_tile_repeat_uv(r.w.u[j], r.w.v[j], rc.trx, rc.try_, &ru, &rv);
        return px_apply_color(nearest_at(t, ru * rc.scx + rc.ofx, rv * rc.scy + rc.ofy), col);
      }
      // continue the walk to the piece holding this pixel
      RepWalk w = r.w;
      int steps;
      for (;;) {
        steps = wr_rep_steps(rc, w);
        if (rel < w.pos + steps * 4 + 4) break;
        wr_rep_advance(rc, w, steps);
      }
      if (rel >= w.pos + steps * 4) {
        // the chunk that may straddle a tile boundary: every lane wraps on its own
        float uj = w.u[j], vj = w.v[j];
        if (steps > 0) { uj = uj + (float)steps * rc.su; vj = vj + (float)steps * rc.sv; }
        float ru, rv;
        wr_tile_repeat_uv(uj, vj, rc.trx, rc.try_, &ru, &rv);
        if (rc.linear) {
          int ix = (int)wr_clamp(ru * rc.scx + rc.ofx, rc.minu, rc.maxu), iy = (int)wr_clamp(rv * rc.scy + rc.ofy, rc.minv, rc.maxv);
          return px_apply_color(wr_texture_linear_rgba8(t, ix, iy), col);
        }
        return px_apply_color(nearest_at(t, ru * rc.scx + rc.ofx, rv * rc.scy + rc.ofy), col);
      }
      // inside a run: the ordinary span code on fract(uv) mapped into the tile's uv rect
      const int p = rel - w.pos;
      if (r.run_pos != w.pos) {
        r.run_pos = w.pos;
        r.run_steps = steps;
        if (rc.linear) {
          TexRow& tr = r.tr;
          tr.mode = TEX_LINEAR;
          tr.body_len = steps * 4;
          tr.filter = rc.filter;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            tr.qu[q] = (w.u[q] - floorf(w.u[q])) * rc.scx + rc.ofx;
            tr.qv[q] = (w.v[q] - floorf(w.v[q])) * rc.scy + rc.ofy;
          }
          tr.ustep = rc.su * rc.scx;
          tr.vstep = rc.sv * rc.scy;
          tr.minu = rc.minu; tr.minv = rc.minv; tr.maxu = rc.maxu; tr.maxv = rc.maxv;
          wr_tex_linear_partition(t, tr, tr.body_len, p & ~3, false);  // per-lane: no warp cooperation
        } else {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            r.nu[q] = (w.u[q] - floorf(w.u[q])) * rc.scx + rc.ofx;
            r.nv[q] = (w.v[q] - floorf(w.v[q])) * rc.scy + rc.ofy;
          }
        }
      }
      if (rc.linear) return px_apply_color(wr_tex_body(t, r.tr, p), col);
      const float fu = wr_repeat_add(r.nu[j], rc.su * rc.scx, p >> 2), fv = wr_repeat_add(r.nv[j], rc.sv * rc.scy, p >> 2);
      return px_apply_color(nearest_at(t, fu, fv), col);
    }
    // fragment path: compute_repeated_uvs (brush_image.glsl:319-352)
    float uv[2];
    wr_interp_at<2>(a, r.o, r.step, rel, uv);
    const float usx = k.g[8] - k.f[4], usy = k.g[9] - k.f[5];
    float lu = uv[0] * r.pd, lv = uv[1] * r.pd, ru, rv;
    if (k.g[6] != 0.0f) {
      lu = wr_max(lu, 0.0f);
      lv = wr_max(lv, 0.0f);
      ru = (lu - floorf(lu)) * usx + k.f[4];
      rv = (lv - floorf(lv)) * usy + k.f[5];
      if (lu >= k.g[10]) ru = k.g[8];
      if (lv >= k.g[11]) rv = k.g[9];
    } else {
      ru = (lu - floorf(lu)) * usx + k.f[4];
      rv = (lv - floorf(lv)) * usy + k.f[5];
    }
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

// depth-run walk (raster.cuh WrRunWalk): interpolants = v_uv; span shader = swgl_commitTexture*
template <> struct WrRun<ImageShader> {
  enum { n = 2 };
  WRD_MEMBER int drawn(const ImageShader::Row& r) { return r.tr.body_len; }
};
template <> struct WrRun<ImageRepeatShader> {
  enum { n = 2 };
  WRD_MEMBER int drawn(const ImageRepeatShader::Row& r) { return r.body_len; }
};
