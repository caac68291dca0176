// texspan.cuh — textured span sampling shared by the quad, image-brush and
// composite shaders: the reference's swgl_commitTexture* family
// (swgl/src/swgl_ext.h:385-612, 775-947) restated per pixel.
//
// SWGL draws the first len&~3 pixels of a span with a span shader that picks a
// filter from the uv step (needsTextureLinear / needsNearestFallback) and then
// runs blendTextureLinearDispatch: a clamped prefix through the fallback
// bilinear filter, an interior run through a specialised filter (FAST: 1 texel
// per pixel with constant fractions; DOWNSCALE: 2 texels per pixel; UPSCALE),
// and the remainder through the fallback again.  TexRow holds that partition,
// computed once per (command,row) per warp; wr_tex_body evaluates one pixel.
// The remaining len&3 pixels run the fragment shader (texture(): wr_tex_fragment).
//
// Exactness: inside FALLBACK / UPSCALE / nearest-fallback runs the reference
// accumulates the (quantised) uv chunk by chunk.  wr_tex_row_setup walks that
// running sum up to the first chunk the tile touches (warp-uniform, once per
// command/row/tile) and wr_tex_body finishes it per pixel (<= 32 additions in
// a 128-pixel tile), so every run is reproduced bit for bit.
#pragma once
#include "blend.cuh"
#include "sample.cuh"
#include "repeat_add.cuh"

enum { TEX_NONE = 0, TEX_LINEAR = 1, TEX_NEAREST_FAST = 2, TEX_NEAREST_FALLBACK = 3, TEX_LINEAR_R8 = 4 };
enum { LF_NEAREST = 0, LF_FALLBACK = 1, LF_UPSCALE = 2, LF_FAST = 3, LF_DOWNSCALE = 4 };

struct TexRow {
  int mode, body_len;
  float u[4], v[4];  // uv lanes of chunk 0 of the span body
  // linear
  int filter, before, inside;
  float qu[4], qv[4], ustep, vstep, minu, minv, maxu, maxv;
  int fcx, fcy, fnext, ffx, ffy;  // FAST/DOWNSCALE: clamped start texel, next-row flag, fractions
  int uiy0;                       // UPSCALE: lane 0's clamped quantised y
  // nearest
  int nix, nry, nminx, nmaxx;     // fast
  int nsolid;                     // fallback: single-texel span
  // running-sum bases: x lanes of the first chunk >= the tile start inside each
  // of the three dispatch segments (prefix fallback / interior / remainder)
  int kb[3];
  float bu[3][4], bv[4];
  int exact;  // bit s: segment s's u sums are exact (wr_sum_exact); bit 3: the v sums
};

// Texel access by texture format.  R8 texels travel in Px.r (the lane the R8
// blend stage reads); matchTextureFormat (swgl_ext.h:143-156) guarantees the
// span paths only ever see source format == target format.
WRD Px wr_tex_linear_any(const TexView& t, int ix, int iy) {
  if (t.fmt == WRCU_FMT_RGBA8) return wr_texture_linear_rgba8(t, ix, iy);
  return Px{0, 0, wr_texture_linear_r8(t, ix, iy), 0};
}
WRD Px wr_tex_load_any(const TexView& t, int cy, int cx) {
  if (t.fmt == WRCU_FMT_RGBA8) return px_unpack(__ldg((const uint32_t*)(t.ptr + (size_t)cy * t.pitch) + cx));
  return Px{0, 0, (int)__ldg(t.ptr + (size_t)cy * t.pitch + cx), 0};
}

// needsTextureLinear (swgl_ext.h:554-587); u0,u1,v0,v1 = lanes 0,1
WRD int wr_needs_texture_linear(const TexView& t, float u0, float u1, float v0, float v1, int span) {
  if (t.w < 2) return LF_NEAREST;
  if (v0 != v1) return LF_FALLBACK;
  float px0 = u0 * (float)t.w, px1 = u1 * (float)t.w, py0 = v0 * (float)t.h;
  int sp = (span & ~127) + 128;
  int scaled = (int)roundf((px1 - px0) * (float)sp);
  int scale = scaled != sp ? (scaled == sp * 2 ? 2 : 1) : 0;
  if (scale) return (px0 < px1 && px1 - px0 <= 1) ? LF_UPSCALE : (scale == 2 ? LF_DOWNSCALE : LF_FALLBACK);
  if ((((int)(px0 * 4.0f + 0.5f)) & 3) != 2 || (((int)(py0 * 4.0f + 0.5f)) & 3) != 2) return LF_FAST;
  return LF_NEAREST;
}

// Set up the span-body partition.  `use_sampler_filter`: swgl_commitTexture
// (image brush, composite) picks linear vs nearest from the sampler's filter;
// swgl_commitTextureLinear (quads) always takes the linear decision tree.
WRD void wr_tex_seq_base(const float* start, float step, int kb, float* out) {
  for (int j = 0; j < 4; j++) {
    float v = start[j];
    v = wr_repeat_add(v, step, kb);
    out[j] = v;
  }
}

// The running-sum bases of one textured row — up to three x segments and the y lanes, four chunk
// lanes each — are 16 independent wr_repeat_add walks.  row_setup runs with the whole warp
// converged on the same (command,row), so lane l < 16 walks one of them and the results are
// broadcast: one walk's worth of instructions instead of sixteen (a 4K-wide scaled span crosses
// ~18 binades per walk by its last tile).
WRD float wr_sel4(const float* a, int j) { return j == 0 ? a[0] : (j == 1 ? a[1] : (j == 2 ? a[2] : a[3])); }
// `coop` must be false wherever lanes arrive with different arguments (per-pixel callers).
WRD void wr_tex_bases(const float* s0, int n0, const float* s1, int n1, const float* s2, int n2, float ustep,
                      const float* sv, int nv, float vstep, float (*bu)[4], float* bv, bool coop) {
  if ((n0 | n1 | n2 | nv) == 0) {  // first tile of a span (every glyph): nothing to walk
    for (int j = 0; j < 4; j++) { bu[0][j] = s0[j]; bu[1][j] = s1[j]; bu[2][j] = s2[j]; bv[j] = sv[j]; }
    return;
  }
#ifndef WRCU_HOSTEMU
  if (!coop)
#endif
  {
    wr_tex_seq_base(s0, ustep, n0, bu[0]);
    wr_tex_seq_base(s1, ustep, n1, bu[1]);
    wr_tex_seq_base(s2, ustep, n2, bu[2]);
    wr_tex_seq_base(sv, vstep, nv, bv);
    return;
  }
#ifndef WRCU_HOSTEMU
  const int l = threadIdx.x & 31, g = (l >> 2) & 3, j = l & 3;
  const float x = g == 0 ? wr_sel4(s0, j) : (g == 1 ? wr_sel4(s1, j) : (g == 2 ? wr_sel4(s2, j) : wr_sel4(sv, j)));
  const int n = g == 0 ? n0 : (g == 1 ? n1 : (g == 2 ? n2 : nv));
  const float val = l < 16 ? wr_repeat_add(x, g == 3 ? vstep : ustep, n) : 0.0f;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    bu[0][q] = __shfl_sync(0xFFFFFFFFu, val, q);
    bu[1][q] = __shfl_sync(0xFFFFFFFFu, val, 4 + q);
    bu[2][q] = __shfl_sync(0xFFFFFFFFu, val, 8 + q);
    bv[q] = __shfl_sync(0xFFFFFFFFu, val, 12 + q);
  }
#endif
}

// Running sums that never round: when the four lane bases and the step are multiples of one power of
// two g and stay below 2^24 g over the tile's 32 chunks, every partial sum is exactly representable, so the
// reference's chunk-by-chunk accumulation equals base + m*step and needs no per-pixel replay
// (integer and half-texel scale factors — 1:1, 2x video, device-pixel-ratio 2 — all land here).
WRD bool wr_sum_exact(const float* base, float step) {
  // any power-of-two grid g works: all values multiples of g and below 2^24 * g.  The step decides
  // almost always (a scale factor that is not a dyadic rational is on no grid): test it first.
  const float s2 = step * 256.0f;
  if (s2 != truncf(s2)) return false;
  const float grids[3] = {256.0f, 16.0f, 2.0f};
  for (int gi = 0; gi < 3; gi++) {
    const float g = grids[gi];
    float s = step * g;
    bool ok = s == truncf(s);
    float amax = 0.0f;
    for (int j = 0; j < 4; j++) {
      float b = base[j] * g;
      ok = ok && b == truncf(b);
      amax = wr_max(amax, fabsf(b));
    }
    if (ok && amax + 34.0f * fabsf(s) < 16777216.0f) return true;
  }
  return false;
}
WRD float wr_sum_at(float base, float step, int m, bool exact) {
  if (exact) return base + (float)m * step;
  for (int s = 0; s < m; s++) base = base + step;
  return base;
}

// blendTextureLinearDispatch's partition (swgl_ext.h:385-448) for a span whose quantised uv lanes,
// steps, clamp bounds and filter are already in `r`: prefix through the fallback filter, interior
// through the selected filter, remainder through the fallback; plus the running-sum bases of the
// first chunk >= tile_rel inside each segment.
WRD_SHARED void wr_tex_linear_partition(const TexView& t, TexRow& r, int body_len, int tile_rel, bool coop = true) {
  const int filter = r.filter;
  r.before = 0;
  r.inside = 0;
  if (filter != LF_FALLBACK) {
    // blendTextureLinearDispatch (swgl_ext.h:385-448)
    float q0 = r.qu[0];
    float beforeDist = wr_max(0.0f, r.minu) - q0;
    if (beforeDist > 0) {
      r.before = min(max((int)ceilf(beforeDist / r.ustep) * 4, 0), body_len);
      q0 = q0 + (float)(r.before / 4) * r.ustep;
    }
    float insideDist = wr_min(r.maxu, (float)((t.w - 4) * 128)) - q0;
    if (r.ustep > 0.0f && insideDist >= r.ustep) {
      int inside = body_len - r.before;
      if (filter == LF_DOWNSCALE) inside = min(((int)(insideDist * (0.5f / 128.0f))) & ~3, inside);
      else if (filter == LF_UPSCALE) inside = min((int)(insideDist / r.ustep) * 4, inside);
      else inside = min(((int)(insideDist * (1.0f / 128.0f))) & ~3, inside);
      r.inside = max(inside, 0);
    }
    if (r.inside > 0) {
      int ix0 = (int)wr_clamp(q0, r.minu, r.maxu), iy0 = (int)wr_clamp(r.qv[0], r.minv, r.maxv);
      r.uiy0 = iy0;
      int tx = ix0 >> 7, ty = iy0 >> 7;
      r.fcx = wr_clamp_coord(tx, t.w - 1);
      r.fcy = wr_clamp_coord(ty, t.h);
      r.fnext = (ty >= 0 && ty < t.h - 1) ? 1 : 0;
      int overread = tx > t.w - 2 ? -1 : 0;
      r.ffx = (int)(short)((((ix0 & (tx >= 0 ? -1 : 0)) | overread) & 0x7F) - overread);
      r.ffy = iy0 & 0x7F;
    }
  }
  // running-sum bases per segment
  {
    float s0[4], s1[4], s2[4];
    for (int j = 0; j < 4; j++) {
      s0[j] = r.qu[j];
      s1[j] = r.before > 0 ? s0[j] + (float)(r.before / 4) * r.ustep : s0[j];
      s2[j] = r.inside > 0 ? s1[j] + (float)(r.inside / 4) * r.ustep : s1[j];
    }
    r.kb[0] = max(0, tile_rel >> 2);
    r.kb[1] = max(0, (tile_rel - r.before) >> 2);
    r.kb[2] = max(0, (tile_rel - r.before - r.inside) >> 2);
    wr_tex_bases(s0, r.before > 0 ? min(r.kb[0], r.before >> 2) : 0,
                 s1, (r.inside > 0 && r.filter == LF_UPSCALE) ? min(r.kb[1], r.inside >> 2) : 0,
                 s2, r.kb[2], r.ustep, r.qv, r.kb[2], r.vstep, r.bu, r.bv, coop);
    r.kb[0] = min(r.kb[0], r.before >> 2);
    r.kb[1] = min(r.kb[1], r.inside >> 2);
    r.exact = body_len <= 32 ? 0 :  // short spans: the replay is at most 8 additions, not worth the test
              (r.before > 0 && wr_sum_exact(r.bu[0], r.ustep) ? 1 : 0) |
              (r.inside > 0 && r.filter == LF_UPSCALE && wr_sum_exact(r.bu[1], r.ustep) ? 2 : 0) |
              (wr_sum_exact(r.bu[2], r.ustep) ? 4 : 0) | (wr_sum_exact(r.bv, r.vstep) ? 8 : 0);
  }
}

WRD_SHARED void wr_tex_row_setup(const TexView& t, const float* bounds, bool use_sampler_filter, int body_len,
                          const float* u, const float* v, int tile_rel, TexRow& r,
                          int target_fmt = WRCU_FMT_RGBA8, bool coop = true) {
  r.body_len = body_len;
  r.mode = TEX_NONE;
  if (body_len == 0 || t.fmt != target_fmt) {
    r.body_len = 0;
    return;
  }
  for (int j = 0; j < 4; j++) { r.u[j] = u[j]; r.v[j] = v[j]; }
  bool linear = !use_sampler_filter || t.filter == WRCU_LINEAR;
  int filter = LF_NEAREST;
  if (linear) {
    filter = wr_needs_texture_linear(t, u[0], u[1], v[0], v[1], body_len);
  } else {
    // needsNearestFallback (swgl_ext.h:860-863)
    float p0x = u[0] * (float)t.w, p1x = u[1] * (float)t.w, p0y = v[0] * (float)t.h, p1y = v[1] * (float)t.h;
    int sp = (body_len & ~127) + 128;
    int scaled = (int)roundf((p1x - p0x) * (float)sp);
    if ((p1y - p0y) * (float)body_len >= 0.5f || scaled != sp) {
      r.mode = TEX_NEAREST_FALLBACK;
      float ustep = 4.0f * (p1x - p0x), vstep = 4.0f * (p1y - p0y);
      r.ustep = ustep; r.vstep = vstep;
      r.minu = bounds[0] * (float)t.w; r.minv = bounds[1] * (float)t.h;
      r.maxu = bounds[2] * (float)t.w; r.maxv = bounds[3] * (float)t.h;
      r.nsolid = ((int)r.minu >= (int)r.maxu || fabsf(ustep) * (float)body_len * 1.0f < 0.5f) &&
                 ((int)r.minv >= (int)r.maxv || fabsf(vstep) * (float)body_len * 1.0f < 0.5f);
      for (int j = 0; j < 4; j++) { r.qu[j] = u[j] * (float)t.w; r.qv[j] = v[j] * (float)t.h; }
      r.kb[2] = r.nsolid ? 0 : max(0, tile_rel >> 2);
      wr_tex_bases(r.qu, 0, r.qu, 0, r.qu, r.kb[2], ustep, r.qv, r.kb[2], vstep, r.bu, r.bv, coop);
      r.exact = body_len <= 32 ? 0 : (wr_sum_exact(r.bu[2], ustep) ? 4 : 0) | (wr_sum_exact(r.bv, vstep) ? 8 : 0);
      return;
    }
  }
  if (filter == LF_NEAREST) {
    // blendTextureNearestFast (swgl_ext.h:476-541)
    r.mode = TEX_NEAREST_FAST;
    r.nix = (int)(u[0] * (float)t.w);
    int iy = (int)(v[0] * (float)t.h);
    int minUx = (int)(bounds[0] * (float)t.w), minUy = (int)(bounds[1] * (float)t.h);
    int maxUx = (int)(bounds[2] * (float)t.w), maxUy = (int)(bounds[3] * (float)t.h);
    r.nry = wr_clamp_coord(min(max(iy, minUy), maxUy), t.h);
    r.nminx = min(max(minUx, 0), t.w - 1);
    r.nmaxx = min(max(maxUx, r.nminx), t.w - 1);
    return;
  }
  r.mode = TEX_LINEAR;
  r.filter = filter;
  for (int j = 0; j < 4; j++) {
    r.qu[j] = wr_linear_quantize(u[j], t.w);
    r.qv[j] = wr_linear_quantize(v[j], t.h);
  }
  r.ustep = 4.0f * (r.qu[1] - r.qu[0]);
  r.vstep = 4.0f * (r.qv[1] - r.qv[0]);
  r.minu = wr_max(wr_linear_quantize(bounds[0], t.w), 0.0f);
  r.minv = wr_max(wr_linear_quantize(bounds[1], t.h), 0.0f);
  r.maxu = wr_max(wr_linear_quantize(bounds[2], t.w), r.minu);
  r.maxv = wr_max(wr_linear_quantize(bounds[3], t.h), r.minv);
  wr_tex_linear_partition(t, r, body_len, tile_rel, coop);
}

// blendTextureLinearR8 (swgl_ext.h:634-650): R8 atlas through the fallback
// bilinear filter, expanded to four lanes (glyph blit into an RGBA8 target).
WRD_SHARED void wr_tex_row_setup_r8(const TexView& t, const float* bounds, int body_len, const float* u, const float* v,
                             int tile_rel, TexRow& r) {
  r.mode = TEX_NONE;
  r.body_len = 0;
  if (body_len == 0 || t.fmt != WRCU_FMT_R8 || t.w < 2) return;
  r.mode = TEX_LINEAR_R8;
  r.body_len = body_len;
  r.before = r.inside = 0;
  for (int j = 0; j < 4; j++) {
    r.qu[j] = wr_linear_quantize(u[j], t.w);
    r.qv[j] = wr_linear_quantize(v[j], t.h);
  }
  r.ustep = 4.0f * (r.qu[1] - r.qu[0]);
  r.vstep = 4.0f * (r.qv[1] - r.qv[0]);
  r.minu = wr_max(wr_linear_quantize(bounds[0], t.w), 0.0f);
  r.minv = wr_max(wr_linear_quantize(bounds[1], t.h), 0.0f);
  r.maxu = wr_max(wr_linear_quantize(bounds[2], t.w), r.minu);
  r.maxv = wr_max(wr_linear_quantize(bounds[3], t.h), r.minv);
  r.kb[2] = max(0, tile_rel >> 2);
  wr_tex_bases(r.qu, 0, r.qu, 0, r.qu, r.kb[2], r.ustep, r.qv, r.kb[2], r.vstep, r.bu, r.bv, true);
  r.exact = body_len <= 32 ? 0 : (wr_sum_exact(r.bu[2], r.ustep) ? 4 : 0) | (wr_sum_exact(r.bv, r.vstep) ? 8 : 0);
}

// Source texel (before colour modulation) of body pixel `rel` (0-based in the span).
WRD_SHARED Px wr_tex_body(const TexView& t, const TexRow& r, int rel) {
  if (r.mode == TEX_NEAREST_FAST) {
    int sx = min(max(r.nix + rel, r.nminx), r.nmaxx);
    return wr_tex_load_any(t, r.nry, sx);
  }
  int j = rel & 3;
  if (r.mode == TEX_LINEAR_R8) {
    const int m = (rel >> 2) - r.kb[2];
    float qu = wr_sum_at(r.bu[2][j], r.ustep, m, r.exact & 4), qv = wr_sum_at(r.bv[j], r.vstep, m, r.exact & 8);
    int rr = wr_texture_linear_r8(t, (int)wr_clamp(qu, r.minu, r.maxu), (int)wr_clamp(qv, r.minv, r.maxv));
    return Px{rr, rr, rr, rr};
  }
  if (r.mode == TEX_NEAREST_FALLBACK) {
    float su = r.bu[2][j], sv = r.bv[j];
    if (!r.nsolid) {
      const int m = (rel >> 2) - r.kb[2];
      su = wr_sum_at(su, r.ustep, m, r.exact & 4);
      sv = wr_sum_at(sv, r.vstep, m, r.exact & 8);
    }
    int ix = (int)wr_clamp(su, r.minu, r.maxu), iy = (int)wr_clamp(sv, r.minv, r.maxv);
    int cx = wr_clamp_coord(ix, t.w), cy = wr_clamp_coord(iy, t.h);
    return wr_tex_load_any(t, cy, cx);
  }
  // TEX_LINEAR
  if (rel < r.before || rel >= r.before + r.inside) {
    // fallback filter (swgl_ext.h:172-184): uv += uv_step per chunk
    float qu, qv;
    if (rel < r.before) {
      qu = wr_sum_at(r.bu[0][j], r.ustep, (rel >> 2) - r.kb[0], r.exact & 1);
      qv = r.qv[j];  // prefix exists only for constant-y filters
    } else {
      int p = rel - r.before - r.inside;
      qu = wr_sum_at(r.bu[2][j], r.ustep, (p >> 2) - r.kb[2], r.exact & 4);
      qv = wr_sum_at(r.bv[j], r.vstep, (p >> 2) - r.kb[2], r.exact & 8);
    }
    return wr_tex_linear_any(t, (int)wr_clamp(qu, r.minu, r.maxu), (int)wr_clamp(qv, r.minv, r.maxv));
  }
  int p = rel - r.before;
  if (r.filter == LF_UPSCALE) {
    float qu = wr_sum_at(r.bu[1][j], r.ustep, (p >> 2) - r.kb[1], r.exact & 2);
    int ix = (p < 4) ? (int)wr_clamp(qu, r.minu, r.maxu) : (int)qu;
    return wr_tex_linear_any(t, ix, r.uiy0);
  }
  // FAST / DOWNSCALE (swgl_ext.h:284-371): integer texel stepping, constant fractions
  int mul = r.filter == LF_DOWNSCALE ? 2 : 1;
  if (t.fmt == WRCU_FMT_R8) {
    const uint8_t* r0 = t.ptr + (size_t)r.fcy * t.pitch + (size_t)(r.fcx + p * mul);
    const uint8_t* r1 = r0 + (r.fnext ? t.pitch : 0);
    int l = wr_lerp7(__ldg(r0), __ldg(r1), r.ffy);
    int rr = wr_lerp7(__ldg(r0 + 1), __ldg(r1 + 1), r.ffy);
    return Px{0, 0, wr_lerp7(l, rr, r.ffx) & 0xFFFF, 0};
  }
  const uint8_t* row0 = t.ptr + (size_t)r.fcy * t.pitch + (size_t)(r.fcx + p * mul) * 4;
  const uint8_t* row1 = row0 + (r.fnext ? t.pitch : 0);
  uint2 a = make_uint2(__ldg((const uint32_t*)row0), __ldg((const uint32_t*)row0 + 1));
  uint2 b = make_uint2(__ldg((const uint32_t*)row1), __ldg((const uint32_t*)row1 + 1));
  int vv[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int a0 = (a.x >> (8 * k)) & 0xFF, a1 = (b.x >> (8 * k)) & 0xFF;
    int b0 = (a.y >> (8 * k)) & 0xFF, b1 = (b.y >> (8 * k)) & 0xFF;
    int l = wr_lerp7(a0, a1, r.ffy);
    int rr = wr_lerp7(b0, b1, r.ffy);
    vv[k] = wr_lerp7(l, rr, r.ffx) & 0xFFFF;
  }
  return Px{vv[0], vv[1], vv[2], vv[3]};
}

// texture(sampler2D, vec2) → float RGBA (texture.h:948-975), uv already clamped
WRD_SHARED void wr_tex_fragment(const TexView& t, float cu, float cv, float* out) {
  if (t.fmt == WRCU_FMT_R8) {
    int rr;
    if (t.filter == WRCU_LINEAR) {
      rr = wr_texture_linear_r8(t, (int)wr_linear_quantize(cu, t.w), (int)wr_linear_quantize(cv, t.h));
    } else {
      int x = wr_clamp_coord((int)(cu * (float)t.w), t.w), y = wr_clamp_coord((int)(cv * (float)t.h), t.h);
      rr = __ldg(t.ptr + (size_t)y * t.pitch + x);
    }
    out[0] = (float)rr * (1.0f / 255.0f);
    out[1] = 0.0f; out[2] = 0.0f; out[3] = 1.0f;
    return;
  }
  if (t.filter == WRCU_LINEAR) {
    Px p = wr_texture_linear_rgba8(t, (int)wr_linear_quantize(cu, t.w), (int)wr_linear_quantize(cv, t.h));
    out[0] = (float)p.r * (1.0f / 255.0f); out[1] = (float)p.g * (1.0f / 255.0f);
    out[2] = (float)p.b * (1.0f / 255.0f); out[3] = (float)p.a * (1.0f / 255.0f);
  } else {
    int x = wr_clamp_coord((int)(cu * (float)t.w), t.w), y = wr_clamp_coord((int)(cv * (float)t.h), t.h);
    uint32_t p = __ldg((const uint32_t*)(t.ptr + (size_t)y * t.pitch) + x);
    out[0] = (float)((p >> 16) & 0xFF) * (1.0f / 255.0f); out[1] = (float)((p >> 8) & 0xFF) * (1.0f / 255.0f);
    out[2] = (float)(p & 0xFF) * (1.0f / 255.0f); out[3] = (float)(p >> 24) * (1.0f / 255.0f);
  }
}
