// shader_gradient.cuh — brush_linear_gradient [ALPHA_PASS]
// (webrender/res/brush_linear_gradient.glsl:18-91, gradient_shared.glsl,
// gradient.glsl:45-61).  Span body: swgl_commitLinearGradientRGBA8
// (swgl/src/swgl_ext.h:1389-1604); tail: sample_gradient in float.
//
// The reference's span routine is a sequential walk along the row: it merges
// LUT entries with equal colour steps into runs, steps a 16-bit colour across
// each run, and samples one chunk per-pixel between runs.  Run boundaries
// depend on float state accumulated along the walk, so the walk itself is
// replayed here (wr_grad_iter) — but a run is usually the whole row (two-stop
// gradients merge all 128 entries), so the replay is one or two iterations.
// The per-entry "can merge with next" test is precomputed by the setup kernel
// into a 130-bit mask, turning the reference's entry-by-entry scan into a
// find-first-zero.
#pragma once
#include "raster.cuh"

// CmdCold for linear gradients:
//  f[0..1] v_scale_dir, f[2] v_start_offset, f[3] v_gradient_repeat,
//  i[0] v_gradient_address, i[1] table passes swgl_validateGradient,
//  i[2] != 0: tileRepeat off (cs_linear_gradient: v_pos is not wrapped to [0,1)),
//  g[0..4] merge mask (bit e: step[e] == step[e+1]), as float bit patterns
#define GRAD_SIZE 128.0f

WRD float wr_fract(float v) { return __fsub_rn(v, floorf(v)); }

// texelFetch in the 1024-texel-wide data texture the LUT lives in; x clamps to
// its row (clamp2D, texture.h:268-275), rows past the end read the last row
WRD float4 wr_grad_texel(const RasterArgs& a, int addr, int dx) {
  if (a.n_gbuf_f <= 0) return make_float4(0, 0, 0, 0);
  int x = (int)((uint32_t)addr % 1024U), yrow = (int)((uint32_t)addr / 1024U);
  x = min(max(x + dx, 0), 1023);
  int rows = (a.n_gbuf_f + 1023) / 1024;
  if (yrow >= rows) yrow = rows - 1;
  int idx = yrow * 1024 + x;
  if (idx >= a.n_gbuf_f) return make_float4(0, 0, 0, 0);  // zero padding of the last row
  return __ldg(a.gbuf_f + idx);
}

WRD bool wr_merge_bit(const CmdCold& k, int e) {
  return (__float_as_uint(k.g[e >> 5]) >> (e & 31)) & 1u;
}
// largest m >= s with bits s..m-1 all set
WRD int wr_merge_run_up(const CmdCold& k, int s) {
  int w = s >> 5, b = s & 31;
  uint32_t v = ~__float_as_uint(k.g[w]) >> b;
  while (v == 0) {
    s += 32 - b;
    if (++w >= 5) return 129;
    b = 0;
    v = ~__float_as_uint(k.g[w]);
  }
  return min(s + (__ffs((int)v) - 1), 129);
}
// smallest m <= s with bits m..s-1 all set
WRD int wr_merge_run_down(const CmdCold& k, int s) {
  int m = s;
  while (m > 0 && wr_merge_bit(k, m - 1)) {
    // whole words of ones are skipped at once
    int e = m - 1, b = e & 31;
    uint32_t word = __float_as_uint(k.g[e >> 5]);
    uint32_t inv = ~word & (b == 31 ? 0xFFFFFFFFu : ((1u << (b + 1)) - 1u));
    if (inv == 0) { m = e - b; continue; }
    m = (e & ~31) + (32 - __clz((int)inv));
    break;
  }
  return m;
}

struct GradWalk {
  float px[4], py[4];  // v_pos lanes at the iteration start
  int x;               // span-relative pixel the iteration starts at
  int span;            // pixels of the span body still to draw
};
struct GradRun {
  int inside;          // whole chunks drawn by colour stepping
  float colorF[4], dC[4];   // BGRA, 0..0xFF00 scale
  uint32_t delta01, delta23;  // deltaColor packed 2x16
};
struct GradRowConst {
  float psx, psy, delta, dcx0, dcx1, dcy0, dcy1;
};

WRD void wr_grad_offsets(const CmdCold& k, const GradWalk& w, float* off, float* rx0, float* ry0) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float rx = k.i[2] ? w.px[j] : wr_fract(w.px[j]), ry = k.i[2] ? w.py[j] : wr_fract(w.py[j]);
    if (j == 0) { *rx0 = rx; *ry0 = ry; }
    float o = rx * k.f[0] + ry * k.f[1] - k.f[2];
    if (k.f[3] != 0.0f) o = wr_fract(o);
    off[j] = o;
  }
}

// One iteration of the `while (span > 0)` loop of commitLinearGradient up to and
// including the colour set-up of its merged run (swgl_ext.h:1440-1557).
// Leaves w untouched; off = offsets at the iteration start.
WRD void wr_grad_iter(const RasterArgs& a, const CmdCold& k, const GradRowConst& rc, const GradWalk& w,
                      GradRun& run, float* off) {
  const float4* stops = a.gbuf_f + k.i[0];
  float chunks = 0.25f * (float)w.span;
  float rx0, ry0;
  wr_grad_offsets(k, w, off, &rx0, &ry0);
  chunks = wr_min(chunks, rc.dcx0 - rx0 * rc.dcx1);
  chunks = wr_min(chunks, rc.dcy0 - ry0 * rc.dcy1);
  const float delta = rc.delta;
  float startEntry;
  int minIndex, maxIndex;
  if (off[0] < 0.0f) {
    startEntry = 0.0f;
    minIndex = maxIndex = 0;
    if (delta > 0.0f) chunks = wr_min(chunks, -off[0] / delta);
  } else if (off[0] < 1.0f) {
    startEntry = 1.0f + off[0] * GRAD_SIZE;
    if (delta < 0.0f) chunks = wr_min(chunks, -off[0] / delta);
    else if (delta > 0.0f) chunks = wr_min(chunks, (1.0f - off[0]) / delta);
    float endEntry = wr_clamp(1.0f + (off[0] + delta * (float)(int)chunks) * GRAD_SIZE, 0.0f, 1.0f + GRAD_SIZE);
    minIndex = (int)startEntry;
    maxIndex = minIndex;
    if (delta > 0.0f) {
      // while (maxIndex + 1 < endEntry && can_merge(maxIndex, maxIndex + 1)) maxIndex++
      int lim = (int)ceilf(endEntry) - 1;  // largest m with m < endEntry
      maxIndex = min(wr_merge_run_up(k, minIndex), max(minIndex, lim));
      chunks = wr_min(chunks, ((float)(maxIndex + 1) - startEntry) / (delta * GRAD_SIZE));
    } else if (delta < 0.0f) {
      // while (minIndex - 1 > endEntry && can_merge(minIndex - 1, minIndex)) minIndex--
      int lim = (int)floorf(endEntry) + 1;  // smallest m with m > endEntry
      minIndex = max(wr_merge_run_down(k, maxIndex), min(maxIndex, lim));
      chunks = wr_min(chunks, ((float)minIndex - startEntry) / (delta * GRAD_SIZE));
    }
  } else {
    startEntry = 1.0f + GRAD_SIZE;
    minIndex = maxIndex = (int)startEntry;
    if (delta < 0.0f) chunks = wr_min(chunks, (1.0f - off[0]) / delta);
  }
  run.inside = 0;
  if (chunks >= 1.0f) {
    run.inside = (int)chunks;
    float4 mn = __ldg(stops + 2 * minIndex);
    float4 mx = __ldg(stops + 2 * maxIndex), ms = __ldg(stops + 2 * maxIndex + 1);
    float minC[4] = {mn.z, mn.y, mn.x, mn.w};
    float maxC[4] = {mx.z + ms.z, mx.y + ms.y, mx.x + ms.x, mx.w + ms.w};
    float inv = 1.0f / (float)(maxIndex + 1 - minIndex);
    uint32_t dl[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      float lo = minC[c] * (float)0xFF00, hi = maxC[c] * (float)0xFF00;
      float range = (hi - lo) * inv;
      run.colorF[c] = lo + range * (startEntry - (float)minIndex) + (float)0x80;
      run.dC[c] = range * (delta * GRAD_SIZE);
      dl[c] = (uint32_t)wr_round_pixel(run.dC[c], 1.0f) & 0xFFFFu;
    }
    run.delta01 = dl[0] | (dl[1] << 16);
    run.delta23 = dl[2] | (dl[3] << 16);
  }
}

// advance the walk over the run's stepped chunks (swgl_ext.h:1578-1592)
WRD void wr_grad_skip_inside(const GradRowConst& rc, GradWalk& w, int inside) {
  w.span -= inside * 4;
  w.x += inside * 4;
  float fi = (float)inside;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    w.px[j] = w.px[j] + rc.psx * fi;
    w.py[j] = w.py[j] + rc.psy * fi;
  }
}
// ...and over the per-pixel sampled chunk that follows (swgl_ext.h:1594-1602)
WRD void wr_grad_skip_sampled(const GradRowConst& rc, GradWalk& w) {
  w.span -= 4;
  w.x += 4;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    w.px[j] = w.px[j] + rc.psx;
    w.py[j] = w.py[j] + rc.psy;
  }
}

// sampleGradient (swgl_ext.h:1350-1371) for one lane
WRD Px wr_grad_sample_entry(const float4* stops, float entry) {
  int index = (int)entry;
  float offset = entry - (float)index;
  float4 s = __ldg(stops + 2 * index), d = __ldg(stops + 2 * index + 1);
  Px o;
  o.r = wr_round_pixel(s.x + d.x * offset, 255.0f) & 0xFFFF;
  o.g = wr_round_pixel(s.y + d.y * offset, 255.0f) & 0xFFFF;
  o.b = wr_round_pixel(s.z + d.z * offset, 255.0f) & 0xFFFF;
  o.a = wr_round_pixel(s.w + d.w * offset, 255.0f) & 0xFFFF;
  return o;
}

// colour of chunk cch, lane j of a stepped run: the 16-bit lanes restart from
// colorF every 64 chunks (swgl_ext.h:1558-1576); seg_done segments are already
// folded into colorF
WRD Px wr_grad_run_color(const GradRun& run, int cch, int j, int seg_done) {
  int seg = cch >> 6, within = cch & 63;
  const float lane_f = 0.25f * (float)j;
  int v[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    float cf = run.colorF[c];
    for (int s = seg_done; s < seg; s++) cf = cf + run.dC[c] * 64.0f;
    float lf = j == 0 ? cf : cf + run.dC[c] * lane_f;
    uint32_t dl = ((c < 2 ? run.delta01 : run.delta23) >> (16 * (c & 1))) & 0xFFFFu;
    uint32_t c16 = ((uint32_t)wr_round_pixel(lf, 1.0f) + (uint32_t)within * dl) & 0xFFFFu;
    v[c] = (int)(c16 >> 8);
  }
  return Px{v[0], v[1], v[2], v[3]};
}

struct GradientShader {
  struct Row {
    float o[2], step[2];
    int body_len;
    GradRowConst rc;
    GradWalk w;     // walk state at the iteration that reaches this tile
    GradRun run;    // that iteration's run
    float off[4];   // offsets at that iteration's start (for inside == 0)
    int seg_done;   // 64-chunk segments already folded into run.colorF
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    int len = c.x1 - c.x0;
    r.body_len = (rgba && len >= 4 && k.i[1] != 0) ? (len & ~3) : 0;
    r.seg_done = 0;
    if (!r.body_len) return;
    GradWalk& w = r.w;
    for (int j = 0; j < 4; j++) {
      float p[2];
      wr_interp_at<2>(a, r.o, r.step, j, p);
      w.px[j] = p[0];
      w.py[j] = p[1];
    }
    w.x = 0;
    w.span = r.body_len;
    GradRowConst& rc = r.rc;
    rc.psx = (w.px[1] - w.px[0]) * 4.0f;
    rc.psy = (w.py[1] - w.py[0]) * 4.0f;
    rc.delta = rc.psx * k.f[0] + rc.psy * k.f[1];
    if (!isfinite(rc.delta)) { r.body_len = 0; return; }
    rc.dcx0 = 0.25f * (float)r.body_len; rc.dcx1 = 0.0f;
    rc.dcy0 = rc.dcx0; rc.dcy1 = 0.0f;
    if (!k.i[2] && rc.psx != 0.0f) {
      float rr = 1.0f / rc.psx;
      rc.dcx0 = (rc.psx >= 0.0f ? 1.0f : 0.0f) * rr;
      rc.dcx1 = 1.0f * rr;
    }
    if (!k.i[2] && rc.psy != 0.0f) {
      float rr = 1.0f / rc.psy;
      rc.dcy0 = (rc.psy >= 0.0f ? 1.0f : 0.0f) * rr;
      rc.dcy1 = 1.0f * rr;
    }
    // walk to the iteration covering the first pixel this tile draws
    int first = max(tx0, (int)c.x0) - (int)c.x0;
    if (first >= r.body_len) return;
    for (;;) {
      wr_grad_iter(a, k, rc, w, r.run, r.off);
      int end = w.x + r.run.inside * 4 + 4;
      if (first < end) break;
      if (r.run.inside) wr_grad_skip_inside(rc, w, r.run.inside);
      wr_grad_skip_sampled(rc, w);
    }
    // fold whole 64-chunk segments before the tile into colorF (sequential float
    // adds, as the reference does between segments)
    if (r.run.inside && first > w.x) {
      int seg = ((first - w.x) >> 2) >> 6;
      for (int s = 0; s < seg; s++)
        for (int ch = 0; ch < 4; ch++) r.run.colorF[ch] = r.run.colorF[ch] + r.run.dC[ch] * 64.0f;
      r.seg_done = seg;
    }
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    int rel = x - c.x0;
    if (rel < r.body_len) {
      const float4* stops = a.gbuf_f + k.i[0];
      // common case: inside the run row_setup stopped at
      int in_end = r.w.x + r.run.inside * 4;
      if (rel < in_end) return wr_grad_run_color(r.run, (rel - r.w.x) >> 2, rel & 3, r.seg_done);
      GradWalk w = r.w;
      GradRun run = r.run;
      float off[4] = {r.off[0], r.off[1], r.off[2], r.off[3]};
      for (;;) {
        in_end = w.x + run.inside * 4;
        if (rel < in_end) return wr_grad_run_color(run, (rel - w.x) >> 2, rel & 3, 0);
        if (run.inside) {
          wr_grad_skip_inside(r.rc, w, run.inside);
          float rx0, ry0;
          wr_grad_offsets(k, w, off, &rx0, &ry0);
        }
        if (rel < in_end + 4) {
          float entry = wr_clamp(off[rel & 3] * GRAD_SIZE + 1.0f, 0.0f, 1.0f + GRAD_SIZE);
          Px s = wr_grad_sample_entry(stops, entry);
          return s;
        }
        wr_grad_skip_sampled(r.rc, w);
        wr_grad_iter(a, k, r.rc, w, run, off);
      }
    }
    // fragment path (brush_linear_gradient.glsl:73-91, gradient.glsl:45-61)
    float p[2];
    wr_interp_at<2>(a, r.o, r.step, rel, p);
    if (!k.i[2]) { p[0] = wr_fract(p[0]); p[1] = wr_fract(p[1]); }
    float offset = (p[0] * k.f[0] + p[1] * k.f[1]) - k.f[2];
    offset = offset - floorf(offset) * k.f[3];
    float xx = wr_clamp(1.0f + offset * GRAD_SIZE, 0.0f, 1.0f + GRAD_SIZE);
    float ei = floorf(xx), ef = xx - ei;
    int addr = k.i[0] + 2 * (int)ei;
    float4 t0 = wr_grad_texel(a, addr, 0), t1 = wr_grad_texel(a, addr, 1);
    Px o;
    o.r = wr_round_pixel((t0.x + t1.x * ef) * 1.0f, 255.0f) & 0xFFFF;
    o.g = wr_round_pixel((t0.y + t1.y * ef) * 1.0f, 255.0f) & 0xFFFF;
    o.b = wr_round_pixel((t0.z + t1.z * ef) * 1.0f, 255.0f) & 0xFFFF;
    o.a = wr_round_pixel((t0.w + t1.w * ef) * 1.0f, 255.0f) & 0xFFFF;
    return o;
  }
};

template <> struct WrRun<GradientShader> {
  enum { n = 2 };
  WRD_MEMBER int drawn(const GradientShader::Row& r) { return r.body_len; }
};

#ifndef WRCU_HOSTEMU
// Strip mode: the Row is the state of the reference's walk along the span (position, current run); a pixel further
// right continues that walk (source() walks forward from it), exactly as the reference does — so the Row of the tile
// to the left serves, and the walk from the span start to this tile is not repeated.
template <> struct WrRowReuse<GradientShader> {
  enum { v = 1 };
  WRD_MEMBER bool ok(const GradientShader::Row&) { return true; }
};
#endif
