// shader_cs_gradient.cuh — the cached gradient render tasks drawn by
// draw_texture_cache_target (renderer/mod.rs:4085-4183):
//   cs_fast_linear_gradient (webrender/res/cs_fast_linear_gradient.glsl)
//   cs_linear_gradient      (cs_linear_gradient.glsl; GradientShader with tileRepeat off)
//   cs_radial_gradient      (cs_radial_gradient.glsl; span: swgl_commitRadialGradientRGBA8,
//                            swgl/src/swgl_ext.h:1628-1837)
//   cs_conic_gradient       (cs_conic_gradient.glsl; fragment shader only)
//
// CmdCold layout for these kinds:
//   linear: f[0..1] v_scale_dir, f[2] v_start_offset          radial: f[0] v_start_radius
//   conic:  f[0..1] v_center, f[2] v_start_offset, f[4] v_angle, f[5] v_offset_scale
//   all:    f[3] v_gradient_repeat, i[0] v_gradient_address, i[1] table valid,
//           i[2] tileRepeat off, g[0..4] merge mask
//   fast linear: g[0..3] vColor0, g[4..7] vColor1
#pragma once
#include "shader_gradient.cuh"
#include "setup_common.cuh"

// float -> int the way the reference's x86 build converts (cvttss2si: NaN and
// out-of-range give INT_MIN)
WRD int wr_f2i_x86(float v) {
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return (int)0x80000000;
  return (int)v;
}

// sample_gradient (gradient.glsl:45-61) for one pixel
WRD void wr_grad_sample_float(const RasterArgs& a, const CmdCold& k, float offset, float* out) {
  offset = offset - floorf(offset) * k.f[3];
  float xx = wr_clamp(1.0f + offset * GRAD_SIZE, 0.0f, 1.0f + GRAD_SIZE);
  float ei = floorf(xx), ef = xx - ei;
  int addr = k.i[0] + 2 * (int)ei;
  float4 t0 = wr_grad_texel(a, addr, 0), t1 = wr_grad_texel(a, addr, 1);
  out[0] = t0.x + t1.x * ef;
  out[1] = t0.y + t1.y * ef;
  out[2] = t0.z + t1.z * ef;
  out[3] = t0.w + t1.w * ef;
}
WRD Px wr_grad_pack(const float* c) {
  Px o;
  o.r = wr_round_pixel(c[0], 255.0f) & 0xFFFF;
  o.g = wr_round_pixel(c[1], 255.0f) & 0xFFFF;
  o.b = wr_round_pixel(c[2], 255.0f) & 0xFFFF;
  o.a = wr_round_pixel(c[3], 255.0f) & 0xFFFF;
  return o;
}
WRD Px wr_grad_fragment(const RasterArgs& a, const CmdCold& k, float offset) {
  float c[4];
  wr_grad_sample_float(a, k, offset, c);
  return wr_grad_pack(c);
}
// ps_quad.glsl:406-417 main() around a gradient pattern_fragment: v_color * sample, .rrrr for mask quads
// (base colour in g[8..11])
WRD Px wr_quad_grad_fragment(const RasterArgs& a, const CmdHot& c, const CmdCold& k, float offset) {
  float smp[4], col[4];
  wr_grad_sample_float(a, k, offset, smp);
#pragma unroll
  for (int ch = 0; ch < 4; ch++) col[ch] = (k.g[8 + ch] * 1.0f) * smp[ch];
  if (c.flags & CMD_OUT_RRRR) col[1] = col[2] = col[3] = col[0];
  return wr_grad_pack(col);
}

// ---- cs_fast_linear_gradient: mix(vColor0, vColor1, vPos), no span shader --------
struct FastLinearShader {
  struct Row {
    float o[1], step[1];
    float base[4][1];
    int kb;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<1>(a, k, c, y, r.o, r.step);
    r.kb = wr_chunk_base<1>(a, r.o, r.step, c, tx0, r.base);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    int rel = x - c.x0;
    float t[1];
    wr_chunk_lane<1>(a, r.base, r.step, r.kb, rel >> 2, rel & 3, t);
    Px o;
    o.r = wr_round_pixel((k.g[4] - k.g[0]) * t[0] + k.g[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel((k.g[5] - k.g[1]) * t[0] + k.g[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel((k.g[6] - k.g[2]) * t[0] + k.g[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel((k.g[7] - k.g[3]) * t[0] + k.g[3], 255.0f) & 0xFFFF;
    return o;
  }
};

// ---- cs_conic_gradient: fragment shader only ---------------------------------------
struct ConicShader {
  struct Row {
    float o[2], step[2];
    float base[4][2];
    int kb;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    r.kb = wr_chunk_base<2>(a, r.o, r.step, c, tx0, r.base);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    int rel = x - c.x0;
    float p[2];
    wr_chunk_lane<2>(a, r.base, r.step, r.kb, rel >> 2, rel & 3, p);
    float cx = p[0] - k.f[0], cy = p[1] - k.f[1];
#ifdef WRCU_HOSTEMU
    float at = atan2f(cy, cx);
#else
    // the reference calls libm's atan2f (glsl.h:2838-2843); the double-precision
    // result rounded once is the correctly rounded value it approximates
    float at = (float)atan2((double)cy, (double)cx);
#endif
    float angle = at + k.f[4];
    float offset = wr_fract(angle / (2.0f * 3.141592653589793f)) * k.f[5] - k.f[2];
    return wr_grad_fragment(a, k, offset);
  }
};

// ---- cs_radial_gradient ---------------------------------------------------------------
// The reference's span routine walks the row chunk by chunk: dot(pos,pos) is a
// second-order running sum, and the merged-stop runs it finds depend on that
// state, so the walk itself is the specification.  row_setup replays it once per
// (command,row,tile), warp-uniformly: every lane advances the same four running
// sums, and only the lane that owns a pixel evaluates its colour (sqrt + LUT or
// ramp).  Chunks left of the tile cost the two additions per lane of the
// recurrence; the walk stops at the tile's right edge.
struct RadialShader {
#ifdef WRCU_HOSTEMU
  enum { NOUT = WRCU_TILE_W };
#else
  enum { NOUT = 4 };
#endif
  struct Row {
    float o[2], step[2];
    int body_len, own0;  // own0: x of the first pixel whose colour `out` holds
    Px out[NOUT];
  };
  WRD_MEMBER void emit(Row& r, int ax, const Px& v) {
    int idx = ax - r.own0;
    if (idx >= 0 && idx < NOUT) r.out[idx] = v;
  }
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    int len = c.x1 - c.x0;
    r.body_len = (rgba && len >= 4 && k.i[1] != 0) ? (len & ~3) : 0;
#ifdef WRCU_HOSTEMU
    r.own0 = tx0;
#else
    r.own0 = tx0 + (threadIdx.x & 31) * 4;
#endif
    if (!r.body_len) return;
    const int span = r.body_len;
    const int first = max(tx0, (int)c.x0) - (int)c.x0;
    const int last = min(span, tx0 + WRCU_TILE_W - (int)c.x0);
    if (first >= span) return;
    const float4* stops = a.gbuf_f + k.i[0];
    const float radius = k.f[0];
    const bool repeat = k.f[3] != 0.0f;
    const float size = GRAD_SIZE;
    float px[4], py[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float p[2];
      wr_interp_at<2>(a, r.o, r.step, j, p);
      px[j] = p[0];
      py[j] = p[1];
    }
    float dx = px[1] - px[0], dy = py[1] - py[0];
    float deltaDelta = dx * dx + dy * dy;
    if (!isfinite(deltaDelta) || !isfinite(radius)) { r.body_len = 0; return; }
    float invDelta, middleT, middleB;
    if (deltaDelta > 0.0f) {
      invDelta = 1.0f / deltaDelta;
      middleT = -(dx * px[0] + dy * py[0]) * invDelta;
      middleB = middleT * middleT - (px[0] * px[0] + py[0] * py[0]) * invDelta;
    } else {
      invDelta = 0.0f;
      middleT = (float)span;
      middleB = 0.0f;
    }
    float mer[2];
    {
      float f[2] = {middleT, (float)span};
#pragma unroll
      for (int j = 0; j < 2; j++) {
        float vx = px[0] + dx * f[j], vy = py[0] + dy * f[j];
        mer[j] = sqrtf(wr_max(vx * vx + vy * vy, 1.0e-12f));
      }
    }
    const float middleRadius = (float)span < middleT ? mer[1] : mer[0];
    const float endRadius = mer[1];
    dx *= 4.0f;
    dy *= 4.0f;
    deltaDelta *= 16.0f;
    float dotPos[4], dotPosDelta[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      dotPos[j] = px[j] * px[j] + py[j] * py[j];
      dotPosDelta[j] = 2.0f * (px[j] * dx + py[j] * dy) + deltaDelta;
    }
    const float deltaDelta2 = 2.0f * deltaDelta;
    for (int t = 0; t < last;) {
      float offset[4];
#pragma unroll
      for (int j = 0; j < 4; j++) offset[j] = sqrtf(wr_max(dotPos[j], 1.0e-12f)) - radius;
      float startRadius = radius;
      if (repeat) {
        startRadius += offset[0];
#pragma unroll
        for (int j = 0; j < 4; j++) offset[j] = wr_fract(offset[j]);
        startRadius -= offset[0];
      }
      float intercept = -1.0f;
      int minIndex = 0, maxIndex = (int)(1.0f + size);
      const bool past = (float)t >= middleT;
      if (offset[0] < 0.0f) {
        maxIndex = minIndex;
        if (past) intercept = radius;
      } else if (offset[0] < 1.0f) {
        minIndex = (int)(1.0f + offset[0] * size);
        maxIndex = minIndex;
        float searchOffset = (past ? endRadius : middleRadius) - startRadius;
        int searchIndex = (int)wr_clamp(1.0f + size * searchOffset, 1.0f, size);
        if (past) {
          // while (maxIndex + 1 <= searchIndex && can_merge(maxIndex, maxIndex + 1)) maxIndex++
          maxIndex = min(wr_merge_run_up(k, minIndex), max(minIndex, searchIndex));
          intercept = (float)(maxIndex + 1);
        } else {
          // while (minIndex - 1 >= searchIndex && can_merge(minIndex - 1, minIndex)) minIndex--
          minIndex = max(wr_merge_run_down(k, maxIndex), min(maxIndex, searchIndex));
          intercept = (float)minIndex;
        }
        intercept = wr_clamp((intercept - 1.0f) / size, 0.0f, 1.0f) + startRadius;
      } else {
        minIndex = maxIndex;
        if (!past) intercept = radius + 1.0f;
      }
      float endT = past ? (float)span : (float)min(span, wr_f2i_x86(middleT));
      if (intercept >= 0.0f) {
        float b = middleB + intercept * intercept * invDelta;
        if (b > 0.0f) {
          b = sqrtf(b);
          endT = wr_min(endT, past ? middleT + b : middleT - b);
        } else {
          endT = wr_min(endT, middleT);
        }
      }
      if ((float)t + 4.0f <= endT) {
        int inside = wr_f2i_x86(endT - (float)t) & ~3;
        float4 mn = __ldg(stops + 2 * minIndex);
        float4 mx = __ldg(stops + 2 * maxIndex), ms = __ldg(stops + 2 * maxIndex + 1);
        const float minC[4] = {mn.z * 255.0f, mn.y * 255.0f, mn.x * 255.0f, mn.w * 255.0f};
        const float maxC[4] = {(mx.z + ms.z) * 255.0f, (mx.y + ms.y) * 255.0f, (mx.x + ms.x) * 255.0f,
                               (mx.w + ms.w) * 255.0f};
        float colorF[4], dCF[4];
        const float sc = size / (float)(maxIndex + 1 - minIndex);
        const float at = startRadius + (float)(minIndex - 1) / size;
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
          dCF[ch] = (maxC[ch] - minC[ch]) * sc;
          colorF[ch] = minC[ch] - dCF[ch] * at;
        }
        for (int e = 0; e < inside; e += 4) {
          const int ax0 = (int)c.x0 + t + e;
          if (ax0 + 4 > r.own0 && ax0 < r.own0 + NOUT) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              float og = sqrtf(dotPos[j]);
              Px s;
              s.b = wr_round_pixel(colorF[0] + dCF[0] * og, 1.0f) & 0xFFFF;
              s.g = wr_round_pixel(colorF[1] + dCF[1] * og, 1.0f) & 0xFFFF;
              s.r = wr_round_pixel(colorF[2] + dCF[2] * og, 1.0f) & 0xFFFF;
              s.a = wr_round_pixel(colorF[3] + dCF[3] * og, 1.0f) & 0xFFFF;
              emit(r, ax0 + j, s);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            dotPos[j] = dotPos[j] + dotPosDelta[j];
            dotPosDelta[j] = dotPosDelta[j] + deltaDelta2;
          }
          if (t + e + 4 >= last) return;  // the tile's pixels are all emitted
        }
        t += inside;
        if (t >= span) break;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          offset[j] = sqrtf(wr_max(dotPos[j], 1.0e-12f)) - radius;
          if (repeat) offset[j] = wr_fract(offset[j]);
        }
      }
      {
        const int ax0 = (int)c.x0 + t;
        if (ax0 + 4 > r.own0 && ax0 < r.own0 + NOUT) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            float entry = wr_clamp(offset[j] * size + 1.0f, 0.0f, 1.0f + size);
            emit(r, ax0 + j, wr_grad_sample_entry(stops, entry));
          }
        }
      }
      t += 4;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        dotPos[j] = dotPos[j] + dotPosDelta[j];
        dotPosDelta[j] = dotPosDelta[j] + deltaDelta2;
      }
    }
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    int rel = x - c.x0;
    if (rel < r.body_len) return r.out[x - r.own0];
    float p[2];
    wr_interp_at<2>(a, r.o, r.step, rel, p);
    float offset = sqrtf(p[0] * p[0] + p[1] * p[1]) - k.f[0];
    if (k.i[3]) return wr_quad_grad_fragment(a, c, k, offset);  // ps_quad_radial_gradient
    return wr_grad_fragment(a, k, offset);
  }
};

// ---- ps_quad_conic_gradient (ps_quad_conic_gradient.glsl:67-92): fragment shader only --
// if_then_else(c, a, b) is mix(b, a, c) (shared.glsl:205): (a - b) * c + b, not a select
WRD float wr_approx_atan2(float y, float x) {
  float ax = fabsf(x), ay = fabsf(y);
  float slope = wr_min(ax, ay) / wr_max(ax, ay);
  float s2 = slope * slope;
  float r = ((-0.0464964749f * s2 + 0.15931422f) * s2 - 0.327622764f) * s2 * slope + slope;
  float t = 1.57079637f - r;
  r = (t - r) * (ay > ax ? 1.0f : 0.0f) + r;
  t = 3.14159274f - r;
  r = (t - r) * (x < 0.0f ? 1.0f : 0.0f) + r;
  return r * copysignf(1.0f, y);
}
struct QuadConicShader {
  struct Row {
    float o[2], step[2];
    float base[4][2];
    int kb;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    r.kb = wr_chunk_base<2>(a, r.o, r.step, c, tx0, r.base);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    int rel = x - c.x0;
    float p[2];
    wr_chunk_lane<2>(a, r.base, r.step, r.kb, rel >> 2, rel & 3, p);
    float angle = wr_approx_atan2(p[1], p[0]) + k.f[4];
    float offset = wr_fract(angle / (2.0f * 3.141592653589793f)) * k.f[5] - k.f[2];
    return wr_quad_grad_fragment(a, c, k, offset);
  }
};

// ---- vertex stage of the four programs ----------------------------------------------------
WRD void wr_setup_cs_gradient_one(const SetupArgs& a, int idx) {
  const float* f = (const float*)(a.instances + (size_t)idx * a.stride);
  const int* fi = (const int*)f;
  const FrameTablesDev& T = a.tabs;
  const int kind = a.kind;
  const float tsx = f[2] - f[0], tsy = f[3] - f[1];
  QuadOut q;
  memset(&q, 0, sizeof q);
  float fc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, gc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int address = 0;
  float rscale = 0.0f;
  switch (kind) {
    case WRCU_KIND_FAST_LINEAR_GRADIENT:
      for (int i = 0; i < 8; i++) gc[i] = f[4 + i];
      break;
    case WRCU_KIND_LINEAR_GRADIENT: {
      float dirx = f[6] - f[4], diry = f[7] - f[5];
      float dd = dirx * dirx + diry * diry;
      float sdx = dirx / dd, sdy = diry / dd;
      fc[2] = f[4] * sdx + f[5] * sdy;
      fc[0] = sdx * tsx;
      fc[1] = sdy * tsy;
      fc[3] = (float)(fi[10] == 1);
      address = fi[11];
      break;
    }
    case WRCU_KIND_RADIAL_GRADIENT: {
      float rd = f[9] - f[8];
      rscale = rd != 0.0f ? 1.0f / rd : 0.0f;
      fc[0] = f[8] * rscale;
      fc[3] = (float)(fi[11] == 1);
      address = fi[12];
      break;
    }
    default: {
      float dd = f[9] - f[8];
      rscale = dd != 0.0f ? 1.0f / dd : 0.0f;
      fc[5] = rscale;
      fc[4] = 3.141592653589793f / 2.0f - f[10];
      fc[2] = f[8] * rscale;
      fc[0] = f[4] * rscale;
      fc[1] = f[5] * rscale;
      fc[3] = (float)(fi[11] == 1);
      address = fi[12];
      break;
    }
  }
  const float axs[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ays[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int v = 0; v < 4; v++) {
    float ax = axs[v], ay = ays[v];
    float px = (f[2] - f[0]) * ax + f[0], py = (f[3] - f[1]) * ay + f[1];
    q.pos[v] = wr_mat_mul(a.tgt.proj, make_float4(px, py, 0.0f, 1.0f));
    switch (kind) {
      case WRCU_KIND_FAST_LINEAR_GRADIENT:
        q.interp[v][0] = (1.0f - 0.0f) * ((ay - ax) * f[12] + ax) + 0.0f;
        break;
      case WRCU_KIND_LINEAR_GRADIENT:
        q.interp[v][0] = ax * f[8];
        q.interp[v][1] = ay * f[9];
        break;
      case WRCU_KIND_RADIAL_GRADIENT:
        q.interp[v][0] = (tsx * ax * f[6] - f[4]) * rscale;
        q.interp[v][1] = (tsy * ay * f[7] - f[5]) * rscale;
        q.interp[v][1] *= f[10];
        break;
      default:
        q.interp[v][0] = tsx * ax * rscale * f[6];
        q.interp[v][1] = tsy * ay * rscale * f[7];
        break;
    }
  }
  q.n_interp = kind == WRCU_KIND_FAST_LINEAR_GRADIENT ? 1 : 2;
  uint32_t merge[5] = {0, 0, 0, 0, 0};
  bool valid = false;
  if (kind != WRCU_KIND_FAST_LINEAR_GRADIENT) valid = wr_grad_validate_merge(T, address, merge);
  float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  wr_pack_color(q, white);
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 8; i++) k->f[i] = fc[i];
    if (kind == WRCU_KIND_FAST_LINEAR_GRADIENT) {
      for (int i = 0; i < 8; i++) k->g[i] = gc[i];
    } else {
      for (int i = 0; i < 5; i++) k->g[i] = __uint_as_float(merge[i]);
    }
    k->i[0] = address;
    k->i[1] = valid ? 1 : 0;
    k->i[2] = 1;
    k->i[3] = 0;
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_cs_gradient)

template <> struct WrRun<RadialShader> {  // ps_quad_radial_gradient
  enum { n = 2 };
  WRD_MEMBER int drawn(const RadialShader::Row& r) { return r.body_len; }
};
