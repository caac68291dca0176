// shader_blur.cuh — cs_blur ALPHA_TARGET / COLOR_TARGET (webrender/res/cs_blur.glsl):
// one direction of the separable Gaussian blur behind box-shadow masks and
// filter: blur().  Span body: blendGaussianBlur (swgl/src/swgl_ext.h:951-981) over
// gaussianBlurHorizontal / Vertical (swgl/src/texture.h:1165-1310) — integer
// accumulation in 16-bit lanes with a float-stepped coefficient; tail: main().
#pragma once
#include "raster.cuh"
#include "setup_common.cuh"

// CmdCold: f[0..3] vUvRect, f[4..5] vOffsetScale, f[6..7] vGaussCoefficients,
//          i[0] vSupport.x, i[1] COLOR_TARGET
WRD int wr_addsat16(int x, int y) {
  int r = (x + y) & 0xFFFF;
  return r < x ? 0xFFFF : r;
}

struct BlurShader {
  struct Row {
    float o[2], step[2];
    int drawn;          // pixels drawn by the span routine (whole chunks)
    int cx0, cy;        // texel of lane 0 at the span start
    int lo, hi, hori;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    r.drawn = 0;
    int len = c.x1 - c.x0;
    bool color = k.i[1] != 0;
    bool fmt_ok = (rgba && color && t.fmt == WRCU_FMT_RGBA8) || (!rgba && !color && t.fmt == WRCU_FMT_R8);
    if (!fmt_ok || len < 4) return;
    int span = len & ~3;
    float sw = (float)t.w, sh = (float)t.h;
    r.cx0 = (int)(r.o[0] * sw);
    r.cy = (int)(r.o[1] * sh);
    int b0 = (int)(k.f[0] * sw), b1 = (int)(k.f[1] * sh), b2 = (int)(k.f[2] * sw), b3 = (int)(k.f[3] * sh);
    r.hori = k.f[4] != 0.0f;
    r.lo = r.hori ? b0 : b1;
    r.hi = r.hori ? b2 : b3;
    int endX = min(min(b2, r.cx0 + span), (int)sw);
    int avail = endX - r.cx0;
    r.drawn = avail >= 4 ? (avail & ~3) : 0;
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool rgba) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    int rel = x - c.x0;
    int v[4] = {0, 0, 0, 0};
    if (rel < r.drawn) {
      const int bpp = t.fmt == WRCU_FMT_RGBA8 ? 4 : 1;
      const int p = rel & 3, ix = r.cx0 + (rel & ~3);
      float coeff = k.f[6] * (float)(1 << 8), coeffStep = k.f[7];
      const float coeffStep2 = k.f[7] * k.f[7];
      const int rx = wr_clamp_coord(ix, t.w - 1), ry = wr_clamp_coord(r.cy, t.h);
      const uint8_t* base = t.ptr + (size_t)ry * t.pitch + (size_t)rx * bpp;
      int c16 = (int)(uint16_t)(coeff + 0.5f);
      int sum[4] = {0, 0, 0, 0};
      for (int ch = 0; ch < bpp; ch++) sum[ch] = ((int)__ldg(base + p * bpp + ch) * c16) & 0xFFFF;
      const int radius = k.i[0];
      if (r.hori) {
        const int leftBound = ix - max(r.lo, 0);
        const int rightBound = min(r.hi, t.w - 1) - ix;
        const int validRadius = min(radius, min(leftBound, rightBound - 3));
        for (int offset = 1; offset <= radius; offset++) {
          coeff *= coeffStep;
          coeffStep *= coeffStep2;
          c16 = (int)(uint16_t)(coeff + 0.5f);
          // lane p's samples were loaded for lane 3 / lane 0 at an earlier offset of the
          // chunk (with THAT offset's clamp), or come from the chunk's first four texels
          const int ro = p + offset, lo_ = offset - p;
          int rpos = ro <= 3 ? ro : (ro - 3 <= validRadius ? ro : min(ro, rightBound));
          int lpos = lo_ <= 0 ? p - offset : (lo_ <= validRadius ? -lo_ : -min(lo_, leftBound));
          const uint8_t* pr = base + (ptrdiff_t)rpos * bpp;
          const uint8_t* pl = base + (ptrdiff_t)lpos * bpp;
          for (int ch = 0; ch < bpp; ch++)
            sum[ch] = wr_addsat16(sum[ch], (((int)__ldg(pr + ch) + (int)__ldg(pl + ch)) * c16) & 0xFFFF);
        }
      } else {
        const int belowBound = r.cy - max(r.lo, 0);
        const int aboveBound = min(r.hi, t.h - 1) - r.cy;
        const int validRadius = min(radius, min(belowBound, aboveBound));
        ptrdiff_t above = 0, below = 0;
        for (int offset = 1; offset <= radius; offset++) {
          if (offset <= validRadius) { above += t.pitch; below -= t.pitch; }
          else {
            if (offset <= aboveBound) above += t.pitch;
            if (offset <= belowBound) below -= t.pitch;
          }
          coeff *= coeffStep;
          coeffStep *= coeffStep2;
          c16 = (int)(uint16_t)(coeff + 0.5f);
          for (int ch = 0; ch < bpp; ch++)
            sum[ch] = wr_addsat16(sum[ch], (((int)__ldg(base + above + p * bpp + ch) + (int)__ldg(base + below + p * bpp + ch)) * c16) & 0xFFFF);
        }
      }
      for (int ch = 0; ch < bpp; ch++) v[ch] = sum[ch] >> 8;
      if (!rgba) return Px{0, 0, v[0], 0};
      return Px{v[0], v[1], v[2], v[3]};
    }
    // fragment main (cs_blur.glsl:132-182): interpolants advance by interp_step * (drawn/4),
    // then one interp_step per tail chunk
    const int trel = rel - r.drawn, j = trel & 3, kc = trel >> 2;
    float uv[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      float st = r.step[i], istep = st * 4.0f;
      float lj = r.o[i];
      for (int s = 0; s < j; s++) lj += st;
      if (r.drawn > 0) lj += istep * ((float)r.drawn * 0.25f);
      lj = wr_repeat_add(lj, istep, kc);
      uv[i] = lj;
    }
    const bool color = k.i[1] != 0;
    float orig[4];
    wr_tex_fragment(t, uv[0], uv[1], orig);
    if (!color) orig[1] = orig[2] = orig[3] = orig[0];
    float gx = k.f[6], gy = k.f[7], gz = k.f[7] * k.f[7];
    float avg[4];
    for (int i = 0; i < 4; i++) avg[i] = orig[i] * gx;
    const int support = min(k.i[0], 300);
    for (int i = 1; i <= support; i += 2) {
      gx *= gy; gy *= gz;
      float sub = gx;
      gx *= gy; gy *= gz;
      sub += gx;
      float ratio = gx / sub;
      float ox = k.f[4] * ((float)i + ratio), oy = k.f[5] * ((float)i + ratio);
      float s0[4], s1[4];
      wr_tex_fragment(t, wr_max(uv[0] - ox, k.f[0]), wr_max(uv[1] - oy, k.f[1]), s0);
      wr_tex_fragment(t, wr_min(uv[0] + ox, k.f[2]), wr_min(uv[1] + oy, k.f[3]), s1);
      if (!color) { s0[1] = s0[2] = s0[3] = s0[0]; s1[1] = s1[2] = s1[3] = s1[0]; }
      for (int q = 0; q < 4; q++) avg[q] += (s0[q] + s1[q]) * sub;
    }
    Px o;
    o.r = wr_round_pixel(avg[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(avg[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(avg[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(avg[3], 255.0f) & 0xFFFF;
    return o;
  }
};

// cs_blur vertex stage (cs_blur.glsl:47-116)
WRD void wr_setup_blur_one(const SetupArgs& a, int idx) {
  const int* iv = (const int*)(a.instances + (size_t)idx * a.stride);
  const float* fv = (const float*)iv;
  const FrameTablesDev& T = a.tabs;
  float4 trg = wr_fetch(T.render_tasks, T.n_render_tasks, iv[0] * 2);
  float4 src = wr_fetch(T.render_tasks, T.n_render_tasks, iv[1] * 2);
  int dir = iv[2];
  float radius = fv[3], rgx = fv[4], rgy = fv[5];
  float tw = (float)a.color0.w, th = (float)a.color0.h;
  int support = (int)ceilf(1.5f * radius) * 2;
  float c0 = 1.0f, c1 = 1.0f;
  if (support > 0) {
    // the reference evaluates exp() with the host libm; the device takes the
    // correctly rounded value through double precision
#ifdef WRCU_HOSTEMU
    float gy = expf(-0.5f / (radius * radius));
#else
    float gy = (float)exp((double)(-0.5f / (radius * radius)));
#endif
    float gx = 1.0f / (sqrtf(2.0f * 3.14159265f) * radius);
    c0 = gx; c1 = gy;
    float cx = gx, cy = gy, cz = gy * gy, total = cx;
    for (int k = 1; k <= support; k += 2) {
      cx *= cy; cy *= cz;
      float sub = cx;
      cx *= cy; cy *= cz;
      sub += cx;
      total += 2.0f * sub;
    }
    c0 /= total;
  }
  QuadOut q;
  memset(&q, 0, sizeof q);
  float uv0x = src.x / tw, uv0y = src.y / th, uv1x = src.z / tw, uv1y = src.w / th;
  const float ax[4] = {0.0f, 1.0f, 1.0f, 0.0f}, ay[4] = {0.0f, 0.0f, 1.0f, 1.0f};
  for (int k = 0; k < 4; k++) {
    float px = (trg.z - trg.x) * ax[k] + trg.x, py = (trg.w - trg.y) * ay[k] + trg.y;
    q.pos[k] = wr_mat_mul(a.tgt.proj, make_float4(px, py, 0.0f, 1.0f));
    q.interp[k][0] = (uv1x - uv0x) * ax[k] + uv0x;
    q.interp[k][1] = (uv1y - uv0y) * ay[k] + uv0y;
  }
  q.n_interp = 2;
  q.flags = CMD_TEXTURED;
  q.col[0] = q.col[1] = q.col[2] = q.col[3] = 255;
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    k->f[0] = (src.x + 0.5f) / tw; k->f[1] = (src.y + 0.5f) / th;
    k->f[2] = (src.x + rgx - 0.5f) / tw; k->f[3] = (src.y + rgy - 0.5f) / th;
    k->f[4] = dir == 0 ? 1.0f / tw : 0.0f;
    k->f[5] = dir == 1 ? 1.0f / th : 0.0f;
    k->f[6] = c0; k->f[7] = c1;
    k->i[0] = support;
    k->i[1] = (a.features & WRCU_FEAT_COLOR_TARGET) ? 1 : 0;
  }
  if (unsupported) {
    atomicAdd(&a.info->unsupported, 1);
    atomicAdd(a.err_counter, 1);
  }
}
WR_SETUP_KERNEL(wr_setup_blur)
