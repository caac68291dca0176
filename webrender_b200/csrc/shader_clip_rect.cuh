// shader_clip_rect.cuh — cs_clip_rectangle [FAST_PATH]
// (webrender/res/cs_clip_rectangle.glsl, clip_shared.glsl, ellipse.glsl,
// transform.glsl get_node_pos): rounded-rect coverage into an alpha target.
//
// Parity target is SWGL's output, which for R8 targets comes from the
// program's span shader swgl_drawSpanR8 (cs_clip_rectangle.glsl:223-495) for
// whole 4-pixel chunks and from the fragment main for the span tail.  The span
// shader partitions each row into clear / start-AA / opaque / end-AA / clear
// runs from per-row scalars; ClipRectShader::row_setup computes those scalars
// once per (command,row) per warp and ::source classifies the pixel, so every
// byte equals the reference's.
#pragma once
#include "raster.cuh"
#include "setup_common.cuh"

// CmdCold.g layout for clip rectangles
//  g[0] mode, g[1] fast, g[2..4] vClipParams, g[5..8] vTransformBounds,
//  g[9..24] vClipCenter_Radius TL,TR,BR,BL, g[25..36) vClipPlane TL,TR,BR,BL (3 each = 12 → g[25..36])
#define CR_MODE 0
#define CR_FAST 1
#define CR_PARAMS 2
#define CR_BOUNDS 5
#define CR_CORNER 9
#define CR_PLANE 25

WRD float cr_distance_aa(float aa_range, float sd) {
  return wr_clamp(0.5f - sd * aa_range, 0.0f, 1.0f);
}
WRD float cr_mix(float x, float y, float a) { return (y - x) * a + x; }
WRD float cr_sd_rounded_box(float px, float py, const float* p) {
  float dx = fabsf(px) - p[0], dy = fabsf(py) - p[1];
  float mx = wr_max(dx, 0.0f), my = wr_max(dy, 0.0f);
  return (sqrtf(mx * mx + my * my) + wr_min(wr_max(dx, dy), 0.0f)) - p[2];
}
WRD float cr_ellipse_approx(float px, float py, float irx, float iry, float scale) {
  float prx = px * irx, pry = py * iry;
  float g = (px * prx + py * pry) - scale;
  float gx = (1.0f + scale) * prx, gy = (1.0f + scale) * pry;
  return g * (1.0f / sqrtf(gx * gx + gy * gy));
}
WRD float cr_sd_rect(float px, float py, const float* b) {
  return wr_max(wr_max(b[0] - px, px - b[2]), wr_max(b[1] - py, py - b[3]));
}
WRD float cr_distance_to_rounded_rect(const float* g, float px, float py) {
  const float* cr = g + CR_CORNER;
  const float* pl = g + CR_PLANE;
  float c0 = 1.0e-6f, c1 = 1.0e-6f, c2 = 1.0f, c3 = 1.0f;
  if (px * pl[0] + py * pl[1] > pl[2]) { c0 = cr[0] - px; c1 = cr[1] - py; c2 = cr[2]; c3 = cr[3]; }
  if (px * pl[3] + py * pl[4] > pl[5]) { c0 = (cr[4] - px) * -1.0f; c1 = (cr[5] - py) * 1.0f; c2 = cr[6]; c3 = cr[7]; }
  if (px * pl[6] + py * pl[7] > pl[8]) { c0 = px - cr[8]; c1 = py - cr[9]; c2 = cr[10]; c3 = cr[11]; }
  if (px * pl[9] + py * pl[10] > pl[11]) { c0 = (cr[12] - px) * 1.0f; c1 = (cr[13] - py) * -1.0f; c2 = cr[14]; c3 = cr[15]; }
  return wr_max(cr_ellipse_approx(c0, c1, c2, c3, 1.0f), cr_sd_rect(px, py, g + CR_BOUNDS));
}

struct ClipRectShader {
  struct Row {
    const float* g;
    float L0[4], step[4];      // vLocalPos at the span start (lane 0) and per-pixel step
    int body_len;              // pixels drawn by the span shader
    bool span_ok, wneg;        // span shader ran / w <= 0 (all clear)
    float w, sx, sy, aa_range; // 1/w, chunk step of local_pos, AA scale
    int r1, r2, r3, r4;        // thresholds in "remaining pixels" space
    bool start_corner_on, end_corner_on;
    float sp[3], ep[3], sc[4], ec[4];  // start/end plane and corner
  };

  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    r.g = k.g;
    // interpolants at the span start (exact running sums of the edge walk)
    wr_row_interp<4>(a, k, c, y, r.L0, r.step);
    int len = c.x1 - c.x0;
    r.body_len = (!rgba && len >= 4) ? (len & ~3) : 0;
    r.span_ok = false;
    r.wneg = false;
    if (r.body_len == 0) return;
    float istep_w = r.step[3] * 4.0f;
    if (istep_w != 0.0f) {  // perspective: span shader bails, fragment path draws everything
      r.body_len = 0;
      return;
    }
    r.span_ok = true;
    float w = r.L0[3];
    if (w <= 0.0f) {
      r.wneg = true;
      return;
    }
    w = 1.0f / w;
    r.w = w;
    const float* g = r.g;
    bool fast = g[CR_FAST] != 0.0f;
    // lanes 0,1 of local_pos at the span start
    float l1x = (r.L0[0] + r.step[0]) * w, l1y = (r.L0[1] + r.step[1]) * w;
    float p0x = r.L0[0] * w, p0y = r.L0[1] * w;
    float sx = (r.step[0] * 4.0f) * w, sy = (r.step[1] * 4.0f) * w;
    r.sx = sx;
    r.sy = sy;
    float step_scale = wr_max(sx * sx + sy * sy, 1.0e-6f);
    float aa_range = 1.0f / (fabsf(l1x - p0x) + fabsf(l1y - p0y));
    r.aa_range = aa_range;
    float aa_margin = 1.0f / sqrtf(aa_range * aa_range * step_scale);
    float rect[4];
    if (fast) {
      rect[0] = -g[CR_PARAMS] - g[CR_PARAMS + 2]; rect[1] = -g[CR_PARAMS + 1] - g[CR_PARAMS + 2];
      rect[2] = g[CR_PARAMS] + g[CR_PARAMS + 2];  rect[3] = g[CR_PARAMS + 1] + g[CR_PARAMS + 2];
    } else {
      rect[0] = g[CR_BOUNDS]; rect[1] = g[CR_BOUNDS + 1]; rect[2] = g[CR_BOUNDS + 2]; rect[3] = g[CR_BOUNDS + 3];
    }
    bool negx = sx < 0.0f, negy = sy < 0.0f;
    float cd0 = (negx ? rect[2] : rect[0]) - p0x, cd1 = (negy ? rect[3] : rect[1]) - p0y;
    float cd2 = (negx ? rect[0] : rect[2]) - p0x, cd3 = (negy ? rect[1] : rect[3]) - p0y;
    float rsx = 1.0f / sx, rsy = 1.0f / sy;
    cd0 = (sx != 0.0f) ? cd0 * rsx : 1.0e6f * (cd0 >= 0.0f ? 1.0f : 0.0f);
    cd1 = (sy != 0.0f) ? cd1 * rsy : 1.0e6f * (cd1 >= 0.0f ? 1.0f : 0.0f);
    cd2 = (sx != 0.0f) ? cd2 * rsx : 1.0e6f * (cd2 >= 0.0f ? 1.0f : 0.0f);
    cd3 = (sy != 0.0f) ? cd3 * rsy : 1.0e6f * (cd3 >= 0.0f ? 1.0f : 0.0f);
    float opaque_start = wr_max(cd0, cd1), opaque_end = wr_min(cd2, cd3);
    float aa_start = opaque_start, aa_end = opaque_end;
    r.sp[0] = r.sp[1] = r.sp[2] = 1.0e6f;
    r.ep[0] = r.ep[1] = r.ep[2] = 1.0e6f;
    r.sc[0] = r.sc[1] = 1.0e6f; r.sc[2] = r.sc[3] = 1.0f;
    r.ec[0] = r.ec[1] = 1.0e6f; r.ec[2] = r.ec[3] = 1.0f;
    float z = g[CR_PARAMS + 2];
    float offset = (g[CR_PARAMS] + g[CR_PARAMS + 1] + z) * z;
#pragma unroll
    for (int i = 0; i < 4; i++) {  // CLIP_CORNER in order TL, TR, BR, BL
      float pl0, pl1, pl2;
      if (fast) {
        pl0 = (i == 0 || i == 3) ? -z : z;
        pl1 = (i < 2) ? -z : z;
        pl2 = offset;
      } else {
        pl0 = g[CR_PLANE + 3 * i]; pl1 = g[CR_PLANE + 3 * i + 1]; pl2 = g[CR_PLANE + 3 * i + 2];
      }
      float dist = (p0x * pl0 + p0y * pl1) - pl2;
      float scale = -(sx * pl0 + sy * pl1);
      if (scale >= 0.0f) {
        if (dist > opaque_start * scale) {
          if (!fast) { r.sc[0] = g[CR_CORNER + 4 * i]; r.sc[1] = g[CR_CORNER + 4 * i + 1]; r.sc[2] = g[CR_CORNER + 4 * i + 2]; r.sc[3] = g[CR_CORNER + 4 * i + 3]; }
          r.sp[0] = pl0; r.sp[1] = pl1; r.sp[2] = pl2;
          float inv_scale = 1.0f / wr_max(scale, 1.0e-6f);
          opaque_start = dist * inv_scale;
          float apex = (0.7071f - 0.5f) * 2.0f * fabsf(pl0 * pl1);
          aa_start = opaque_start - apex * inv_scale;
        }
      } else if (dist > opaque_end * scale) {
        if (!fast) { r.ec[0] = g[CR_CORNER + 4 * i]; r.ec[1] = g[CR_CORNER + 4 * i + 1]; r.ec[2] = g[CR_CORNER + 4 * i + 2]; r.ec[3] = g[CR_CORNER + 4 * i + 3]; }
        r.ep[0] = pl0; r.ep[1] = pl1; r.ep[2] = pl2;
        float inv_scale = 1.0f / wr_min(scale, -1.0e-6f);
        opaque_end = dist * inv_scale;
        float apex = (0.7071f - 0.5f) * 2.0f * fabsf(pl0 * pl1);
        aa_end = opaque_end - apex * inv_scale;
      }
    }
    aa_margin = wr_max(aa_margin - wr_max(aa_start - aa_end, 0.0f), 0.0f);
    aa_start -= aa_margin;
    aa_end += aa_margin;
    float fl = (float)r.body_len;
    int aa_start_len = (int)wr_clamp(fl - 4.0f * floorf(aa_start), 0.0f, fl);
    int opaque_start_len = (int)wr_clamp(fl - 4.0f * ceilf(opaque_start), 0.0f, fl);
    int opaque_end_len = (int)wr_clamp(fl - 4.0f * floorf(opaque_end), 0.0f, fl);
    int aa_end_len = (int)wr_clamp(fl - 4.0f * ceilf(aa_end), 0.0f, fl);
    // the span shader's run sequence on the "remaining" counter
    r.r1 = min(r.body_len, aa_start_len);
    r.r2 = min(r.r1, opaque_start_len);
    r.r3 = min(r.r2, opaque_end_len);
    r.r4 = min(r.r3, aa_end_len);
    r.start_corner_on = !fast && r.sp[0] < 1.0e5f;
    r.end_corner_on = !fast && r.ep[0] < 1.0e5f;
  }

  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int y, bool rgba) {
    (void)a; (void)y;
    const float* g = r.g;
    float mode = g[CR_MODE];
    bool fast = g[CR_FAST] != 0.0f;
    int rel = x - c.x0;
    float v;
    if (rel < r.body_len) {
      if (r.wneg) {
        v = 0.0f;
      } else {
        int j = rel & 3;
        int R = r.body_len - (rel & ~3);  // `remaining` when this pixel's chunk is committed
        if (R > r.r1 || R <= r.r4) {
          v = mode;
        } else if (R <= r.r2 && R > r.r3) {
          v = 1.0f - mode;
        } else {
          // AA chunk: rebuild the lane's local_pos with the span shader's additions
          float lx = r.L0[0], ly = r.L0[1];
          for (int s = 0; s < j; s++) { lx += r.step[0]; ly += r.step[1]; }  // init_interp lanes
          lx *= r.w;
          ly *= r.w;
          int num_aa = r.body_len - r.r1;
          if (num_aa > 0) { float kf = (float)(num_aa / 4); lx += kf * r.sx; ly += kf * r.sy; }
          bool in_start = R > r.r2;
          int chunks_start = in_start ? (r.r1 - R) / 4 : (r.r1 - r.r2) / 4;
          for (int s = 0; s < chunks_start; s++) { lx += r.sx; ly += r.sy; }
          bool use_corner;
          const float *pl, *cn;
          if (in_start) {
            use_corner = r.start_corner_on; pl = r.sp; cn = r.sc;
          } else {
            int num_opaque = r.r2 - r.r3;
            if (num_opaque > 0) { float kf = (float)(num_opaque / 4); lx += kf * r.sx; ly += kf * r.sy; }
            int chunks_end = (r.r3 - R) / 4;
            for (int s = 0; s < chunks_end; s++) { lx += r.sx; ly += r.sy; }
            use_corner = r.end_corner_on; pl = r.ep; cn = r.ec;
          }
          float dd;
          if (use_corner && (lx * pl[0] + ly * pl[1] > pl[2]))
            dd = cr_ellipse_approx(lx - cn[0], ly - cn[1], cn[2], cn[3], 1.0f);
          else
            dd = fast ? cr_sd_rounded_box(lx, ly, g + CR_PARAMS) : cr_sd_rect(lx, ly, g + CR_BOUNDS);
          float alpha = cr_distance_aa(r.aa_range, dd);
          v = cr_mix(alpha, 1.0f - alpha, mode);
        }
      }
    } else {
      // fragment main (cs_clip_rectangle.glsl:170-199) for the span tail / RGBA8 targets
      int trel = rel - r.body_len;
      int j = trel & 3, kc = trel >> 2;
      float adv = (float)r.body_len * 0.25f;  // step_interp_inputs(drawn)
      float L[2][4], Lj[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float st = r.step[i], istep = st * 4.0f;
        float l0 = r.L0[i], l1 = l0 + st, lj = l0;
        for (int s = 0; s < j; s++) lj += st;  // init_interp: lanes accumulate sequentially
        if (r.body_len > 0) { float d = istep * adv; l0 += d; l1 += d; lj += d; }
        for (int s = 0; s < kc; s++) { l0 += istep; l1 += istep; lj += istep; }  // one run() per tail chunk
        L[0][i] = l0; L[1][i] = l1; Lj[i] = lj;
      }
      float ljx = Lj[0], ljy = Lj[1], ljw = Lj[3];
      float p0x = L[0][0] / L[0][3], p0y = L[0][1] / L[0][3];
      float p1x = L[1][0] / L[1][3], p1y = L[1][1] / L[1][3];
      float aa_range = 1.0f / (fabsf(p1x - p0x) + fabsf(p1y - p0y));
      float px = ljx / ljw, py = ljy / ljw;
      float dist = fast ? cr_sd_rounded_box(px, py, g + CR_PARAMS) : cr_distance_to_rounded_rect(g, px, py);
      float alpha = cr_distance_aa(aa_range, dist);
      float fa = cr_mix(alpha, 1.0f - alpha, mode);
      v = ljw > 0.0f ? fa : 0.0f;
    }
    int r8 = wr_round_pixel(v, 255.0f) & 0xFFFF;
    if (!rgba) return Px{0, 0, r8, 0};
    return Px{wr_round_pixel(0.0f, 255.0f), wr_round_pixel(0.0f, 255.0f), r8, wr_round_pixel(1.0f, 255.0f)};
  }
};
