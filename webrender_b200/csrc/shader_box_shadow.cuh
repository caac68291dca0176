// shader_box_shadow.cuh — cs_clip_box_shadow
// (webrender/res/cs_clip_box_shadow.glsl:59-137 + span shader 150-323).
//
// The reference's R8 span shader walks a row as: solid (clip mode) before the
// shadow rect, then repeatedly [one per-fragment chunk | one texture or solid
// run inside the current nine-patch sector], then solid after the rect.  Lane
// positions accumulate in float along that walk, so row_setup replays the walk
// (a handful of iterations: one per nine-patch column) up to the iteration
// that reaches this tile and caches that iteration's run; source() finishes
// the walk for pixels further right.
#pragma once
#include "raster.cuh"
#include "shader_clip_rect.cuh"  // cr_mix

// CmdCold: f[0..3] vEdge, f[4..7] vUvBounds, g[0..3] vUvBounds_NoClamp,
//          g[4..7] vTransformBounds, g[8] vClipMode.x
struct BsLanes {
  float ulx[4], uly[4], lx[4], ly[4];  // uv_linear and local_pos lanes
};
struct BsConst {
  float usx, usy, sx, sy;  // chunk steps
  int shadow_start_len, shadow_end_len, os[4];
};
struct BsRun {
  int n;        // pixels the run draws (0 = none)
  int adv;      // chunks the lanes advance afterwards (num_inside / 4)
  int center;   // 1: solid value, 0: texture span
  float val;
  float ub[4];  // uv_bounds of the sector
  float u[4], v[4];
};

WRD void bs_map_uv(const CmdCold& k, float ulx, float uly, float* u, float* v) {
  float ux = wr_clamp(ulx, 0.0f, k.f[0]), uy = wr_clamp(uly, 0.0f, k.f[1]);
  ux += wr_max(0.0f, ulx - k.f[2]);
  uy += wr_max(0.0f, uly - k.f[3]);
  *u = cr_mix(k.g[0], k.g[2], ux);
  *v = cr_mix(k.g[1], k.g[3], uy);
}
WRD float bs_in_rect(const CmdCold& k, float px, float py) {
  float sx = (px >= k.g[4] ? 1.0f : 0.0f) - (px >= k.g[6] ? 1.0f : 0.0f);
  float sy = (py >= k.g[5] ? 1.0f : 0.0f) - (py >= k.g[7] ? 1.0f : 0.0f);
  return sx * sy;
}
WRD float bs_texel(const TexView& t, float u, float v) {
  float t4[4];
  wr_tex_fragment(t, u, v, t4);
  return t4[0];
}
// the fragment shader body for one lane (cs_clip_box_shadow.glsl:122-137 / 242-253)
WRD float bs_fragment(const RasterArgs& a, const CmdCold& k, float ulx, float uly, float lx, float ly) {
  float u, v;
  bs_map_uv(k, ulx, uly, &u, &v);
  u = wr_clamp(u, k.f[4], k.f[6]);
  v = wr_clamp(v, k.f[5], k.f[7]);
  float in_rect = bs_in_rect(k, lx, ly);
  float texel = bs_texel(a.color0, u, v);
  float mode = k.g[8];
  float alpha = cr_mix(texel, 1.0f - texel, mode);
  return cr_mix(mode, alpha, in_rect);
}
WRD void bs_advance(BsLanes& l, const BsConst& c, float kf) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    l.ulx[j] += kf * c.usx;
    l.uly[j] += kf * c.usy;
    l.lx[j] += kf * c.sx;
    l.ly[j] += kf * c.sy;
  }
}
WRD void bs_step(BsLanes& l, const BsConst& c) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    l.ulx[j] += c.usx;
    l.uly[j] += c.usy;
    l.lx[j] += c.sx;
    l.ly[j] += c.sy;
  }
}
// The sector run that follows a per-fragment chunk; R1 = remaining pixels after
// that chunk, l = lanes after it (cs_clip_box_shadow.glsl:268-316).
WRD void bs_run(const RasterArgs& a, const CmdCold& k, const BsConst& c, const BsLanes& l, int R1, BsRun& run) {
  int num_inside = R1 - 4 - c.shadow_end_len;
  run.ub[0] = k.f[4]; run.ub[1] = k.f[5]; run.ub[2] = k.f[6]; run.ub[3] = k.f[7];
  if (R1 >= c.os[1]) {
    num_inside = min(num_inside, R1 - c.os[1]);
  } else if (R1 >= c.os[3]) {
    num_inside = min(num_inside, R1 - c.os[3]);
    run.ub[1] = run.ub[3] = wr_clamp(cr_mix(k.g[1], k.g[3], k.f[1]), k.f[5], k.f[7]);
  }
  if (R1 >= c.os[0]) {
    num_inside = min(num_inside, R1 - c.os[0]);
  } else if (R1 >= c.os[2]) {
    num_inside = min(num_inside, R1 - c.os[2]);
    run.ub[0] = run.ub[2] = wr_clamp(cr_mix(k.g[0], k.g[2], k.f[0]), k.f[4], k.f[6]);
  }
  run.n = 0;
  run.adv = 0;
  run.center = 0;
  if (num_inside <= 0) return;
  run.n = min(num_inside, R1);
  run.adv = num_inside / 4;
#pragma unroll
  for (int j = 0; j < 4; j++) bs_map_uv(k, l.ulx[j], l.uly[j], &run.u[j], &run.v[j]);
  if (run.ub[0] == run.ub[2] && run.ub[1] == run.ub[3]) {
    run.center = 1;
    float texel = bs_texel(a.color0, wr_clamp(run.u[0], run.ub[0], run.ub[2]), wr_clamp(run.v[0], run.ub[1], run.ub[3]));
    run.val = cr_mix(texel, 1.0f - texel, k.g[8]);
  }
}

struct BoxShadowShader {
  struct Row {
    float L0[6], step[6];
    int body_len;
    bool wneg;
    float w;
    BsConst c;
    BsLanes ln;  // lanes at the start of the cached iteration
    int R;       // remaining pixels at the start of the cached iteration
    BsRun run;   // its sector run
    TexRow tr;   // texture partition of that run, based at the tile start
  };

  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& cm, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[cm.cold];
    wr_row_interp<6>(a, k, cm, y, r.L0, r.step);
    int len = cm.x1 - cm.x0;
    r.body_len = (!rgba && len >= 4) ? (len & ~3) : 0;
    r.wneg = false;
    if (r.body_len == 0) return;
    if (r.step[3] * 4.0f != 0.0f) {  // perspective: the span shader bails
      r.body_len = 0;
      return;
    }
    float w = r.L0[3];
    if (w <= 0.0f) {
      r.wneg = true;
      return;
    }
    w = 1.0f / w;
    r.w = w;
    BsLanes& l = r.ln;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float p[6];
      wr_interp_at<6>(a, r.L0, r.step, j, p);
      l.ulx[j] = p[4] * w; l.uly[j] = p[5] * w;
      l.lx[j] = p[0] * w;  l.ly[j] = p[1] * w;
    }
    BsConst& c = r.c;
    c.usx = (r.step[4] * 4.0f) * w; c.usy = (r.step[5] * 4.0f) * w;
    c.sx = (r.step[0] * 4.0f) * w;  c.sy = (r.step[1] * 4.0f) * w;
    float ul0x = l.ulx[0], ul0y = l.uly[0], p0x = l.lx[0], p0y = l.ly[0];
    const float* tb = k.g + 4;
    float cd0 = (c.sx < 0.0f ? tb[2] : tb[0]) - p0x, cd1 = (c.sy < 0.0f ? tb[3] : tb[1]) - p0y;
    float cd2 = (c.sx < 0.0f ? tb[0] : tb[2]) - p0x, cd3 = (c.sy < 0.0f ? tb[1] : tb[3]) - p0y;
    float rsx = 1.0f / c.sx, rsy = 1.0f / c.sy;
    cd0 = (c.sx != 0.0f) ? cd0 * rsx : 1.0e6f * (cd0 >= 0.0f ? 1.0f : 0.0f);
    cd1 = (c.sy != 0.0f) ? cd1 * rsy : 1.0e6f * (cd1 >= 0.0f ? 1.0f : 0.0f);
    cd2 = (c.sx != 0.0f) ? cd2 * rsx : 1.0e6f * (cd2 >= 0.0f ? 1.0f : 0.0f);
    cd3 = (c.sy != 0.0f) ? cd3 * rsy : 1.0e6f * (cd3 >= 0.0f ? 1.0f : 0.0f);
    float shadow_start = wr_max(cd0, cd1), shadow_end = wr_min(cd2, cd3);
    float fl = (float)r.body_len;
    c.shadow_start_len = (int)wr_clamp(fl - 4.0f * floorf(shadow_start), 0.0f, fl);
    c.shadow_end_len = (int)wr_clamp(fl - 4.0f * ceilf(shadow_end), 0.0f, fl);
    const float* e = k.f;
    float od[4];
    od[0] = (c.usx < 0.0f ? e[2] : e[0]) - ul0x;
    od[1] = (c.usy < 0.0f ? e[3] : e[1]) - ul0y;
    od[2] = (c.usx < 0.0f ? e[0] : e[2]) - ul0x;
    od[3] = (c.usy < 0.0f ? e[1] : e[3]) - ul0y;
    float rux = 1.0f / c.usx, ruy = 1.0f / c.usy;
    od[0] = (c.usx != 0.0f) ? od[0] * rux : 1.0e6f * (od[0] >= 0.0f ? 1.0f : 0.0f);
    od[1] = (c.usy != 0.0f) ? od[1] * ruy : 1.0e6f * (od[1] >= 0.0f ? 1.0f : 0.0f);
    od[2] = (c.usx != 0.0f) ? od[2] * rux : 1.0e6f * (od[2] >= 0.0f ? 1.0f : 0.0f);
    od[3] = (c.usy != 0.0f) ? od[3] * ruy : 1.0e6f * (od[3] >= 0.0f ? 1.0f : 0.0f);
#pragma unroll
    for (int i = 0; i < 4; i++) c.os[i] = (int)wr_clamp(fl - 4.0f * floorf(od[i]), (float)c.shadow_end_len, fl);
    // the solid section before the shadow rect
    r.R = r.body_len;
    if (r.R > c.shadow_start_len) {
      int num_before = r.R - c.shadow_start_len;
      bs_advance(l, c, (float)(num_before / 4));
      r.R = c.shadow_start_len;
    }
    // walk to the iteration that covers the first body pixel this tile draws
    int first = max(tx0, (int)cm.x0) - (int)cm.x0;
    int Rp = r.body_len - (first & ~3);
    r.run.n = 0;
    r.tr.mode = TEX_NONE;
    r.tr.body_len = 0;
    for (;;) {
      if (r.R <= 0) return;
      BsLanes nl = l;
      bs_step(nl, c);
      int R1 = r.R - 4;
      if (R1 <= c.shadow_end_len) { r.run.n = 0; r.run.adv = 0; return; }
      bs_run(a, k, c, nl, R1, r.run);
      int R2 = R1 - r.run.n;
      if (Rp > R2) break;  // this iteration reaches the tile
      if (r.run.adv) bs_advance(nl, c, (float)r.run.adv);
      l = nl;
      r.R = R2;
    }
    if (r.run.n && !r.run.center) {
      int run_start = r.body_len - (r.R - 4);  // span-relative pixel the run starts at
      wr_tex_row_setup(a.color0, r.run.ub, false, r.run.n, r.run.u, r.run.v, max(first - run_start, 0), r.tr,
                       WRCU_FMT_R8);
    }
  }

  WRD_MEMBER int run_pixel(const RasterArgs& a, const CmdCold& k, const BsRun& run, const TexRow& tr, int rel_run) {
    if (run.center) return wr_round_pixel(run.val, 255.0f) & 0xFFFF;
    int v = wr_tex_body(a.color0, tr, rel_run).r;
    return k.g[8] != 0.0f ? ((255 - v) & 0xFFFF) : v;
  }

  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& cm, const Row& r, int x, int, bool rgba) {
    const CmdCold& k = a.cold[cm.cold];
    const float mode = k.g[8];
    int rel = x - cm.x0;
    int r8;
    if (rel < r.body_len) {
      if (r.wneg) {
        r8 = wr_round_pixel(0.0f, 255.0f);
      } else {
        const BsConst& c = r.c;
        int j = rel & 3;
        int Rp = r.body_len - (rel & ~3);
        if (Rp > c.shadow_start_len) {
          r8 = wr_round_pixel(mode, 255.0f) & 0xFFFF;
        } else {
          BsLanes l = r.ln;
          int R = r.R;
          bool cached = true;
          for (;;) {
            if (R <= 0 || Rp > R) { r8 = wr_round_pixel(mode, 255.0f) & 0xFFFF; break; }  // not reachable
            if (Rp == R) {
              r8 = wr_round_pixel(bs_fragment(a, k, l.ulx[j], l.uly[j], l.lx[j], l.ly[j]), 255.0f) & 0xFFFF;
              break;
            }
            bs_step(l, c);
            int R1 = R - 4;
            if (R1 <= c.shadow_end_len) { r8 = wr_round_pixel(mode, 255.0f) & 0xFFFF; break; }
            int rel_run = rel - (r.body_len - R1);
            if (cached) {
              if (Rp > R1 - r.run.n) { r8 = run_pixel(a, k, r.run, r.tr, rel_run); break; }
              if (r.run.adv) bs_advance(l, c, (float)r.run.adv);
              R = R1 - r.run.n;
              cached = false;
            } else {
              BsRun run;
              bs_run(a, k, c, l, R1, run);
              if (Rp > R1 - run.n) {
                TexRow tr;
                tr.mode = TEX_NONE;
                if (!run.center)
                  wr_tex_row_setup(a.color0, run.ub, false, run.n, run.u, run.v, rel_run, tr, WRCU_FMT_R8, false);
                r8 = run_pixel(a, k, run, tr, rel_run);
                break;
              }
              if (run.adv) bs_advance(l, c, (float)run.adv);
              R = R1 - run.n;
            }
          }
        }
      }
    } else {
      // fragment main for the span tail / non-R8 targets
      int trel = rel - r.body_len;
      int j = trel & 3, kc = trel >> 2;
      float adv = (float)r.body_len * 0.25f;
      float Lj[6];
#pragma unroll
      for (int i = 0; i < 6; i++) {
        float st = r.step[i], istep = st * 4.0f;
        float lj = r.L0[i];
        for (int s = 0; s < j; s++) lj += st;
        if (r.body_len > 0) lj += istep * adv;
        for (int s = 0; s < kc; s++) lj += istep;
        Lj[i] = lj;
      }
      float wq = Lj[3];
      float v = bs_fragment(a, k, Lj[4] / wq, Lj[5] / wq, Lj[0] / wq, Lj[1] / wq);
      v = wq > 0.0f ? v : 0.0f;
      r8 = wr_round_pixel(v, 255.0f) & 0xFFFF;
    }
    if (!rgba) return Px{0, 0, r8, 0};
    return Px{r8, r8, r8, r8};
  }
};
