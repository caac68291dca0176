// shader_mix_blend.cuh — brush_mix_blend [ALPHA_PASS] (webrender/res/
// brush_mix_blend.glsl): CSS mix-blend-mode of a picture (sColor1) against a
// read-back backdrop (sColor0).  Fragment path only.
#pragma once
#include "raster.cuh"
#include "shader_opacity.cuh"  // wr_image_quad_uv

// CmdCold: f[0..3] v_src_uv_sample_bounds, f[4..7] v_backdrop_uv_sample_bounds,
//          g[0] v_perspective, g[1] gl_FragCoord.w, i[0] v_op
// interpolants: [v_src_uv.xy, v_backdrop_uv.xy]
WRD float mb_lum(const float* c) { return c[0] * 0.3f + c[1] * 0.59f + c[2] * 0.11f; }
WRD float mb_sat(const float* c) {
  return wr_max(c[0], wr_max(c[1], c[2])) - wr_min(c[0], wr_min(c[1], c[2]));
}
WRD void mb_clip_color(float* C) {
  float L = mb_lum(C);
  float n = wr_min(C[0], wr_min(C[1], C[2]));
  float x = wr_max(C[0], wr_max(C[1], C[2]));
  if (n < 0.0f)
    for (int i = 0; i < 3; i++) C[i] = L + (((C[i] - L) * L) / (L - n));
  if (x > 1.0f)
    for (int i = 0; i < 3; i++) C[i] = L + (((C[i] - L) * (1.0f - L)) / (x - L));
}
WRD void mb_set_lum(const float* C, float l, float* out) {
  float dl = l - mb_lum(C);
  for (int i = 0; i < 3; i++) out[i] = C[i] + dl;
  mb_clip_color(out);
}
WRD void mb_set_sat_inner(float& Cmin, float& Cmid, float& Cmax, float s) {
  if (Cmax > Cmin) {
    Cmid = (((Cmid - Cmin) * s) / (Cmax - Cmin));
    Cmax = s;
  } else {
    Cmid = 0.0f;
    Cmax = 0.0f;
  }
  Cmin = 0.0f;
}
WRD void mb_set_sat(float* C, float s) {
  if (C[0] <= C[1]) {
    if (C[1] <= C[2]) mb_set_sat_inner(C[0], C[1], C[2], s);
    else if (C[0] <= C[2]) mb_set_sat_inner(C[0], C[2], C[1], s);
    else mb_set_sat_inner(C[2], C[0], C[1], s);
  } else {
    if (C[0] <= C[2]) mb_set_sat_inner(C[1], C[0], C[2], s);
    else if (C[1] <= C[2]) mb_set_sat_inner(C[1], C[2], C[0], s);
    else mb_set_sat_inner(C[2], C[1], C[0], s);
  }
}
WRD void mb_hard_light(const float* Cb, const float* Cs, float* out) {
  for (int i = 0; i < 3; i++) {
    float m = Cb[i] * (2.0f * Cs[i]);
    float s2 = 2.0f * Cs[i] - 1.0f;
    float sc = Cb[i] + s2 - (Cb[i] * s2);
    out[i] = (sc - m) * (Cs[i] >= 0.5f ? 1.0f : 0.0f) + m;
  }
}
WRD float mb_color_dodge(float Cb, float Cs) {
  if (Cb == 0.0f) return 0.0f;
  if (Cs == 1.0f) return 1.0f;
  return wr_min(1.0f, Cb / (1.0f - Cs));
}
WRD float mb_color_burn(float Cb, float Cs) {
  if (Cb == 1.0f) return 1.0f;
  if (Cs == 0.0f) return 0.0f;
  return 1.0f - wr_min(1.0f, (1.0f - Cb) / Cs);
}
WRD float mb_soft_light(float Cb, float Cs) {
  if (Cs <= 0.5f) return Cb - (1.0f - 2.0f * Cs) * Cb * (1.0f - Cb);
  float D = Cb <= 0.25f ? ((16.0f * Cb - 12.0f) * Cb + 4.0f) * Cb : __fsqrt_rn(Cb);
  return Cb + (2.0f * Cs - 1.0f) * (D - Cb);
}

struct MixBlendShader {
  struct Row {
    float o[4], step[4];
    float pd;
    float base[4][4];
    int kb;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<4>(a, k, c, y, r.o, r.step);
    r.pd = (1.0f - k.g[1]) * k.g[0] + k.g[1];
    r.kb = wr_chunk_base<4>(a, r.o, r.step, c, tx0, r.base);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    int rel = x - c.x0;
    float L[4];
    wr_chunk_lane<4>(a, r.base, r.step, r.kb, rel >> 2, rel & 3, L);
    float Cb[4], Cs[4], t[3], result[4];
    wr_tex_fragment(a.color0, wr_clamp(L[2], k.f[4], k.f[6]), wr_clamp(L[3], k.f[5], k.f[7]), Cb);
    float pd = r.pd;
    if (a.persp) {  // gl_FragCoord.w varies per sample
      const float fw = wr_persp_zw(*a.persp, 1, rel);
      pd = (1.0f - fw) * k.g[0] + fw;
    }
    wr_tex_fragment(a.color1, wr_clamp(L[0] * pd, k.f[0], k.f[2]), wr_clamp(L[1] * pd, k.f[1], k.f[3]), Cs);
    if (Cb[3] != 0.0f) for (int i = 0; i < 3; i++) Cb[i] /= Cb[3];
    if (Cs[3] != 0.0f) for (int i = 0; i < 3; i++) Cs[i] /= Cs[3];
    result[0] = 1.0f; result[1] = 1.0f; result[2] = 0.0f; result[3] = 1.0f;
    switch (k.i[0] & 0xFF) {
      case 1: for (int i = 0; i < 3; i++) result[i] = Cb[i] * Cs[i]; break;
      case 3: mb_hard_light(Cs, Cb, result); break;
      case 4: for (int i = 0; i < 3; i++) result[i] = wr_min(Cs[i], Cb[i]); break;
      case 5: for (int i = 0; i < 3; i++) result[i] = wr_max(Cs[i], Cb[i]); break;
      case 6: for (int i = 0; i < 3; i++) result[i] = mb_color_dodge(Cb[i], Cs[i]); break;
      case 7: for (int i = 0; i < 3; i++) result[i] = mb_color_burn(Cb[i], Cs[i]); break;
      case 8: mb_hard_light(Cb, Cs, result); break;
      case 9: for (int i = 0; i < 3; i++) result[i] = mb_soft_light(Cb[i], Cs[i]); break;
      case 10: for (int i = 0; i < 3; i++) result[i] = fabsf(Cb[i] - Cs[i]); break;
      case 12: t[0] = Cs[0]; t[1] = Cs[1]; t[2] = Cs[2]; mb_set_sat(t, mb_sat(Cb)); mb_set_lum(t, mb_lum(Cb), result); break;
      case 13: t[0] = Cb[0]; t[1] = Cb[1]; t[2] = Cb[2]; mb_set_sat(t, mb_sat(Cs)); mb_set_lum(t, mb_lum(Cb), result); break;
      case 14: mb_set_lum(Cs, mb_lum(Cb), result); break;
      case 15: mb_set_lum(Cb, mb_lum(Cs), result); break;
      default: break;
    }
    for (int i = 0; i < 3; i++) result[i] = (1.0f - Cb[3]) * Cs[i] + Cb[3] * result[i];
    result[3] = Cs[3];
    for (int i = 0; i < 3; i++) result[i] *= result[3];
    Px o;
    o.r = wr_round_pixel(result[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(result[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(result[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(result[3], 255.0f) & 0xFFFF;
    return o;
  }
};

// brush_mix_blend vertex stage (brush_mix_blend.glsl:25-83)
WRD void wr_setup_brush_mix_blend_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  QuadOut q;
  BrushVS vs;
  memset(&q, 0, sizeof q);
  wr_brush_vertex(a, aData, 3, q, vs);
  const FrameTablesDev& T = a.tabs;
  const float* lr = vs.ph.lr;
  float persp = (vs.brush_flags & 1) ? 1.0f : 0.0f;
  float bounds[2][4];
  for (int which = 0; which < 2; which++) {  // 0: source (sColor1), 1: backdrop (sColor0)
    int addr = which == 0 ? vs.ph.user_data[2] : vs.ph.user_data[1];
    const TexView& t = which == 0 ? a.color1 : a.color0;
    float4 r0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, addr);
    float itw = 1.0f / (float)t.w, ith = 1.0f / (float)t.h;
    for (int k = 0; k < 4; k++) {
      float fx = (vs.local_pos[k].x - lr[0]) / (lr[2] - lr[0]);
      float fy = (vs.local_pos[k].y - lr[1]) / (lr[3] - lr[1]);
      wr_image_quad_uv(T, addr, fx, fy);
      float ux = (r0.z - r0.x) * fx + r0.x, uy = (r0.w - r0.y) * fy + r0.y;
      float pf = which == 0 ? (1.0f - vs.world_pos[k].w) * persp + vs.world_pos[k].w : 1.0f;
      q.interp[k][2 * which] = ux * itw * pf;
      q.interp[k][2 * which + 1] = uy * ith * pf;
    }
    bounds[which][0] = (r0.x + 0.5f) * itw; bounds[which][1] = (r0.y + 0.5f) * ith;
    bounds[which][2] = (r0.z - 0.5f) * itw; bounds[which][3] = (r0.w - 0.5f) * ith;
  }
  q.n_interp = 4;
  q.flags |= CMD_TEXTURED;
  float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  wr_pack_color(q, white);
  float fw = 1.0f / q.pos[0].w;
  if (!isfinite(fw)) fw = 0.0f;
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    for (int i = 0; i < 4; i++) { k->f[i] = bounds[0][i]; k->f[4 + i] = bounds[1][i]; }
    k->g[0] = persp;
    k->g[1] = fw;
    k->i[0] = vs.ph.user_data[0];
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_brush_mix_blend)

template <> struct WrRun<MixBlendShader> {
  enum { n = 4 };
  WRD_MEMBER int drawn(const MixBlendShader::Row&) { return 0; }
};
