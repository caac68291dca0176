// shader_blend.cuh — brush_blend [ALPHA_PASS] (webrender/res/brush_blend.glsl
// + blend.glsl): CSS filter ops on an off-screen picture.  Fragment path only
// (the program has no span shader).  pow() is the reference's vector
// approximation approx_pow2(approx_log2(x) * y) (swgl/src/glsl.h:776-799),
// restated in integer/float ops so results match bit for bit.
#pragma once
#include "raster.cuh"
#include "shader_opacity.cuh"  // wr_image_quad_uv

// CmdCold: f[0..3] v_uv_sample_bounds, f[4] v_perspective, f[5] gl_FragCoord.w, f[6] v_amount,
//          i[0] v_op, i[1] v_table_address, g[0..15] v_color_mat (column major),
//          g[16..19] v_funcs, g[20..23] v_color_offset
WRD float wr_glsl_floor(float v) {  // Float floor(Float), glsl.h:693-696
  float roundtrip = (float)(int)v;
  return roundtrip - (roundtrip > v ? 1.0f : 0.0f);
}
WRD float wr_approx_log2(float x) {
  uint32_t bits = __float_as_uint(x);
  float e = (float)bits * (1.0f / (1 << 23));
  float m = __uint_as_float((bits & 0x007fffffu) | 0x3f000000u);
  return e - 124.225514990f - 1.498030302f * m - 1.725879990f / (0.3520887068f + m);
}
WRD float wr_approx_pow2(float x) {
  float f = x - wr_glsl_floor(x);
  float sc = x + 121.274057500f - 1.490129070f * f + 27.728023300f / (4.84252568f - f);
  int r = (int)(1.0f * (1 << 23) * sc + 0.5f);
  return __uint_as_float((uint32_t)r);
}
WRD float wr_glsl_pow(float x, float y) {
  return (x == 0.0f || x == 1.0f) ? x : wr_approx_pow2(wr_approx_log2(x) * y);
}

// CalculateFilter (blend.glsl:196-237) on an un-premultiplied colour
WRD void wr_calculate_filter(const RasterArgs& a, const CmdCold& k, float* color, float& alpha) {
  const float amount = k.f[6];
  switch (k.i[0]) {
    case 0: for (int c = 0; c < 3; c++) color[c] = wr_clamp(color[c] * amount - 0.5f * amount + 0.5f, 0.0f, 1.0f); break;
    case 3: for (int c = 0; c < 3; c++) color[c] = ((1.0f - color[c]) - color[c]) * amount + color[c]; break;
    case 6: for (int c = 0; c < 3; c++) color[c] = wr_clamp(color[c] * amount, 0.0f, 1.0f); break;
    case 8:
      for (int c = 0; c < 3; c++) {
        float c1 = color[c] / 12.92f;
        float c2 = wr_glsl_pow(color[c] / 1.055f + 0.055f / 1.055f, 2.4f);
        color[c] = color[c] <= 0.04045f ? c1 : c2;
      }
      break;
    case 9:
      for (int c = 0; c < 3; c++) {
        float c1 = color[c] * 12.92f;
        float c2 = 1.055f * wr_glsl_pow(color[c], 1.0f / 2.4f) - 0.055f;
        color[c] = color[c] <= 0.0031308f ? c1 : c2;
      }
      break;
    case 11: {
      float ca[4] = {color[0], color[1], color[2], alpha};
      int offset = 0;
      for (int i = 0; i < 4; i++) {
        switch ((int)k.g[16 + i]) {
          case 1:
          case 2: {
            int kk = (int)wr_glsl_floor(ca[i] * 255.0f + 0.5f);
            float4 t = wr_fetch(a.gpu_cache, a.n_gpu_cache, k.i[1] + offset + kk / 4);
            int sel = kk % 4;
            float v = sel == 0 ? t.x : sel == 1 ? t.y : sel == 2 ? t.z : t.w;
            ca[i] = wr_clamp(v, 0.0f, 1.0f);
            offset += 64;
            break;
          }
          case 3: {
            float4 t = wr_fetch(a.gpu_cache, a.n_gpu_cache, k.i[1] + offset);
            ca[i] = wr_clamp(t.x * ca[i] + t.y, 0.0f, 1.0f);
            offset += 1;
            break;
          }
          case 4: {
            float4 t = wr_fetch(a.gpu_cache, a.n_gpu_cache, k.i[1] + offset);
            ca[i] = wr_clamp(t.x * wr_glsl_pow(ca[i], t.y) + t.z, 0.0f, 1.0f);
            offset += 1;
            break;
          }
          default: break;
        }
      }
      color[0] = ca[0]; color[1] = ca[1]; color[2] = ca[2]; alpha = ca[3];
      break;
    }
    case 10:
      color[0] = k.g[20]; color[1] = k.g[21]; color[2] = k.g[22];
      alpha = k.g[23];
      break;
    default: {
      const float* m = k.g;
      float vin[4] = {color[0], color[1], color[2], alpha}, r[4];
      for (int c = 0; c < 4; c++)
        r[c] = wr_clamp((m[c] * vin[0] + m[4 + c] * vin[1] + m[8 + c] * vin[2] + m[12 + c] * vin[3]) + k.g[20 + c], 0.0f, 1.0f);
      color[0] = r[0]; color[1] = r[1]; color[2] = r[2]; alpha = r[3];
    }
  }
}

struct BlendShader {
  struct Row {
    float o[2], step[2];
    float pd;
    float base[4][2];
    int kb;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    r.pd = (1.0f - k.f[5]) * k.f[4] + k.f[5];
    r.kb = wr_chunk_base<2>(a, r.o, r.step, c, tx0, r.base);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    int rel = x - c.x0;
    float uv[2];
    wr_chunk_lane<2>(a, r.base, r.step, r.kb, rel >> 2, rel & 3, uv);
    float Cs[4];
    float pd = r.pd;
    if (a.persp) {  // gl_FragCoord.w varies per sample
      const float fw = wr_persp_zw(*a.persp, 1, rel);
      pd = (1.0f - fw) * k.f[4] + fw;
    }
    wr_tex_fragment(a.color0, wr_clamp(uv[0] * pd, k.f[0], k.f[2]), wr_clamp(uv[1] * pd, k.f[1], k.f[3]), Cs);
    float alpha = Cs[3];
    float color[3];
    for (int i = 0; i < 3; i++) color[i] = alpha != 0.0f ? Cs[i] / alpha : Cs[i];
    wr_calculate_filter(a, k, color, alpha);
    Px o;
    o.r = wr_round_pixel(alpha * color[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(alpha * color[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(alpha * color[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(alpha * 1.0f, 255.0f) & 0xFFFF;
    return o;
  }
};

// SetupFilterParams (blend.glsl:26-91); m is column major
WRD void wr_setup_filter_params(const FrameTablesDev& T, int op, float amount, int addr, float* m, float* offset,
                                int* table_address) {
  float lumR = 0.2126f, lumG = 0.7152f, lumB = 0.0722f;
  float oR = 1.0f - lumR, oG = 1.0f - lumG, oB = 1.0f - lumB;
  float inv = 1.0f - amount;
#define WR_COL(i, a_, b_, c_, d_) m[4 * i] = (a_); m[4 * i + 1] = (b_); m[4 * i + 2] = (c_); m[4 * i + 3] = (d_);
  if (op == 1) {
    WR_COL(0, lumR + oR * inv, lumR - lumR * inv, lumR - lumR * inv, 0.0f)
    WR_COL(1, lumG - lumG * inv, lumG + oG * inv, lumG - lumG * inv, 0.0f)
    WR_COL(2, lumB - lumB * inv, lumB - lumB * inv, lumB + oB * inv, 0.0f)
    WR_COL(3, 0.0f, 0.0f, 0.0f, 1.0f)
  } else if (op == 2) {
    // the reference calls libm's cosf/sinf on the host CPU; the correctly rounded
    // double-precision result is the closest the device can get to it
#ifdef WRCU_HOSTEMU
    float cc = cosf(amount), ss = sinf(amount);
#else
    float cc = (float)cos((double)amount), ss = (float)sin((double)amount);
#endif
    WR_COL(0, lumR + oR * cc - lumR * ss, lumR - lumR * cc + 0.143f * ss, lumR - lumR * cc - oR * ss, 0.0f)
    WR_COL(1, lumG - lumG * cc - lumG * ss, lumG + oG * cc + 0.140f * ss, lumG - lumG * cc + lumG * ss, 0.0f)
    WR_COL(2, lumB - lumB * cc + oB * ss, lumB - lumB * cc - 0.283f * ss, lumB + oB * cc + lumB * ss, 0.0f)
    WR_COL(3, 0.0f, 0.0f, 0.0f, 1.0f)
  } else if (op == 4) {
    WR_COL(0, inv * lumR + amount, inv * lumR, inv * lumR, 0.0f)
    WR_COL(1, inv * lumG, inv * lumG + amount, inv * lumG, 0.0f)
    WR_COL(2, inv * lumB, inv * lumB, inv * lumB + amount, 0.0f)
    WR_COL(3, 0.0f, 0.0f, 0.0f, 1.0f)
  } else if (op == 5) {
    WR_COL(0, 0.393f + 0.607f * inv, 0.349f - 0.349f * inv, 0.272f - 0.272f * inv, 0.0f)
    WR_COL(1, 0.769f - 0.769f * inv, 0.686f + 0.314f * inv, 0.534f - 0.534f * inv, 0.0f)
    WR_COL(2, 0.189f - 0.189f * inv, 0.168f - 0.168f * inv, 0.131f + 0.869f * inv, 0.0f)
    WR_COL(3, 0.0f, 0.0f, 0.0f, 1.0f)
  } else if (op == 7) {
    for (int c = 0; c < 4; c++) {
      float4 v = wr_fetch(T.gpu_cache, T.n_gpu_cache, addr + c);
      m[4 * c] = v.x; m[4 * c + 1] = v.y; m[4 * c + 2] = v.z; m[4 * c + 3] = v.w;
    }
    float4 o = wr_fetch(T.gpu_cache, T.n_gpu_cache, addr + 4);
    offset[0] = o.x; offset[1] = o.y; offset[2] = o.z; offset[3] = o.w;
  } else if (op == 11) {
    *table_address = addr;
  } else if (op == 10) {
    float4 o = wr_fetch(T.gpu_cache, T.n_gpu_cache, addr);
    offset[0] = o.x; offset[1] = o.y; offset[2] = o.z; offset[3] = o.w;
  }
#undef WR_COL
}

// brush_blend vertex stage (brush_blend.glsl:43-89)
WRD void wr_setup_brush_blend_one(const SetupArgs& a, int idx) {
  int4 aData = *(const int4*)(a.instances + (size_t)idx * a.stride);
  QuadOut q;
  BrushVS vs;
  memset(&q, 0, sizeof q);
  wr_brush_vertex(a, aData, 3, q, vs);
  const FrameTablesDev& T = a.tabs;
  int src = vs.ph.user_data[0];
  float4 r0 = wr_fetch(T.gpu_cache, T.n_gpu_cache, src);
  float itw = 1.0f / (float)a.color0.w, ith = 1.0f / (float)a.color0.h;
  const float* lr = vs.ph.lr;
  float persp = (vs.brush_flags & 1) ? 1.0f : 0.0f;
  for (int k = 0; k < 4; k++) {
    float fx = (vs.local_pos[k].x - lr[0]) / (lr[2] - lr[0]);
    float fy = (vs.local_pos[k].y - lr[1]) / (lr[3] - lr[1]);
    wr_image_quad_uv(T, src, fx, fy);
    float ux = (r0.z - r0.x) * fx + r0.x, uy = (r0.w - r0.y) * fy + r0.y;
    float mm = (1.0f - vs.world_pos[k].w) * persp + vs.world_pos[k].w;
    q.interp[k][0] = ux * itw * mm;
    q.interp[k][1] = uy * ith * mm;
  }
  q.n_interp = 2;
  q.flags |= CMD_TEXTURED;
  float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  wr_pack_color(q, white);
  float amount = (float)vs.ph.user_data[2] / 65536.0f;
  int mode = vs.ph.user_data[1];
  int op = mode & 0xffff;
  float m[16], offset[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int i = 0; i < 16; i++) m[i] = 0.0f;
  int table_address = 0;
  wr_setup_filter_params(T, op, amount, vs.ph.user_data[2], m, offset, &table_address);
  float fw = 1.0f / q.pos[0].w;
  if (!isfinite(fw)) fw = 0.0f;
  int unsupported = 0;
  bool ok = wr_emit_quad(a, idx, q, &unsupported);
  if (ok) {
    CmdCold* k = &a.cold[idx];
    k->f[0] = (r0.x + 0.5f) * itw; k->f[1] = (r0.y + 0.5f) * ith;
    k->f[2] = (r0.z - 0.5f) * itw; k->f[3] = (r0.w - 0.5f) * ith;
    k->f[4] = persp; k->f[5] = fw; k->f[6] = amount;
    k->i[0] = op; k->i[1] = table_address;
    for (int i = 0; i < 16; i++) k->g[i] = m[i];
    k->g[16] = (float)((mode >> 28) & 0xf); k->g[17] = (float)((mode >> 24) & 0xf);
    k->g[18] = (float)((mode >> 20) & 0xf); k->g[19] = (float)((mode >> 16) & 0xf);
    for (int i = 0; i < 4; i++) k->g[20 + i] = offset[i];
  }
  wr_finish_setup(a, unsupported);
}
WR_SETUP_KERNEL(wr_setup_brush_blend)

// no span shader: every chunk of a run goes through run()
template <> struct WrRun<BlendShader> {
  enum { n = 2 };
  WRD_MEMBER int drawn(const BlendShader::Row&) { return 0; }
};
