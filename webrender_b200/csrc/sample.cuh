// sample.cuh — texture sampling with the reference's arithmetic
// (swgl/src/texture.h): 7-bit bilinear fractions, int16 lerps, clamped rows.
// Samplers read linear HBM buffers through the read-only path; no CUDA texture
// units (their filtering arithmetic differs from SWGL's).
#pragma once
#include "blend.cuh"
#include "wrcu_internal.h"

WRD int wr_clamp_coord(int coord, int limit) {  // texture.h:73-75
  return min(max(coord, 0), limit - 1);
}

WRD int wr_lerp7(int a, int b, int f) {
  // a + (((b - a) * f) >> 7) in int16 lanes (texture.h:493-497)
  return (int)(short)(a + (int)(short)(((int)(short)((b - a) * f)) >> 7));
}

// textureLinearUnpackedRGBA8 for one lane (texture.h:1027-1075): (ix,iy) is the
// coordinate quantised to 1/128 texel (linearQuantize, texture.h:427-431).
WRD Px wr_texture_linear_rgba8(const TexView& t, int ix, int iy) {
  int x = ix >> 7, y = iy >> 7;
  int cx = wr_clamp_coord(x, t.w - 1);
  int cy = wr_clamp_coord(y, t.h);
  const uint8_t* row0 = t.ptr + (size_t)cy * t.pitch + (size_t)cx * 4;
  const uint8_t* row1 = row0 + ((y >= 0 && y < t.h - 1) ? t.pitch : 0);
  int overread = x > t.w - 2 ? -1 : 0;  // computeFracX, texture.h:468-471
  int fx = (int)(short)((((ix & (x >= 0 ? -1 : 0)) | overread) & 0x7F) - overread);
  int fy = iy & 0x7F;
  uint2 a = make_uint2(__ldg((const uint32_t*)row0), __ldg((const uint32_t*)row0 + 1));
  uint2 b = make_uint2(__ldg((const uint32_t*)row1), __ldg((const uint32_t*)row1 + 1));
  int v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int a0 = (a.x >> (8 * k)) & 0xFF, a1 = (b.x >> (8 * k)) & 0xFF;
    int b0 = (a.y >> (8 * k)) & 0xFF, b1 = (b.y >> (8 * k)) & 0xFF;
    int l = wr_lerp7(a0, a1, fy);
    int r = wr_lerp7(b0, b1, fy);
    v[k] = wr_lerp7(l, r, fx) & 0xFFFF;
  }
  return Px{v[0], v[1], v[2], v[3]};
}

// textureLinearUnpackedR8 for one lane (texture.h:542-574)
WRD int wr_texture_linear_r8(const TexView& t, int ix, int iy) {
  int x = ix >> 7, y = iy >> 7;
  int cx = wr_clamp_coord(x, t.w - 1);
  int cy = wr_clamp_coord(y, t.h);
  const uint8_t* row0 = t.ptr + (size_t)cy * t.pitch + (size_t)cx;
  const uint8_t* row1 = row0 + ((y >= 0 && y < t.h - 1) ? t.pitch : 0);
  int overread = x > t.w - 2 ? -1 : 0;
  int fx = (int)(short)((((ix & (x >= 0 ? -1 : 0)) | overread) & 0x7F) - overread);
  int fy = iy & 0x7F;
  int a0 = __ldg(row0), a1 = __ldg(row1), b0 = __ldg(row0 + 1), b1 = __ldg(row1 + 1);
  int l = wr_lerp7(a0, a1, fy);
  int r = wr_lerp7(b0, b1, fy);
  return wr_lerp7(l, r, fx) & 0xFFFF;
}

// linearQuantize(P, 128, sampler) for one axis: (uv * size) * 128 + (0.5 - 64)
WRD float wr_linear_quantize(float uv, int size) {
  return __fadd_rn(__fmul_rn(__fmul_rn(uv, (float)size), 128.0f), 0.5f - 0.5f * 128.0f);
}
