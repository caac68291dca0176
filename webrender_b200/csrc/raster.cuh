// raster.cuh — tile-resident raster kernels.
//
// The reference draws a batch instance by instance: for each instance every
// row, every span, every 4-pixel chunk is read-modify-written in the
// framebuffer (swgl/src/rasterize.h:783-1054, blend.h:416).  On a GPU that
// shape would serialise on memory: N overlapping instances = N passes over
// the same bytes.  Here the loop nest is inverted: one CTA owns a 128x8-pixel
// tile of the render target, keeps its pixels in registers (4 pixels/thread),
// streams the batch's command list through shared memory IN BATCH ORDER (order
// matters for blending), applies fragment + blend stage per pixel, and writes
// the tile once.  DRAM sees each target byte at most once in and once out per
// batch, however many layers the batch stacks.
//
// Thread mapping: 256 threads = 8 warps; warp w owns tile row w (128 px =
// 512 B contiguous for RGBA8 → one fully coalesced 16-byte vector access per
// lane); lane l owns pixels [4l, 4l+4).
#pragma once
#include "blend.cuh"
#include "cmd.cuh"
#include "sample.cuh"
#include "texspan.cuh"
#include "wrcu_internal.h"

struct RasterArgs {
  TargetDev tgt;
  const CmdHot* hot;
  const CmdCold* cold;
  const BatchInfo* info;
  int n;
  int blend;        // wrcu_blend
  int depth_mode;   // wrcu_depth
  Px blend_color;   // glBlendColor in lane order
  TexView color0;   // sColor0
  TexView color1;   // sColor1 (brush_mix_blend source)
  int fast_eligible;  // host-side part of the solid-premult fast-path test
  const float4* gbuf_f;  // gpu_buffer_f (gradient LUTs)
  int n_gbuf_f;
  const float4* gpu_cache;  // component-transfer tables
  int n_gpu_cache;
};

#define CHUNK_CMDS 256

// AA weight of pixel x for the current command (DO_AA, blend.h:433-445, with
// the span set-up of aa_span, rasterize.h:546-557).  Chunks of 4 start at the
// span start c.x0.
WRD int wr_aa_weight(const CmdHot& c, const CmdCold& k, int x) {
  int j = (x - c.x0) & 3;
  int xc = x - j;
  int opaque = max((int)c.aa_right_start - (int)c.aa_left_end - 3, 0);
  int off = xc - c.aa_left_end;
  if (off >= 0 && off < opaque) return 256;
  float offs = (float)(c.aa_left_end + j);
  float left = __fadd_rn(k.aa_l0, __fmul_rn(offs, k.aa_ls));
  float right = __fadd_rn(k.aa_r0, __fmul_rn(offs, k.aa_rs));
  float fo = (float)off;
  float dist = wr_clamp(wr_min(__fadd_rn(left, __fmul_rn(k.aa_ls, fo)),
                               __fadd_rn(right, __fmul_rn(k.aa_rs, fo))),
                        0.0f, 256.0f);
  return wr_round_pixel(dist, 1.0f);
}

// Interpolants of a screen-axis-aligned quad.  The reference walks the left and
// right edges row by row, adding the per-row slope each time (Edge::nextRow,
// rasterize.h:880-884), then derives the span's start value and per-pixel step
// (rasterize.h:1003-1017).  wr_row_interp reproduces that running sum exactly:
// it starts from the Edge constructor's value at the first row and adds the
// slope (y - y0) times — warp-uniform work, once per (command,row).
template <int N>
WRD void wr_row_interp(const CmdCold& k, const CmdHot& c, int y, float* o, float* step) {
  float y0c = (float)c.y0 + 0.5f;
  float dy = __fsub_rn(y0c, k.yt);
  float stepScale = __fdiv_rn(1.0f, __fsub_rn(k.xr, k.xl));
  if (!isfinite(stepScale)) stepScale = 0.0f;
  float x0f = __fsub_rn(__fadd_rn((float)c.x0, 0.5f), k.xl);
  int rows = y - c.y0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    float sl = __fmul_rn(__fsub_rn(k.i_lb[i], k.i_lt[i]), k.yscale);
    float sr = __fmul_rn(__fsub_rn(k.i_rb[i], k.i_rt[i]), k.yscale);
    float li = __fadd_rn(k.i_lt[i], __fmul_rn(dy, sl));
    float ri = __fadd_rn(k.i_rt[i], __fmul_rn(dy, sr));
    for (int r = 0; r < rows; r++) {
      li = __fadd_rn(li, sl);
      ri = __fadd_rn(ri, sr);
    }
    float st = __fmul_rn(__fsub_rn(ri, li), stepScale);
    step[i] = st;
    o[i] = __fadd_rn(li, __fmul_rn(st, x0f));
  }
}

// Value of interpolant lanes at pixel x of the span: lane j of chunk k.  Chunk
// offset accumulates sequentially as init_interp does (glsl.h:3083-3088), then
// the chunk advance is one multiply-add (step_interp_inputs(drawn) after a
// span body; exact for the 1:1 mappings that dominate).
template <int N>
WRD void wr_interp_at(const float* o, const float* step, int rel, float* out) {
  int j = rel & 3;
  float kf = (float)(rel >> 2);
#pragma unroll
  for (int i = 0; i < N; i++) {
    float v = o[i];
    for (int s = 0; s < j; s++) v = __fadd_rn(v, step[i]);         // init_interp lanes first,
    v = __fadd_rn(v, __fmul_rn(__fmul_rn(step[i], 4.0f), kf));     // then interp_step * chunks
    out[i] = v;
  }
}

// Fragment-path interpolants: the reference advances varyings once per 4-pixel
// chunk (v += interp_step, glsl-to-cxx step_interp_inputs), a running sum.
// wr_chunk_base walks that sum to the first chunk a tile touches (warp-uniform,
// once per command/row/tile); wr_chunk_lane finishes the walk for one pixel
// (at most 32 more additions inside a 128-pixel tile).
template <int N>
WRD int wr_chunk_base(const float* o, const float* step, const CmdHot& c, int tx0, float (*base)[N]) {
  int kb = max(0, (max(tx0, (int)c.x0) - (int)c.x0) >> 2);
#pragma unroll
  for (int i = 0; i < N; i++) {
    float is = __fmul_rn(step[i], 4.0f);
    float v = o[i];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      // init_interp: lane j of chunk 0 = lane j-1 + step (glsl.h:3083-3088);
      // every later chunk adds interp_step to each lane
      float l = v;
      for (int s = 0; s < kb; s++) l = __fadd_rn(l, is);
      base[j][i] = l;
      v = __fadd_rn(v, step[i]);
    }
  }
  return kb;
}
// lane j (0..3) of chunk k >= kb, given the lanes of chunk kb in base
template <int N>
WRD void wr_chunk_lane(const float (*base)[N], const float* step, int kb, int k, int j, float* out) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    float is = __fmul_rn(step[i], 4.0f);
    float v = base[j][i];
    for (int s = kb; s < k; s++) v = __fadd_rn(v, is);
    out[i] = v;
  }
}

// ---- fragment stage: ps_quad_textured -------------------------------------------
// (webrender/res/ps_quad_textured.glsl:39-64, ps_quad.glsl:406-417).  The first
// len&~3 pixels of a span are drawn by swgl_drawSpanRGBA8 (solid commit or
// swgl_commitTextureLinearColorRGBA8), the rest by the fragment shader.
struct QuadShader {
  struct Row {
    float o[2], step[2];
    TexRow tr;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    r.tr.mode = TEX_NONE;
    r.tr.body_len = 0;
    if (!(c.flags & CMD_TEXTURED)) return;
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(k, c, y, r.o, r.step);
    int len = c.x1 - c.x0;
    int body_len = (rgba && len >= 4 && !(c.flags & CMD_OUT_RRRR)) ? (len & ~3) : 0;
    float u[4], v[4];
    for (int j = 0; j < 4; j++) {
      float uv[2];
      wr_interp_at<2>(r.o, r.step, j, uv);
      u[j] = uv[0];
      v[j] = uv[1];
    }
    wr_tex_row_setup(a.color0, k.f, false, body_len, u, v, max(tx0, (int)c.x0) - (int)c.x0, r.tr);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool rgba) {
    Px col{c.col[0], c.col[1], c.col[2], c.col[3]};
    const int len = c.x1 - c.x0;
    int rel = x - c.x0;
    if (!(c.flags & CMD_TEXTURED)) {
      // swgl_drawSpanRGBA8 commits v_color for the span body even for mask quads;
      // only the fragment-shader tail applies .rrrr.  R8 targets have no span
      // shader: always the fragment path.
      if (c.flags & CMD_OUT_RRRR) {
        int body_len0 = (rgba && len >= 4) ? (len & ~3) : 0;
        if (rel >= body_len0) col.b = col.g = col.a = col.r;
      }
      return col;
    }
    const TexView& t = a.color0;
    if (rel < r.tr.body_len) return px_apply_color(wr_tex_body(t, r.tr, rel), col);
    // fragment path: fs_sample_color0 (sample_color0.glsl:25-31), v_color == 1
    const CmdCold& k = a.cold[c.cold];
    float uv[2];
    wr_interp_at<2>(r.o, r.step, rel, uv);
    float tex[4];
    wr_tex_fragment(t, wr_clamp(uv[0], k.f[0], k.f[2]), wr_clamp(uv[1], k.f[1], k.f[3]), tex);
    Px o;
    if (c.flags & CMD_OUT_RRRR) {
      int rr = wr_round_pixel(tex[0], 255.0f) & 0xFFFF;
      o = Px{rr, rr, rr, rr};
    } else {
      o.r = wr_round_pixel(tex[0], 255.0f) & 0xFFFF;
      o.g = wr_round_pixel(tex[1], 255.0f) & 0xFFFF;
      o.b = wr_round_pixel(tex[2], 255.0f) & 0xFFFF;
      o.a = wr_round_pixel(tex[3], 255.0f) & 0xFFFF;
    }
    return o;
  }
};

// One pixel of one command through depth test, fragment stage, AA/mask
// modifiers and the blend stage.  Shared by the tile kernel and (tests only) the
// host emulation.  px = packed destination pixel (RGBA8) or value (R8).
template <class S, int FMT>
WRD void wr_shade_pixel(const RasterArgs& a, const CmdHot& c, const typename S::Row& row, int xx, int y,
                        bool use_depth, uint32_t& px, uint32_t& zb, bool& dirty, bool& zdirty) {
  if (use_depth) {
    if (!(c.z <= zb)) return;  // GL_LEQUAL
    if (a.depth_mode == WRCU_DEPTH_TEST_WRITE) { zb = c.z; zdirty = true; }
  }
  Px src = S::source(a, c, row, xx, y, FMT == WRCU_FMT_RGBA8);
  if (a.blend != WRCU_BLEND_NONE) {
    if (c.flags & (CMD_AA | CMD_MASK)) {
      const CmdCold& k = a.cold[c.cold];
      // Order of the two source modifiers: the blend stage applies AA then the
      // clip mask (blend.h:447-461); commit_masked_solid_span (swgl_ext.h:10-24)
      // — whole chunks of solid spans — folds the mask into the colour first.
      int len = c.x1 - c.x0;
      bool mask_first = (c.flags & CMD_SPAN_SOLID) && (xx - c.x0) < (len >= 4 ? (len & ~3) : 0);
      int mk = 255;
      if (c.flags & CMD_MASK) mk = __ldg(k.mask_ptr + (size_t)(y - k.cmy) * k.mask_pitch + (xx - k.cmx));
      if ((c.flags & CMD_MASK) && mask_first) {
        if (FMT == WRCU_FMT_RGBA8) src = px_scale255(src, mk);
        else src.r = wr_muldiv255(src.r, mk);
      }
      if (c.flags & CMD_AA) {
        int aa = wr_aa_weight(c, k, xx);
        if (FMT == WRCU_FMT_RGBA8) src = px_scale256(src, aa);
        else src.r = wr_muldiv256(src.r, aa);
      }
      if ((c.flags & CMD_MASK) && !mask_first) {
        if (FMT == WRCU_FMT_RGBA8) src = px_scale255(src, mk);
        else src.r = wr_muldiv255(src.r, mk);
      }
    }
    if (FMT == WRCU_FMT_RGBA8) {
      int key = a.blend;
      Px kc = a.blend_color;
      if (c.flags & (CMD_DROP_SHADOW | CMD_SUBPIXEL_TEXT)) {  // SWGL_CLIP_FLAG_BLEND_OVERRIDE (rasterize.h:410-413)
        const CmdCold& k = a.cold[c.cold];
        key = (c.flags & CMD_DROP_SHADOW) ? WRCU_BLEND__DROP_SHADOW : WRCU_BLEND__SUBPIXEL_TEXT;
        kc = Px{k.i[0] & 0xFFFF, (k.i[0] >> 16) & 0xFFFF, k.i[1] & 0xFFFF, (k.i[1] >> 16) & 0xFFFF};
      }
      px = px_pack(wr_blend_rgba8(key, src, px_unpack(px), kc));
    }
    else px = wr_pack16(wr_blend_r8(a.blend, src.r, (int)px));
  } else {
    if (FMT == WRCU_FMT_RGBA8) px = px_pack(src);
    else px = wr_pack16(src.r);
  }
  dirty = true;
}

#ifdef WRCU_HOSTEMU
// tests only: the same per-pixel code, pixel by pixel, commands in batch order
template <class S, int FMT>
static void wr_raster(const RasterArgs& a) {
  const BatchInfo bi = *a.info;
  if (a.fast_eligible && bi.simple) return;
  for (int y = 0; y < a.tgt.h; y++) {
    uint8_t* rowp = a.tgt.color + (size_t)y * a.tgt.color_pitch;
    uint32_t* zrow = a.tgt.depth ? (uint32_t*)((uint8_t*)a.tgt.depth + (size_t)y * a.tgt.depth_pitch) : nullptr;
    const bool use_depth = a.depth_mode != WRCU_DEPTH_OFF && zrow != nullptr;
    for (int i = 0; i < a.n; i++) {
      const CmdHot c = a.hot[i];
      if (y < c.y0 || y >= c.y1) continue;
      typename S::Row row;
      for (int tx0 = (c.x0 / WRCU_TILE_W) * WRCU_TILE_W; tx0 < c.x1; tx0 += WRCU_TILE_W) {
      S::row_setup(a, c, y, tx0, FMT == WRCU_FMT_RGBA8, row);
      for (int xx = max((int)c.x0, tx0); xx < min((int)c.x1, tx0 + WRCU_TILE_W); xx++) {
        uint32_t px = FMT == WRCU_FMT_RGBA8 ? ((uint32_t*)rowp)[xx] : rowp[xx];
        uint32_t zb = use_depth ? zrow[xx] : 0;
        bool dirty = false, zdirty = false;
        wr_shade_pixel<S, FMT>(a, c, row, xx, y, use_depth, px, zb, dirty, zdirty);
        if (dirty) { if (FMT == WRCU_FMT_RGBA8) ((uint32_t*)rowp)[xx] = px; else rowp[xx] = (uint8_t)px; }
        if (zdirty) zrow[xx] = zb;
      }
      }
    }
  }
}
static void wr_raster_solid_premult(const RasterArgs& a) {
  const BatchInfo bi = *a.info;
  if (!bi.simple) return;
  for (int i = 0; i < a.n; i++) {
    const CmdHot c = a.hot[i];
    uint32_t srb = (uint32_t)c.col[0] | ((uint32_t)c.col[2] << 16), sga = (uint32_t)c.col[1] | ((uint32_t)c.col[3] << 16);
    uint32_t cc = 255u - c.col[3];
    for (int y = c.y0; y < c.y1; y++) {
      uint32_t* rowp = (uint32_t*)(a.tgt.color + (size_t)y * a.tgt.color_pitch);
      for (int x = c.x0; x < c.x1; x++) {
        uint32_t p = rowp[x];
        uint32_t rb = wr_premult_over_pair(p & 0x00FF00FFu, srb, cc);
        uint32_t ga = wr_premult_over_pair((p >> 8) & 0x00FF00FFu, sga, cc);
        rowp[x] = rb | (ga << 8);
      }
    }
  }
}
#else
// ---- the generic tile kernel (any command kind via the shader policy S, any
// blend key).  S::row_setup computes per-(command,row) constants once per warp
// (all 32 lanes of a warp share the row, so the work is warp-uniform);
// S::source returns the fragment stage's output for one pixel as 16-bit lanes.
template <class S, int FMT>
__global__ void __launch_bounds__(WRCU_THREADS)
wr_raster(RasterArgs a) {
  __shared__ CmdHot sh[CHUNK_CMDS];
  const int tx0 = blockIdx.x * WRCU_TILE_W, ty0 = blockIdx.y * WRCU_TILE_H;
  const BatchInfo bi = *a.info;
  if (a.fast_eligible && bi.simple) return;  // handled by wr_raster_solid_premult
  // whole-CTA early out: tile outside the batch's bounding box
  if (tx0 >= bi.bx1 || tx0 + WRCU_TILE_W <= bi.bx0 || ty0 >= bi.by1 || ty0 + WRCU_TILE_H <= bi.by0)
    return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = tx0 + lane * 4, y = ty0 + warp;
  const bool row_ok = y < a.tgt.h;
  uint32_t px[4] = {0, 0, 0, 0};
  uint32_t zb[4] = {0, 0, 0, 0};
  bool loaded = false, dirty = false, zdirty = false;
  uint8_t* rowp = a.tgt.color + (size_t)y * a.tgt.color_pitch;
  uint32_t* zrow = a.tgt.depth ? (uint32_t*)((uint8_t*)a.tgt.depth + (size_t)y * a.tgt.depth_pitch) : nullptr;
  const bool use_depth = a.depth_mode != WRCU_DEPTH_OFF && zrow != nullptr;

  for (int base = 0; base < a.n; base += CHUNK_CMDS) {
    __syncthreads();
    int m = min(CHUNK_CMDS, a.n - base);
    if (threadIdx.x < m) sh[threadIdx.x] = a.hot[base + threadIdx.x];
    __syncthreads();
    for (int i = 0; i < m; i++) {
      const CmdHot c = sh[i];
      if (!row_ok || y < c.y0 || y >= c.y1) continue;          // warp-uniform
      if (c.x1 <= tx0 || c.x0 >= tx0 + WRCU_TILE_W) continue;  // CTA-uniform
      if (!loaded) {
        // lazy tile load: first command that touches this row
        loaded = true;
        if (FMT == WRCU_FMT_RGBA8) {
          uint4 v = *(const uint4*)(rowp + (size_t)x * 4);
          px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
        } else {
          uint32_t v = *(const uint32_t*)(rowp + x);
          px[0] = v & 0xFF; px[1] = (v >> 8) & 0xFF; px[2] = (v >> 16) & 0xFF; px[3] = v >> 24;
        }
        if (use_depth) {
          uint4 v = *(const uint4*)(zrow + x);
          zb[0] = v.x; zb[1] = v.y; zb[2] = v.z; zb[3] = v.w;
        }
      }
      typename S::Row row;
      S::row_setup(a, c, y, tx0, FMT == WRCU_FMT_RGBA8, row);
#pragma unroll
      for (int p = 0; p < 4; p++) {
        int xx = x + p;
        if (xx < c.x0 || xx >= c.x1) continue;
        wr_shade_pixel<S, FMT>(a, c, row, xx, y, use_depth, px[p], zb[p], dirty, zdirty);
      }
    }
  }
  if (dirty) {
    if (FMT == WRCU_FMT_RGBA8) {
      *(uint4*)(rowp + (size_t)x * 4) = make_uint4(px[0], px[1], px[2], px[3]);
    } else {
      *(uint32_t*)(rowp + x) = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
    }
  }
  if (zdirty) *(uint4*)(zrow + x) = make_uint4(zb[0], zb[1], zb[2], zb[3]);
}

// ---- specialised hot kernel: solid quads, premultiplied-alpha over, no depth,
// no mask/AA (config B, the alpha-blend brush pass).  Pixels stay unpacked as
// (rb, ga) lane pairs across the whole command list; per pixel-layer the work
// is 2 x (IMAD, PRMT, IADD, VMIN).  Commands whose colour lanes exceed 255 or
// that carry mask/AA flags are not routed here (the host checks the batch).
__global__ void __launch_bounds__(WRCU_THREADS)
wr_raster_solid_premult(RasterArgs a) {
  __shared__ CmdHot sh[CHUNK_CMDS];
  const int tx0 = blockIdx.x * WRCU_TILE_W, ty0 = blockIdx.y * WRCU_TILE_H;
  const BatchInfo bi = *a.info;
  if (!bi.simple) return;  // mixed batch → wr_raster_quads
  if (tx0 >= bi.bx1 || tx0 + WRCU_TILE_W <= bi.bx0 || ty0 >= bi.by1 || ty0 + WRCU_TILE_H <= bi.by0)
    return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = tx0 + lane * 4, y = ty0 + warp;
  const bool row_ok = y < a.tgt.h;
  uint8_t* rowp = a.tgt.color + (size_t)y * a.tgt.color_pitch + (size_t)x * 4;
  uint32_t rb[4], ga[4];
  bool dirty = false;
  {
    uint4 v = row_ok ? *(const uint4*)rowp : make_uint4(0, 0, 0, 0);
    uint32_t pv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int p = 0; p < 4; p++) {
      rb[p] = pv[p] & 0x00FF00FFu;
      ga[p] = (pv[p] >> 8) & 0x00FF00FFu;
    }
  }
  for (int base = 0; base < a.n; base += CHUNK_CMDS) {
    __syncthreads();
    int m = min(CHUNK_CMDS, a.n - base);
    if (threadIdx.x < m) sh[threadIdx.x] = a.hot[base + threadIdx.x];
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < m; i++) {
      const uint4 h0 = *(const uint4*)&sh[i];         // rect, flags, z
      const uint2 h1 = *(const uint2*)&sh[i].col[0];  // colour lanes
      int cx0 = (short)(h0.x & 0xFFFF), cy0 = (short)(h0.x >> 16);
      int cx1 = (short)(h0.y & 0xFFFF), cy1 = (short)(h0.y >> 16);
      if (y < cy0 || y >= cy1) continue;
      // h1.x = B | G<<16, h1.y = R | A<<16  → pairs (B,R) and (G,A)
      uint32_t srb = __byte_perm(h1.x, h1.y, 0x5410);  // B | R<<16
      uint32_t sga = __byte_perm(h1.x, h1.y, 0x7632);  // G | A<<16
      uint32_t cc = 255u - (h1.y >> 16);
      if (cx0 <= x && cx1 >= x + 4) {
#pragma unroll
        for (int p = 0; p < 4; p++) {
          rb[p] = wr_premult_over_pair(rb[p], srb, cc);
          ga[p] = wr_premult_over_pair(ga[p], sga, cc);
        }
        dirty = true;
      } else if (cx1 > x && cx0 < x + 4) {
#pragma unroll
        for (int p = 0; p < 4; p++) {
          if (x + p >= cx0 && x + p < cx1) {
            rb[p] = wr_premult_over_pair(rb[p], srb, cc);
            ga[p] = wr_premult_over_pair(ga[p], sga, cc);
          }
        }
        dirty = true;
      }
    }
  }
  if (dirty && row_ok) {
    uint4 v;
    v.x = rb[0] | (ga[0] << 8);
    v.y = rb[1] | (ga[1] << 8);
    v.z = rb[2] | (ga[2] << 8);
    v.w = rb[3] | (ga[3] << 8);
    *(uint4*)rowp = v;
  }
}
#endif  // !WRCU_HOSTEMU
