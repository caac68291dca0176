// shader_text.cuh — ps_text_run (webrender/res/ps_text_run.glsl:278-337): glyph
// blit from the glyph atlas (R8 coverage expanded to RGBA8, or RGBA8 colour /
// subpixel glyphs) modulated by the text colour.
// CmdCold: f[0..3] v_uv_bounds; g[0..3] v_color, g[4..6] v_mask_swizzle, g[7] DUAL variant
#pragma once
#include "raster.cuh"

struct TextShader {
  struct Row {
    float o[2], step[2];
    TexRow tr;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool rgba, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    wr_row_interp<2>(a, k, c, y, r.o, r.step);
    int len = c.x1 - c.x0;
    bool span_ok = k.g[4] == 0.0f || k.g[4] == 1.0f;  // swgl_drawSpanRGBA8 guard on v_mask_swizzle.x
    int body_len = (rgba && len >= 4 && span_ok) ? (len & ~3) : 0;
    float u[4], v[4];
    for (int j = 0; j < 4; j++) {
      float uv[2];
      wr_interp_at<2>(a, r.o, r.step, j, uv);
      u[j] = uv[0];
      v[j] = uv[1];
    }
    int tile_rel = max(tx0, (int)c.x0) - (int)c.x0;
    if (k.g[7] == 0.0f && a.color0.fmt == WRCU_FMT_R8) wr_tex_row_setup_r8(a.color0, k.f, body_len, u, v, tile_rel, r.tr);
    else wr_tex_row_setup(a.color0, k.f, false, body_len, u, v, tile_rel, r.tr);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const CmdCold& k = a.cold[c.cold];
    const TexView& t = a.color0;
    int rel = x - c.x0;
    if (rel < r.tr.body_len) {
      Px col{c.col[0], c.col[1], c.col[2], c.col[3]};  // 255 lanes for the DUAL variant (NoColor)
      return px_apply_color(wr_tex_body(t, r.tr, rel), col);
    }
    float uv[2];
    wr_interp_at<2>(a, r.o, r.step, rel, uv);
    float mask[4], col[4];
    wr_tex_fragment(t, wr_clamp(uv[0], k.f[0], k.f[2]), wr_clamp(uv[1], k.f[1], k.f[3]), mask);
    if (k.g[6] != 0.0f) mask[1] = mask[2] = mask[3] = mask[0];
    if (k.g[7] == 0.0f) {
      float al = mask[3];
      for (int i = 0; i < 3; i++) mask[i] = mask[i] * k.g[4] + al * k.g[5];
    }
    for (int i = 0; i < 4; i++) col[i] = k.g[i] * mask[i] * 1.0f;
    Px o;
    o.r = wr_round_pixel(col[0], 255.0f) & 0xFFFF;
    o.g = wr_round_pixel(col[1], 255.0f) & 0xFFFF;
    o.b = wr_round_pixel(col[2], 255.0f) & 0xFFFF;
    o.a = wr_round_pixel(col[3], 255.0f) & 0xFFFF;
    return o;
  }
};

#ifndef WRCU_HOSTEMU
template <> struct WrMinCtas<TextShader> { enum { v = 3 }; };
template <> struct WrNarrowSpans<TextShader> { enum { v = 1 }; };  // glyph rows: ~12 pixels of a 128-pixel warp row
#endif

#ifndef WRCU_HOSTEMU
// ---- glyph-major text -----------------------------------------------------------------------------
// A text run is thousands of ~10x10-pixel quads: in the tile kernel every (glyph, row, tile) is a warp-wide
// step with a third of the lanes busy, and the glyphs of a tile are a chain of dependent fetches (command ->
// row table -> texels) taken one after the other although they touch different pixels.  Here a WARP takes a
// GLYPH: its lanes are the glyph's pixels — 32/width rows at a time — each running the same row set-up and
// fragment code as the tile kernel on its own (row, pixel); thousands of glyphs are in flight at once.
//
// Order matters where glyphs overlap (text blending is not commutative bit for bit).  A warp first finds the
// EARLIER commands of the batch whose rect meets its glyph's (tile bins, or for small batches a scan of the hot
// records) and waits until each is RESOLVED: drawn here (CMD_DONE, set with release semantics after the pixels)
// or left to the ordered tile kernel (CMD_ORDERED).  Glyphs are handed out by a ticket counter, so every earlier
// glyph is held by a warp that is already running and the wait cannot deadlock.  A glyph this kernel cannot draw —
// transformed (CMD_GENERAL / CMD_CLIP_DIST), failing depth samples (depth runs), a box over 64 pixels — or that
// overlaps an earlier ORDERED one is itself flagged CMD_ORDERED; the tile kernel afterwards draws exactly those, in
// batch order (RasterArgs::glyph_major) — nothing when the count (BatchInfo::n_ordered) is zero.
#define WR_GLYPH_THREADS 128
__device__ __forceinline__ uint32_t wr_ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <int FMT>
__global__ void __launch_bounds__(WR_GLYPH_THREADS, 4) wr_raster_glyphs(RasterArgs a) {
  wr_pdl_launch_dependents();
  wr_pdl_wait();  // commands and bins (the set-up launch), the pixels under the glyphs (whatever drew last)
  const int lane = threadIdx.x & 31;
  CmdHot* const hot = const_cast<CmdHot*>(a.hot);
  GenRow grow;
  a.gen = &grow;
  a.persp = nullptr;
  a.rl = nullptr;
  a.lane_rows = 1;
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(const_cast<int*>(&a.info->glyph_ticket), 1);
    i = __shfl_sync(0xFFFFFFFFu, i, 0);
    if (i >= a.n) return;
    const CmdHot c = a.hot[i];
    if (c.x1 <= c.x0 || c.y1 <= c.y0) continue;  // (never waited for: an empty rect meets nothing)
    const CmdCold& k = a.cold[c.cold];
    const int W = (int)c.x1 - (int)c.x0, H = (int)c.y1 - (int)c.y0;
    bool ordered = (c.flags & (CMD_GENERAL | CMD_CLIP_DIST | CMD_PERSP)) != 0 || W > 64 || H > 64;
    if (!ordered && a.depth_mode != WRCU_DEPTH_OFF) {
      // depth-tested (never written by text): drawable here when every sample of every row passes
      if (a.depth_mode != WRCU_DEPTH_TEST || !a.fail_pool || k.fail_off < 0) ordered = true;
      else {
        bool f = false;
        for (int r = lane; r < H; r += 32) f = f || __ldg(a.fail_pool + (size_t)k.fail_off + (size_t)r * (k.fail_w + 1)) != 0u;
        ordered = __any_sync(0xFFFFFFFFu, f);
      }
    }
    const int tx_a = max((int)c.x0, 0) / WRCU_TILE_W, tx_b = (min((int)c.x1, a.tgt.w) + WRCU_TILE_W - 1) / WRCU_TILE_W;
    const int ty_a = max((int)c.y0, 0) / WRCU_TILE_H, ty_b = (min((int)c.y1, a.tgt.h) + WRCU_TILE_H - 1) / WRCU_TILE_H;
    // earlier commands whose rect meets this one's: wait for each to be resolved
    bool after_ordered = false;
    auto wait_if_met = [&](int j) {
      const short4 o = *(const short4*)&a.hot[j];  // x0, y0, x1, y1 (written by the set-up launch only)
      if (!(o.z > o.x && o.w > o.y && o.x < c.x1 && o.z > c.x0 && o.y < c.y1 && o.w > c.y0)) return;
      uint32_t f = 0;
      for (int spin = 0; spin < (1 << 22); spin++) {  // (bounded: a lost flag must not hang the GPU)
        f = wr_ld_acquire(&a.hot[j].flags);
        if (f & (CMD_DONE | CMD_ORDERED)) break;
        __nanosleep(64);
      }
      if (!(f & CMD_DONE)) after_ordered = true;
    };
    if (a.tile_mask) {
      const int wend = i >> 5;
      for (int ty = ty_a; ty < ty_b; ty++)
        for (int tx = tx_a; tx < tx_b; tx++) {
          const uint32_t* tm = a.tile_mask + (size_t)(ty * a.bin_tiles_x + tx) * a.bin_words;
          for (int wi = lane; wi <= wend; wi += 32) {
            uint32_t word = __ldg(tm + wi) | __ldg(a.wide_mask + wi);
            if (wi == wend) word &= (1u << (i & 31)) - 1u;
            while (word) {
              const int b = __ffs((int)word) - 1;
              word &= word - 1;
              wait_if_met(wi * 32 + b);
            }
          }
        }
    } else {
      for (int j = lane; j < i; j += 32) wait_if_met(j);
    }
    ordered = ordered || __any_sync(0xFFFFFFFFu, after_ordered);  // (the barrier also orders the lanes' acquires before the pixel reads)
    if (ordered) {
      if (a.tile_ord) {
        const int nx = tx_b - tx_a, nt = nx * (ty_b - ty_a);
        for (int t = lane; t < nt; t += 32) {
          const int tid = (ty_a + t / nx) * a.bin_tiles_x + tx_a + t % nx;
          atomicOr(&a.tile_ord[tid >> 5], 1u << (tid & 31));
        }
      }
      __syncwarp();
      if (lane == 0) {
        atomicAdd(const_cast<int*>(&a.info->n_ordered), 1);
        __threadfence();
        atomicOr(&hot[i].flags, (uint32_t)CMD_ORDERED);
      }
      continue;
    }
    // ---- draw: lanes = pixels, 32 / wpad rows per step.  Destination pixels are read through L2 (__ldcg):
    // a neighbouring glyph of another SM may have written the same 128-byte line a moment ago.
    const int wpad = W <= 4 ? 4 : (W <= 8 ? 8 : (W <= 16 ? 16 : 32)), rpp = 32 / wpad;
    const int lx = lane & (wpad - 1), ly = lane / wpad;
    for (int xb = 0; xb < W; xb += 32)
      for (int yb = 0; yb < H; yb += rpp) {
        const int x = (int)c.x0 + xb + lx, y = (int)c.y0 + yb + ly;
        if (x >= c.x1 || y >= c.y1) continue;
        TextShader::Row row;
        TextShader::row_setup(a, c, y, (int)c.x0, FMT == WRCU_FMT_RGBA8, row);
        uint32_t px, zb = 0;
        bool dirty = false, zdirty = false;
        if (FMT == WRCU_FMT_RGBA8) {
          uint32_t* dp = (uint32_t*)(a.tgt.color + (size_t)y * a.tgt.color_pitch) + x;
          px = __ldcg(dp);
          wr_shade_pixel<TextShader, FMT>(a, c, row, x, y, false, px, zb, dirty, zdirty);
          if (dirty) *dp = px;
        } else {
          uint8_t* dp = a.tgt.color + (size_t)y * a.tgt.color_pitch + x;
          px = __ldcg(dp);
          wr_shade_pixel<TextShader, FMT>(a, c, row, x, y, false, px, zb, dirty, zdirty);
          if (dirty) *dp = (uint8_t)px;
        }
      }
    __syncwarp();
    if (lane == 0) {
      __threadfence();  // the glyph's pixels before its flag
      atomicOr(&hot[i].flags, (uint32_t)CMD_DONE);
    }
  }
}
#endif

template <> struct WrRun<TextShader> {
  enum { n = 2 };
  WRD_MEMBER int drawn(const TextShader::Row& r) { return r.tr.body_len; }
};
