// shader_quad_mask.cuh — ps_quad_mask [FAST_PATH] (webrender/res/ps_quad_mask.glsl):
// rounded-rect coverage multiplied into an off-screen colour/alpha target by
// handle_clips (renderer/mod.rs:2278).  The program has no span shader, so SWGL
// runs ps_quad.glsl's fragment main for every 4-pixel chunk; the varying
// vClipLocalPos advances by one interp_step per chunk.
#pragma once
#include "shader_clip_rect.cuh"

struct QuadMaskShader {
  struct Row {
    const float* g;
    float base[4][4], step[4];
    int kb;
  };
  WRD_MEMBER void row_setup(const RasterArgs& a, const CmdHot& c, int y, int tx0, bool, Row& r) {
    const CmdCold& k = a.cold[c.cold];
    r.g = k.g;
    float o[4];
    wr_row_interp<4>(a, k, c, y, o, r.step);
    r.kb = wr_chunk_base<4>(a, o, r.step, c, tx0, r.base);
  }
  WRD_MEMBER Px source(const RasterArgs& a, const CmdHot& c, const Row& r, int x, int, bool) {
    const float* g = r.g;
    int rel = x - c.x0, kc = rel >> 2, j = rel & 3;
    float L0[4], L1[4], Lj[4];
    wr_chunk_lane<4>(a, r.base, r.step, r.kb, kc, 0, L0);
    wr_chunk_lane<4>(a, r.base, r.step, r.kb, kc, 1, L1);
    wr_chunk_lane<4>(a, r.base, r.step, r.kb, kc, j, Lj);
    float p0x = L0[0] / L0[3], p0y = L0[1] / L0[3];
    float p1x = L1[0] / L1[3], p1y = L1[1] / L1[3];
    float aa_range = 1.0f / (fabsf(p1x - p0x) + fabsf(p1y - p0y));
    float px = Lj[0] / Lj[3], py = Lj[1] / Lj[3];
    bool fast = g[CR_FAST] != 0.0f;
    float dist = fast ? cr_sd_rounded_box(px, py, g + CR_PARAMS) : cr_distance_to_rounded_rect(g, px, py);
    float alpha = cr_distance_aa(aa_range, dist);
    float fa = cr_mix(alpha, 1.0f - alpha, g[CR_MODE]);
    int v = wr_round_pixel(fa, 255.0f) & 0xFFFF;
    return Px{v, v, v, v};
  }
};
