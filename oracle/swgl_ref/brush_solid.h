// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "brush_solid" and
// "brush_solid ALPHA_PASS" (webrender/res/brush_solid.glsl).  Under SWGL the
// two variants compile to the same code: antialias_brush() and do_clip() are
// constant 1.0 (brush.glsl:226-232, prim_shared.glsl:218-221).
#pragma once

template <int VARIANT>
struct brush_solid_vert_t : BrushVertBase<brush_solid_vert_t<VARIANT>> {
  typedef brush_solid_vert_t Self;
  static const int VECS_PER_SPECIFIC_BRUSH = 1;
  typedef typename PrimVertBase::VertexInfo VertexInfo;
  typedef WrCommon::RectWithEndpoint RectWithEndpoint;
  typedef WrCommon::PictureTask PictureTask;

  vec4_scalar v_color;
  struct InterpOutputs {};

  brush_solid_vert_t() { this->init_vertex_abi(); }

  // brush_solid.glsl:22-38
  void brush_vs(VertexInfo&, int prim_address, RectWithEndpoint, RectWithEndpoint,
                ivec4_scalar prim_user_data, int, mat4_scalar, PictureTask&, int, vec4_scalar) {
    vec4_scalar color = this->fetch_from_gpu_cache_1(prim_address);
    float opacity = float(prim_user_data.x) / 65535.0f;
    v_color = color * opacity;
  }

  ALWAYS_INLINE void store_interp_outputs(char*, size_t) {}
  using PrimVertBase::load_attribs;
  WR_VERTEX_ABI(brush_solid)
};

template <int VARIANT>
struct brush_solid_frag_t : FragmentShaderImpl, brush_solid_vert_t<VARIANT> {
  typedef brush_solid_frag_t Self;
  typedef typename brush_solid_vert_t<VARIANT>::InterpOutputs InterpOutputs;

  static void read_interp_inputs(FragmentShaderImpl*, const void*, const void*) {}
  ALWAYS_INLINE void step_interp_inputs(int = 4) {}

  // brush_solid.glsl:42-48 + brush.glsl main
  void main() {
    vec4_scalar color = this->v_color;
    if (VARIANT == 1) color *= 1.0f;  // antialias_brush(), then do_clip()
    this->gl_FragColor = color;
  }
  void swgl_drawSpanRGBA8() { swgl_commitSolidRGBA8(this->v_color); }
  void swgl_drawSpanR8() { swgl_commitSolidR8(this->v_color.x); }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  static int draw_span_R8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, R8);
  }
  WR_FRAGMENT_ABI()
  brush_solid_frag_t() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
    this->draw_span_R8_func = &draw_span_R8;
  }
};

typedef brush_solid_frag_t<0> brush_solid_frag;
typedef brush_solid_frag_t<1> brush_solid_ALPHA_PASS_frag;
WR_PROGRAM(brush_solid, "brush_solid")
WR_PROGRAM(brush_solid_ALPHA_PASS, "brush_solid ALPHA_PASS")
